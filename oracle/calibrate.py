"""TEST INFRASTRUCTURE ONLY.  Records BatchNorm calibration statistics for the synthetic weights:

    python oracle/calibrate.py [seed]     ->  centertrack_b200/data/bn_calib_seed<seed>.npz

Runs the oracle network once on one synthetic 128x160 frame pair with every BN adopting the batch
statistics of its input (DLA34Oracle.calibrate), layer by layer, and stores running_mean/var.  The
product never calls this; it only reads the committed .npz (centertrack_b200/synthetic.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, '..'))

import ct_oracle as co                                 # noqa: E402
from centertrack_b200 import synthetic as syn          # noqa: E402
from centertrack_b200.model import create_model        # noqa: E402
from centertrack_b200.opts import opts                 # noqa: E402


def main(seed=317):
  stats = {}
  for node in ('dcn', 'conv', 'gcn'):                   # the trunk statistics are shared; every node kind adds its own BNs
    opt = opts().init(['tracking', '--pre_hm', '--dla_node', node])
    template = create_model(opt.arch, opt.heads, opt.head_conv, opt=opt).state_dict()
    sd = syn.make_state_dict(template, seed, calibrated=False)
    orc = co.DLA34Oracle(sd, opt.heads, dla_node=node)
    orc.calibrate = True
    img, pre, hm = syn.synthetic_inputs(1, 128, 160, seed=seed + 1)
    orc.feats(img, pre, hm)
    for k in sd:
      if (k.endswith('running_mean') or k.endswith('running_var')) and k not in stats:
        stats[k] = orc.sd[k].numpy()
  os.makedirs(os.path.dirname(syn.calib_path(seed)), exist_ok=True)
  np.savez_compressed(syn.calib_path(seed), **stats)
  print('wrote', syn.calib_path(seed), len(stats), 'arrays')


if __name__ == '__main__':
  main(int(sys.argv[1]) if len(sys.argv) > 1 else 317)
