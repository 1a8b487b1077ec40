"""One eager step of the benchmarked plan (B frames x HxW, bf16, device tracking) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_step \
        python tools/profile_step.py [--config coco_tracking] [--batch 32] [--precision bf16]
It prints the op list in launch order (index, kind, name, engine, a_mode) so that tools/ncu_summary.py can label the
launches of the capture."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
  sys.path.insert(0, p)

from bench import CONFIGS, K                                   # noqa: E402
from centertrack_b200 import synthetic as wt                    # noqa: E402
from centertrack_b200.runner import StreamRunner                # noqa: E402
from helpers import make_model                                  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='coco_tracking')
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--precision', default='bf16')
args = ap.parse_args()
H, W = CONFIGS[args.config][:2]
dev = torch.device('cuda')
opt, model, sd = make_model(args.config)
model = model.to(dev)
runner = StreamRunner(model, args.batch, H, W, K=K, precision=args.precision, device=dev, opt=opt, device_tracking=True,
                      use_graph=False)
img, pre, hm = wt.synthetic_inputs(2, H, W, seed=317)
for s in range(3):
  runner.load_device_inputs(img[s & 1:(s & 1) + 1].expand(args.batch, 3, H, W).contiguous().to(dev), None, s)
for _ in range(3):                 # warm: first-frame path, then two steady-state steps (tracks exist)
  runner.step_device()
torch.cuda.synchronize()
print('OPS')
print(0, 'memset+render', 'ct_render_tracks')
for i, (kind, pl, name) in enumerate(runner.eng.ops):
  print(i + 1, kind, name, getattr(pl, 'engine', ''), getattr(pl, 'a_mode', ''))
print(len(runner.eng.ops) + 1, 'decode', 'ct_decode')
print(len(runner.eng.ops) + 2, 'track', 'ct_track_step')
sys.stdout.flush()
torch.cuda.cudart().cudaProfilerStart()
runner.step_device()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
