"""One eager step (no CUDA graph) of the bf16 hot path for ncu captures:
   ncu --set full --clock-control none --import-source on -k regex:conv_ -s N -c M -o gpurun_out/prof python tools/profile_step.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from centertrack_b200 import synthetic as syn          # noqa
from centertrack_b200.decode import generic_decode     # noqa
from helpers import make_model                         # noqa

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device('cuda')
opt, model, sd = make_model('coco_tracking')
model = model.to(dev)
eng = model.engine_for(B, 512, 512, dev, 'bf16')
eng.set_fused_activations(True)
img, pre, hm = syn.synthetic_inputs(1, 512, 512)
g = torch.Generator().manual_seed(0)
x = (img + 0.05 * torch.randn(B, 3, 512, 512, generator=g)).to(dev)
p = (pre + 0.05 * torch.randn(B, 3, 512, 512, generator=g)).to(dev)
h = hm.expand(B, 1, 512, 512).contiguous().to(dev)
for _ in range(steps):
  out = dict(eng.forward(x, p, h))
  generic_decode(out, K=100)
torch.cuda.synchronize()
names = [(k, n) for k, pl, n in eng.ops]
print('ops per step:', len(names))
for i, (k, n) in enumerate(names):
  print(i, k, n)
