"""bench.py pieces that run without a GPU: the `--impl reference` arm's JSON line (the oracle port on the host cores,
one 64x96-free full-size frame per step is too slow here, so the contract is checked on the smallest legal run) and
the stock-PyTorch context leg on the CPU device (same code path as on the GPU, minus the CUDA synchronisations)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                     capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1, r.stdout
  d = json.loads(lines[0])
  assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
  assert d['metric'] == 'frames/sec (device-timed) DLA-34 512x512'
  assert d['value'] > 0 and d['steps'] == 1
  assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
  assert d['e2e'] == {'value': d['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_stock_pytorch_context_leg_runs_on_the_cpu_device():
  sys.path.insert(0, ROOT)
  import bench
  from centertrack_b200 import synthetic as wt
  from helpers import make_model
  opt, model, sd = make_model('coco_tracking')
  saved = bench.K
  bench.K = 40
  try:
    r = bench.stock_pytorch_leg(sd, opt.heads, 2, 64, 96, torch.device('cpu'), wt, steps=1, warmup=0)
  finally:
    bench.K = saved
  assert r['kind'] == 'port' and r['fp32'] > 0 and r['bf16_autocast'] > 0 and r['frames_per_step'] == 2
