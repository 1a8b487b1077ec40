"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
  if p not in sys.path:
    sys.path.insert(0, p)

from centertrack_b200 import synthetic as wt   # noqa: E402

TASKS = {'coco_tracking': ['tracking'], 'mot': ['tracking', '--num_classes', '1', '--input_h', '544',
                                                 '--input_w', '960'],
         'nuscenes_ddd': ['tracking,ddd'], 'coco_pose': ['tracking,multi_pose']}


def make_opt(cfg, extra=()):
  from centertrack_b200.opts import opts
  return opts().init(TASKS[cfg] + ['--pre_hm'] + list(extra))


def make_model(cfg, seed=317, extra=()):
  """B200 model with the deterministic synthetic weights (same tensors the goldens were made with)."""
  from centertrack_b200.model import create_model
  opt = make_opt(cfg, extra)
  m = create_model(opt.arch, opt.heads, opt.head_conv, opt=opt)
  sd = wt.make_state_dict(m.state_dict(), seed)
  m.load_state_dict(sd)
  return opt, m, sd


def decode_inputs(kind, B, C, H, W, seed):
  """Must stay identical to oracle/gen_golden.py::decode_inputs."""
  rng = np.random.RandomState(seed)
  out = {'hm': (1. / (1. + np.exp(-(2 * rng.randn(B, C, H, W) - 4.6)))).astype(np.float32),
         'reg': rng.rand(B, 2, H, W).astype(np.float32),
         'wh': (rng.randn(B, 2, H, W) * 6).astype(np.float32),
         'tracking': (rng.randn(B, 2, H, W) * 3).astype(np.float32)}
  if kind == 'ddd':
    out.update({'dep': (rng.rand(B, 1, H, W) * 60).astype(np.float32),
                'rot': rng.randn(B, 8, H, W).astype(np.float32),
                'dim': (rng.rand(B, 3, H, W) * 4).astype(np.float32),
                'amodel_offset': rng.randn(B, 2, H, W).astype(np.float32)})
  if kind == 'pose':
    out.update({'hps': (rng.randn(B, 34, H, W) * 6).astype(np.float32),
                'hm_hp': (1. / (1. + np.exp(-(2 * rng.randn(B, 17, H, W) - 3.0)))).astype(np.float32),
                'hp_offset': rng.rand(B, 2, H, W).astype(np.float32)})
  if kind == 'mot':
    out['ltrb_amodal'] = (rng.randn(B, 4, H, W) * 8).astype(np.float32)
  return out


DECODE_CASES = [('coco', 1, 80, 128, 128, 100, 11), ('mot', 1, 1, 136, 240, 100, 12),
                ('ddd', 1, 10, 112, 200, 100, 13), ('pose', 1, 1, 128, 128, 100, 14),
                ('coco', 2, 80, 32, 32, 50, 15)]
