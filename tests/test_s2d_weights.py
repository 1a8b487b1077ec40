"""Space-to-depth restatements used by the bf16 engine for the 16-channel 512x512 layers (engine.py): the regrouped
weights must reproduce the original convolutions exactly (float64, CPU)."""
import torch
import torch.nn.functional as F

from centertrack_b200.engine import s2d_weights_3x3_s1, s2d_weights_3x3_s2


def _s2d(x):            # [B, C, H, W] -> [B, (sy, sx, C), H/2, W/2]
  B, C, H, W = x.shape
  return x.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(B, 4 * C, H // 2, W // 2)


def test_stride1_conv_on_the_space_to_depth_grid():
  g = torch.Generator().manual_seed(5)
  x = torch.randn(2, 16, 12, 20, generator=g, dtype=torch.float64)
  w = torch.randn(16, 16, 3, 3, generator=g, dtype=torch.float64)
  b = torch.randn(16, generator=g, dtype=torch.float64)
  ref = _s2d(F.conv2d(x, w, b, 1, 1))
  got = F.conv2d(_s2d(x), s2d_weights_3x3_s1(w), b.repeat(4), 1, 1)
  assert torch.allclose(got, ref, rtol=0, atol=1e-12)
  assert float((s2d_weights_3x3_s1(w) != 0).double().mean()) == 0.25


def test_stride2_conv_as_2x2_over_the_space_to_depth_input():
  g = torch.Generator().manual_seed(6)
  x = torch.randn(2, 16, 12, 20, generator=g, dtype=torch.float64)
  w = torch.randn(32, 16, 3, 3, generator=g, dtype=torch.float64)
  b = torch.randn(32, generator=g, dtype=torch.float64)
  ref = F.conv2d(x, w, b, 2, 1)
  got = F.conv2d(F.pad(_s2d(x), (1, 0, 1, 0)), s2d_weights_3x3_s2(w), b)
  assert got.shape == ref.shape and torch.allclose(got, ref, rtol=0, atol=1e-12)
