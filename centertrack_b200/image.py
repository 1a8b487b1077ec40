"""Host-side geometry used around the hot path (stays numpy, as in the reference):
utils/image.py:20-26 (transform_preds_with_trans), 37-70 (get_affine_transform), 73-76
(affine_transform), 105-125 (gaussian_radius), 128-154 (gaussian2D / draw_umich_gaussian)."""
import numpy as np


def _solve_affine(src, dst):
  # 2x3 matrix M with M @ [x, y, 1]^T = dst for three point pairs (what cv2.getAffineTransform solves)
  A = np.concatenate([np.asarray(src, np.float64), np.ones((3, 1))], axis=1)
  return np.linalg.solve(A, np.asarray(dst, np.float64)).T


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32),
                         inv=0):
  if not isinstance(scale, (np.ndarray, list)):
    scale = np.array([scale, scale], dtype=np.float32)
  scale = np.asarray(scale, dtype=np.float32)
  center = np.asarray(center, dtype=np.float32)
  src_w, dst_w, dst_h = scale[0], output_size[0], output_size[1]
  a = np.pi * rot / 180
  sn, cs = np.sin(a), np.cos(a)
  src_dir = np.array([0 * cs - (src_w * -0.5) * sn, 0 * sn + (src_w * -0.5) * cs])
  dst_dir = np.array([0, dst_w * -0.5], np.float32)
  src = np.zeros((3, 2), dtype=np.float32)
  dst = np.zeros((3, 2), dtype=np.float32)
  src[0] = center + scale * shift
  src[1] = center + src_dir + scale * shift
  dst[0] = [dst_w * 0.5, dst_h * 0.5]
  dst[1] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
  for pts in (src, dst):
    d = pts[0] - pts[1]
    pts[2] = pts[1] + np.array([-d[1], d[0]], dtype=np.float32)
  return _solve_affine(dst, src) if inv else _solve_affine(src, dst)


def transform_preds_with_trans(coords, trans):
  t = np.ones((coords.shape[0], 3), np.float32)
  t[:, :2] = coords
  return np.dot(trans, t.transpose()).transpose()[:, :2]


def affine_transform(pt, t):
  return np.dot(t, np.array([pt[0], pt[1], 1.], dtype=np.float32).T)[:2]


def gaussian_radius(det_size, min_overlap=0.7):
  h, w = det_size
  b1 = h + w
  c1 = w * h * (1 - min_overlap) / (1 + min_overlap)
  r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
  b2 = 2 * (h + w)
  c2 = (1 - min_overlap) * w * h
  r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
  a3 = 4 * min_overlap
  b3 = -2 * min_overlap * (h + w)
  c3 = (min_overlap - 1) * w * h
  r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
  return min(r1, r2, r3)


def gaussian2D(shape, sigma=1):
  m, n = [(ss - 1.) / 2. for ss in shape]
  y, x = np.ogrid[-m:m + 1, -n:n + 1]
  h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
  h[h < np.finfo(h.dtype).eps * h.max()] = 0
  return h


def draw_umich_gaussian(heatmap, center, radius, k=1):
  diameter = 2 * radius + 1
  g = gaussian2D((diameter, diameter), sigma=diameter / 6)
  x, y = int(center[0]), int(center[1])
  height, width = heatmap.shape[0:2]
  left, right = min(x, radius), min(width - x, radius + 1)
  top, bottom = min(y, radius), min(height - y, radius + 1)
  mh = heatmap[y - top:y + bottom, x - left:x + right]
  mg = g[radius - top:radius + bottom, radius - left:radius + right]
  if min(mg.shape) > 0 and min(mh.shape) > 0:
    np.maximum(mh, mg * k, out=mh)
  return heatmap
