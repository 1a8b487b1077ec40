"""Output-grid -> image coordinates and per-detection dicts: utils/post_process.py:12-91 and
utils/ddd_utils.py:91-136 of the reference (host side of Detector.post_process; numpy)."""
import numpy as np

from .image import get_affine_transform, transform_preds_with_trans


def get_alpha(rot):
  # rot: (B, 8) [bin1_cls0, bin1_cls1, bin1_sin, bin1_cos, bin2_cls0, bin2_cls1, bin2_sin, bin2_cos]
  idx = rot[:, 1] > rot[:, 5]
  alpha1 = np.arctan2(rot[:, 2], rot[:, 3]) + (-0.5 * np.pi)
  alpha2 = np.arctan2(rot[:, 6], rot[:, 7]) + (0.5 * np.pi)
  return alpha1 * idx + alpha2 * (1 - idx)


def unproject_2d_to_3d(pt_2d, depth, P):
  z = depth - P[2, 3]
  x = (pt_2d[0] * depth - P[0, 3] - P[0, 2] * z) / P[0, 0]
  y = (pt_2d[1] * depth - P[1, 3] - P[1, 2] * z) / P[1, 1]
  return np.array([x, y, z], dtype=np.float32).reshape(3)


def alpha2rot_y(alpha, x, cx, fx):
  rot_y = alpha + np.arctan2(x - cx, fx)
  if rot_y > np.pi:
    rot_y -= 2 * np.pi
  if rot_y < -np.pi:
    rot_y += 2 * np.pi
  return rot_y


def ddd2locrot(center, alpha, dim, depth, calib):
  loc = unproject_2d_to_3d(center, depth, calib)
  loc[1] += dim[0] / 2
  return loc, alpha2rot_y(alpha, center[0], calib[0, 2], calib[0, 0])


def generic_post_process(opt, dets, c, s, h, w, num_classes, calibs=None, height=-1, width=-1):
  if 'scores' not in dets:
    return [{}], [{}]
  ret = []
  for i in range(len(dets['scores'])):
    preds = []
    trans = get_affine_transform(c[i], s[i], 0, (w, h), inv=1).astype(np.float32)
    scores = dets['scores'][i]
    # scores are sorted descending, so the reference's `break` at the first score < out_thresh
    # keeps a prefix: transform that prefix in one shot
    below = np.nonzero(scores < opt.out_thresh)[0]
    n = int(below[0]) if len(below) else len(scores)
    if n == 0:
      ret.append(preds)
      continue
    cts = dets['cts'][i][:n]
    ct_img = transform_preds_with_trans(cts.reshape(-1, 2), trans)
    trk = bbox = hps = None
    if 'tracking' in dets:
      trk = transform_preds_with_trans((dets['tracking'][i][:n] + cts).reshape(-1, 2), trans) - ct_img
    if 'bboxes' in dets:
      bbox = transform_preds_with_trans(dets['bboxes'][i][:n].reshape(-1, 2), trans).reshape(n, 4)
    if 'hps' in dets:
      hps = transform_preds_with_trans(dets['hps'][i][:n].reshape(-1, 2), trans).reshape(n, -1)
    for j in range(n):
      item = {'score': scores[j], 'class': int(dets['clses'][i][j]) + 1, 'ct': ct_img[j]}
      if trk is not None:
        item['tracking'] = trk[j]
      if bbox is not None:
        item['bbox'] = bbox[j]
      if hps is not None:
        item['hps'] = hps[j]
      if 'dep' in dets and len(dets['dep'][i]) > j:
        item['dep'] = dets['dep'][i][j]
      if 'dim' in dets and len(dets['dim'][i]) > j:
        item['dim'] = dets['dim'][i][j]
      if 'rot' in dets and len(dets['rot'][i]) > j:
        item['alpha'] = get_alpha(dets['rot'][i][j:j + 1])[0]
      if 'rot' in dets and 'dep' in dets and 'dim' in dets and len(dets['dep'][i]) > j:
        if 'amodel_offset' in dets and len(dets['amodel_offset'][i]) > j:
          ct_output = dets['bboxes'][i][j].reshape(2, 2).mean(axis=0)
          amodel_ct_output = ct_output + dets['amodel_offset'][i][j]
          ct = transform_preds_with_trans(amodel_ct_output.reshape(1, 2), trans).reshape(2).tolist()
        else:
          b = item['bbox']
          ct = [(b[0] + b[2]) / 2, (b[1] + b[3]) / 2]
        item['ct'] = ct
        item['loc'], item['rot_y'] = ddd2locrot(ct, item['alpha'], item['dim'], item['dep'], calibs[i])
      preds.append(item)
    for key in ('nuscenes_att', 'velocity'):
      if key in dets:
        for j in range(len(preds)):
          preds[j][key] = dets[key][i][j]
    ret.append(preds)
  return ret
