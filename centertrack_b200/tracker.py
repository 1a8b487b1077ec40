"""Displacement association with the reference's surface (utils/tracker.py:6-138):
Tracker(opt).init_track / step / reset, attributes .tracks / .id_count; mutates and returns the
same result dicts.  Greedy matching (the default, and what `ct_track_step` runs on the device), `--hungarian`
(tracker.py:52-72; the reference imports sklearn's removed `linear_assignment`, here scipy's
`linear_sum_assignment` solves the same problem) and the MOT `--public_det` birth rule (tracker.py:83-103)."""
import numpy as np

_BLOCKED = 1e18      # cost of a forbidden (detection, track) pair
_TAKEN = 1e16        # anything at or above this is not a match


def greedy_assignment(dist):
  """tracker.py:129-138: detections, in score order, take their nearest track that is still free.
  `dist` [N, M] is consumed (matched columns are blocked in place); returns int32 pairs [n, 2]."""
  pairs = []
  if dist.shape[1] > 0:
    for det, row in enumerate(dist):
      trk = int(row.argmin())
      if row[trk] < _TAKEN:
        dist[:, trk] = _BLOCKED
        pairs.append((det, trk))
  return np.array(pairs, np.int32).reshape(-1, 2)


def hungarian_assignment(cost):
  """tracker.py:52-55,63-72: minimum-cost assignment over the gated cost matrix (blocked pairs clamped to exactly
  _BLOCKED so that they all weigh the same); pairs the solver was forced to take through a blocked cell are handed
  back as `rejected` -- the reference appends those detections / tracks to its unmatched lists AFTER the naturally
  unmatched ones, which fixes the order new ids are given in."""
  from scipy.optimize import linear_sum_assignment
  cost[cost > _BLOCKED] = _BLOCKED
  rows, cols = linear_sum_assignment(cost)
  keep = cost[rows, cols] <= _TAKEN if len(rows) else np.zeros(0, bool)
  pairs = np.stack([rows, cols], 1).astype(np.int64).reshape(-1, 2)
  return pairs[keep], pairs[~keep], pairs


def _box_area(box):
  return (box[2] - box[0]) * (box[3] - box[1])


class Tracker(object):

  def __init__(self, opt):
    self.opt = opt
    self.reset()

  def reset(self):
    self.id_count = 0
    self.tracks = []

  def _birth(self, item):
    """A detection confident enough to start a track gets the next id (tracker.py:18-21,108-113)."""
    self.id_count += 1
    item.update(tracking_id=self.id_count, age=1, active=1)
    return item

  def init_track(self, results):
    for item in results:
      if item['score'] <= self.opt.new_thresh:
        continue
      self._birth(item)
      if 'ct' not in item:
        x0, y0, x1, y1 = item['bbox'][:4]
        item['ct'] = [(x0 + x1) / 2, (y0 + y1) / 2]
      self.tracks.append(item)

  def _gated_cost(self, results):
    """Squared distance between each detection's predicted previous centre (ct + tracking) and each track centre;
    pairs farther apart than either box's area, or of different classes, are blocked (tracker.py:28-49)."""
    n, m = len(results), len(self.tracks)
    pred = np.array([np.asarray(r['ct']) + np.asarray(r['tracking']) for r in results], np.float32).reshape(n, 2)
    prev = np.array([t['ct'] for t in self.tracks], np.float32).reshape(m, 2)
    det_area = np.array([_box_area(r['bbox']) for r in results], np.float32)
    trk_area = np.array([_box_area(t['bbox']) for t in self.tracks], np.float32)
    det_cls = np.array([r['class'] for r in results], np.int32)
    trk_cls = np.array([t['class'] for t in self.tracks], np.int32)
    cost = ((prev[None, :, :] - pred[:, None, :]) ** 2).sum(axis=2)            # [n, m] float32
    blocked = (cost > trk_area[None, :]) | (cost > det_area[:, None]) | (det_cls[:, None] != trk_cls[None, :])
    return cost + blocked * _BLOCKED

  def _public_births(self, results, pred, unmatched, det_area, public_det):
    """tracker.py:83-103 (MOT public-detection protocol): a track may only start on a provided detection.  Each public
    detection, in order, claims the unmatched detection whose predicted previous centre is nearest to it, if that is
    closer than the detection's own box area; a claimed detection is out of the pool whether or not it is confident
    enough to be born."""
    pub = np.array([d['ct'] for d in public_det], np.float32).reshape(1, -1, 2)
    cost = ((pred.reshape(-1, 1, 2) - pub) ** 2).sum(axis=2)                   # [n, p] float32
    free = np.zeros(len(results), bool)
    free[list(unmatched)] = True
    cost[~free] = _BLOCKED
    born = []
    for j in range(cost.shape[1]):
      i = int(cost[:, j].argmin())
      if cost[i, j] < det_area[i]:
        cost[i, :] = _BLOCKED
        if results[i]['score'] > self.opt.new_thresh:
          born.append(self._birth(results[i]))
    return born

  def step(self, results, public_det=None):
    n, m = len(results), len(self.tracks)
    cost = self._gated_cost(results)
    if getattr(self.opt, 'hungarian', False):
      pairs, rejected, taken = hungarian_assignment(cost)
    else:
      pairs = taken = greedy_assignment(cost.copy())
      rejected = pairs[:0]
    out = []
    for det, trk in pairs:                       # continued tracks inherit the id; `active` counts the streak
      item, old = results[det], self.tracks[trk]
      item.update(tracking_id=old['tracking_id'], age=1, active=old['active'] + 1)
      out.append(item)
    fresh = sorted(set(range(n)) - set(taken[:, 0].tolist())) + rejected[:, 0].tolist()
    lost = sorted(set(range(m)) - set(taken[:, 1].tolist())) + rejected[:, 1].tolist()
    if getattr(self.opt, 'public_det', False) and fresh:
      pred = np.array([np.asarray(r['ct']) + np.asarray(r['tracking']) for r in results], np.float32).reshape(n, 2)
      det_area = np.array([_box_area(r['bbox']) for r in results], np.float32)
      out.extend(self._public_births(results, pred, fresh, det_area, public_det))
    else:
      out.extend(self._birth(results[i]) for i in fresh if results[i]['score'] > self.opt.new_thresh)
    for i in lost:                               # unmatched tracks coast while younger than max_age
      old = self.tracks[i]
      if old['age'] < self.opt.max_age:
        old['age'] += 1
        old['active'] = 0
        out.append(old)
    self.tracks = out
    return out
