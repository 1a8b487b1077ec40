"""Greedy displacement association with the reference's surface (utils/tracker.py:6-138):
Tracker(opt).init_track / step / reset, attributes .tracks / .id_count; mutates and returns the
same result dicts.  Hungarian matching and the MOT public-detection mode are out of scope (not in any
benchmark config) and raise."""
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


def _box_area(box):
  return (box[2] - box[0]) * (box[3] - box[1])


class Tracker(object):

  def __init__(self, opt):
    if getattr(opt, 'hungarian', False) or getattr(opt, 'public_det', False):
      raise NotImplementedError('--hungarian / --public_det are outside the B200 hot-path scope')
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

  def step(self, results, public_det=None):
    pairs = greedy_assignment(self._gated_cost(results))
    out = []
    for det, trk in pairs:                       # continued tracks inherit the id; `active` counts the streak
      item, old = results[det], self.tracks[trk]
      item.update(tracking_id=old['tracking_id'], age=1, active=old['active'] + 1)
      out.append(item)
    fresh = set(range(len(results))) - set(pairs[:, 0].tolist())
    out.extend(self._birth(results[i]) for i in sorted(fresh) if results[i]['score'] > self.opt.new_thresh)
    lost = set(range(len(self.tracks))) - set(pairs[:, 1].tolist())
    for i in sorted(lost):                       # unmatched tracks coast while younger than max_age
      old = self.tracks[i]
      if old['age'] < self.opt.max_age:
        old['age'] += 1
        old['active'] = 0
        out.append(old)
    self.tracks = out
    return out
