"""Greedy displacement association with the reference's surface (utils/tracker.py:6-138):
Tracker(opt).init_track / step / reset, attributes .tracks / .id_count; mutates and returns the
same result dicts.  Hungarian matching and the MOT public-detection mode are out of scope (not in any
benchmark config) and raise."""
import numpy as np


def greedy_assignment(dist):
  """tracker.py:129-138: detections in score order take their nearest still-free track."""
  matched = []
  if dist.shape[1] == 0:
    return np.array(matched, np.int32).reshape(-1, 2)
  for i in range(dist.shape[0]):
    j = dist[i].argmin()
    if dist[i][j] < 1e16:
      dist[:, j] = 1e18
      matched.append([i, j])
  return np.array(matched, np.int32).reshape(-1, 2)


class Tracker(object):

  def __init__(self, opt):
    self.opt = opt
    if getattr(opt, 'hungarian', False) or getattr(opt, 'public_det', False):
      raise NotImplementedError('--hungarian / --public_det are outside the B200 hot-path scope')
    self.reset()

  def init_track(self, results):
    for item in results:
      if item['score'] > self.opt.new_thresh:
        self.id_count += 1
        item['active'] = 1
        item['age'] = 1
        item['tracking_id'] = self.id_count
        if 'ct' not in item:
          bbox = item['bbox']
          item['ct'] = [(bbox[0] + bbox[2]) / 2, (bbox[1] + bbox[3]) / 2]
        self.tracks.append(item)

  def reset(self):
    self.id_count = 0
    self.tracks = []

  def step(self, results, public_det=None):
    N, M = len(results), len(self.tracks)
    dets = np.array([np.asarray(d['ct']) + np.asarray(d['tracking']) for d in results],
                    np.float32).reshape(N, 2)
    area = lambda b: (b[2] - b[0]) * (b[3] - b[1])
    track_size = np.array([area(t['bbox']) for t in self.tracks], np.float32)
    track_cat = np.array([t['class'] for t in self.tracks], np.int32)
    item_size = np.array([area(r['bbox']) for r in results], np.float32)
    item_cat = np.array([r['class'] for r in results], np.int32)
    tracks = np.array([t['ct'] for t in self.tracks], np.float32).reshape(M, 2)
    dist = ((tracks.reshape(1, -1, 2) - dets.reshape(-1, 1, 2)) ** 2).sum(axis=2)   # N x M
    invalid = ((dist > track_size.reshape(1, M)) + (dist > item_size.reshape(N, 1)) +
               (item_cat.reshape(N, 1) != track_cat.reshape(1, M))) > 0
    dist = dist + invalid * 1e18
    matches = greedy_assignment(dist.copy())
    matched_d, matched_t = set(matches[:, 0].tolist()), set(matches[:, 1].tolist())
    ret = []
    for m in matches:
      track = results[m[0]]
      track['tracking_id'] = self.tracks[m[1]]['tracking_id']
      track['age'] = 1
      track['active'] = self.tracks[m[1]]['active'] + 1
      ret.append(track)
    for i in range(N):
      if i in matched_d:
        continue
      track = results[i]
      if track['score'] > self.opt.new_thresh:
        self.id_count += 1
        track['tracking_id'] = self.id_count
        track['age'] = 1
        track['active'] = 1
        ret.append(track)
    for i in range(M):
      if i in matched_t:
        continue
      track = self.tracks[i]
      if track['age'] < self.opt.max_age:
        track['age'] += 1
        track['active'] = 0
        ret.append(track)
    self.tracks = ret
    return ret
