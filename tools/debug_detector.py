"""Debug helper: Detector.run vs the oracle pipeline on one synthetic frame (fp32 engine)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle'))
import numpy as np, torch
import ct_oracle as co
from helpers import make_model
from centertrack_b200.detector import Detector
from centertrack_b200 import detector as D
opt, model, sd = make_model('coco_tracking', extra=['--b200_precision', 'fp32', '--track_thresh', '0.02', '--new_thresh', '0.02', '--input_h', '128', '--input_w', '160'])
opt.load_model = ''
D.create_model = lambda *a, **k: model
det = Detector(opt)
rng = np.random.RandomState(0)
f = rng.randint(0, 255, (120, 160, 3)).astype(np.uint8)
images, meta = det.pre_process(f, 1.0)
orc = co.DLA34Oracle(sd, opt.heads)
phm = np.zeros((1, 1, 128, 160), np.float32)
ref = orc.forward(images, images, torch.from_numpy(phm))
eng = det.model.engine_for(1, 128, 160, torch.device('cuda'), 'fp32')
out = eng.forward(images.cuda().contiguous(), images.cuda().contiguous(), torch.from_numpy(phm).cuda())
for k in ref:
  r = ref[k].numpy(); g = out[k].float().cpu().numpy()
  print(k, 'max|ref|', np.abs(r).max(), 'max err', np.abs(r - g).max(), 'mean err', np.abs(r - g).mean())
for name in ['stem', 'base.level2', 'base.level5', 'feat']:
  pass
out2 = eng.forward(images.cuda().contiguous(), images.cuda().contiguous(), None)
for k in ref:
  r = ref[k].numpy(); g = out2[k].float().cpu().numpy()
  print('pre_hm=None', k, 'max err', np.abs(r - g).max())
ret = det.run(f)
got = ret['results']
o = co.sigmoid_output(ref)
dets = co.generic_decode({k: v for k, v in o.items()}, opt.K)
dets = {k: v for k, v in dets.items() if not k.startswith('_')}
res = co.generic_post_process(dets, [meta['c']], [meta['s']], meta['out_height'], meta['out_width'], opt.out_thresh, [meta['calib']])[0]
res = [r for r in res if r['score'] > opt.out_thresh]
print(len(got), len(res), 'out_thresh', opt.out_thresh)
for i in range(min(len(got), len(res), 100)):
  a, b = got[i], res[i]
  flag = '' if np.abs(np.asarray(a['bbox']) - np.asarray(b['bbox'])).max() < 0.05 else '  <<<'
  if flag or i < 5:
    print(i, a['score'], b['score'], a['class'], b['class'], a['ct'], b['ct'], np.round(a['bbox'], 2), np.round(b['bbox'], 2), flag)
