"""Per-dataset constants that Detector.__init__ reads off the reference's dataset classes
(detector.py:38-47): generic_dataset.py:21-52 plus datasets/{coco,coco_hp,mot,nuscenes,kitti_tracking,
kitti,crowdhuman}.py.  The dataset classes themselves (loading / eval) are out of scope."""
import numpy as np

_MEAN = np.array([0.40789654, 0.44719302, 0.47026115], dtype=np.float32).reshape(1, 1, 3)
_STD = np.array([0.28863828, 0.27408164, 0.27809835], dtype=np.float32).reshape(1, 1, 3)
_FLIP = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]


class DatasetInfo(object):
  mean, std = _MEAN, _STD
  rest_focal_length = 1200
  num_joints = 17
  flip_idx = _FLIP

  def __init__(self, name, default_resolution, num_categories, **kw):
    self.name = name
    self.default_resolution = list(default_resolution)
    self.num_categories = num_categories
    for k, v in kw.items():
      setattr(self, k, v)


dataset_factory = {
    'coco': DatasetInfo('coco', [512, 512], 80),
    'coco_hp': DatasetInfo('coco_hp', [512, 512], 1),
    'mot': DatasetInfo('mot', [544, 960], 1),
    'crowdhuman': DatasetInfo('crowdhuman', [512, 512], 1),  # the reference class leaves num_categories unset (crowdhuman.py:13: `num_classes`)
    'nuscenes': DatasetInfo('nuscenes', [448, 800], 10),
    'kitti': DatasetInfo('kitti', [384, 1280], 3),
    'kitti_tracking': DatasetInfo('kitti_tracking', [384, 1280], 3),
    'custom': DatasetInfo('custom', [-1, -1], 1),           # custom_dataset.py:8-9: class-level defaults before --num_classes
}


def get_dataset(name):
  return dataset_factory[name]
