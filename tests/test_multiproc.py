"""N>1 host logic on CPU: two gloo processes shard 5 streams, gather their record buffers and rebuild
the per-stream order (the only exchange of the path)."""
import os
import socket
import subprocess
import sys

import torch

from centertrack_b200.sharding import gather_records, owner_of_stream, streams_of_rank


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  p = s.getsockname()[1]
  s.close()
  return p


def test_stream_sharding_and_record_gather_world2():
  port = _free_port()
  script = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mp_worker.py')
  ps = [subprocess.Popen([sys.executable, script, str(r), '2', str(port), '5'], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True) for r in range(2)]
  outs = [p.communicate(timeout=180)[0] for p in ps]
  for p, o in zip(ps, outs):
    assert p.returncode == 0, o
  assert 'RESULT 0 1 [0, 2, 4]' in outs[0] and 'RESULT 1 1 [1, 3]' in outs[1]


def test_stream_maps():
  assert streams_of_rank(10, 3, 8) == [3] and streams_of_rank(3, 5, 8) == []
  assert [owner_of_stream(s, 4) for s in range(6)] == [0, 1, 2, 3, 0, 1]
  r = torch.arange(6.).view(1, 2, 3)
  assert gather_records(r).shape == (1, 1, 2, 3)
