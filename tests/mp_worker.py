"""Worker for tests/test_multiproc.py: python mp_worker.py <rank> <world> <port> <n_streams>."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from centertrack_b200.sharding import gather_records, merge_stream_results, streams_of_rank


def main():
  rank, world, port, n_streams = (int(a) for a in sys.argv[1:5])
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  mine = streams_of_rank(n_streams, rank, world)
  B = (n_streams + world - 1) // world
  rec = torch.zeros(B, 4, 3)
  for b, s in enumerate(mine):
    rec[b] = float(s)                      # stand-in for the decode records of stream s
  allg = gather_records(rec, dist)
  merged = merge_stream_results(allg, n_streams, world)
  ok = all(float(merged[s][0, 0]) == float(s) for s in range(n_streams))
  on0 = gather_records(rec, dist, dst=0)
  ok &= (on0 is not None) == (rank == 0)
  if rank == 0:
    ok &= torch.equal(on0, allg)
  dist.barrier()
  dist.destroy_process_group()
  print('RESULT', rank, int(ok), mine)
  sys.exit(0 if ok else 1)


if __name__ == '__main__':
  main()
