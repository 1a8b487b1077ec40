"""Multi-GPU plumbing: the path shards by VIDEO STREAM -- frame t of a stream needs frame
t-1's image and tracker state (detector.py:99-110,148), different streams are independent.  One process
per GPU holds a full replica of the weights (~40 MB bf16) and owns streams s with s % world == rank;
there is no data-path collective.  The only exchange is the optional gather of the fixed-size record
buffers [B,K,F] to every rank / rank 0 (26 KB per frame)."""
import torch


def streams_of_rank(n_streams, rank, world):
  """Round-robin stream -> rank map (stream s lives on GPU s mod world)."""
  return list(range(rank, n_streams, world))


def owner_of_stream(stream, world):
  return stream % world


def gather_records(records, dist=None, dst=None):
  """records: [B,K,F] float32 tensor (CUDA with nccl, CPU with gloo), same shape on every rank.
  Returns [world,B,K,F] on every rank (dst=None, all_gather) or on rank dst only (else None)."""
  if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
    return records.unsqueeze(0)
  world = dist.get_world_size()
  if dst is None:
    out = torch.empty((world * records.shape[0],) + tuple(records.shape[1:]), dtype=records.dtype,
                      device=records.device)                      # concatenation along dim 0
    dist.all_gather_into_tensor(out, records.contiguous())
    return out.view((world,) + tuple(records.shape))
  bufs = [torch.empty_like(records) for _ in range(world)] if dist.get_rank() == dst else None
  dist.gather(records.contiguous(), bufs, dst=dst)
  return torch.stack(bufs) if bufs is not None else None


def merge_stream_results(gathered, n_streams, world):
  """[world, B, K, F] (rank r's row b = its b-th owned stream) -> list indexed by global stream id."""
  out = [None] * n_streams
  for r in range(world):
    for b, s in enumerate(streams_of_rank(n_streams, r, world)):
      out[s] = gathered[r, b]
  return out
