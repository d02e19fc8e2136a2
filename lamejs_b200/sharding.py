"""Static sharding of independent streams over ranks + the final byte gather.

Streams (lamejs Mp3Encoder instances) never exchange data while encoding, so the multi-GPU plan is: stream j goes to
rank j % world (round-robin, like config C4), every rank encodes its shard with no collective on the data path, and the
encoded bytes are gathered to rank 0 at the end.  CBR without reservoir makes every stream's byte count a closed form of
its sample count (mp3b200_stream_bytes), so each rank knows all sizes up front and the gather needs no size exchange.
Backend-agnostic: NCCL on GPUs (bench.py), gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def shard_streams(nstreams, world, rank):
    """Indices of the streams rank `rank` encodes (round-robin)."""
    return list(range(rank, nstreams, world))


def shard_layout(stream_bytes, world):
    """Per rank: (stream indices, byte offset of each of its streams inside the rank's packed buffer, total)."""
    out = []
    for r in range(world):
        idx = shard_streams(len(stream_bytes), world, r)
        offs = np.concatenate([[0], np.cumsum([stream_bytes[i] for i in idx])]).astype(np.int64)
        out.append((idx, offs[:-1], int(offs[-1])))
    return out


def gather_encoded(packed, stream_bytes, group=None, dst=0):
    """`packed`: this rank's encoded streams back to back (uint8 tensor on the backend's device).  Returns on `dst` the
    list of per-stream byte tensors in global stream order, None elsewhere.  One collective: a padded gather."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    layout = shard_layout(stream_bytes, world)
    cap = max(l[2] for l in layout)
    buf = torch.zeros(cap, dtype=torch.uint8, device=packed.device)
    buf[: packed.numel()] = packed
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    result = [None] * len(stream_bytes)
    for r, (idx, offs, _) in enumerate(layout):
        for i, o in zip(idx, offs):
            result[i] = bufs[r][int(o): int(o) + int(stream_bytes[i])]
    return result
