"""N>1 host logic on CPU: world_size-2 gloo run of the static stream sharding + final byte gather.  The per-rank
"encoder" here is the CPU oracle (test infrastructure) standing in for the GPU so the collective plumbing, ordering
and closed-form sizes are exercised without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from synth import make_signal


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lens, q):
    import oracle_lib as O
    import lamejs_b200 as M
    from lamejs_b200.sharding import gather_encoded, shard_streams

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [M.stream_bytes(1, 44100, 128, n) for n in lens]
    mine = shard_streams(len(lens), world, rank)
    chunks = []
    for j in mine:
        x = make_signal("octave", lens[j], 44100, seed=j)[0]
        data = O.encode_stream(1, 44100, 128, x, None)[0]
        assert len(data) == sizes[j]
        chunks.append(np.frombuffer(data, dtype=np.uint8))
    packed = torch.from_numpy(np.concatenate(chunks).copy()) if chunks else torch.zeros(0, dtype=torch.uint8)
    res = gather_encoded(packed, sizes)
    if rank == 0:
        q.put([bytes(t.numpy().tobytes()) for t in res])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(oracle):
    lens = [3000, 1152 * 5, 10, 1152 * 9 + 77, 2000]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for j, n in enumerate(lens):
        x = make_signal("octave", n, 44100, seed=j)[0]
        assert got[j] == oracle.encode_stream(1, 44100, 128, x, None)[0], j


def test_shard_layout_is_a_partition():
    from lamejs_b200.sharding import shard_layout, shard_streams

    sizes = [5, 7, 11, 13, 17, 19, 23]
    for world in (1, 2, 4, 8):
        seen = []
        for r, (idx, offs, tot) in enumerate(shard_layout(sizes, world)):
            assert idx == shard_streams(len(sizes), world, r)
            assert tot == sum(sizes[i] for i in idx)
            seen += idx
        assert sorted(seen) == list(range(len(sizes)))
