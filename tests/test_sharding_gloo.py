"""CPU, world_size 2 over gloo: the multi-GPU host logic -- contiguous block ranges per rank and
the 256-bin histogram all-reduce of shared-table mode (the only collective on the path)."""
import importlib
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, block, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    huf = importlib.import_module("huffman-avx512_b200")
    from _cases import biased
    from _libs import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = np.frombuffer(biased(n, seed=77), dtype=np.uint8)
    lo, hi = huf.sharded.shard_bytes(n, block, rank, world)
    shard = data[lo:hi]
    # on a GPU box k_histogram produces the local histogram; here the checker stands in for it
    hist = torch.from_numpy(Oracle().histogram(shard.tobytes()).astype(np.int64))
    huf.sharded.allreduce_histogram(hist)
    q.put((rank, lo, hi, hist.numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything(huf):
    for n_blocks in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 8):
            ranges = [huf.sharded.shard_blocks(n_blocks, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_blocks
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            per = -(-n_blocks // world) if n_blocks else 0
            assert all(b1 - b0 <= per for b0, b1 in ranges)
    assert huf.sharded.shard_bytes(1000, 128, 1, 2) == (512, 1000)


def test_histogram_allreduce_world2():
    from _cases import biased
    from _libs import Oracle
    n, block, world = 1_000_003, 16384, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, block, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = Oracle().histogram(biased(n, seed=77)).astype(np.int64).tolist()
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n
    assert res[0][3] == want and res[1][3] == want  # every rank holds the global histogram
