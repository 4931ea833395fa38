"""CPU: the N>1 launcher logic of bench.py (rank-sharded streams + gather of fixed-size detection records)
with world_size 2 over gloo -- no GPU, no data-path collective other than the final gather."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank owns its own stream (weak scaling): different seeds -> different frames
    frames = bench.synth_stream(1000 + rank, 2)
    rec = torch.zeros((2, bench.MAX_DET, 7), dtype=torch.float32)
    rec[:, 0, 0] = float(frames[:, 0, 0, 0].sum())           # a rank-dependent fingerprint
    rec[:, 0, 6] = rank
    gathered = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(gathered, rec)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                   # max-over-ranks timing rule
    if rank == 0:
        out.put(([g[:, 0, 6].tolist() for g in gathered], float(t.item()), [float(g[0, 0, 0]) for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_and_max_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ranks, tmax, fp = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ranks == [[0.0, 0.0], [1.0, 1.0]]
    assert tmax == 11.0
    assert fp[0] != fp[1]          # the two ranks really processed different streams
