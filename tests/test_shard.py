"""Multi-GPU host logic on CPU: world_size-2 gloo job, shard by system index, final host gather.
The per-shard compute here is the host build of the device integrator (tests/hostsim) standing in
for the GPU — what is under test is the partition + gather, which must reproduce the unsharded
result bit for bit (systems are independent, no data-path collective)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from nbody_ai_b200 import _abi, engine, workloads
from conftest import ROOT


def test_shard_range_partition():
    for n in (0, 1, 7, 64, 70001):
        for ws in (1, 2, 3, 8):
            parts = [engine.shard_range(n, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_shard_balanced_partition(built):
    """Cost-balanced shards: a partition with equal counts (+-1) and near-equal cost, heavy systems spread."""
    par = workloads.c2_params(400, 6, seed=5)
    cost = engine.system_cost(par, 365.0, 8760)
    assert cost.min() >= 3.0 and cost.max() >= 7.0          # C4 systems (2 + 1) ... several SBABC3 steps per hour
    for ws in (1, 2, 3, 8):
        parts = [engine.shard_balanced(cost, r, ws) for r in range(ws)]
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(len(par)))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1 and all(np.all(np.diff(p) > 0) for p in parts)
        tot = np.array([cost[p].sum() for p in parts])
        assert tot.max() / tot.min() < 1.01
        heavy = np.array([(cost[p] >= 7.0).sum() for p in parts])
        assert heavy.max() - heavy.min() <= 1
    # contiguous shards of the same batch are visibly less even
    tot = np.array([cost[slice(*engine.shard_range(len(par), r, 8))].sum() for r in range(8)])
    assert tot.max() / tot.min() > 1.01


def _host_products(par):
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    plan = engine.Plan(par, 40.0, 961, substeps=1, flags=_abi.F_MEANS)
    buf = engine.Buffers(plan)
    inp, out = buf.as_structs()
    assert hs.hostsim_run(C.byref(plan.cfg), C.byref(inp), C.byref(out)) == 0
    B, Np = plan.B, plan.Np
    first_tt = np.array([[buf.tt_of(b, j)[0] if buf.n_transits(b, j) else np.nan for j in range(Np)] for b in range(B)])
    return {"n_tt": buf.n_tt.reshape(B, Np).copy(), "mean_a": buf.mean_a.reshape(B, Np).copy(),
            "mean_e": buf.mean_e.reshape(B, Np).copy(), "first_tt": first_tt, "status": buf.status.copy()}


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    par = workloads.c2_params(3, 2, seed=21)           # 9 systems: uneven split over 2 ranks
    lo, hi = engine.shard_range(len(par), rank, world)
    local = _host_products(par[lo:hi])
    full = engine.gather_results(local, len(par))
    # the same with cost-balanced (non-contiguous) shards
    mine = engine.shard_balanced(engine.system_cost(par, 40.0, 961), rank, world)
    full_b = engine.gather_results(_host_products(par[mine]), len(par), index=mine)
    if rank == 0:
        q.put(({k: v for k, v in full.items()}, {k: v for k, v in full_b.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(built):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, got_balanced = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _host_products(workloads.c2_params(3, 2, seed=21))
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
        assert np.array_equal(got_balanced[k], want[k], equal_nan=True), k


def _shared_worker(rank, world, port, q, tag):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    hs = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
    par = workloads.c2_params(3, 2, seed=21)
    mine = engine.shard_balanced(engine.system_cost(par, 40.0, 961), rank, world)
    pool = engine.SharedHostPool("%s_%d" % (tag, rank), 8 << 20, register=False)       # no CUDA here: plain shared memory
    for step in range(2):                         # the pool is reused from step to step
        pool.reset()
        plan = engine.Plan(par[mine], 40.0, 961, substeps=1, flags=_abi.F_MEANS, pool=pool)
        buf = pool.buffers_for(plan)
        inp, out = buf.as_structs()
        assert hs.hostsim_run(C.byref(plan.cfg), C.byref(inp), C.byref(out)) == 0
        pool.publish(plan, buf, extra={"index": mine.astype(np.int64)})
        dist.barrier()                            # the final host gather: every shard is now readable in place
        if rank == 0:
            n_tt = np.zeros((len(par), 2), dtype=np.int32)
            mean_a, first_tt = np.zeros((len(par), 2)), np.full((len(par), 2), np.nan)
            for r in range(world):
                sh = engine.SharedHostPool.attach("%s_%d" % (tag, r), ack=True)
                assert sh["seq"] == step + 1
                idx = sh["index"]
                n_tt[idx] = sh["n_tt"].reshape(-1, 2)
                mean_a[idx] = sh["mean_a"].reshape(-1, 2)
                for q_, b in enumerate(idx):
                    for j in range(2):
                        if sh["n_tt"][q_ * 2 + j]:
                            first_tt[b, j] = sh["tt"][sh["tt_offsets"][q_ * 2 + j]]
                engine.SharedHostPool.acknowledge(sh)
                del sh
        assert pool.wait_consumed()               # the owner may reuse its segment once the reader has acknowledged
        dist.barrier()
    if rank == 0:
        q.put({"n_tt": n_tt, "mean_a": mean_a, "first_tt": first_tt})
    dist.barrier()
    pool.close()
    dist.destroy_process_group()


def test_shared_memory_gather(built):
    """engine.SharedHostPool: each rank's results in its own shared segment, read in place by rank 0 after a barrier —
    the zero-copy final host gather of a one-node job; bit-identical to the unsharded run."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    tag = "nbt_test_%d" % os.getpid()
    procs = [ctx.Process(target=_shared_worker, args=(r, 2, port, q, tag)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _host_products(workloads.c2_params(3, 2, seed=21))
    for k in got:
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    assert not [f for f in os.listdir("/dev/shm") if f.startswith(tag)]
