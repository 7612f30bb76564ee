"""world_size-2 gloo test of the N>1 host logic (sharding, all-gather, inverse permutation) on CPU.
The per-shard evaluator is the CPU oracle here (tests may use it); on the GPU box bench.py runs the same
`evaluate_sharded` with `Evaluator.eval_batch` over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "model-predictive-control-tuning_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import copy
    from mpcgpu import shell3x3, synthetic_population
    from mpcgpu.distributed import evaluate_sharded
    from oracle import oracle as orc
    p = shell3x3(2)
    p = copy.copy(p); p.nit = 60; p.r = p.r[:60].copy(); p.v = p.v[:60].copy(); p.yref = p.yref[:, :60].copy()
    op = orc.OracleProblem(p)
    N, Nu, dl, lm = synthetic_population(p, 37, seed=8)     # odd size: ragged shards
    ev = lambda a, b, c, d, mode: orc.eval_batch(op, a, b, c, d, mode, 1)[0]
    full = evaluate_sharded(ev, N, Nu, dl, lm, "gam")
    fv = evaluate_sharded(ev, N, Nu, dl, lm, "vns")
    if rank == 0:
        ref = orc.eval_batch(op, N, Nu, dl, lm, "gam", 1)[0]
        refv = orc.eval_batch(op, N, Nu, dl, lm, "vns", 1)[0]
        q.put((bool(np.array_equal(full, ref)), bool(np.array_equal(fv, refv)), full.shape, fv.shape))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluation_matches_single_rank_bit_for_bit():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = q.get(timeout=300)
    for p in procs: p.join(timeout=60)
    assert res[0] and res[1], "gathered fitness differs from the single-rank result"
    assert res[2] == (37, 3) and res[3] == (37,)


def test_shard_indices_partition():
    from mpcgpu.distributed import shard_indices, work_estimate
    rng = np.random.default_rng(0)
    w = rng.random(101)
    parts = [shard_indices(101, 4, r, w) for r in range(4)]
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(101))
    loads = [w[p].sum() for p in parts]
    assert max(loads) / min(loads) < 1.15     # sorted round-robin balances the estimated work


def test_python_work_key_equals_the_library_key():
    """Both multi-GPU front ends must deal the same shards: mpcgpu.distributed.work_estimate (torchrun path) against
    mpcgpu_work_estimate (the key mpcgpu_create_multi and the launch order use; pure host code, callable without a GPU),
    for a plant with hard MV limits only and for one with soft output bands."""
    import ctypes as C
    import mpcgpu
    from mpcgpu import _capi
    from mpcgpu.distributed import work_estimate
    lib = _capi.load_library()
    for prob, soft in ((mpcgpu.shell3x3(2), False), (mpcgpu.shell7x5(), True)):
        ps, keep = _capi.make_problem_struct(prob)
        N, Nu, dl, lm = mpcgpu.synthetic_population(prob, 300, seed=2)
        N = np.ascontiguousarray(N, dtype=np.int32); Nu = np.ascontiguousarray(Nu, dtype=np.int32)
        dl = np.ascontiguousarray(dl, dtype=np.float64); lm = np.ascontiguousarray(lm, dtype=np.float64)
        w = np.zeros(len(N))
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        assert lib.mpcgpu_work_estimate(C.byref(ps), len(N), P(N), P(Nu), P(dl), P(lm), P(w)) == 0
        wp = work_estimate(N, Nu, dl, lm, dead_max=int(prob.plant.d.max()), soft=soft)
        np.testing.assert_allclose(w, wp, rtol=1e-12, atol=1e-12)
        assert np.array_equal(np.argsort(-w, kind="stable"), np.argsort(-wp, kind="stable"))
