"""Multi-GPU entry points of the C-ABI on real devices (skipped with fewer than two GPUs): one process per GPU, NCCL
communicator per context bootstrapped with gple_comm_unique_id / gple_ctx_comm_init, no torch.distributed on the data path.

  * gple_evolve_sharded on 2 ranks returns, on every rank, exactly (bit for bit) what gple_evolve returns on one GPU, for even
    and uneven partitions and all three elements;
  * gple_allgather_points / gple_allreduce_sum on uneven blocks;
  * one process holding contexts on two devices trains and predicts on both (per-device kernel attributes)."""
import os
import tempfile
import time

import numpy as np
import pytest

from gaussian_process_liouville_equation_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


needs_two = pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
THETA_C = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])
CENTRE = (-0.5, syn.P0)


def _inputs(n, q):
    sets = [syn.training_set(61, e, n, CENTRE) for e in range(3)]
    pts = []
    for e in range(3):
        Xe, ye = syn.extra_points(61, e, sets[e][0], q + e, CENTRE)  # three different sizes: uneven blocks
        pts.append(syn.points_aos(Xe, ye))
    return sets, pts


def _models(sets, ctx):
    from gaussian_process_liouville_equation_b200 import complex_kernel, kernel

    return [kernel.TrainingKernel(syn.theta_real(), sets[0], ctx=ctx), complex_kernel.TrainingComplexKernel(THETA_C, sets[1], ctx=ctx), kernel.TrainingKernel(syn.theta_real(), sets[2], ctx=ctx)]


def _worker(rank, world, id_file, n, q, out_dir):
    from gaussian_process_liouville_equation_b200 import _lib as L
    from gaussian_process_liouville_equation_b200 import dynamics

    ctx = L.Context(rank)
    if rank == 0:
        uid = L.comm_unique_id()
        with open(id_file + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(id_file + ".tmp", id_file)
    else:
        for _ in range(6000):
            if os.path.exists(id_file):
                break
            time.sleep(0.01)
        uid = open(id_file, "rb").read()
    ctx.comm_init(rank, world, uid)
    assert ctx.comm_info() == (rank, world)
    sets, pts = _inputs(n, q)
    g = _models(sets, ctx)
    out = dynamics.evolve_sharded(1, pts, syn.MASS, 1.0, g, ctx=ctx)
    # all-gather of an uneven partition and a sum over the ranks
    total = 37
    full = np.arange(total * 4, dtype=np.float64).reshape(total, 4)
    lo, hi = L.partition(total, rank, world)
    mine = np.full_like(full, -1.0)
    mine[lo:hi] = full[lo:hi]
    ctx.allgather_points(mine, total)
    s = np.array([1.0 + rank, 10.0 * (rank + 1)])
    ctx.allreduce_sum(s)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), e0=out[0], e1=out[1], e2=out[2], gathered=mine, summed=s)


@needs_two
@pytest.mark.parametrize("q", [4096, 1001])
def test_sharded_evolve_on_two_gpus_equals_one_gpu_bit_for_bit(q):
    import torch.multiprocessing as mp

    from gaussian_process_liouville_equation_b200 import _lib as L
    from gaussian_process_liouville_equation_b200 import dynamics

    n, world = 300, 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, os.path.join(d, "id"), n, q, d), nprocs=world, join=True)
        ranks = [np.load(os.path.join(d, f"rank{r}.npz")) for r in range(world)]
    sets, pts = _inputs(n, q)
    ctx = L.default_context()
    whole = dynamics.evolve(1, pts, syn.MASS, 1.0, _models(sets, ctx), ctx=ctx)
    for r in ranks:
        for e in range(3):
            assert np.array_equal(r[f"e{e}"], whole[e]), (e, np.abs(r[f"e{e}"] - whole[e]).max())
        assert np.array_equal(r["gathered"], np.arange(37 * 4, dtype=np.float64).reshape(37, 4))
        assert np.array_equal(r["summed"], np.array([3.0, 30.0]))


@needs_two
def test_one_process_with_contexts_on_two_devices():
    """gple_ctx_create(device) for two devices in ONE process: the opt-in shared-memory attributes are per device."""
    from gaussian_process_liouville_equation_b200 import _lib as L
    from gaussian_process_liouville_equation_b200 import kernel

    X, y = syn.training_set(62, 0, 700, CENTRE)
    Xq, _ = syn.extra_points(62, 0, X, 3000, CENTRE)
    res = []
    for dev in (0, 1):
        ctx = L.Context(dev)
        k = kernel.TrainingKernel(syn.theta_real(), (X, y), ctx=ctx)
        p = kernel.PredictiveKernel(Xq, k)
        res.append((k.get_error(), k.get_population(), p.get_prediction().copy(), p.get_variance().copy()))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])


def test_single_context_without_communicator_is_rank_0_of_1():
    from gaussian_process_liouville_equation_b200 import _lib as L

    ctx = L.default_context()
    assert ctx.comm_info() == (0, 1)
    assert L.partition(10, 0, 1) == (0, 10)
    a = np.arange(8.0).reshape(2, 4)
    ctx.allgather_points(a, 2)  # no-op
    assert np.array_equal(a, np.arange(8.0).reshape(2, 4))
