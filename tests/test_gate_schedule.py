"""The automatic schedule of the staged bound (csrc/gpr.cu, gate_schedule; include/gple_b200.h, gple_gate_schedule_automatic):
host-only logic of the library, checked without a GPU."""
import ctypes as C
import importlib.util
import os

import pytest

from gaussian_process_liouville_equation_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_known_schedules():
    assert L.gate_schedule_automatic(False, 16) == [(1, 0), (2, 0), (5, 0), (10, 0)]  # C2: N = 2048
    assert L.gate_schedule_automatic(True, 16) == [(1, 0), (2, 1), (5, 1), (10, 2), (16, 6)]
    assert L.gate_schedule_automatic(False, 32) == [(1, 0), (3, 0), (8, 0), (21, 0)]  # C4: N = 4096
    assert L.gate_schedule_automatic(False, 128) == [(1, 0), (2, 0), (6, 0), (14, 0), (35, 0), (84, 0)]  # north star: N = 16384
    assert L.gate_schedule_automatic(False, 1) == [] and L.gate_schedule_automatic(True, 1) == []  # one block: nothing to stage
    assert L.gate_schedule_automatic(False, 0) == []


@pytest.mark.parametrize("complex_element", [False, True])
def test_schedule_is_cumulative_and_leaves_a_last_stage(complex_element):
    for blocks in range(1, 260):
        s = L.gate_schedule_automatic(complex_element, blocks)
        assert len(s) <= 15
        re = im = 0
        for r, i in s:
            assert 0 <= r <= blocks and 0 <= i <= (blocks if complex_element else 0)
            assert r >= re and i >= im and (r, i) != (re, im)  # every stage adds tiles
            re, im = r, i
        assert re < blocks or im < (blocks if complex_element else 0)  # the implied last stage is never empty
        if blocks >= 2:
            assert s[0] == (1, 0)  # the first block alone, no Im tile (each costs a product over all Re columns)


def test_matches_the_offline_model():
    """profiles/gate_schedule_sim.py restates the rule for the dynamic programme; both must agree."""
    spec = importlib.util.spec_from_file_location("gate_schedule_sim", os.path.join(ROOT, "profiles", "gate_schedule_sim.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    for blocks in (2, 3, 4, 6, 8, 16, 24, 32, 64, 100, 128, 200):
        for nb in (1, 2):
            want, re, im = [], 0, 0
            for r, i in sim.automatic(blocks, nb):
                r, i = max(re, min(r, blocks)), max(im, min(i, blocks if nb == 2 else 0))
                if (r > re or i > im) and (r < blocks or i < (blocks if nb == 2 else 0)):
                    want.append((r, i))
                    re, im = r, i
            assert L.gate_schedule_automatic(nb == 2, blocks) == want


def test_capacity_and_bad_arguments():
    lib = L.load()
    re, im = (C.c_int * 2)(), (C.c_int * 2)()
    n = lib.gple_gate_schedule_automatic(0, 128, C.cast(re, C.c_void_p), C.cast(im, C.c_void_p), 2)
    assert n == 6 and list(re) == [1, 2]  # the count is the full one, only `capacity` entries are written
    assert lib.gple_gate_schedule_automatic(0, 128, None, None, 0) == 6
    assert lib.gple_gate_schedule_automatic(0, -1, None, None, 0) < 0
    assert lib.gple_gate_schedule_automatic(0, 16, None, None, 4) < 0


def _load_sim():
    spec = importlib.util.spec_from_file_location("gate_schedule_sim", os.path.join(ROOT, "profiles", "gate_schedule_sim.py"))
    sim = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sim)
    return sim


def test_offline_model_restates_the_reference_kernels():
    """profiles/gate_schedule_sim.py builds the (composite) covariances itself; they must be the oracle's."""
    import numpy as np

    from gaussian_process_liouville_equation_b200 import synthetic as syn
    from oracle import oracle as orc

    sim = _load_sim()
    X, y = syn.training_set(5, 1, 96, (0.0, syn.P0))
    Xq = syn.extra_points(5, 1, X, 40, (0.0, syn.P0))[0]
    tr = np.array([1.0, 0.9 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 2e-2])
    tc = np.array([1.0, 1.2, 0.8 * syn.SIGMA_X, 1.1 * syn.SIGMA_P, 0.7, 1.1 * syn.SIGMA_X, 0.9 * syn.SIGMA_P, 2e-2])
    K, ks, prior, noise = sim.real_model(X, Xq, tr)
    assert np.abs(K - orc.kernel_real(X, X, tr, True, False)).max() < 1e-14 and np.abs(ks - orc.kernel_real(Xq, X, tr, False, False)).max() < 1e-14
    assert prior == pytest.approx(K[0, 0], rel=1e-15) and noise == pytest.approx((tr[0] * tr[3]) ** 2, rel=1e-15)
    Kc, kc, prior_c, noise_c = sim.complex_model(X, Xq, tc)
    Ko, Kto = orc.kernel_complex(X, X, tc, True, False)
    ref = 0.5 * np.block([[(Ko + Kto).real, (-Ko + Kto).imag], [(Ko + Kto).imag, (Ko - Kto).real]])  # covariance of [Re y; Im y]
    assert np.abs(Kc - ref).max() < 1e-14 and prior_c == pytest.approx(Ko[0, 0].real, rel=1e-15)
    ko, kto = orc.kernel_complex(Xq, X, tc, False, False)
    rows = 0.5 * np.block([[(ko + kto).real, (-ko + kto).imag], [(ko + kto).imag, (ko - kto).real]])
    assert np.abs(kc[0::2] - rows[: len(Xq)]).max() < 1e-14 and np.abs(kc[1::2] - rows[len(Xq):]).max() < 1e-14


def test_staged_bound_never_contradicts_the_reference_gate():
    """The three bounds of the bound-gated variance, evaluated on the host along the library's automatic
    schedule, against the reference's own cutoff (kernel.h:301-332 through the oracle): whatever a stage decides is what the full
    computation gives -- gate exactly 1 or exactly 0 -- and the undecided queries are the ones with a gate below 1."""
    import numpy as np
    import scipy.linalg as sl

    from gaussian_process_liouville_equation_b200 import synthetic as syn
    from oracle import oracle as orc

    n, centre = 640, (0.0, syn.P0)  # five 128-blocks: stages after 1 and 3 blocks, then the rest
    X, y = syn.training_set(9, 0, n, centre)
    Xq = syn.extra_points(9, 0, X, 3000, centre)[0]
    th = syn.theta_real()
    k = orc.TrainingKernel(th, X, y)
    out = k.predict(Xq)
    with np.errstate(divide="ignore", invalid="ignore"):
        # the reference's cutoff factor: the prediction is in rescaled labels (RescaleMaximum / max |y|, kernel.cpp:279-280), the
        # cutoff prediction is scaled back
        gate = np.where(out["pred"] != 0.0, out["cutoff"] * (10.0 / np.abs(y).max()) / out["pred"], np.nan)
    K = np.asarray(k.K)
    f = orc.kernel_real(Xq, X, th, False, False) @ np.asarray(k.v)
    Z = sl.solve_triangular(np.linalg.cholesky(K), orc.kernel_real(Xq, X, th, False, False).T, lower=True)
    prior, noise = K[0, 0], (th[0] * th[3]) ** 2
    f2 = f * f
    one, zero = f2 >= 4.0 * prior, f2 <= 0.5 * noise
    open_ = ~(one | zero)
    schedule = L.gate_schedule_automatic(False, n // 128)
    assert schedule == [(1, 0), (3, 0)]
    q = (Z ** 2).reshape(n // 128, 128, -1).sum(1)
    seen = 0
    acc = np.zeros(len(Xq))
    for re_end, _ in schedule:
        acc = acc + q[seen:re_end].sum(0)
        seen = re_end
        decided = open_ & (f2 >= 4.0 * (prior - acc))
        one |= decided
        open_ &= ~decided
    assert one.sum() > 1000 and open_.sum() > 50 and zero.sum() >= 0
    assert np.all(np.abs(gate[one] - 1.0) < 1e-12)  # decided 1 by a bound: the full computation says 1
    assert np.all(np.nan_to_num(np.abs(gate[zero]), nan=0.0) == 0.0)  # decided 0 by the noise floor
    # among the queries no bound decides are all those whose gate is below 1 and above 0: only the exact variance gives it
    var = prior - (acc + q[seen:].sum(0))
    band = (f2 < 4.0 * var) & (f2 > var)
    assert band.sum() > 10 and np.all(open_[band]) and np.all((gate[band] > 0.0) & (gate[band] < 1.0))
    assert np.abs(var - out["var"]).max() < 1e-9 * prior
