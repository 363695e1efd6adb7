"""Offline model of the staged bound of the bound-gated variance (csrc/gpr.cu, gate_schedule): where to put the stage
boundaries.  Pure numpy / scipy on the host, no GPU, no library code: the kernel formulas of real_spec / complex_spec are
restated below.  For a sample of the bench's queries it computes Z = L^-1 k*^T once, the per-128-block partial sums of Z^2, and
from them the fraction of queries still undecided after any cumulative tile set ("survival").  The cost of a schedule is
   sum over stages of  survival(before the stage) x [ sum over the stage's tiles of (tile + 1)  +  KSTAR x columns generated ],
in units of one 128 x 128 x 128 tile product per row (tile t of the triangular product costs t + 1; KSTAR = 0.26 is the measured
cost of generating 128 columns of K* relative to one tile product).  A dynamic programme over the lattice of cumulative
(Re blocks, Im blocks) gives the best schedule for a number of stages; the automatic rule of gate_schedule() is evaluated
next to it.  Predicted -21 % on the C2 workload against the two-boundary schedule, measured -19 % (profiles/r02_gate_schedule.md).

usage: python profiles/gate_schedule_sim.py N [queries per element] [real|complex|both]"""
import functools
import math
import os
import sys

import numpy as np
import scipy.linalg as sl

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (workload constants only)
from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402

KSTAR = 0.26


def gauss(A, B, mag2, lx, lp):
    d = ((A[:, None, 0] - B[None, :, 0]) / lx) ** 2 + ((A[:, None, 1] - B[None, :, 1]) / lp) ** 2
    return mag2 * np.exp(-0.5 * d)


def real_model(X, Xq, th):
    K = gauss(X, X, th[0] ** 2, th[1], th[2]) + (th[0] * th[3]) ** 2 * np.eye(len(X))
    return K, gauss(Xq, X, th[0] ** 2, th[1], th[2]), th[0] ** 2 * (1 + th[3] ** 2), (th[0] * th[3]) ** 2


def complex_model(X, Xq, th):
    """composite [Re; Im] process (complex_spec): RR = s2 sr^2 G(lr), II = s2 si^2 G(li), RI = s2 sc^2 G(lc), noise / 2 on each part"""
    s2, sr, lr, si, li = th[0] ** 2, th[1], th[2:4], th[4], th[5:7]
    ss = lr ** 2 + li ** 2
    lc = np.sqrt(ss / 2)
    sc2 = sr * si * np.prod(2 * lr * li / ss)
    hn = 0.5 * s2 * th[7] ** 2

    def blocks(A, B):
        return gauss(A, B, s2 * sr * sr, *lr), gauss(A, B, s2 * si * si, *li), gauss(A, B, s2 * sc2, *lc)

    rr, ii, ri = blocks(X, X)
    N = len(X)
    K = np.block([[rr + hn * np.eye(N), ri], [ri.T, ii + hn * np.eye(N)]])
    qr, qi, qc = blocks(Xq, X)
    ks = np.stack([np.hstack([qr, qc]), np.hstack([qc, qi])], axis=1).reshape(2 * len(Xq), 2 * N)  # rows (Re, Im) per query
    return K, ks, s2 * (sr * sr + si * si) + 2 * hn, s2 * th[7] ** 2


def survival(K, ks, label, prior, noise, nb):
    n = K.shape[0]
    T, Th = n // 128, n // 128 // nb
    L = np.linalg.cholesky(K)
    Z = sl.solve_triangular(L, ks.T, lower=True, check_finite=False)
    f = ks @ sl.cho_solve((L, True), label)
    Q = ks.shape[0] // nb
    f2 = (f * f).reshape(Q, nb).sum(1)
    qq = (Z ** 2).reshape(T, 128, Q, nb).sum((1, 3))
    open0 = ~((f2 >= 4 * prior) | (f2 <= 0.5 * noise))
    cre = np.concatenate([np.zeros((1, Q)), np.cumsum(qq[:Th], axis=0)])
    cim = np.concatenate([np.zeros((1, Q)), np.cumsum(qq[Th:], axis=0)]) if nb == 2 else np.zeros((1, Q))
    alive = np.array([[(open0 & ~(f2 >= 4 * (prior - cre[a] - cim[b]))).mean() for b in range(len(cim))] for a in range(Th + 1)])
    alive[0, 0] = open0.mean()
    return alive, Th


def schedule_cost(alive, Th, nb, stages):
    Ti = Th if nb == 2 else 0
    a = b = 0
    total = 0.0
    for c, d in list(stages) + [(Th, Ti)]:
        c, d = max(a, min(c, Th)), max(b, min(d, Ti))
        tiles = list(range(a, c)) + [Th + j for j in range(b, d)]
        if tiles:
            total += nb * alive[a, b] * (sum(t + 1 for t in tiles) + KSTAR * (max(tiles) + 1))
        a, b = c, d
    return total


def automatic(Th, nb):
    """gate_schedule() of csrc/gpr.cu"""
    out, late = [], 21 * Th // 32
    if late >= 1:
        out.append((1, 0))
        k = max(1, round(math.log(late) / math.log(2.4)))
        for i in range(1, k + 1):
            b = max(1, round(late ** (i / k)))
            out.append((b, max(1, b // 5)))
    if nb == 2:
        out.append((Th, max(1, 3 * Th // 8)))
    return out


def best_schedules(alive, Th, nb, max_stages):
    Ti = Th if nb == 2 else 0
    step_re = max(1, Th // 32)  # coarser lattice for many blocks
    res = sorted(set(list(range(0, min(Th, 8))) + list(range(0, Th + 1, step_re)) + [Th]))
    ims = sorted(set(list(range(0, min(Ti, 4) + 1)) + list(range(0, Ti + 1, max(1, Ti // 8))) + [Ti]))

    def stage(a, b, c, d):
        tiles = list(range(a, c)) + [Th + j for j in range(b, d)]
        return nb * alive[a, b] * (sum(t + 1 for t in tiles) + KSTAR * (max(tiles) + 1))

    @functools.lru_cache(None)
    def best(a, b, k):
        if (a, b) == (Th, Ti):
            return 0.0, ()
        if k == 1:
            return stage(a, b, Th, Ti), ()
        out = None
        for c in res:
            for d in ims:
                if c < a or d < b or (c, d) == (a, b) or (c, d) == (Th, Ti):
                    continue
                rest, path = best(c, d, k - 1)
                tot = stage(a, b, c, d) + rest
                if out is None or tot < out[0]:
                    out = (tot, ((c, d),) + path)
        direct = (stage(a, b, Th, Ti), ())
        return direct if out is None or direct[0] <= out[0] else out

    return [(k,) + best(0, 0, k) for k in range(1, max_stages + 1)]


def main():
    N = int(sys.argv[1])
    Qs = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    sets = [syn.training_set(2, e, N, bench.CENTRE) for e in range(3)]
    Xq = np.concatenate([syn.extra_points(2, e, sets[e][0], Qs, bench.CENTRE)[0] for e in range(3)])
    jobs = []
    if which in ("real", "both"):
        X, y = sets[0]
        K, ks, prior, noise = real_model(X, Xq, bench.THETA_R)
        jobs.append(("real element rho00", K, ks, 10.0 / np.abs(y).max() * y.real, prior, noise, 1))
    if which in ("complex", "both"):
        X, y = sets[1]
        K, ks, prior, noise = complex_model(X, Xq, np.asarray(bench.THETA_C))
        s = 10.0 / np.abs(y).max()
        jobs.append(("complex element rho10", K, ks, np.concatenate([s * y.real, s * y.imag]), prior, noise, 2))
    for name, K, ks, label, prior, noise, nb in jobs:
        alive, Th = survival(K, ks, label, prior, noise, nb)
        Ti = Th if nb == 2 else 0
        print(f"## {name}, N = {N}, {len(Xq)} queries: open after the cheap bounds {alive[0, 0]:.4f}, after every tile {alive[Th, Ti]:.4f}")
        print("   floor (undecidable rows x full product):", round(nb * alive[Th, Ti] * (nb * Th) * (nb * Th + 1) / 2, 2))
        t4 = max(2, min(8, Th // 4))
        print("   round-1 schedule  ", [(t4, max(1, t4 // 4))], round(schedule_cost(alive, Th, nb, [(t4, max(1, t4 // 4))]), 2))
        print("   three-stage       ", [(t4, max(1, t4 // 4)), (5 * Th // 8, max(1, t4 // 4))], round(schedule_cost(alive, Th, nb, [(t4, max(1, t4 // 4)), (5 * Th // 8, max(1, t4 // 4))]), 2))
        print("   automatic (round 2)", automatic(Th, nb), round(schedule_cost(alive, Th, nb, automatic(Th, nb)), 2))
        for k, cost, path in best_schedules(alive, Th, nb, 7 if Th <= 32 else 6):
            print(f"   best with {k} stage(s): {cost:.2f}", list(path))


if __name__ == "__main__":
    main()
