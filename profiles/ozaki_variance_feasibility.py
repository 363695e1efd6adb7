"""Feasibility of the variance GEMM on the int8 tensor path ("FP64 products on the low-precision tensor pipes").

Host-only numerical experiment, no GPU: Z = K* W^T (W = L^-1, lower triangular) is computed by error-free splitting of both
operands into signed 7-bit slices (Ozaki scheme): every row of K* and every row of W is scaled by a power of two to |x| < 1 and cut
into s slices of 7 bits; a slice product is an integer GEMM whose int32 accumulation is exact for K <= 2^31 / 127^2 = 133000, so
the only error is the truncation of the operands after 7 s bits below the row maximum.  Slice pairs (i, j) with i + j >= s are
below that truncation and are dropped: s (s + 1) / 2 integer GEMMs.  The script reports, for the C2 real element at reduced N, the
error of the posterior variance k** - sum Z^2 (relative to k**) against an 80-bit reference, per number of slices, next to
the plain FP64 product -- and the int8 throughput a B200 would need to beat the DMMA kernel.

usage: python profiles/ozaki_variance_feasibility.py [N] [queries] [real|complex]"""
import os
import sys

import numpy as np
import scipy.linalg as sl

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gaussian_process_liouville_equation_b200 import synthetic as syn  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 512
KIND = sys.argv[3] if len(sys.argv) > 3 else "real"
BITS = 7


def slices(A, s):
    """rows scaled to |x| < 1 by a power of two, then s signed slices of BITS bits: A = 2^e * sum_k S_k 2^(-BITS (k + 1)) + remainder"""
    e = np.ceil(np.log2(np.abs(A).max(axis=1, keepdims=True) + 1e-300)) + 1
    R = A / 2.0 ** e  # exact
    out = []
    for _ in range(s):
        R = R * 2.0 ** BITS  # exact
        S = np.trunc(R)
        out.append(S)  # small integers held in float64: their products and sums below stay exact (127^2 K < 2^53)
        R = R - S  # exact
    return e, out


def ozaki_product(A, B, s):
    """A (m x k) times B^T (n x k) from slice products, accumulated in float64 from the highest-order pairs down"""
    ea, Sa = slices(A, s)
    eb, Sb = slices(B, s)
    Z = np.zeros((A.shape[0], B.shape[0]))
    pairs = 0
    for order in range(2 * s - 2, -1, -1):  # small terms first
        if order >= s:
            continue
        acc = np.zeros(Z.shape)
        for i in range(order + 1):
            j = order - i
            if i < s and j < s:
                acc += Sa[i] @ Sb[j].T  # what one int8 tensor GEMM with int32 accumulation computes exactly
                pairs += 1
        Z += acc * 2.0 ** (-BITS * (order + 2))
    return Z * 2.0 ** ea * 2.0 ** eb.T, pairs


def main():
    e = 0 if KIND == "real" else 1
    X, y = syn.training_set(2, e, N, bench.CENTRE)
    Xq = syn.extra_points(2, e, X, Q, bench.CENTRE)[0]

    def gauss(A, B, mag2, lx, lp):
        d = ((A[:, None, 0] - B[None, :, 0]) / lx) ** 2 + ((A[:, None, 1] - B[None, :, 1]) / lp) ** 2
        return mag2 * np.exp(-0.5 * d)

    if KIND == "real":
        sf, lx, lp, sn = bench.THETA_R
        K = gauss(X, X, sf ** 2, lx, lp) + (sf * sn) ** 2 * np.eye(N)
        ks = gauss(Xq, X, sf ** 2, lx, lp)
    else:  # composite [Re; Im] process of the widely-linear complex element (complex_spec of csrc/gpr.cu)
        th = np.asarray(bench.THETA_C)
        s2, sr, lr, si, li = th[0] ** 2, th[1], th[2:4], th[4], th[5:7]
        ss = lr ** 2 + li ** 2
        lc, sc2, hn = np.sqrt(ss / 2), sr * si * np.prod(2 * lr * li / ss), 0.5 * s2 * th[7] ** 2
        blocks = lambda A, B: (gauss(A, B, s2 * sr * sr, *lr), gauss(A, B, s2 * si * si, *li), gauss(A, B, s2 * sc2, *lc))  # noqa: E731
        rr, ii, ri = blocks(X, X)
        K = np.block([[rr + hn * np.eye(N), ri], [ri.T, ii + hn * np.eye(N)]])
        qr, qi, qc = blocks(Xq, X)
        ks = np.vstack([np.hstack([qr, qc]), np.hstack([qc, qi])])  # Re rows, then Im rows
    prior = K[0, 0] if KIND == "real" else K[0, 0] + K[N, N]
    W = sl.solve_triangular(np.linalg.cholesky(K), np.eye(len(K)), lower=True)
    ref = (ks.astype(np.longdouble) @ W.T.astype(np.longdouble))  # 64-bit mantissa
    fold = (lambda v: v) if KIND == "real" else (lambda v: v[: len(v) // 2] + v[len(v) // 2:])  # variance of a complex query: Re + Im rows
    var_ref = prior - fold((ref * ref).sum(1))
    z64 = ks @ W.T
    var64 = prior - fold((z64 * z64).sum(1))
    print(f"# Ozaki splitting of the variance GEMM: {KIND} element, N = {N}, {Q} queries, max |W| = {np.abs(W).max():.3g}, cond(K) = {np.linalg.cond(K):.3g}\n")
    print("| product | integer GEMMs | max error of the variance / k** | max |dZ| |")
    print("|---|---:|---:|---:|")
    print(f"| FP64 (what DMMA computes) | - | {float(np.abs(var64 - var_ref).max() / prior):.2e} | {float(np.abs(z64 - ref).max()):.2e} |")
    for s in (6, 7, 8, 9, 10, 11):
        Z, pairs = ozaki_product(ks, W, s)
        var = prior - fold((Z * Z).sum(1))
        print(f"| {s} slices of {BITS} bits | {pairs} | {float(np.abs(var - var_ref).max() / prior):.2e} | {float(np.abs(Z - ref).max()):.2e} |")
    print("\nBreak-even against the DMMA kernel (32 TFLOP/s executed): an int8 rate of 32 x (integer GEMMs) TOP/s; the nominal dense")
    print("int8 tensor rate of a B200 is 4500 TOP/s.")


if __name__ == "__main__":
    main()
