"""TEST / BASELINE INFRASTRUCTURE ONLY.  A best-effort CPU implementation of the same path on the host's BLAS / LAPACK (numpy + scipy:
potrf, potri, batched GEMM variance), next to the reference-shaped oracle port (per-query GEMV variance, solve(Identity) inverse,
unblocked pivoted LDL^T) -- SURVEY.md section 8d asks for both CPU numbers.  It follows the same formulas as the product's composite
form and is checked against the oracle in tests/test_oracle_real.py::test_best_effort_cpu_matches_the_oracle."""
import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular


def gauss(XL, XR, mag, lx, lp):
    dx = (XL[:, None, 0] - XR[None, :, 0]) / lx
    dp = (XL[:, None, 1] - XR[None, :, 1]) / lp
    return mag * mag * np.exp(-0.5 * (dx * dx + dp * dp))


def complex_blocks(theta):
    """(sigma_R, l_R), (sigma_I, l_I) and the derived correlation kernel (complex_kernel.cpp:142-157)"""
    _, sr, lrx, lrp, si, lix, lip, _ = theta
    lr, li = np.array([lrx, lrp]), np.array([lix, lip])
    ss = lr * lr + li * li
    lc = np.sqrt(ss / 2.0)
    sc = np.sqrt(sr * si * np.prod(2.0 * lr * li / ss))
    return (sr, lr), (si, li), (sc, lc)


class RealModel:
    """kernel.cpp:244-335 with LAPACK: Cholesky, v, diag(K^-1), LOOCV error, population"""

    def __init__(self, theta, X, y):
        sf, lx, lp, sn = theta
        self.theta, self.X = theta, X
        lab = y.real
        self.rescale = 10.0 / np.abs(lab).max()
        lab = lab * self.rescale
        K = gauss(X, X, sf, lx, lp) + (sf * sn) ** 2 * np.eye(len(X))
        self.c = cho_factor(K, lower=True, overwrite_a=True, check_finite=False)
        self.v = cho_solve(self.c, lab, check_finite=False)
        Linv = solve_triangular(self.c[0], np.eye(len(X)), lower=True, check_finite=False)
        self.Linv = Linv
        kinv_diag = np.einsum("ij,ij->j", Linv, Linv)
        self.error = float(np.sum((self.v / kinv_diag) ** 2))
        self.population = float(2.0 * np.pi * sf * sf * lx * lp * self.v.sum() / self.rescale)
        self.prior = sf * sf * (1.0 + sn * sn)

    def predict(self, Xq):
        sf, lx, lp, _ = self.theta
        Ks = gauss(Xq, self.X, sf, lx, lp)
        f = Ks @ self.v
        Z = Ks @ self.Linv.T
        var = self.prior - np.einsum("ij,ij->i", Z, Z)
        return f, var


class ComplexModel:
    """complex_kernel.cpp:221-286 as the composite [Re f; Im f] process with LAPACK"""

    def __init__(self, theta, X, y):
        s, sn = theta[0], theta[7]
        self.theta, self.X = theta, X
        self.rescale = 10.0 / np.abs(y).max()
        lab = np.concatenate([y.real, y.imag]) * self.rescale
        (sr, lr), (si, li), (sc, lc) = complex_blocks(theta)
        n = len(X)
        eye = 0.5 * sn * sn * np.eye(n)
        KR, KI, KC = gauss(X, X, sr, *lr), gauss(X, X, si, *li), gauss(X, X, sc, *lc)
        Cov = s * s * np.block([[KR + eye, KC], [KC, KI + eye]])
        self.c = cho_factor(Cov, lower=True, overwrite_a=True, check_finite=False)
        self.w = cho_solve(self.c, lab, check_finite=False)
        self.Linv = solve_triangular(self.c[0], np.eye(2 * n), lower=True, check_finite=False)
        M_diag = np.einsum("ij,ij->j", self.Linv, self.Linv)
        cross = np.einsum("ij,ij->j", self.Linv[:, :n], self.Linv[:, n:])
        # complex_kernel.cpp:270-286 in composite form (P_ii, Q_ii from the diagonals of M's blocks)
        P = 0.25 * (M_diag[:n] + M_diag[n:])
        Q = 0.25 * (M_diag[:n] - M_diag[n:]) - 0.5j * cross
        v = 0.5 * (self.w[:n] + 1j * self.w[n:])
        self.error = float(np.sum(np.abs((P * v - np.conj(Q * v)) / (P * P - np.abs(Q) ** 2)) ** 2))
        self.prior = s * s * (sr * sr + si * si + sn * sn)

    def predict(self, Xq):
        s = self.theta[0]
        (sr, lr), (si, li), (sc, lc) = complex_blocks(self.theta)
        kR, kI, kC = gauss(Xq, self.X, sr, *lr), gauss(Xq, self.X, si, *li), gauss(Xq, self.X, sc, *lc)
        cr, ci = s * s * np.hstack([kR, kC]), s * s * np.hstack([kC, kI])
        f = cr @ self.w + 1j * (ci @ self.w)
        Zr, Zi = cr @ self.Linv.T, ci @ self.Linv.T
        var = self.prior - np.einsum("ij,ij->i", Zr, Zr) - np.einsum("ij,ij->i", Zi, Zi)
        return f, var
