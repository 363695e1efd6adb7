"""Host-side mirror of the reference's gple/opt.h / opt.cpp: bounds, log transforms, the NLopt callbacks
(loose functions and constraints) and the optimisation driver.

The callbacks keep NLopt's shapes (gple/opt.cpp:441, 594, 644, 844, 879): objective(x, grad) -> value with
`grad is None` meaning "no gradient", vector constraint(x, want_grad) -> (residuals, jacobian).  All numerics
behind them run on the GPU through the C-ABI (gple_loose_function, gple_train_*).

NLopt itself is not available in this environment, so the driver uses the equivalent algorithms of
scipy.optimize: Nelder-Mead for the per-element stage (reference: LN_NELDERMEAD, opt.h:51-52), SLSQP with
equality constraints for the diagonal / full stages (reference: AUGLAG_EQ around LD_SLSQP, opt.cpp:333-336,
opt.h:53) and DIRECT-L for the global fallback (reference: GN_DIRECT_L, opt.h:54).  As SURVEY.md section 7 notes,
parity with NLopt is on outcomes (loss, constraint residuals), not on iterates; the deterministic contract is the
callbacks, which are parity-tested against the CPU restatement in tests/.
"""
from __future__ import annotations

import numpy as np

from . import dynamics
from . import predict as pr

AverageTolerance = 0.05  # opt.h:13
InitialMagnitude = 1.0  # opt.cpp:25
InitialNoise = 1e-2  # opt.cpp:27
_BIG = np.finfo(np.float64).max


def make_normal(d):
    """opt.cpp:420-431"""
    return np.where(np.isfinite(d), d, _BIG) if isinstance(d, np.ndarray) else (d if np.isfinite(d) else _BIG)


def calculate_kernel_bounds(lb, ub):
    """opt.cpp:33-60: magnitude and noise pinned, characteristic lengths in [lb, ub]"""
    return (np.array([InitialMagnitude, lb[0], lb[1], InitialNoise]), np.array([InitialMagnitude, ub[0], ub[1], InitialNoise]))


def calculate_complex_kernel_bounds(lb, ub):
    """opt.cpp:66-104"""
    lo = np.array([InitialMagnitude, InitialMagnitude / 10.0, lb[0], lb[1], InitialMagnitude / 10.0, lb[0], lb[1], InitialNoise])
    hi = np.array([InitialMagnitude, InitialMagnitude * 10.0, ub[0], ub[1], InitialMagnitude * 10.0, ub[0], ub[1], InitialNoise])
    return lo, hi


def _log_indices(n):
    # opt.cpp:109-145: complex kernel -> sub-magnitudes (1, 4) and noise (7); real kernel -> noise (3)
    return [1, 4, 7] if n == 8 else [3]


def local_parameter_to_global(param):
    """opt.cpp:109-145"""
    r = np.array(param, dtype=np.float64)
    idx = _log_indices(len(r))
    r[idx] = np.log(r[idx])
    return r


def global_parameter_to_local(param):
    """opt.cpp:197-232"""
    r = np.array(param, dtype=np.float64)
    idx = _log_indices(len(r))
    r[idx] = np.exp(r[idx])
    return r


def local_gradient_to_global(param_local, grad_local):
    """opt.cpp:155-191: d/d(ln p) = p d/dp"""
    r = np.array(grad_local, dtype=np.float64)
    idx = _log_indices(len(r))
    r[idx] = r[idx] * np.asarray(param_local)[idx]
    return r


class Callbacks:
    """The NLopt callbacks of opt.cpp bound to training / extra training sets.  `backend` (tests only) swaps the
    numerics provider; the default is the CUDA library."""

    def __init__(self, TrainingSets, ExtraTrainingSets, Energies=None, TotalEnergy=None, Purity=None, backend=None):
        self.ts, self.ets = TrainingSets, ExtraTrainingSets
        self.Energies, self.TotalEnergy, self.Purity = Energies, TotalEnergy, Purity
        self.backend = backend
        self.numevals = 0

    def loose_function(self, x, element, want_grad=False):
        """opt.cpp:441-482 for one element"""
        self.numevals += 1
        lf = self.backend.loose_function if self.backend else dynamics.loose_function
        return lf(np.asarray(x, dtype=np.float64), self.ts[element], self.ets[element], grad=want_grad)

    def loose_function_global_wrapper(self, x, element, want_grad=False):
        """opt.cpp:489-497"""
        xl = global_parameter_to_local(x)
        if not want_grad:
            return self.loose_function(xl, element)
        v, g = self.loose_function(xl, element, True)
        return v, local_gradient_to_global(xl, g)

    def _sum_loose(self, x, elements, slices, want_grad):
        err, grad = 0.0, np.zeros(len(x))
        for e, sl in zip(elements, slices):
            if self.ts[e] is None:
                continue
            if want_grad:
                v, g = self.loose_function(x[sl], e, True)
                grad[sl] = g
            else:
                v = self.loose_function(x[sl], e)
            err += v
        err = make_normal(err)
        return (err, make_normal(grad)) if want_grad else err

    def diagonal_loose(self, x, want_grad=False):
        """opt.cpp:594-617: x = [theta_00, theta_11]"""
        return self._sum_loose(np.asarray(x), pr.DIAGONAL, (slice(0, 4), slice(4, 8)), want_grad)

    def full_loose(self, x, want_grad=False):
        """opt.cpp:844-870: x = [theta_00, theta_10, theta_11]"""
        return self._sum_loose(np.asarray(x), (0, 1, 2), pr.ELEMENT_SLICES, want_grad)

    def _kernels(self, params, want_grad):
        return pr.TrainingKernels(params, self.ts, False, True, want_grad, self.backend)

    def diagonal_constraints(self, x, num_constraints, want_grad=False):
        """opt.cpp:644-719: residuals [population - 1, energy - E0, (purity - purity0)] and the Jacobian (m x 8)"""
        x = np.asarray(x, dtype=np.float64)
        self.numevals += 1
        k = self._kernels([x[0:4], np.zeros(8), x[4:8]], want_grad)  # construct_all_parameters_from_diagonal (:622-636)
        res = [k.calculate_population() - 1.0, k.calculate_total_energy_average(self.Energies) - self.TotalEnergy]
        if num_constraints == 3:
            res.append(k.calculate_purity() - self.Purity)
        res = make_normal(np.array(res))
        if not want_grad:
            return res
        jac = [k.population_derivative(), k.total_energy_derivative(self.Energies)]
        if num_constraints == 3:
            pd = k.purity_derivative()
            jac.append(np.concatenate([pd[0:4], pd[12:16]]))
        return res, make_normal(np.array(jac))

    def full_constraints(self, x, want_grad=False):
        """opt.cpp:879-929: residuals [population - 1, energy - E0, purity - purity0] and the Jacobian (3 x 16)"""
        x = np.asarray(x, dtype=np.float64)
        self.numevals += 1
        k = self._kernels([x[sl] for sl in pr.ELEMENT_SLICES], want_grad)
        res = make_normal(np.array([k.calculate_population() - 1.0, k.calculate_total_energy_average(self.Energies) - self.TotalEnergy, k.calculate_purity() - self.Purity]))
        if not want_grad:
            return res
        jac = np.zeros((3, pr.NumTotalParameters))
        jac[0, 0:4], jac[0, 12:16] = np.split(k.population_derivative(), 2)
        jac[1, 0:4], jac[1, 12:16] = np.split(k.total_energy_derivative(self.Energies), 2)
        jac[2] = k.purity_derivative()
        return res, make_normal(jac)


class Optimization:
    """opt.h:17-105.  optimize(density, extra_points) -> (error, steps, type) and updates the stored parameters."""

    Default, LocalPrevious, LocalInitial, Global = range(4)
    RelativeTolerance, AbsoluteTolerance, InitialStepSize = 1e-5, 1e-15, 0.5  # opt.cpp:342-355

    def __init__(self, sigma_r0, mass, pes_model, InitialTotalEnergy, InitialPurity, backend=None, max_global_evals=2000):
        self.TotalEnergy, self.Purity, self.mass, self.pes_model = InitialTotalEnergy, InitialPurity, mass, pes_model
        s = np.asarray(sigma_r0, dtype=np.float64)
        self.InitialKernelParameter = np.array([InitialMagnitude, s[0], s[1], InitialNoise])  # opt.cpp:286-305
        self.InitialComplexKernelParameter = np.array([InitialMagnitude, InitialMagnitude, s[0], s[1], InitialMagnitude, s[0], s[1], InitialNoise])  # :306-332
        self.ParameterVectors = self._initial()
        self.backend, self.max_global_evals = backend, max_global_evals

    def _initial(self):
        return [self.InitialKernelParameter.copy(), self.InitialComplexKernelParameter.copy(), self.InitialKernelParameter.copy()]

    def get_parameters(self):
        return self.ParameterVectors

    # ---- stages --------------------------------------------------------------------------------------
    def _nelder_mead(self, cb, e, x0, bounds, is_global=False):
        from scipy.optimize import direct, minimize

        lo, hi = bounds
        free = hi > lo
        if is_global:
            xg0, lg, hg = local_parameter_to_global(x0), local_parameter_to_global(lo), local_parameter_to_global(hi)

            def fg(z):
                x = xg0.copy()
                x[free] = z
                return cb.loose_function_global_wrapper(x, e)

            res = direct(fg, list(zip(lg[free], hg[free])), locally_biased=True, maxfun=self.max_global_evals, f_min_rtol=self.RelativeTolerance)
            x = xg0.copy()
            x[free] = res.x
            return x, float(res.fun), int(res.nfev)

        def f(z):
            x = x0.copy()
            x[free] = z
            return cb.loose_function(x, e)

        z0 = x0[free]
        n = len(z0)
        simplex = np.vstack([z0] + [np.clip(z0 + self.InitialStepSize * np.eye(n)[i], lo[free], hi[free]) for i in range(n)])
        for i in range(n):  # a clipped vertex may coincide with z0: step the other way
            if np.allclose(simplex[i + 1], z0):
                simplex[i + 1] = np.clip(z0 - self.InitialStepSize * np.eye(n)[i] * 0.5, lo[free], hi[free])
        res = minimize(f, z0, method="Nelder-Mead", bounds=list(zip(lo[free], hi[free])),
                       options=dict(xatol=self.RelativeTolerance * max(1e-12, np.abs(z0).max()), fatol=0.0, maxfev=400 * n, initial_simplex=simplex, adaptive=False))
        x = x0.copy()
        x[free] = res.x
        return x, float(res.fun), int(res.nfev)

    def optimize_elementwise(self, cb, params, bounds, is_global=False):
        """opt.cpp:517-587"""
        total, steps = 0.0, []
        for e in range(3):
            if cb.ts[e] is None:
                steps.append(0)
                continue
            params[e], err, nev = self._nelder_mead(cb, e, np.asarray(params[e], dtype=np.float64), bounds[e], is_global)
            total += err
            steps.append(nev)
        return total, steps

    def _slsqp(self, fun, con, x0, lo, hi):
        from scipy.optimize import minimize

        free = hi > lo

        def embed(z):
            x = x0.copy()
            x[free] = z
            return x

        def f(z):
            v, g = fun(embed(z), True)
            return v, g[free]

        cons = dict(type="eq", fun=lambda z: con(embed(z), False), jac=lambda z: con(embed(z), True)[1][:, free])
        res = minimize(f, x0[free], jac=True, method="SLSQP", bounds=list(zip(lo[free], hi[free])), constraints=[cons],
                       options=dict(ftol=self.RelativeTolerance * 1e-3, maxiter=200))
        return embed(res.x), float(res.fun)

    def optimize_diagonal(self, cb, params, bounds, with_purity):
        """opt.cpp:730-800"""
        x0 = np.concatenate([params[0], params[2]])
        lo, hi = np.concatenate([bounds[0][0], bounds[2][0]]), np.concatenate([bounds[0][1], bounds[2][1]])
        m = 3 if with_purity else 2
        n0 = cb.numevals
        x, err = self._slsqp(cb.diagonal_loose, lambda x, g: cb.diagonal_constraints(x, m, g), x0, lo, hi)
        params[0], params[2] = x[0:4].copy(), x[4:8].copy()
        return err, cb.numevals - n0

    def optimize_full(self, cb, params, bounds):
        """opt.cpp:940-1015"""
        x0 = np.concatenate(params)
        lo, hi = np.concatenate([b[0] for b in bounds]), np.concatenate([b[1] for b in bounds])
        n0 = cb.numevals
        x, err = self._slsqp(cb.full_loose, cb.full_constraints, x0, lo, hi)
        for e, sl in enumerate(pr.ELEMENT_SLICES):
            params[e] = x[sl].copy()
        return err, cb.numevals - n0

    # ---- driver (opt.cpp:1019-1392) ------------------------------------------------------------------------
    def optimize(self, density, extra_points):
        ts, ets = pr.construct_training_sets(density), pr.construct_training_sets(extra_points)
        obs = self.backend.observable_sums if self.backend else dynamics.observable_sums

        def sums(e, pes_index):
            return obs(self.pes_model, np.asarray(density[e], dtype=np.float64), self.mass, pes_index)

        Energies = np.array([sums(e, i)[7] / sums(e, i)[0] if ts[e] is not None else 0.0 for i, e in enumerate(pr.DIAGONAL)])  # predict.cpp:182-190
        bounds = []
        for e in range(3):  # opt.cpp:1026-1052
            if ts[e] is not None:
                o, npts = sums(e, 0), len(density[e])
                sd = np.sqrt(o[5:7] / npts - (o[3:5] / npts) ** 2)  # predict.cpp:109-126
                lb, ub = sd / np.sqrt(npts), 2.0 * sd
            else:
                lb, ub = np.full(2, 0.01), np.full(2, 1.0)
            bounds.append(calculate_complex_kernel_bounds(lb, ub) if e == 1 else calculate_kernel_bounds(lb, ub))
        cb = Callbacks(ts, ets, Energies, self.TotalEnergy, self.Purity, self.backend)
        offdiag = ts[1] is not None

        def move_into_bounds(pv):
            for e in range(3):
                pv[e] = np.clip(pv[e], bounds[e][0], bounds[e][1])

        def do_optimize(pv, opt_type):  # opt.cpp:1101-1198
            for e in range(3):
                pv[e] = np.array(pv[e], dtype=np.float64)
                pv[e][0] = InitialMagnitude
            move_into_bounds(pv)
            err, steps = self.optimize_elementwise(cb, pv, bounds)
            if offdiag:
                _, ds = self.optimize_diagonal(cb, pv, bounds, with_purity=False)
                err, fs = self.optimize_full(cb, pv, bounds)
                steps += [ds, fs]
            else:
                err, ds = self.optimize_diagonal(cb, pv, bounds, with_purity=True)
                steps += [ds, 0]
            k = pr.TrainingKernels(pv, ts, False, False, False, self.backend)
            for e in range(3):  # opt.cpp:1179-1195
                if k[e] is not None:
                    pv[e][0] = k[e].get_magnitude()
            return [err, steps, opt_type]

        def check_averages(pv):  # opt.cpp:1200-1270
            k = pr.TrainingKernels(pv, ts, False, True, False, self.backend)

            def beyond(calc, ref):
                err = abs(calc / ref - 1.0)
                return 0.0 if err < AverageTolerance else err

            return np.array([beyond(k.calculate_population(), 1.0), beyond(k.calculate_total_energy_average(Energies), self.TotalEnergy), beyond(k.calculate_purity(), self.Purity)])

        def compare_and_overwrite(result, check, result_new, check_new, pv_new):  # opt.cpp:1272-1318
            better = int(np.sum((check_new < check) & (check > 2.0 * AverageTolerance)))
            worse = int(np.sum((check_new > check) & (check_new > 2.0 * AverageTolerance)))
            if better > worse or (better == worse and (check_new.sum() < check.sum() or result_new[0] < result[0])):
                self.ParameterVectors = pv_new
                result[0] = result_new[0]
                result[1] = [a + b for a, b in zip(result[1], result_new[1])]
                result[2] = result_new[2]
                return check_new
            return check

        # 1. previous parameters
        result = do_optimize(self.ParameterVectors, self.LocalPrevious)
        check = check_averages(self.ParameterVectors)
        if not check.any():
            return tuple(result), check
        # 2. initial parameters
        pv = self._initial()
        r2 = do_optimize(pv, self.LocalInitial)
        check = compare_and_overwrite(result, check, r2, check_averages(pv), pv)
        if not check.any():
            return tuple(result), check
        # 3. global search per element (log space), then local polish
        pv = self._initial()
        move_into_bounds(pv)
        _, gsteps = self.optimize_elementwise(cb, pv, bounds, is_global=True)
        pv = [global_parameter_to_local(p) if ts[e] is not None else p for e, p in enumerate(pv)]
        r3 = do_optimize(pv, self.Global)
        r3[1] = [a + b for a, b in zip(r3[1], gsteps + [0, 0])]
        check = compare_and_overwrite(result, check, r3, check_averages(pv), pv)
        return tuple(result), check
