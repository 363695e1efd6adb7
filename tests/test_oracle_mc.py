"""Pins the oracle's Metropolis sampler (oracle/gple_oracle_mc.hpp): the Philox4x32-10 known-answer vectors of Random123,
the chain autocorrelation against a direct numpy evaluation, detailed-balance statistics of the walk, and the host-side
tuning logic (mc.py) run on the oracle-backed sampler."""
import numpy as np
import pytest

import oracle_backend
from gaussian_process_liouville_equation_b200 import mc
from gaussian_process_liouville_equation_b200 import synthetic as syn
from oracle import oracle as orc

ANALYTIC = ((syn.X0, syn.P0), (syn.SIGMA_X, syn.SIGMA_P), (0.8, 0.6), (0.0, 0.4))


def u53(hi, lo):
    return float(((hi >> 5) << 26) | (lo >> 6)) / 2.0 ** 53


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10: (ctr, key) -> output words"""
    # ctr = 0, key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8  (block 0 of chain 0, step 0, stream 0, seed 0)
    u = orc.philox_draws(0, 0, 0, 0)
    assert u[0] == u53(0x6627E8D5, 0xE169C58D) and u[1] == u53(0xBC57AC4C, 0x9B00DBD8)
    # ctr = ffffffff x 4, key = ffffffff x 2 -> 408f276d 41c83b0e a20bc7c6 6d5451fd  (block 1: stream << 1 | 1 = ffffffff)
    u = orc.philox_draws(0xFFFFFFFFFFFFFFFF, 0x7FFFFFFF, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFF)
    assert u[2] == u53(0x408F276D, 0x41C83B0E)
    # ctr = 243f6a88 85a308d3 13198a2e 03707344, key = a4093822 299f31d0 -> d16cfe09 94fdcceb 5001e420 24126ea1
    u = orc.philox_draws(0x299F31D0A4093822, 0x03707344 >> 1, 0x85A308D3243F6A88, 0x13198A2E)
    assert u[0] == u53(0xD16CFE09, 0x94FDCCEB) and u[1] == u53(0x5001E420, 0x24126EA1)


def test_chain_semantics_and_autocorrelation():
    g = syn.rng(80, 0)
    pts = np.zeros((40, 4))
    pts[:, 0] = syn.X0 + syn.SIGMA_X * g.standard_normal(40)
    pts[:, 1] = syn.P0 + syn.SIGMA_P * g.standard_normal(40)
    s = oracle_backend.Sampler(7, analytic=ANALYTIC)
    out, acc, chains = s.chains(pts, 60, 0.5, 0, 0, want_chain=True)
    assert chains.shape == (40, 61, 2) and np.array_equal(chains[:, 0], pts[:, :2]) and np.array_equal(chains[:, -1], out[:, :2])
    moved = (np.abs(np.diff(chains, axis=1)).sum(axis=2) > 0).sum(axis=1)
    assert np.array_equal(moved, np.rint(acc * 60).astype(int))  # the acceptance ratio counts the moves
    assert np.abs(np.diff(chains, axis=1)).max() <= 0.5  # uniform displacement in (-d, d) per dimension
    rho = mc.initial_distribution(*ANALYTIC[:2], out[:, :2], 0, 0, ANALYTIC[2], ANALYTIC[3])
    assert np.allclose(out[:, 2] + 1j * out[:, 3], rho, rtol=1e-13)  # relabelled with the density at the last state
    # same seed / stream -> same chains; another stream -> different chains
    again, _, _ = oracle_backend.Sampler(7, analytic=ANALYTIC).chains(pts, 60, 0.5, 0, 0)
    other, _, _ = s.chains(pts, 60, 0.5, 0, 0)
    assert np.array_equal(again, out) and not np.array_equal(other, out)
    ac = s.autocorrelation(chains)
    c = chains - chains.mean(axis=1, keepdims=True)
    ref = np.array([np.mean([(c[k, :61 - j] * c[k, j:]).sum() / (61 - j) for k in range(40)]) for j in range(30)])
    assert np.allclose(ac, ref, rtol=1e-11, atol=1e-14)


def test_walk_keeps_the_target_distribution():
    """Started from exact samples of |rho10| (a Gaussian), the walk must leave mean and spread unchanged (detailed balance)."""
    g = syn.rng(81, 0)
    n = 4000
    pts = np.zeros((n, 4))
    pts[:, 0] = syn.X0 + syn.SIGMA_X * g.standard_normal(n)
    pts[:, 1] = syn.P0 + syn.SIGMA_P * g.standard_normal(n)
    out, acc, _ = oracle_backend.Sampler(11, analytic=ANALYTIC).chains(pts, 100, 0.6, 1, 0)
    assert 0.15 < acc.mean() < 0.9
    se = 4.0 / np.sqrt(n)
    assert abs(out[:, 0].mean() - syn.X0) < se * syn.SIGMA_X and abs(out[:, 1].mean() - syn.P0) < se * syn.SIGMA_P
    assert out[:, 0].std() == pytest.approx(syn.SIGMA_X, rel=0.06) and out[:, 1].std() == pytest.approx(syn.SIGMA_P, rel=0.06)
    assert np.allclose(np.angle(out[:, 2] + 1j * out[:, 3]), 0.4)  # phase[row] - phase[col] of element (1, 0)

def test_monte_carlo_selection_logic_on_the_oracle_sampler():
    g = syn.rng(82, 0)
    n = 24
    pts = np.zeros((n, 4))
    pts[:, 0] = syn.X0 + syn.SIGMA_X * g.standard_normal(n)
    pts[:, 1] = syn.P0 + syn.SIGMA_P * g.standard_normal(n)
    params = [mc.MCParameters(), mc.MCParameters(), mc.MCParameters()]
    out = mc.monte_carlo_selection([pts, None, None], params, oracle_backend.Sampler(5, analytic=((syn.X0, syn.P0), (syn.SIGMA_X, syn.SIGMA_P), (1.0, 0.0), (0.0, 0.0))))
    assert out[1] is None and out[2] is None and out[0].shape == (n, 4)
    assert params[0].get_max_displacement() in mc.PossibleDisplacement and 0.05 <= params[0].get_max_displacement() <= 5.0
    assert 1 <= params[0].get_num_MC_steps() < 1000
    rho = mc.initial_distribution((syn.X0, syn.P0), (syn.SIGMA_X, syn.SIGMA_P), out[0][:, :2], 0, 0)
    assert np.allclose(out[0][:, 2], rho.real, rtol=1e-13) and np.all(out[0][:, 3] == 0.0)
