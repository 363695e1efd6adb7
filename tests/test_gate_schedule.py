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
