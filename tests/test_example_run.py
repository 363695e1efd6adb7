"""examples/mqcle_run.cpp: the reference's main loop (gple/main.cpp:19-202) on the C++ host headers -- Metropolis selection,
optimisation, evolve / is_very_small / new-element selection / re-optimisation / model rebuild per tick.  CPU: it compiles and
links.  GPU: a short run far from the crossing conserves population, energy and purity and never populates the other
elements; a run started at the crossing populates them (new_element_point_selection) and keeps the trace near 1."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "mqcle_run")
LIBDIR = os.path.join(ROOT, "gaussian_process_liouville_equation_b200")


def build():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", "-pthread", os.path.join(ROOT, "examples", "mqcle_run.cpp"), "-o", EXE, f"-L{LIBDIR}", "-lgple_b200", f"-Wl,-rpath,{LIBDIR}"])


def parse(out):
    ticks, info = [], {}
    for line in out.strip().splitlines():
        w = line.split()
        if not w or w[0] == "NCCL":  # the NCCL banner ("NCCL version ...") goes to stdout on rank 0
            continue
        if w[0] == "tick":
            ticks.append([float(v) for v in w[2:]])
        else:
            info[w[0]] = [float(v) for v in w[1:]]
    return ticks, info


def run(*args):
    return parse(subprocess.run([EXE, *map(str, args)], capture_output=True, text=True, check=True, timeout=600).stdout)


def run_ranks(world, *args):
    """`world` copies of the example, one per GPU, bootstrapped through a file (GPLE_COMM_FILE); returns rank 0's report"""
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        procs = []
        for r in range(world):
            env = dict(os.environ, GPLE_RANK=str(r), GPLE_WORLD_SIZE=str(world), GPLE_LOCAL_DEVICE=str(r), GPLE_COMM_FILE=os.path.join(d, "nccl_id"))
            procs.append(subprocess.Popen([EXE, *map(str, args)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=900) for p in procs]
        for p, (o, e) in zip(procs, outs):
            assert p.returncode == 0, e[-2000:]
    return parse(outs[0][0])


def test_main_loop_example_compiles_and_links():
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_main_loop_far_from_the_crossing():
    build()
    ticks, info = run(96, 5, 0, 1, 7)
    assert len(ticks) == 6 and info["elements"] == [96.0, 0.0, 0.0]
    pop0, e0, pur0 = ticks[0]
    assert abs(pop0 - 1.0) < 0.1 and abs(pur0 - 1.0) < 0.15
    for pop, e, pur in ticks:
        assert abs(pop - pop0) < 0.05 and abs(e / e0 - 1.0) < 0.02 and abs(pur - pur0) < 0.1
    assert 0.05 <= info["displacement"][0] <= 5.0 and 1 <= info["mc_steps"][0] < 1000


@pytest.mark.gpu
def test_main_loop_at_the_crossing_populates_the_other_elements():
    build()
    ticks, info = run(64, 3, 0, 1, 7, -0.3)
    assert len(ticks) == 4
    assert info["elements"][0] == 64.0 and info["elements"][1] == 64.0 and info["elements"][2] == 64.0
    assert info["optimisations"][0] >= 2
    for pop, e, pur in ticks:
        assert 0.7 < pop < 1.3


def _gpu_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2])
def test_closed_loop_on_two_gpus_is_bit_identical_to_one_gpu(world):
    """The closed loop of gple/main.cpp:135-202 through the C++ host on `world` GPUs (evolve shard -> all-gather inside the
    library -> construct_training_sets -> TrainingKernels rebuild, with new-element selection and re-optimisation on the way):
    every point of every tick and every printed observable equals the one-GPU run bit for bit."""
    build()
    args = (64, 4, 2, 1, 7, -0.3)
    one, info1 = run(*args)
    many, infoN = run_ranks(world, *args)
    assert infoN["ranks"] == [float(world)] and info1["ranks"] == [1.0]
    assert one == many
    hashes = [k for k in info1 if k.startswith("hash")]
    assert len(hashes) == 4 and all(info1[k] == infoN[k] for k in hashes)
    assert info1["elements"] == infoN["elements"] and info1["optimisations"] == infoN["optimisations"]
