"""(Named to run last in the GPU suite: it compiles a host program.)  AmrHydro::timeStepFAS as C++ host code (suhmo_b200/host/suhmo_amrhydro.hpp) on the GPU against the CPU oracle's independent
restatement (oracle/picard_amr.py): the oracle runs whole time steps here and writes inputs and outcomes to a fixture
(tests/amr_timestep_fixture.py); the C++ program tests/cpp/timestep_host.cpp replays it on the device and asserts the same Picard
iteration counts, V-cycle counts, convergence measures and BIT-identical head and gap height.  Without a GPU: the program compiles
with -Wall -Wextra -Werror, links, and reads back what the writer wrote."""
import os
import subprocess

import pytest

from suhmo_b200 import build
from suhmo_b200 import synthetic as syn
from tests.amr_timestep_fixture import write_fixture
from tests.problem import amr_hierarchy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "timestep_host")


def compile_host():
    build.build()
    libdir = os.path.join(ROOT, "suhmo_b200", "lib")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "timestep_host.cpp"), "-L", libdir, "-lsuhmo_gpu", f"-Wl,-rpath,{libdir}", "-o", EXE])
    return EXE


def one_level(name="C2"):
    """a single-level problem at the configuration's native size (exec/1_convergence_distributed: 64 x 16, periodic in y), two boxes"""
    cfg = syn.config(name, 1)
    cfg.max_box_size = 32
    return cfg, [syn.domain_split(cfg.nx, cfg.ny, 32, cfg.block_factor)]


def _rg(var, val_min, val_max):
    return dict(var=var, val_min=val_min, val_max=val_max, fill_ratio=0.7, tags_grow=1, grow_dir=(0, 0), block_factor=8, nesting_radius=2,
                max_box_size=32, max_level=2)


CASES = {
    "C5_3lev_first_steps": lambda: (amr_hierarchy("C5"), dict(cur_step=1, nsteps=2)),     # m_cur_step < 2: more than two Picard iterations
    "C5_3lev_late_steps": lambda: (amr_hierarchy("C5"), dict(cur_step=60, nsteps=2)),     # m_cur_step >= 50: eps_PicardIte, bottom 16
    "C4_2lev_valley": lambda: (amr_hierarchy("C4"), dict(cur_step=10, nsteps=1)),         # ice mask < 0, masked gradients
    # AmrHydro::regrid between the steps: new shapes on both refined levels, then the finest level disappears, level 1 moves, the finest level comes back
    "C5_regrid_between_steps": lambda: (amr_hierarchy("C5"), dict(cur_step=5, nsteps=5, regrid_before={1: _rg("Pi", 5.0e6, 1e30), 2: _rg("Pi", 0.0, 1.6e6),
                                                                                                       3: _rg("Pi", 5.0e6, 1e30), 4: _rg("Pi", 5.0e6, 1e30)})),
    "C2_1lev_implicit_gap": lambda: (one_level("C2"), dict(cur_step=3, nsteps=2, impl_diff=True)),
}


def test_cpp_timestep_compiles_and_reads_its_fixture(tmp_path):
    exe = compile_host()
    (cfg, lv), kw = CASES["C5_3lev_first_steps"]()
    path = str(tmp_path / "ts.bin")
    reports = write_fixture(path, cfg, lv, **kw)
    r = subprocess.run([exe, path, "--parse-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"3 levels, {64 * 64 + 4 * 32 * 32 + 3 * 32 * 32} cells, 2 steps from m_cur_step 1" in r.stdout, r.stdout
    assert r.stdout.rstrip().endswith(" ".join(str(x["picard_iterations"]) for x in reports)), r.stdout
    assert reports[0]["picard_iterations"] > 3       # the reference's "more than two iterations" rule of the first steps
    (cfg, lv), kw = CASES["C5_regrid_between_steps"]()
    reports = write_fixture(path, cfg, lv, **kw)
    assert [len(x["boxes"]) for x in reports] == [3, 3, 2, 2, 3] and reports[1]["boxes"] != reports[0]["boxes"], [x["boxes"] for x in reports]
    r = subprocess.run([exe, path, "--parse-only"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "5 steps from m_cur_step 5" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_cpp_timestep_bit_exact(tmp_path, case):
    exe = compile_host()
    (cfg, lv), kw = CASES[case]()
    path = str(tmp_path / "ts.bin")
    write_fixture(path, cfg, lv, **kw)
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "timestep_host: OK" in r.stdout and "0 + 0 differ" in r.stdout, r.stdout
