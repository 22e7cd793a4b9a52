"""input.hydro reader (reader part): the C++ header suhmo_b200/host/suhmo_inputs.hpp and its independent Python
twin suhmo_b200/inputs.py must understand the committed sample and -- where the reference tree is mounted (this container only;
nothing here runs on the GPU box) -- every input.hydro the reference ships, identically."""
import glob
import json
import os
import subprocess

import pytest

from suhmo_b200 import inputs, synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "inputs_dump")
SAMPLE = os.path.join(ROOT, "tests", "data", "input.sample.hydro")
REF_INPUTS = sorted(glob.glob("/root/reference/exec/**/input*.hydro", recursive=True))


@pytest.fixture(scope="module")
def exe():
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "inputs_dump.cpp"), "-o", EXE])
    return EXE


def cpp_read(exe, path, step):
    return json.loads(subprocess.run([exe, path, str(step)], capture_output=True, text=True, check=True).stdout)


def agree(cpp, py, step):
    head, gap = inputs.solver_blocks(step)
    for k in ("problem_type", "domain_size", "num_cells", "dx", "is_periodic", "bc", "params", "picard", "moulins", "mesh", "controls"):
        assert cpp[k] == py[k], (k, cpp[k], py[k])
    assert cpp["head_solver"] == head and cpp["gap_solver"] == gap


def test_sample_input(exe):
    py = inputs.read(SAMPLE)
    for step in (0, 49, 50, 3000):
        agree(cpp_read(exe, SAMPLE, step), py, step)
    assert py["bc"] == {"lo_type": [0, 0], "hi_type": [1, 0], "lo_val": [0.25, 0.0], "hi_val": [-0.5, 0.0]}   # y is periodic: no y values read
    assert py["params"]["use_NL"] == 1 and py["params"]["bcoeff_otf"] == 1
    assert py["picard"]["use_ImplDiff"] == 0            # the later definition wins
    assert py["mesh"]["ref_ratios"] == [2, 2] and py["mesh"]["max_base_grid_size"] == 16   # defaults to max_box_size
    assert len(py["moulins"]) == 2 and py["moulins"][1] == [40.5, 4.5, 12.0, 2.0]
    # what sg::AmrHydroControls::setParams hands the C++ driver class (regrid / run controls, tagging variables in file order)
    assert py["controls"]["tag_vars"] == [["meltingRate", 0.02, 1.0e8, 6, 0]] and py["controls"]["domain0"] == [0, 0, 31, 7]
    assert py["controls"]["fixed_dt"] == 3600.0 and py["controls"]["regrid_interval"] == 10000000 and py["controls"]["eps_PicardIte"] == 1.0e-4
    cfg = inputs.to_config(py)
    c1 = syn.config("C1", 1)
    for k in ("ibc", "nx", "ny", "domain_size", "periodic", "bc_lo", "bc_hi", "max_box_size", "block_factor", "A", "omega", "nu", "H", "slope"):
        assert getattr(cfg, k) == getattr(c1, k), k
    prm, bc, pic = inputs.to_ctypes(py)
    assert prm.use_NL == 1 and bc.hi_type[0] == 1 and pic.n_moulins == 2


def test_nl_switches_follow_the_reference_nesting(exe, tmp_path):
    """use_NL is only read under use_fas, bcoeff_otf only under use_NL (src/AmrHydro.cpp:876-881)"""
    text = open(SAMPLE).read().replace("solver.use_fas = true", "solver.use_fas = false")
    p = tmp_path / "in.hydro"
    p.write_text(text)
    py = inputs.read(str(p))
    assert py["params"]["use_NL"] == 0 and py["params"]["bcoeff_otf"] == 0
    agree(cpp_read(exe, str(p), 0), py, 0)


def test_missing_get_key_aborts(exe, tmp_path):
    text = "\n".join(l for l in open(SAMPLE).read().splitlines() if not l.startswith("suhmo.A "))
    p = tmp_path / "in.hydro"
    p.write_text(text)
    with pytest.raises(KeyError):
        inputs.read(str(p))
    r = subprocess.run([exe, str(p)], capture_output=True, text=True)
    assert r.returncode != 0 and "suhmo.A" in r.stderr


@pytest.mark.skipif(not REF_INPUTS, reason="reference tree not mounted (GPU box)")
def test_every_reference_input(exe):
    assert len(REF_INPUTS) >= 5
    kinds = set()
    for path in REF_INPUTS:
        py = inputs.read(path)
        agree(cpp_read(exe, path, 0), py, 0)
        cfg = inputs.to_config(py)
        assert cfg.nx > 0 and cfg.ny > 0 and cfg.dx[0] > 0
        kinds.add(py["problem_type"])
    assert {"basic", "sqrt", "valley"} <= kinds
