"""The C++ host layer (suhmo_b200/host/suhmo_gpu.hpp) over the C ABI: compiled with g++ everywhere; on the GPU box the
program drives a two-level head solve the way AmrHydro::SolveForHead_nl drives the reference and asserts BIT equality of the head and
the residual history with the oracle fixture tests/golden/host_smoke_2lev.bin, then drives the remaining wrappers on the device."""
import os
import subprocess

import pytest

from suhmo_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_smoke")


def compile_host():
    build.build()
    libdir = os.path.join(ROOT, "suhmo_b200", "lib")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "host_smoke.cpp"), "-L", libdir, "-lsuhmo_gpu", f"-Wl,-rpath,{libdir}", "-o", EXE]
    subprocess.check_call(cmd)
    return EXE


def test_cpp_host_layer_compiles_and_links():
    exe = compile_host()
    assert os.path.exists(exe)


def test_whole_cpp_mirror_links_without_a_gpu():
    """every wrapper of suhmo_gpu.hpp (operator surface, callbacks, Picard-body kernels, tagging/regrid, gap solver) and
    suhmo_inputs.hpp compiles with -Wall -Wextra -Werror and links; the program only asks the library for its version"""
    build.build()
    libdir = os.path.join(ROOT, "suhmo_b200", "lib")
    exe = os.path.join(ROOT, "tests", "cpp", "api_touch")
    subprocess.check_call(["g++", "-std=c++14", "-O0", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "api_touch.cpp"), "-L", libdir, "-lsuhmo_gpu", f"-Wl,-rpath,{libdir}", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "free functions referenced" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_host_layer_solves_two_levels():
    exe = compile_host()
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "host_smoke_2lev.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_smoke: OK" in r.stdout and "0 differ" in r.stdout
