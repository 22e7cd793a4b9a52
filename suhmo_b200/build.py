"""Builds libsuhmo_gpu.so in-tree (suhmo_b200/lib/) with nvcc for sm_100a.  No JIT, no torch extension."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SRC = [os.path.join(_HERE, "csrc", f) for f in ("sg_api.cu", "sg_kernels.cuh", "sg_twin.cuh", "sg_general.cuh", "sg_picard.cuh", "sg_general_host.inc",
                                                 "sg_picard_host.inc", "sg_regrid.inc", "sg_linear.cuh", "sg_linear_host.inc", "sg_nccl.h")]
HDR = os.path.join(ROOT, "include", "suhmo_gpu.h")
SO = os.path.join(_HERE, "lib", "libsuhmo_gpu.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def nccl_include():
    for cand in ("/usr/include",):
        if os.path.exists(os.path.join(cand, "nccl.h")):
            return cand
    try:
        import nvidia.nccl
        return os.path.join(list(nvidia.nccl.__path__)[0], "include")
    except Exception:
        return None


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in SRC + [HDR])


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libsuhmo_gpu.so cannot be built and there is no CPU fallback")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS
    inc = nccl_include()
    if inc and inc != "/usr/include":
        cmd += ["-I", inc]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", SO, SRC[0], "-ldl"]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
