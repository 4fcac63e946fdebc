"""Build recipe for libfftconv_cuda.so (sm_100a only, in-tree).

`python -m fftconv_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfftconv_cuda.so")
MICROBENCH = os.path.join(LIBDIR, "microbench")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # the CUDA runtime as a shared object (ldconfig knows libcudart.so.12 in this image): nothing
    # that ships with the snapshot then carries the runtime's own symbol-name tables, and a process
    # that already holds a runtime (torch) shares it.  FFTCONV_B200_STATIC_CUDART=1 restores nvcc's default.
    *([] if os.environ.get("FFTCONV_B200_STATIC_CUDART") else ["-cudart", "shared"]),
]

# translation units of libfftconv_cuda.so (compiled in parallel, objects under lib/obj/)
UNITS = ["fftconv_cuda.cu", "fftconv_cuda_w1k.cu", "fftconv_cuda_hr2c.cu", "fftconv_cuda_mgpu.cu"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources() -> list[str]:
    out = [os.path.join(ROOT, "include", "fftconv_cuda.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def _unit_deps(unit: str) -> list[str]:
    """Sources a translation unit is rebuilt for: itself, the public header, and the transitive
    closure of its `#include "..."` lines under csrc/."""
    import re

    hdr = os.path.join(ROOT, "include", "fftconv_cuda.h")
    seen, todo = set(), [unit]
    while todo:
        f = todo.pop()
        if f in seen:
            continue
        seen.add(f)
        path = os.path.join(CSRC, f)
        if not os.path.exists(path):
            continue
        with open(path) as fh:
            for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', fh.read(), flags=re.M):
                if os.path.exists(os.path.join(CSRC, inc)):
                    todo.append(inc)
    return [hdr] + [os.path.join(CSRC, f) for f in sorted(seen)]


def build_lib(force: bool = False, verbose: bool = False) -> str:
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for unit in UNITS:
        obj = os.path.join(objdir, unit.replace(".cu", ".o"))
        objs.append(obj)
        if not force and _newer(obj, _unit_deps(unit)):
            continue
        cmd = ["nvcc", *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, unit)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd)))
    for cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    if force or procs or not _newer(LIB, objs):
        subprocess.run(["nvcc", *NVCC_FLAGS, "-shared", "-o", LIB, *objs, "-ldl"], check=True)
    return LIB


def build_microbench(force: bool = False) -> str:
    src = os.path.join(CSRC, "microbench.cu")
    if not os.path.exists(src):
        return ""
    if not force and _newer(MICROBENCH, [src]):
        return MICROBENCH
    subprocess.run(["nvcc", *NVCC_FLAGS, "-o", MICROBENCH, src], check=True)
    return MICROBENCH


REF_MAIN = "/root/reference/py/main.cpp"
REF_BINDING_DIR = os.path.join(HERE, "refbinding", "pyfftconv")
CPP_TEST = os.path.join(ROOT, "tests", "cpp", "test_headers")


def _headers() -> list[str]:
    inc = os.path.join(ROOT, "include")
    return [os.path.join(inc, "fftconv_cuda.h")] + [
        os.path.join(inc, "fftconv", f) for f in sorted(os.listdir(os.path.join(inc, "fftconv")))]


def build_ref_binding(force: bool = False) -> str:
    """Compile the REFERENCE's pybind11 module, unchanged and where it lies, against
    this repo's headers.  Only possible where /root/reference exists."""
    import sysconfig

    if not os.path.exists(REF_MAIN):
        return ""
    import pybind11

    target = os.path.join(REF_BINDING_DIR, "_pyfftconv" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not force and _newer(target, [REF_MAIN, LIB] + _headers()):
        return target
    cmd = ["g++", "-O2", "-std=c++20", "-shared", "-fPIC",
           f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}",
           f"-I{os.path.join(ROOT, 'include')}", REF_MAIN,
           f"-L{LIBDIR}", "-lfftconv_cuda", "-Wl,-rpath,$ORIGIN/../../lib", "-o", target]
    subprocess.run(cmd, check=True)
    return target


def build_cpp_test(force: bool = False) -> str:
    src = CPP_TEST + ".cpp"
    if not force and _newer(CPP_TEST, [src, LIB] + _headers()):
        return CPP_TEST
    cmd = ["g++", "-O2", "-std=c++20", f"-I{os.path.join(ROOT, 'include')}", src,
           f"-L{LIBDIR}", "-lfftconv_cuda", "-Wl,-rpath,$ORIGIN/../../fftconv_b200/lib", "-o", CPP_TEST]
    subprocess.run(cmd, check=True)
    return CPP_TEST


CPP_BENCH = os.path.join(ROOT, "tests", "cpp", "bench_c1")


def build_cpp_bench(force: bool = False) -> str:
    src = CPP_BENCH + ".cpp"
    if not force and _newer(CPP_BENCH, [src, LIB] + _headers()):
        return CPP_BENCH
    cmd = ["g++", "-O2", "-std=c++20", f"-I{os.path.join(ROOT, 'include')}", src,
           f"-L{LIBDIR}", "-lfftconv_cuda", "-Wl,-rpath,$ORIGIN/../../fftconv_b200/lib", "-o", CPP_BENCH]
    subprocess.run(cmd, check=True)
    return CPP_BENCH


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_microbench(force)
    build_cpp_test(force)
    build_cpp_bench(force)
    build_ref_binding(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB)
