"""Builds the in-tree CUDA libraries for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d):
    out = []
    for r, _, fs in os.walk(d):
        out += [os.path.join(r, f) for f in fs if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp"))]
    return out


def build_all(force=False, verbose=False):
    libdir = os.path.join(_HERE, "lib")
    os.makedirs(libdir, exist_ok=True)
    csrc = os.path.join(_HERE, "csrc")
    deps = _sources(csrc) + _sources(os.path.join(ROOT, "include"))
    targets = [("libexab200.so", ["exab200_capi.cu"], [])]
    if os.path.exists(os.path.join(csrc, "host_sim.cu")):
        targets.append(("libexahost.so", ["host_sim.cu"], ["-L" + libdir, "-lexab200", "-Xlinker", "-rpath=$ORIGIN"]))
    for name, srcs, extra in targets:
        out = os.path.join(libdir, name)
        if force or _stale(out, deps):
            cmd = [NVCC] + FLAGS + ["-o", out] + [os.path.join(csrc, s) for s in srcs] + extra
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    # the application driver (`mechanics -opt options.toml`): host-only C++ on top of the two libraries
    main_src = os.path.join(csrc, "mechanics_main.cpp")
    if os.path.exists(main_src):
        out = os.path.join(libdir, "mechanics")
        if force or _stale(out, deps + [os.path.join(libdir, "libexahost.so")]):
            cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-o", out, main_src, "-L" + libdir, "-lexahost",
                   "-lexab200", "-Wl,-rpath,$ORIGIN"]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
    return libdir
