"""In-tree build of libhnh_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build().

The library is placed next to this file (distributed_sddmm_b200/libhnh_b200.so): git-ignored,
but it travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libhnh_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-fopenmp",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def sources() -> list[str]:
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "**", "*.h*"), recursive=True) + \
        glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "**", "*.h*"), recursive=True) + [__file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and not _stale():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
        if verbose:
            sys.stderr.write(out.decode())
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", OUT, *objs, "-lnccl", "-lgomp", "-ldl"]
    subprocess.check_call(link)
    # stand-alone C++ driver with the reference's command line (tools/bench_er.cpp)
    exe = os.path.join(HERE, "bench_er")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-ccbin", "/usr/bin/g++", "-Wno-deprecated-gpu-targets",
                           "-I" + os.path.join(ROOT, "include"), os.path.join(HERE, "tools", "bench_er.cpp"),
                           "-o", exe, "-L" + HERE, "-lhnh_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN"])
    build_reference_driver()
    return OUT


def build_reference_driver() -> str | None:
    """Drop-in proof: compile the REFERENCE's own bench_erdos_renyi.cpp, unchanged (fed through stdin so
    that its `#include "benchmark_dist.hpp"` resolves to include/hnh/compat), against this library.  Only
    where /root/reference exists; the binary (git-ignored) travels to the GPU box."""
    src = "/root/reference/bench_erdos_renyi.cpp"
    if not os.path.exists(src):
        return None
    exe = _compile_reference_main(src, "bench_er_reference_main")
    # the reference's two other drivers (MatrixMarket file input; sweep over R), same recipe
    # ... and its self-check program (fingerprints of sddmmA / spmmA / spmmB on dummyInitialize inputs, then a GAT pass)
    for other, name in (("bench_file.cpp", "bench_file_reference_main"), ("bench_heatmap.cpp", "bench_heatmap_reference_main"),
                        ("scratch.cpp", "scratch_reference_main")):
        path = os.path.join("/root/reference", other)
        if os.path.exists(path):
            try:
                _compile_reference_main(path, name)
            except subprocess.CalledProcessError as e:  # the headline driver above is the required one
                sys.stderr.write(f"build: {other} did not compile against include/hnh/compat ({e})\n")
    # the reference's benchmark HARNESS as well: bench_erdos_renyi.cpp + its own benchmark_dist.cpp (benchmark_algorithm:
    # algorithm selection, inputs, trial loop, FLOP model, JSON record), both unchanged, on this library's classes.  The
    # executable's benchmark_algorithm takes precedence over the library's.
    try:
        _compile_reference_main(src, "bench_er_reference_harness", extra=["/root/reference/benchmark_dist.cpp"])
    except subprocess.CalledProcessError as e:
        sys.stderr.write(f"build: the reference's benchmark_dist.cpp did not compile against include/hnh/compat ({e})\n")
    # the reference's ALS (als_conjugate_gradients.cpp, unchanged: Eigen expressions, the host loop of scale_matrix_rows,
    # raw MPI_Allreduce) under this repo's small driver, on this library's classes (host-access mode, hnh/runtime.h)
    try:
        _compile_reference_main(os.path.join(HERE, "tools", "als_reference_main.cpp"), "als_reference_main",
                                extra=["/root/reference/als_conjugate_gradients.cpp"])
    except subprocess.CalledProcessError as e:
        sys.stderr.write(f"build: the reference's als_conjugate_gradients.cpp did not compile against include/hnh/compat ({e})\n")
    return exe


def _compile_reference_main(src: str, name: str, extra: list[str] | None = None) -> str:
    exe = os.path.join(HERE, name)
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = []
    for i, path in enumerate([src] + list(extra or [])):
        obj = os.path.join(HERE, "build", f"{name}.{i}.o")
        with open(path, "rb") as f:  # through stdin: the source's directory must not shadow the compat headers
            subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-x", "c++", "-c", "-o", obj,
                                   "-I" + os.path.join(ROOT, "include", "hnh", "compat"), "-I" + os.path.join(ROOT, "include"),
                                   "-I/usr/local/cuda/include", "-"], stdin=f, cwd="/tmp")
        objs.append(obj)
    # sources that instantiate the algorithm classes (header-inline code) call the CUDA runtime directly
    cudart = ["-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread"] if extra else []
    subprocess.check_call(["/usr/bin/g++", "-o", exe, *objs, "-L" + HERE, "-lhnh_b200", "-Wl,-rpath,$ORIGIN", *cudart])
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
