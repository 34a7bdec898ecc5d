"""Builds field_interpolation_b200/libfi_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch
extension machinery: the library is a plain C-ABI shared object)."""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libfi_b200.so")
OBJ = os.path.join(HERE, "_build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--extended-lambda",
          "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", os.path.join(HERE, "..", "include")]
# assembly.cu and isosurface.cu restate the reference's scalar fp32 arithmetic bit for bit: no FMA contraction there.
PER_FILE = {"assembly.cu": ["-fmad=false"], "isosurface.cu": ["-fmad=false"]}
SOURCES = ["abi.cu", "assembly.cu", "sort_scan.cu", "stencil.cu", "stencil_fast.cu", "stencil_tma.cu", "stencil_2d.cu", "solver.cu", "mg.cu", "errormap.cu", "isosurface.cu", "dist.cu"]


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "fi_b200.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    jobs = []
    for name in srcs:
        src, obj = os.path.join(CSRC, name), os.path.join(OBJ, name[:-3] + ".o")
        if force or _stale(src, obj):
            cmd = [NVCC, *COMMON, *PER_FILE.get(name, []), "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((name, cmd))
    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for name, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- {name}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {name}")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_host()
    return OUT


HOST_SRC = [os.path.join(HERE, "host", f) for f in ("field_interpolation.cpp", "sparse_linear.cpp", "iso_surface.cpp")]
HOST_OUT = os.path.join(HERE, "libfield_interpolation.so")
CXX = os.environ.get("CXX", "g++")


def build_host(force: bool = False) -> str:
    """The C++ drop-in API (include/field_interpolation/*.hpp): host C++14 over the C ABI, linked to libfi_b200.so."""
    deps = HOST_SRC + [os.path.join(HERE, "host", "structured.hpp"), OUT,
                       os.path.join(HERE, "..", "include", "fi_b200.h"),
                       os.path.join(HERE, "..", "include", "field_interpolation", "field_interpolation.hpp"),
                       os.path.join(HERE, "..", "include", "field_interpolation", "sparse_linear.hpp"),
                       os.path.join(HERE, "..", "include", "field_interpolation", "iso_surface.hpp"),
                       os.path.join(HERE, "..", "include", "emilib", "marching_squares.hpp")]
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_OUT) for d in deps):
        return HOST_OUT
    cmd = [CXX, "-std=c++14", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-o", HOST_OUT, *HOST_SRC,
           "-L", HERE, "-lfi_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("host library build failed")
    return HOST_OUT


def build_cpp_driver(src: str, out: str) -> str:
    """Compiles a C++ program against the drop-in headers (what a user of the reference would do)."""
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(HOST_OUT)):
        return out
    cmd = [CXX, "-std=c++14", "-O2", "-Wall", "-o", out, src, "-I", os.path.join(HERE, "..", "include"), "-L", HERE,
           "-lfield_interpolation", "-lfi_b200", f"-Wl,-rpath,{HERE}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("driver build failed")
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
