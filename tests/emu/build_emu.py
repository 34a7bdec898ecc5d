"""TEST INFRASTRUCTURE.  Builds tests/emu/_build/libfi_emu.so: the library's .cu sources compiled with g++ against the
CPU functional emulator of the CUDA execution model (cuda_emu.hpp), for kernel-logic tests in the GPU-less build container."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "field_interpolation_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libfi_emu.so")
# every source of the library (stencil_tma.cu: TMA loads become synchronous box copies, see its FI_B200_EMU hooks)
CU = ["abi.cu", "assembly.cu", "sort_scan.cu", "stencil.cu", "stencil_fast.cu", "stencil_tma.cu", "stencil_2d.cu", "solver.cu", "mg.cu", "errormap.cu", "isosurface.cu", "dist.cu"]
CPP = ["cuda_emu.cpp", "emu_glue.cpp"]


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in CU] + [os.path.join(HERE, f) for f in CPP]
    deps = srcs + [os.path.join(HERE, "cuda_emu.hpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # -ffp-contract=off: the emulated kernels keep the scalar fp32 order of the -fmad=false CUDA build
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-w",
           "-include", os.path.join(HERE, "cuda_emu.hpp"), "-I", "/usr/local/cuda/include", "-I", os.path.join(ROOT, "include"), "-x", "c++", *srcs, "-o", OUT]
    import concurrent.futures as cf
    objdir = os.path.join(os.path.dirname(OUT), "obj")
    os.makedirs(objdir, exist_ok=True)
    base = cmd[:cmd.index("-x")]
    base = [c for c in base if c != "-shared"]

    def one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        hdrs = [d for d in deps if d.endswith((".hpp", ".cuh"))] + [os.path.join(ROOT, "include", "fi_b200.h")]
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in [src] + hdrs):
            return obj, None
        return obj, subprocess.run([*base, "-x", "c++", "-c", src, "-o", obj], capture_output=True, text=True)

    objs = []
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for obj, r in ex.map(one, srcs):
            if r is not None and r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("emulator build failed")
            objs.append(obj)
    r = subprocess.run([base[0], "-shared", "-Wl,-Bsymbolic", "-o", OUT, *objs, "-ldl", "-lpthread", "-lrt"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("emulator link failed")
    return OUT


FAKE_NCCL_DIR = os.path.join(HERE, "_build", "fake_nccl")


def build_fake_nccl(force: bool = False) -> str:
    """tests/emu/fake_nccl.cpp -> _build/fake_nccl/libnccl.so.2; returns the directory for LD_LIBRARY_PATH."""
    src, out = os.path.join(HERE, "fake_nccl.cpp"), os.path.join(FAKE_NCCL_DIR, "libnccl.so.2")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(FAKE_NCCL_DIR, exist_ok=True)
        r = subprocess.run([os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-I", "/usr/local/cuda/include",
                            src, "-o", out, "-lrt", "-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("fake nccl build failed")
    return FAKE_NCCL_DIR


if __name__ == "__main__":
    build_fake_nccl(force="-f" in sys.argv)
    print(build(force="-f" in sys.argv))
