"""Build librcwa_b200.so (sm_100a, in-tree) and, for the CPU test-suite only, the host emulation
library of the phase-structured single-CTA kernels.

    python -m torcwa_b200.build            # product library (nvcc)
    python -m torcwa_b200.build --emu      # tests/_emu/librcwa_emu.so (g++ -DRCWA_EMU)
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "librcwa_b200.so")
EMU_LIB = os.path.join(ROOT, "tests", "_emu", "librcwa_emu.so")
CU_FILES = ["zgemm.cu", "tc_gemm.cu", "convmat.cu", "assemble.cu", "lu.cu", "hess.cu", "eig.cu", "api.cu"]
EMU_FILES = ["lu.cu", "eig.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _sources():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "rcwa_b200.h"))
    return hdrs


def build(force=False, verbose=False, defines=(), suffix=""):
    """Compile every CUDA translation unit for sm_100a and link librcwa_b200.so in-tree.
    `defines` / `suffix` build an experimental variant (e.g. defines=("HB_NB=64",), suffix="_hb64") next to it;
    torcwa_b200._lib loads it when RCWA_B200_LIB names it."""
    srcs = [os.path.join(CSRC, f) for f in CU_FILES]
    LIB = os.path.join(HERE, "librcwa_b200%s.so" % suffix)
    stamp = LIB + ".stamp"
    dig = _digest(srcs + _sources()) + "".join(defines)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    bdir = os.path.join(HERE, "build" + suffix)
    os.makedirs(bdir, exist_ok=True)
    nvcc = _nvcc()

    def one(src):
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(one, srcs))
    r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


def build_emu(force=False):
    """Host emulation of the single-CTA kernels (TEST INFRASTRUCTURE; never loaded by the product)."""
    srcs = [os.path.join(CSRC, f) for f in EMU_FILES]
    os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
    stamp = EMU_LIB + ".stamp"
    dig = _digest(srcs + _sources())
    if not force and os.path.exists(EMU_LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return EMU_LIB
    cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DRCWA_EMU", "-x", "c++"] + srcs + ["-o", EMU_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
