"""Compile csrc/*.cu into lib/libebfi_b200.so for sm_100a with nvcc (no torch involved).

    python ebfi-be_b200/build.py [--force] [--verbose]

The library has a plain C ABI (include/ebfi_b200.h) and links the CUDA runtime
statically, so it can be dlopen'ed from PyTorch, ctypes, cgo or anything else.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libebfi_b200.so")
SELFTEST_LIB = os.path.join(LIBDIR, "libebfi_b200_selftest.so")      # hardware probes, test-only (include/ebfi_b200_selftest.h)
SELFTEST_SOURCES = {"tcgemm_selftest.cu"}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
         "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--expt-relaxed-constexpr",
         "-Xptxas", "-warn-spills"]
if os.environ.get("EBFI_DEBUG_HANG"):       # debugging build: mbarrier waits trap with a location instead of hanging
    FLAGS.append("-DEBFI_DEBUG_HANG")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def deps():
    out = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "ebfi_b200.h"))
    out.append(os.path.abspath(__file__))
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [d for d in deps() if not d.endswith(".cu")]
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or any(
            os.path.getmtime(d) > os.path.getmtime(obj) for d in [src] + hdrs)
        if stale:
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0 or "warning" in out or "spill" in out.lower()):
            print(f"--- {os.path.basename(src)}\n{out}", file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    st_objs = [o for o in objs if os.path.basename(o)[:-2] + ".cu" in SELFTEST_SOURCES]
    objs = [o for o in objs if o not in st_objs]
    if procs or not os.path.exists(LIB) or not os.path.exists(SELFTEST_LIB):
        arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
        subprocess.run([NVCC, "-shared", "-o", LIB] + objs + arch, check=True)
        subprocess.run([NVCC, "-shared", "-o", SELFTEST_LIB] + st_objs + arch +
                       ["-L", LIBDIR, "-lebfi_b200", "-Xlinker", "-rpath=$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
