"""Build oracle/_ref/ from the UNMODIFIED reference sources where they lie.

TEST INFRASTRUCTURE. Produces, when /root/reference exists (this container; the
GPU box only ever uses the prebuilt files that travel with the snapshot):

  oracle/_ref/dcn_cpu/_ext*.so              reference CPU DCNv2
        sources: models/DCNv2/src/cpu/dcn_v2_cpu.cpp, dcn_v2_im2col_cpu.cpp
        (+ oracle/ref_shim/dcn_cpu_bind.cpp, a 2-function pybind TU, and the
        TH/TH.h shim that maps THArgCheck onto TORCH_CHECK)
  oracle/_ref/fac_cuda/kernelconv2d_cuda*.so  reference FAC CUDA op for sm_100a
        sources: models/FAC/kernelconv2d/KernelConv2D_cuda.cpp, KernelConv2D_kernel.cu
  oracle/_ref/dcn_cuda/_ext_cuda_ref*.so      reference DCNv2 CUDA kernels for sm_100a
        sources: models/DCNv2/src/cuda/dcn_v2_im2col_cuda.cu
        (+ oracle/ref_shim/dcn_cuda_driver.cu: the reference's own host driver
        dcn_v2_cuda.cu needs THCState and cannot be compiled as shipped)

Nothing is copied out of /root/reference; the outputs are git-ignored binaries.
The reference's own build system (setup.py) is not run.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EBFI_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _torch_flags(cuda):
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths("cuda" if cuda else "cpu")] + \
          [f"-I{sysconfig.get_paths()['include']}"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    libs = ["-L" + libdir, "-Wl,-rpath," + libdir, "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python"]
    if cuda:
        libs += ["-lc10_cuda", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart"]
    defs = ["-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    return inc, libs, defs


def _suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def _fresh(target, deps):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(d) for d in deps)


def build_dcn_cpu():
    src = os.path.join(REF, "models/DCNv2/src")
    srcs = [os.path.join(src, "cpu/dcn_v2_cpu.cpp"), os.path.join(src, "cpu/dcn_v2_im2col_cpu.cpp"),
            os.path.join(SHIM, "dcn_cpu_bind.cpp")]
    d = os.path.join(OUT, "dcn_cpu"); os.makedirs(d, exist_ok=True)
    target = os.path.join(d, "_ext" + _suffix())
    if _fresh(target, srcs):
        return target
    inc, libs, defs = _torch_flags(False)
    _run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w", "-DTORCH_EXTENSION_NAME=_ext"] + defs +
         [f"-I{SHIM}", f"-I{src}"] + inc + srcs + ["-o", target] + libs)
    return target


def build_fac_cuda():
    src = os.path.join(REF, "models/FAC/kernelconv2d")
    srcs = [os.path.join(src, "KernelConv2D_cuda.cpp"), os.path.join(src, "KernelConv2D_kernel.cu")]
    d = os.path.join(OUT, "fac_cuda"); os.makedirs(d, exist_ok=True)
    target = os.path.join(d, "kernelconv2d_cuda" + _suffix())
    if _fresh(target, srcs):
        return target
    inc, libs, defs = _torch_flags(True)
    _run(["nvcc", "-O3", "-shared", "-std=c++17", "-w", "-Xcompiler", "-fPIC", "-x", "cu",
          "-DTORCH_EXTENSION_NAME=kernelconv2d_cuda"] + ARCH + defs + [f"-I{src}"] + inc + srcs +
         ["-o", target] + [a if not a.startswith("-Wl,") else "-Xlinker=" + a[4:] for a in libs])
    return target


def build_dcn_cuda():
    src = os.path.join(REF, "models/DCNv2/src")
    srcs = [os.path.join(src, "cuda/dcn_v2_im2col_cuda.cu"), os.path.join(SHIM, "dcn_cuda_driver.cu")]
    d = os.path.join(OUT, "dcn_cuda"); os.makedirs(d, exist_ok=True)
    target = os.path.join(d, "_ext_cuda_ref" + _suffix())
    if _fresh(target, srcs):
        return target
    inc, libs, defs = _torch_flags(True)
    _run(["nvcc", "-O3", "-shared", "-std=c++17", "-w", "-Xcompiler", "-fPIC",
          "-DTORCH_EXTENSION_NAME=_ext_cuda_ref"] + ARCH + defs + [f"-I{SHIM}", f"-I{src}"] + inc + srcs +
         ["-o", target] + [a if not a.startswith("-Wl,") else "-Xlinker=" + a[4:] for a in libs])
    return target


def main(which=("dcn_cpu", "fac_cuda", "dcn_cuda")):
    if not os.path.isdir(REF):
        print(f"{REF} not present: keeping whatever prebuilt files exist under {OUT}")
        return 0
    rc = 0
    for name in which:
        try:
            print("built", globals()["build_" + name]())
        except subprocess.CalledProcessError as e:
            print(f"FAILED {name}: {e}", file=sys.stderr)
            rc = 1
    return rc


if __name__ == "__main__":
    sys.exit(main(tuple(sys.argv[1:]) or ("dcn_cpu", "fac_cuda", "dcn_cuda")))
