"""CPU: the C-ABI library loads, exports every symbol include/ebfi_b200.h declares, and the host
side fails loudly (no CPU fallback). No kernel is launched here."""
import ctypes
import os
import re
import sys

import pytest
import torch

from conftest import ROOT

import ebfi_be_b200
from ebfi_be_b200 import _lib as L


def _declared_symbols(header="ebfi_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ebfi_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(L.LIB_PATH), "run `python ebfi-be_b200/build.py`"
    assert os.path.commonpath([ROOT, L.LIB_PATH]) == ROOT


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(L.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ebfi_b200.h but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes prototype in _lib.SIGNATURES"
    assert sorted(L.SIGNATURES) == declared


def test_selftest_probes_live_in_their_own_library():
    """The tcgen05 probes are test infrastructure: declared in include/ebfi_b200_selftest.h, exported by
    libebfi_b200_selftest.so only."""
    main, st = ctypes.CDLL(L.LIB_PATH), ctypes.CDLL(L.SELFTEST_LIB_PATH)
    declared = [s for s in _declared_symbols("ebfi_b200_selftest.h") if s.startswith("ebfi_selftest_")]
    assert sorted(L.SELFTEST_SIGNATURES) == declared and len(declared) == 4
    for name in declared:
        assert hasattr(st, name) and not hasattr(main, name)
    assert not any(s.startswith("ebfi_selftest_") for s in _declared_symbols())


def test_abi_version_and_error_string():
    lib = L.load()
    assert lib.ebfi_abi_version() == 3
    assert isinstance(lib.ebfi_last_error(), bytes)


def test_geometry_is_validated_without_a_gpu():
    lib = L.load()
    ho, wo = L.c_int(), L.c_int()
    g = L.DcnGeom(1, 64, 256, 256, 64, 3, 3, 1, 1, 1, 1, 1, 1, 8)
    assert lib.ebfi_dcnv2_output_size(g, ho, wo) == 0 and (ho.value, wo.value) == (256, 256)
    g = L.DcnGeom(1, 8, 12, 10, 8, 3, 3, 2, 2, 1, 1, 1, 1, 4)
    assert lib.ebfi_dcnv2_output_size(g, ho, wo) == 0 and (ho.value, wo.value) == (6, 5)
    bad = L.DcnGeom(1, 6, 8, 8, 4, 3, 3, 1, 1, 1, 1, 1, 1, 4)          # 6 % 4 != 0
    assert lib.ebfi_dcnv2_output_size(bad, ho, wo) == -1
    assert b"deformable_group" in lib.ebfi_last_error()
    assert lib.ebfi_dcnv2_backward_workspace_bytes(L.DcnGeom(1, 64, 256, 256, 64, 3, 3, 1, 1, 1, 1, 1, 1, 8)) > 0
    assert lib.ebfi_fac_backward_workspace_bytes(4, 64, 256, 256, 5) > 0


def test_null_and_bad_arguments_return_error_codes():
    lib = L.load()
    assert lib.ebfi_fac_forward(None, None, None, None, 1, 1, 8, 8, 5) == -1
    assert lib.ebfi_fac_forward(None, None, None, None, 1, 1, 8, 8, 4) == -1      # even K
    assert b"odd" in lib.ebfi_last_error()
    assert lib.ebfi_events_to_voxel(None, None, None, None, None, 7, 10, 5, 4, 4, None, 0) == -1


def test_ops_refuse_cpu_tensors():
    from ebfi_be_b200 import dcn_v2, encodings, kernelconv2d
    x = torch.randn(1, 2, 6, 6)
    with pytest.raises(NotImplementedError):                       # KernelConv2D.py:38-39
        kernelconv2d.KernelConv2DFunction.apply(x, torch.randn(1, 18, 4, 4), 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):          # dcn_v2_cuda.cu:38
        dcn_v2.dcn_v2_conv(x, torch.zeros(1, 18, 6, 6), torch.ones(1, 9, 6, 6),
                           torch.randn(2, 2, 3, 3), torch.zeros(2), 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        encodings.events_to_image(torch.zeros(3), torch.zeros(3), torch.ones(3), (4, 4))
    # the widenings refuse CPU tensors the same way (nothing in the product path falls back to the host)
    from ebfi_be_b200 import frame_ops, modification
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        dcn_v2.dcn_v2_conv_packed(x, torch.zeros(1, 27, 6, 6), torch.randn(2, 2, 3, 3), torch.zeros(2), 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        encodings.events_raw_to_stack(torch.zeros(4, dtype=torch.int16), torch.zeros(4, dtype=torch.int16),
                                      torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.int8), 4, (4, 4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        modification.kernelconv_fac_fused(torch.randn(1, 32, 8, 8), torch.randn(1, 32, 8, 8),
                                          torch.randn(800, 64, 3, 3), torch.zeros(800), 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        frame_ops.Frame2Lap(torch.rand(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        frame_ops.Frame2DCP(torch.rand(1, 3, 8, 8))


def test_product_code_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under ebfi-be_b200/ may import, load or execute it."""
    pkg = os.path.join(ROOT, "ebfi-be_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|oracle/_ref|libebfi_oracle|oracle\.oracle", text, re.M):
                    offenders.append(os.path.relpath(os.path.join(dirpath, f), ROOT))
    assert offenders == []


def test_module_parameters_match_reference_layout():
    from ebfi_be_b200 import dcn_v2
    m = dcn_v2.DCN_sep(64, 64, 3, stride=1, padding=1, dilation=1, deformable_groups=8)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    # dcn_v2.py:118-119,203-205 -> checkpoint compatibility
    assert shapes == {"weight": (64, 64, 3, 3), "bias": (64,),
                      "conv_offset_mask.weight": (216, 64, 3, 3), "conv_offset_mask.bias": (216,)}
    assert float(m.conv_offset_mask.weight.abs().sum()) == 0.0 and float(m.bias.abs().sum()) == 0.0
    assert float(m.weight.abs().max()) <= 1.0 / (64 * 9) ** 0.5


def test_psroi_pooling_is_declared_out_of_scope():
    _ext, _ = ebfi_be_b200.install_shims()
    with pytest.raises(NotImplementedError):
        _ext.dcn_v2_psroi_pooling_forward()


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not present")
def test_reference_wrappers_import_unchanged_on_our_shims():
    """models/DCNv2/dcn_v2.py and models/FAC/kernelconv2d/KernelConv2D.py import `_ext` /
    `kernelconv2d_cuda` at module load; with the shims installed they import as they are."""
    import importlib.util
    ebfi_be_b200.install_shims()
    for rel, attr in (("models/DCNv2/dcn_v2.py", "DCN_sep"),
                      ("models/FAC/kernelconv2d/KernelConv2D.py", "KernelConv2D")):
        spec = importlib.util.spec_from_file_location("ref_" + attr, os.path.join("/root/reference", rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert hasattr(mod, attr)
    sys.modules.pop("_ext", None), sys.modules.pop("kernelconv2d_cuda", None)


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/ebfi_b200.h compiles as C99 and as C++17 (no torch, no CUDA headers), and a C program linked against
    libebfi_b200.so resolves the entry points — what a cgo / JNI / pybind maintainer relies on."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "ebfi_b200.h"\n'
                   "int main(void) {\n"
                   "    ebfi_dcn_geom g = {1, 64, 256, 256, 64, 3, 3, 1, 1, 1, 1, 1, 1, 8, 0};\n"
                   "    int ho = 0, wo = 0;\n"
                   "    if (ebfi_abi_version() != EBFI_ABI_VERSION) return 1;\n"
                   "    if (ebfi_dcnv2_output_size(&g, &ho, &wo) != EBFI_OK || ho != 256 || wo != 256) return 2;\n"
                   "    return ebfi_kpn_fused_workspace_bytes(4, 64, 64, 256, 256, 5) > 0 ? 0 : 3;\n"
                   "}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, "-c", str(src), "-o", str(tmp_path / "abi.o")], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-x", "c++", "-c", str(src), "-o", str(tmp_path / "abi_cxx.o")], check=True)
    exe = tmp_path / "abi"
    subprocess.run(["gcc", str(tmp_path / "abi.o"), "-L", libdir, "-lebfi_b200", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_dp_comm_validates_without_a_gpu():
    """The peer-memory all-reduce rejects malformed communicators before touching the device."""
    lib = L.load()
    assert lib.ebfi_dp_comm_bytes(36928) >= 512 + 2 * 36928 * 4
    bad = L.DpComm(9, 0, (L.c_void * 8)(), 1 << 20)
    assert lib.ebfi_dp_allreduce_sum(None, bad, None, 0, None, 0) == -1 and b"world" in lib.ebfi_last_error()
    small = L.DpComm(1, 0, (L.c_void * 8)(256), 256)
    assert lib.ebfi_dp_allreduce_sum(None, small, None, 0, None, 0) == -1 and b"symmetric" in lib.ebfi_last_error()
    assert lib.ebfi_dp_allreduce_sum(None, None, None, 0, None, 0) == -1
