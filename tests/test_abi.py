"""The C-ABI library loads and exports every symbol include/vslam_b200.h declares; struct layouts match; without a
GPU every compute entry fails loudly (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, _has_gpu


def header_functions():
    src = open(os.path.join(ROOT, "include", "vslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vslam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vslam_b200 import _abi

    lib = _abi.lib()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"libvslam_b200.so does not export {name}"
    assert sorted(_abi.SYMBOLS) == declared, "vslam_b200/_abi.py SYMBOLS out of sync with include/vslam_b200.h"


def test_struct_layouts():
    from vslam_b200 import _abi

    assert _abi.KP_DTYPE.itemsize == 28 and _abi.DMATCH_DTYPE.itemsize == 16  # cv::KeyPoint / cv::DMatch
    assert C.sizeof(_abi.Camera) == 40
    assert C.sizeof(_abi.Config) == 4 * 8 + 40 + 4 * 3 + 4 * 8 + 4  # incl. tail padding to 8
    assert C.sizeof(_abi.FrameOut) == 8 + 7 * 8 + 4 * 4 + 8
    assert C.sizeof(_abi.SearchParams) == 56


def test_invalid_config_is_rejected_before_touching_the_device():
    from vslam_b200 import _abi

    lib = _abi.lib()
    h = C.c_void_p()
    cfg = _abi.Config()
    cfg.abi_version = 999
    assert lib.vslam_create(C.byref(cfg), C.byref(h)) == _abi.ERR_INVALID_ARGUMENT
    assert b"ABI version" in lib.vslam_last_error(None)
    cfg.abi_version = _abi.ABI_VERSION
    cfg.width, cfg.height, cfg.n_features = 8, 8, 100
    assert lib.vslam_create(C.byref(cfg), C.byref(h)) == _abi.ERR_INVALID_ARGUMENT
    assert not h.value


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from vslam_b200 import CameraModel, StereoProcessor, VslamError, synth

    with pytest.raises(VslamError) as e:
        StereoProcessor(CameraModel(**synth.EUROC_CAMERA), 1200)
    assert e.value.status == 2 and "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orb-slam3-rust_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl", ".hpp", ".cpp", ".rs")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "orb_oracle" not in txt and "liborb_oracle" not in txt and "cv2" not in txt.replace("cv2.abi3", ""), \
                    f"{f} references the oracle / cv2"
                # the classic mode cites its CPU model (oracle/classic_model.cpp) in comments as its SPECIFICATION; no line
                # of product code may include, import, link or load anything from oracle/
                for line in txt.splitlines():
                    code = line.split("//")[0].split("#")[0] if not line.lstrip().startswith(("#include", "import", "from")) else line
                    if "oracle" in code:
                        assert not any(k in code for k in ("#include", "import ", "CDLL", "dlopen", "-l", "from ")), \
                            f"{f}: product code depends on oracle/: {line.strip()}"
