"""The C-ABI shared library loads and exports every symbol include/vsf.h declares
(no compute calls: this runs without a GPU), and fails loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vsf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vsf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import vision_slam_frontend_b200 as vsf
    from vision_slam_frontend_b200 import capi
    lib = ctypes.CDLL(vsf.library_path())
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/vsf.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == decl
    assert vsf.load_library().vsf_version() == b"0.1.0"


def test_struct_layouts_match_opencv_and_reference():
    from vision_slam_frontend_b200 import capi
    assert capi.DMATCH_DTYPE.itemsize == 16        # cv::DMatch
    assert capi.KEYPOINT_DTYPE.itemsize == 28      # cv::KeyPoint
    assert capi.FEATURE_MATCH_DTYPE.itemsize == 16  # slam_types::FeatureMatch


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import vision_slam_frontend_b200 as vsf
    with pytest.raises(vsf.VsfError) as e:
        vsf.Context()
    assert e.value.code == 3      # VSF_ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vision_slam_frontend_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle's", "").replace("the oracle", "") or \
                    "import oracle" not in text and "from oracle" not in text, f
                assert "import cv2" not in text, f
