"""The C++ Frontend mirror must also compile when OpenCV's and Eigen's own headers are on the
include path (cv_shim.h then steps aside).  Neither library is installed in this image, so
tests/stubs/ holds declaration-only stand-ins with the REAL libraries' shapes - type codes as
macros, cv::Mat(rows, cols, type[, data, step]), Mat::step a MatStep, Eigen types without public
storage - and every frontend source is compiled -fsyntax-only against them: code that only works
against the shim (cv::CV_8U, a two-argument cv::Mat, Vector3f::v) fails here."""
import glob
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frontend_sources_compile_against_real_header_shapes():
    fdir = os.path.join(ROOT, "vision_slam_frontend_b200", "csrc", "frontend")
    srcs = sorted(glob.glob(os.path.join(fdir, "*.cc")))
    assert len(srcs) >= 5
    for src in srcs:
        cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", fdir,
               "-I", os.path.join(ROOT, "tests", "stubs"), "-I", "/usr/local/cuda/include", src]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, src + "\n" + p.stderr[-3000:]


def test_the_stubs_are_really_selected():
    fdir = os.path.join(ROOT, "vision_slam_frontend_b200", "csrc", "frontend")
    probe = '#include "cv_shim.h"\n#if !defined(VSF_HAVE_OPENCV) || !defined(VSF_HAVE_EIGEN)\n#error shim selected\n#endif\n'
    p = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", fdir, "-I",
                        os.path.join(ROOT, "tests", "stubs"), "-"], input=probe, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
