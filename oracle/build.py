"""Compile the native oracle files into oracle/_build/ (git-ignored).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  `oracle/_ref/` (a build of
the reference's own sources) does not exist for this project: the reference
path needs OpenCV C++ 3.2.0, Eigen, glog, gflags and ROS, none of which are in
the image, so it is unbuildable here ("Oracle").
"""
from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_build")


def _stale(target: str, *sources: str) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def ensure_built() -> dict:
    os.makedirs(_OUT, exist_ok=True)
    knn_src = os.path.join(_HERE, "oracle_knn.c")
    knn_so = os.path.join(_OUT, "liboracle_knn.so")
    if _stale(knn_so, knn_src):
        subprocess.check_call(
            ["gcc", "-O3", "-mpopcnt", "-fopenmp", "-shared", "-fPIC",
             "-o", knn_so + ".tmp", knn_src])
        os.replace(knn_so + ".tmp", knn_so)
    sort_src = os.path.join(_HERE, "stdsort_oracle.cc")
    sort_so = os.path.join(_OUT, "liboracle_stdsort.so")
    if _stale(sort_so, sort_src):
        subprocess.check_call(
            ["g++", "-O2", "-std=c++11", "-shared", "-fPIC",
             "-o", sort_so + ".tmp", sort_src])
        os.replace(sort_so + ".tmp", sort_so)
    return {"knn": knn_so, "stdsort": sort_so}


if __name__ == "__main__":
    print(ensure_built())
