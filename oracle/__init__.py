"""CPU oracle for the vision_slam_frontend matching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`vision_slam_frontend_b200/`) may import, link or execute anything under
`oracle/`.  The only permitted users are `tests/`, `__graft_entry__.smoke()`
and the CPU legs of `bench.py` (`cpu_baseline`, `--impl reference`).

Parity status
-------------
The reference (`/root/reference`, ut-amrl/vision_slam_frontend) ships no test,
golden vector or known-answer fixture for this path, and
its arithmetic lives in a third-party dependency that is not vendored:
OpenCV, pinned `EXACT 3.2.0` at `CMakeLists.txt:21` (call sites
`src/slam_frontend.cc:525-527` knnMatch, `:152-156` triangulatePoints).  The
C++ reference itself cannot be built here (no OpenCV C++ / Eigen / glog / ROS).

What pins this oracle instead: the same OpenCV *functions* are importable in
this image as the Python wheel `opencv-python-headless 4.13.0`
(`cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch`, `cv2.triangulatePoints`).
`tests/golden/make_golden.py` ran them here and froze their outputs as
fixtures; `tests/test_oracle_golden.py` checks the restatements in this
directory against those fixtures (and against live cv2 when importable).
So: **pinned against OpenCV 4.13 outputs, unpinned against the reference's
own tests (it has none) and against OpenCV 3.2.0 (not obtainable here).**

Modules
-------
restate.py        numpy restatement of the reference glue + the published
                  BFMatcher / triangulatePoints algorithms
cv2_ref.py        thin wrappers over the real cv2 functions (stand-in for the
                  un-vendored dependency; used to make goldens and as the
                  multi-threaded CPU baseline)
oracle_knn.c      plain-C restatement of k=2 Hamming kNN + ratio (fast checker
                  for the full-size configs)
stdsort_oracle.cc libstdc++ std::sort restatement of GetFeatureMatches' sort +
                  best_percent cut (the order is libstdc++-specific, quirk Q1)
build.py          compiles the two native files into oracle/_build/
"""
