#!/bin/bash
# build libvsf_cuda with extra -D flags into vision_slam_frontend_b200/build/variants/<name>.so
# usage: tools/build_variant.sh name -DVSF_TC_BUCKET=16 ...
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=vision_slam_frontend_b200/build/variants; mkdir -p $out/$name
objs=""
for f in knn2_kernel knn2_tc_kernel stereo_kernels sort_kernel aux_kernels vsf_api; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp "$@" -I include -I vision_slam_frontend_b200/csrc -c vision_slam_frontend_b200/csrc/$f.cu -o $out/$name/$f.o &
  objs="$objs $out/$name/$f.o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $objs -cudart static -Xcompiler -fopenmp
echo $out/$name.so
