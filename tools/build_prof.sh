#!/bin/sh
# Builds the instrumented copy of the library used by tools/potrf_prof.py (clock64 stamps inside potrf_inv_kernel,
# -DSD_PROFILE_POTRF) into superviseddescent_b200/lib_prof/.  Not part of the product build.
set -e
cd "$(dirname "$0")/../superviseddescent_b200"
mkdir -p build_prof lib_prof
for f in sd_api sd_hog sd_linalg sd_gram_tc sd_model sd_comm sd_rank sd_cg; do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -DSD_PROFILE_POTRF \
       -Xcompiler -fPIC,-fvisibility=hidden -I ../include -I csrc -c csrc/$f.cu -o build_prof/$f.o &
done
wait
nvcc -shared -o lib_prof/libsd_b200.so build_prof/*.o -gencode arch=compute_100a,code=sm_100a -lcudart -ldl
echo lib_prof/libsd_b200.so
