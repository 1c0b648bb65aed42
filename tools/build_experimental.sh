#!/bin/sh
# Compile-check of the DRAFT CTA-pair (cta_group::2) SYRK kernel that sits in sd_gram_tc.cu under
# #ifdef SD_EXPERIMENTAL_2CTA.  It is not part of the product build and has not been run on hardware; this only
# proves that the PTX forms it uses (cluster barriers, remote mbarrier arrive, multicast commit, UTCHMMA.2CTA)
# assemble for sm_100a.
set -e
cd "$(dirname "$0")/../superviseddescent_b200"
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -DSD_EXPERIMENTAL_2CTA \
     -Xcompiler -fPIC,-fvisibility=hidden -I ../include -I csrc -c csrc/sd_gram_tc.cu -o build/sd_gram_tc_experimental.o
cuobjdump -sass build/sd_gram_tc_experimental.o | grep -c "UTCHMMA.2CTA"
