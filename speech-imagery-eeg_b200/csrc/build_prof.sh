#!/bin/bash
# Debug build with role-level cycle counters in the tcgen05 kernels (-DIGN_TC_PROFILE) -> ../lib/libign_b200_prof.so
# Use: IGN_B200_LIB=speech-imagery-eeg_b200/lib/libign_b200_prof.so python tools/tc_bwd_profile.py
set -e
cd "$(dirname "$0")"
mkdir -p build_prof ../lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
for f in api instnorm shapelet_simt shapelet_dx shapelet_tc shapelet_tc_bwd gate regulariser; do
  $NVCC -O3 -std=c++17 -lineinfo -DIGN_TC_PROFILE -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -c $f.cu -o build_prof/$f.o &
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libign_b200_prof.so build_prof/*.o -lcudart_static -lpthread -ldl -lrt
