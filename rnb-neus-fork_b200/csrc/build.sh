#!/bin/bash
# Build librnb_b200.so for sm_100a (in-tree; the .so travels to the GPU box with the repo snapshot).
set -e
cd "$(dirname "$0")"
OUT=${RNB_OUT:-../rnb_b200/librnb_b200.so}
SRCS="api.cu sdf_chain.cu pack.cu $(ls dw_gemm.cu render.cu albedo.cu nerf.cu mc.cu adam.cu wnorm.cu 2>/dev/null || true)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
     -Xcompiler -fvisibility=hidden -o "$OUT" $SRCS "$@"
echo "built $OUT"
