#!/bin/bash
# Build a variant of the library with extra -D flags:  tools/build_variant.sh NAME "-DHG_EPI_WARPS=16 ..."
#   (ALL=1 recompiles every source with the flags, default only hg_gemm.cu)
# -> variants/libhgb200_NAME.so ; run anything with HG_LIB_PATH=$PWD/variants/libhgb200_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
B=variants/_build_$name
mkdir -p $B
C=humanoid_gym_with_comments_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
if [ -n "$ALL" ]; then
  for f in hg_runtime hg_gemm hg_gae hg_misc hg_heads hg_optim hg_sym hg_ring; do nvcc $F $@ -c $C/$f.cu -o $B/$f.o & done
  nvcc $F -fmad=false $@ -c $C/hg_env.cu -o $B/hg_env.o &
  nvcc $F -fmad=false $@ -c $C/hg_env_aux.cu -o $B/hg_env_aux.o &
  wait
  objs=$(ls $B/*.o)
else
  nvcc $F $@ -c $C/hg_gemm.cu -o $B/hg_gemm.o
  objs="$B/hg_gemm.o $(ls $C/_build/*.o | grep -v hg_gemm.o)"
fi
nvcc -shared -cudart static -o variants/libhgb200_$name.so $objs 2>/dev/null
echo built variants/libhgb200_$name.so
