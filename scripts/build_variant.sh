#!/bin/bash
# build_variant.sh NAME "EXTRA_NVFLAGS"  -> variants/NAME.so  (kernel tuning experiments; only lm_small_inst_d is rebuilt)
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../mir_optim_b200/csrc"
mkdir -p ../../variants/build_$NAME
B=../../variants/build_$NAME
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
$NV $FLAGS -c lm_small_inst_d.cu -o $B/lm_small_inst_d.o > $B/ptxas.log 2>&1
grep -A2 "ModelGauss4" $B/ptxas.log | grep -E "registers|spill" | head -4
$NV -shared -cudart static -o ../../variants/$NAME.so $B/lm_small_inst_d.o build/lm_small_inst_s.o build/lm_batched.o build/runtime.o build/todo_stubs.o build/shim.o build/peaks.o -ldl -lpthread
