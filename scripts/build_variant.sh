#!/bin/bash
# build_variant.sh NAME "EXTRA_NVFLAGS"  -> variants/NAME.so  (kernel tuning experiments; only the batched-LM instantiations are rebuilt)
set -e
NAME=$1; FLAGS=$2
cd "$(dirname "$0")/../mir_optim_b200/csrc"
mkdir -p ../../variants/build_$NAME
B=../../variants/build_$NAME
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
$NV $FLAGS -c lm_small_inst_d.cu -o $B/lm_small_inst_d.o > $B/ptxas_d.log 2>&1 &
$NV $FLAGS -c lm_small_inst_s.cu -o $B/lm_small_inst_s.o > $B/ptxas_s.log 2>&1 &
wait
grep -A3 "lm_tpp_kernelINS_11ModelGauss4IdLb1EEEdLb0\|lm_tpp_kernelINS_11ModelSumExpIdLi8ELb1EEEdLb1" $B/ptxas_d.log | grep -E "registers|spill" | head -4
OBJS=$(ls build/*.o | grep -v lm_small_inst_)
$NV -shared -cudart static -o ../../variants/$NAME.so $B/lm_small_inst_d.o $B/lm_small_inst_s.o $OBJS -ldl -lpthread
