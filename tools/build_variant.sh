#!/bin/bash
# build_variant.sh NAME "<extra nvcc flags>": builds scratch/libmrt_NAME.so from the current sources (kernel experiments)
set -e
cd "$(dirname "$0")/../micro_raytracer_b200/csrc"
mkdir -p ../../scratch/v_$1
FL="-O3 -std=c++17 -lineinfo -ftz=true -prec-div=false -prec-sqrt=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $2"
nvcc $FL -Xptxas -v -c mrt_kernels.cu -o ../../scratch/v_$1/k.o 2> ../../scratch/v_$1/ptxas.log
nvcc $FL -c mrt_api.cu -o ../../scratch/v_$1/a.o
make -s mrt_jit_src.inc
nvcc $FL -c mrt_jit.cu -o ../../scratch/v_$1/j.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/libmrt_$1.so ../../scratch/v_$1/k.o ../../scratch/v_$1/a.o ../../scratch/v_$1/j.o -ldl
grep -A3 "path_kernel_paramILj0E" ../../scratch/v_$1/ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1"
