#!/bin/bash
# build_variant.sh NAME "<extra nvcc flags>": builds scratch/libmrt_NAME.so from the current sources (kernel experiments,
# e.g.  build_variant.sh precise "-DMRT_PRECISE=1 -prec-div=true -prec-sqrt=true"); select it with MRT_LIB=scratch/libmrt_NAME.so
set -e
cd "$(dirname "$0")/../micro_raytracer_b200/csrc"
D=../../scratch/v_$1
mkdir -p $D
FL="-O3 -std=c++17 -lineinfo -ftz=true -prec-div=false -prec-sqrt=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $2"
make -s mrt_jit_src.inc
nvcc $FL -Xptxas -v -c mrt_kernels.cu -o $D/k.o 2> $D/ptxas.log &
nvcc $FL -c mrt_api.cu -o $D/a.o &
nvcc $FL -c mrt_scene.cu -o $D/s.o &
nvcc $FL -c mrt_jit.cu -o $D/j.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../scratch/libmrt_$1.so $D/k.o $D/a.o $D/s.o $D/j.o -ldl -lpthread
grep -A3 "path_kernel_paramILj0E" $D/ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1"
