#!/bin/bash
# build_variant.sh TAG FLAGS... : link swmhd_b200/libswmhd_TAG.so with substage_rb.cu compiled with extra FLAGS
# (development aid for A/B runs on one GPU box: SWMHD_LIB=$PWD/swmhd_b200/libswmhd_TAG.so)
set -e
TAG=$1; shift
D=swmhd_b200/csrc; O=$D/_obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c $D/substage_rb.cu -o $O/substage_rb_$TAG.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o swmhd_b200/libswmhd_$TAG.so $O/substage_strict.o $O/substage_fast.o $O/substage_rb_$TAG.o $O/aux_kernels.o $O/swmhd_api.o -cudart static
echo swmhd_b200/libswmhd_$TAG.so
