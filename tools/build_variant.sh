#!/bin/bash
# build_variant.sh TAG FLAGS... : link swmhd_b200/libswmhd_TAG.so with substage_rb.cu compiled with extra FLAGS
# (development aid for A/B runs on one GPU box: SWMHD_LIB=$PWD/swmhd_b200/libswmhd_TAG.so).
# SRC=path/to/other_substage_rb.cu builds the variant from another source (e.g. `git show REV:swmhd_b200/csrc/substage_rb.cu`).
set -e
TAG=$1; shift
D=swmhd_b200/csrc; O=$D/_obj
SRC=${SRC:-$D/substage_rb.cu}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I $D "$@" -c $SRC -o $O/substage_rb_$TAG.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o swmhd_b200/libswmhd_$TAG.so $O/substage_strict.o $O/substage_fast.o $O/substage_rb_$TAG.o $O/aux_kernels.o $O/swmhd_api.o $O/nccl_dyn.o -cudart static -ldl
echo swmhd_b200/libswmhd_$TAG.so
