#!/bin/bash
# usage: mkvar.sh <tag> <extra nvcc defines...>   builds var/libarcrad_<tag>.so with alternative solver objects
set -e
cd /root/repo/wrfchem-arc-interactions_b200/csrc
tag=$1; shift
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fno-strict-aliasing"
mkdir -p var/$tag
$NV -fmad=false "$@" -c sw_solve.cu -o var/$tag/sw_solve.o &
$NV "$@" -c lw_solve.cu -o var/$tag/lw_solve.o &
$NV "$@" -c aer_optics.cu -o var/$tag/aer_optics.o &
wait
$NV -shared -o var/libarcrad_$tag.so api.o prep.o var/$tag/sw_solve.o var/$tag/lw_solve.o tables.o var/$tag/aer_optics.o aer_tables.o -lcudart_static -lpthread -ldl -lrt
echo built var/libarcrad_$tag.so
