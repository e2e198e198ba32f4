#!/bin/bash
# Development aid: builds libsbn_b200/lib/libsbn_<name>.so with extra nvcc flags (-D...),
# selected at run time with SBNB_LIBRARY=<path>.   tools/build_variant.sh <name> [flags...]
name=$1; shift
here=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p /tmp/variant_$name
cd $here/libsbn_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3 "$@" -c engine.cu -o /tmp/variant_$name/engine.o &&
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libsbn_$name.so /tmp/variant_$name/engine.o ../lib/gp_engine.o ../lib/site_pattern.o ../lib/model.o ../lib/tree_program.o ../lib/rooted.o -lcudart -lcudadevrt
