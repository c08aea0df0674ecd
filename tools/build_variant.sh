#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> <extra nvcc flags...>
# Rebuilds ONE translation unit with extra -D flags and links lattigo_b200/lib/variants/<name>.so against the other
# (already built) objects; select it at run time with LGPU_SO_PATH. Development A/B only.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p lattigo_b200/lib/variants
obj=lattigo_b200/lib/variants/$name.$(basename ${src%.cu}).o
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fno-fast-math --fmad=false "$@" -c lattigo_b200/csrc/$src -o $obj
others=$(ls lattigo_b200/lib/obj/*.o | grep -v "/$(basename ${src%.cu}).o")
/usr/local/cuda/bin/nvcc -shared -o lattigo_b200/lib/variants/$name.so $obj $others -cudart static -gencode arch=compute_100a,code=sm_100a
echo lattigo_b200/lib/variants/$name.so
