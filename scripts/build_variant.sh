#!/bin/bash
# Diagnostics: builds a variant of the library (extra -D flags) as build/libtmg_<name>.so for A/B runs (TMG_B200_LIB=...).
# usage: scripts/build_variant.sh <name> [nvcc flags...]
set -e
cd "$(dirname "$0")/.."
mkdir -p build
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -I include "$@" \
     -DTMG_BUILD_ID_STR="\"variant-$name\"" -o build/libtmg_$name.so tile_match_gym_b200/csrc/tmg_b200.cu
echo built build/libtmg_$name.so
