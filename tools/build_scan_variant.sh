#!/bin/bash
# Experiment builds of the TMA-streamed forward scan only: tools/build_scan_variant.sh <name> <extra nvcc flags...>
# -> audio-mamba-aum_b200/aum_b200/lib/libaum_b200_<name>.so (every other object comes from the default build);
# select at run time with AUM_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../audio-mamba-aum_b200/csrc"
name=$1; shift
mkdir -p build_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     --expt-relaxed-constexpr -Xptxas -v "$@" -c scan_fwd_tma.cu -o build_$name/scan_fwd_tma.o 2> build_$name/scan_fwd_tma.ptxas.log
objs=$(ls build/*.o | grep -v scan_fwd_tma.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../aum_b200/lib/libaum_b200_$name.so $objs build_$name/scan_fwd_tma.o
echo built ../aum_b200/lib/libaum_b200_$name.so
