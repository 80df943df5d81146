#!/bin/bash
# Build tuning variants of libqcd_b200.so HERE (no GPU needed) into build/variants/<name>.so; a GPU call then
# compares them by QCD_B200_LIB=build/variants/<name>.so without paying nvcc time on the box.
#   scripts/variants.sh name1 "-DFLAG=1 ..." name2 "..." ...
cd "$(dirname "$0")/.." || exit 1
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -I include"
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  ( nvcc $FLAGS $extra -Xptxas=-v -o build/variants/$name.so qcd_b200/csrc/qcd_kernels.cu 2> build/variants/$name.log \
    && echo "built $name: $(grep -A2 'replay_reg_kernelILi12ELb1' build/variants/$name.log | grep -E 'spill|Used' | tr '\n' ' ')" \
    || { echo "FAILED $name"; tail -5 build/variants/$name.log; } ) &
done
wait
