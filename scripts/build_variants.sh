#!/bin/bash
# builds libfastf_gpu variants with different -D flags into fastf_b200/_build/variants/ (A/B on the GPU box via FASTF_GPU_LIB)
# usage: build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
cd "$(dirname "$0")/.."
mkdir -p fastf_b200/_build/variants
rm -f fastf_b200/_build/variants/*.so
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DFASTF_SRC_HASH=\"variant\" $flags -o fastf_b200/_build/variants/$name.so fastf_b200/csrc/capi.cu fastf_b200/csrc/sharded.cu -ldl &
done
wait
ls -la fastf_b200/_build/variants/
