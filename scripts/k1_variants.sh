#!/bin/bash
# Build A/B variants of the K1 epilogue into jegal_b200/csrc/build/variants/<name>.so (run here, no GPU):
#   scripts/k1_variants.sh name "-DFLAG ..." [name2 "flags2" ...]
# then on the GPU box: JEGAL_B200_LIB=jegal_b200/csrc/build/variants/<name>.so python scripts/k1_ab.py
set -e
cd "$(dirname "$0")/../jegal_b200/csrc"
mkdir -p build/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
    -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c simpool.cu -o build/variants/$name.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so \
    build/api.o build/variants/$name.o build/prep.o build/topk.o build/grouped.o build/exchange.o -cudart static
  echo "built build/variants/$name.so ($flags)"
done
