#!/bin/bash
# development aid: build several -D variants of the library side by side for scripts/ab_time.py
# usage: scripts/build_variants.sh name1:"-DFOO" name2:"-DBAR -DBAZ" ...   ->  build/lib_<name>.so
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -shared -Xcompiler -fPIC"
cd "$(dirname "$0")/.." && mkdir -p build
for v in "$@"; do
  n="${v%%:*}"; d="${v#*:}"
  ( /usr/local/cuda/bin/nvcc $F $d -o build/lib_$n.so fthmc_b200/csrc/fthmc_capi.cu > build/lib_$n.log 2>&1 || echo "FAILED $n" ) &
done
wait
ls -la build/lib_*.so
