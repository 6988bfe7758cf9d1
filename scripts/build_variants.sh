#!/bin/bash
# Build tuning variants of the library (different CTA sizes => different register budgets) into
# scripts/_build/; scripts/tune.py picks one through PCTRANS_MSDA_LIB.
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/_build
for spec in "$@"; do   # spec = FWD:BWD  e.g. 1024:512
  f=${spec%%:*}; b=${spec##*:}
  out=scripts/_build/libpctrans_msda_f${f}_b${b}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
       --cudart shared -ccbin /usr/bin/g++ -DPCT_FWD_THREADS=$f -DPCT_BWD_THREADS=$b \
       -o $out pctrans_b200/csrc/msda_capi.cu &
done
wait
ls -la scripts/_build/*.so
