#!/bin/bash
# Build tuning variants of the library into scripts/_build/ (picked through PCTRANS_MSDA_LIB).
#   scripts/build_variants.sh name1:"-DX=1 -DY=2" name2:"-DZ=3" ...
set -e
cd "$(dirname "$0")/.."
mkdir -p scripts/_build
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
       --cudart shared -ccbin /usr/bin/g++ $defs \
       -o scripts/_build/libpctrans_msda_$name.so pctrans_b200/csrc/msda_capi.cu pctrans_b200/csrc/msda_host.cu pctrans_b200/csrc/linear_tc.cu &
done
wait
ls scripts/_build/*.so
