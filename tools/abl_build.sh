#!/bin/bash
# Development helper: build variants of libeavsr_b200.so into build/abl/lib_<name>.so.  Each argument is
# "name:flags", e.g. "nofence:-DEAVSR_ABL=1" or "deep:-DEAVSR_WIN_NSB=3 -DEAVSR_WIN_NOB=4".  Only dcn_fwd_tc.cu is
# recompiled (deform_groups = 8 kernels only); every other object comes from eavsr_b200/csrc/_obj.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/abl
OBJ=eavsr_b200/csrc/_obj
for a in "$@"; do
  name="${a%%:*}"; flags="${a#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       -Xptxas -v -DEAVSR_ONLY_DG8 $flags -c eavsr_b200/csrc/dcn_fwd_tc.cu -o build/abl/dcn_fwd_tc_$name.o 2> build/abl/log_$name.txt &
done
wait
for a in "$@"; do
  name="${a%%:*}"
  nvcc -shared -o build/abl/lib_$name.so build/abl/dcn_fwd_tc_$name.o $(ls $OBJ/*.o | grep -v dcn_fwd_tc.o) -lcudart 2>/dev/null
  rm -f build/abl/dcn_fwd_tc_$name.o
  echo "built build/abl/lib_$name.so"
done
