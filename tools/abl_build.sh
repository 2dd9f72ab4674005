#!/bin/bash
# Development helper: build variants of libeavsr_b200.so into build/abl/lib_<name>.so.  Each argument is
# "name:flags", e.g. "nofence:-DEAVSR_ABL=1" or "deep:-DEAVSR_WIN_NSB=3 -DEAVSR_WIN_NOB=4".  Only dcn_fwd_tc.cu is
# recompiled (deform_groups = 8 kernels only); every other object comes from eavsr_b200/csrc/_obj.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/abl
OBJ=eavsr_b200/csrc/_obj
SRC=${SRC:-dcn_fwd_tc}     # which translation unit to rebuild (SRC=conv3x3_tc tools/abl_build.sh ...)
for a in "$@"; do
  name="${a%%:*}"; flags="${a#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       -Xptxas -v -DEAVSR_ONLY_DG8 $flags -c eavsr_b200/csrc/$SRC.cu -o build/abl/${SRC}_$name.o 2> build/abl/log_$name.txt &
done
wait
for a in "$@"; do
  name="${a%%:*}"
  nvcc -shared -o build/abl/lib_$name.so build/abl/${SRC}_$name.o $(ls $OBJ/*.o | grep -v $SRC.o) -lcudart 2>/dev/null
  rm -f build/abl/${SRC}_$name.o
  echo "built build/abl/lib_$name.so"
done
