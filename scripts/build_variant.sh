#!/bin/bash
# A/B helper: rebuild ONE translation unit with extra -D flags and link it with the in-tree objects of the others.
# usage: build_variant.sh <out.so> <file.cu> [-DFLAG=...]...     (run after the normal build; use with BC_LIB_PATH)
set -e
OUT=$1; SRC=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OBJ=$ROOT/audiotokenization_b200/csrc/obj
TMP=$(mktemp -d)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v "$@" \
     -c "$ROOT/audiotokenization_b200/csrc/$SRC" -o "$TMP/variant.o" 2>&1 | grep -E "Used|spill" | sort | uniq -c
OTHERS=$(ls $OBJ/*.o | grep -v "/$(basename $SRC .cu).o")
nvcc -shared -o "$OUT" $TMP/variant.o $OTHERS -cudart static 2>/dev/null
rm -rf "$TMP"; ls -la "$OUT"
