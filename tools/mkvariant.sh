#!/bin/bash
# development tool: compile a kernel variant into variants/libwgk_<name>.so (run with WGK_LIB=variants/libwgk_<name>.so)
#   tools/mkvariant.sh name [-DWGK_... ...]
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -shared -Xcompiler -fPIC --cudart static \
  -Xptxas -v "$@" -o variants/libwgk_$name.so watergap2_b200/csrc/wgk_api.cu > variants/$name.ptxas 2>&1 || { tail -20 variants/$name.ptxas; exit 1; }
for k in _ZN3wgk15k_cells_pre_tpcE9WgkParamsiii _ZN3wgk13k_river_levelE9WgkParamsii _ZN3wgk12k_tail_chunkE9WgkParamsiii; do
  grep -A3 "Compiling entry function '$k'" variants/$name.ptxas | tr '\n' ' ' | sed 's/ptxas info    ://g; s/  */ /g'; echo
done
