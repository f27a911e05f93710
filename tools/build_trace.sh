#!/bin/bash
# debug builds of the library with per-warp clock stamps (never shipped):
#   bash tools/build_trace.sh          -> tools/_ab/libtrace.so     (-DARGUS_SCHUR_TRACE, tools/schur_trace.py)
#   bash tools/build_trace.sh lin      -> tools/_ab/liblintrace.so  (-DARGUS_LIN_TRACE, tools/lin_trace.py)
mkdir -p tools/_ab
if [ "$1" = "lin" ]; then def=-DARGUS_LIN_TRACE; out=tools/_ab/liblintrace.so; else def=-DARGUS_SCHUR_TRACE; out=tools/_ab/libtrace.so; fi
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $def -o $out \
  argus_gui_b200/csrc/ba_solver.cu argus_gui_b200/csrc/dlt_kernels.cu argus_gui_b200/csrc/pair_kernels.cu argus_gui_b200/csrc/fit_kernels.cu argus_gui_b200/csrc/microbench.cu
