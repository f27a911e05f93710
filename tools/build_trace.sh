#!/bin/bash
# debug build of the library with the Schur kernel's per-warp clock stamps (tools/schur_trace.py); never shipped
mkdir -p tools/_ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DARGUS_SCHUR_TRACE -o tools/_ab/libtrace.so \
  argus_gui_b200/csrc/ba_solver.cu argus_gui_b200/csrc/dlt_kernels.cu argus_gui_b200/csrc/pair_kernels.cu argus_gui_b200/csrc/fit_kernels.cu argus_gui_b200/csrc/microbench.cu
