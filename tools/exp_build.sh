#!/bin/bash
# Builds build/ab/exp.so = the library WITH the timing-experiment kernel variants (-DSRCDSP_TIMING_EXPERIMENTS):
# SRCDSP_TC_DEBUG / SRCDSP_UPT_DEBUG then select instantiations that skip work (wrong results) or count wait
# cycles.  The shipped library (python -m srcdsp_b200.build) has none of them.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static \
     -DSRCDSP_TIMING_EXPERIMENTS -o build/ab/exp.so srcdsp_b200/csrc/*.cu
ls -la build/ab/exp.so
