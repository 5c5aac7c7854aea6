#!/bin/bash
# A/B tuning variants of libmcx.so (results in profiles/r01_history.md): stage timers of one search per library.
#   build:  mkdir -p variants; cd microbecensus_b200/csrc; for v in GAP_REFILL=16 SEG_W=8; do \
#             nvcc <NVFLAGS of the Makefile> -DMCX_$v -shared -o ../../variants/libmcx_${v/=/_}.so mcx.cu; done
#           (macros: MCX_PROBE_POS, MCX_PROBE_NT, MCX_SEED_NT, MCX_WALK_NT, MCX_GAP_NT, MCX_GAP_REFILL, MCX_SEG_W, MCX_FRAMES_NT)
#   run:    gpurun -- 'bash tools/sweep_variants.sh [reads] [read length]'     (variants/ is git-ignored but travels)
mkdir -p gpurun_out
out=gpurun_out/sweep_variants.txt; : > $out
for lib in microbecensus_b200/libmcx.so variants/*.so; do
  [ -f "$lib" ] || continue
  echo "== $lib" >> $out
  MCX_LIB=$PWD/$lib timeout 120 python tools/prof_run.py ${1:-2000000} ${2:-100} 4 >> $out 2>&1
done
cat $out
