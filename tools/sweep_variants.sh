#!/bin/bash
# A/B block-size variants of libmcx.so built under variants/ (see profiles/r01_history.md): stage timers of one 2M x 100 bp search each
out=gpurun_out/sweep_variants.txt; : > $out
for lib in microbecensus_b200/libmcx.so variants/*.so; do
  echo "== $lib" >> $out
  MCX_LIB=$PWD/$lib timeout 120 python tools/prof_run.py ${1:-2000000} ${2:-100} 4 >> $out 2>&1
done
cat $out
