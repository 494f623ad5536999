#!/bin/bash
# F1 kernel variants (variants/lib_*.so built by mono_vifi_b200.build with defines): microbenchmark of each, one JSON line per variant
out=${1:-gpurun_out/f1_variants.txt}
: > $out
for v in variants/lib_*.so; do
  MVF_LIB=$PWD/$v timeout 120 python tools/f1_microbench.py 12 192 640 20 --coherent >> $out 2>&1 || echo "FAILED $v" >> $out
done
cat $out
