#!/bin/bash
# GPU-box helper: per-kernel timings of every tuning build under build_variants/ (tools/variant_bench.py)
mkdir -p gpurun_out
for lib in build_variants/libdfr_*.so; do
  DFR_LIBRARY=$PWD/$lib timeout 300 python tools/variant_bench.py ${1:-1048576} 5 20 2>&1 | tail -2
done | tee gpurun_out/variants.txt
