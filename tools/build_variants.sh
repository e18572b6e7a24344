#!/bin/bash
# builds window-kernel variants into quant_iron_b200/lib/variants/ (experiments only)
set -e
cd "$(dirname "$0")/.."
mkdir -p quant_iron_b200/lib/variants
for mb in 3 4 5 6; do
  out=quant_iron_b200/lib/variants/libqiron_mb${mb}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DQI_WINDOW_MIN_BLOCKS=${mb} -Iinclude -Iquant_iron_b200/csrc -shared -o $out \
     quant_iron_b200/csrc/{engine,state,gates,window,pauli,measure,shard}.cu -lcudart &
done
wait
ls -la quant_iron_b200/lib/variants/
