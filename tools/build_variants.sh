#!/bin/bash
# builds window-kernel variants into quant_iron_b200/lib/variants/ (experiments only)
# usage: tools/build_variants.sh name1:"-DFLAG ..." name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p quant_iron_b200/lib/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  out=quant_iron_b200/lib/variants/libqiron_${name}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
     $flags -Iinclude -Iquant_iron_b200/csrc -shared -o $out \
     quant_iron_b200/csrc/{engine,state,gates,window,pauli,pauli_window,measure,shard}.cu -lcudart &
done
wait
ls quant_iron_b200/lib/variants/
