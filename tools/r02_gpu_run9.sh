#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02i
mkdir -p "$OUT"
python tools/prof_qft.py 30 3 > "$OUT/qft30.txt" 2>&1
python tools/prof_qft.py 33 1 > "$OUT/qft33.txt" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 6 -c 2 -o "$OUT/tile_qft30" python tools/prof_qft.py 30 1 > "$OUT/ncu_qft.log" 2>&1
ls -la "$OUT"
