#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02zd
mkdir -p "$OUT"
timeout 900 python tools/prof_qft_restore.py 33 > "$OUT/qft33_restore.txt" 2>&1
timeout 600 python tools/prof_qft_restore.py 30 > "$OUT/qft30_restore.txt" 2>&1
ls -la "$OUT"
