#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02zh
mkdir -p "$OUT"
timeout 900 python -m pytest tests -q -m "gpu and not large" -k "prob or sampl or measure or collapse or golden or ref_ported" > "$OUT/pytest_measure.log" 2>&1
echo "exit $?" >> "$OUT/pytest_measure.log"
timeout 600 python tools/prof_measure.py 30 > "$OUT/measure30.txt" 2>&1
ls -la "$OUT"
