#!/bin/bash
for R in 701; do
for a in 0 128 256 383 263 271; do
  echo "== ISDFB_ABLATE=$a R=$R"
  ISDFB_ABLATE=$a python tools/kernel_time.py bf16x3 $R 2>&1 | grep -E "chain ms"
  ISDFB_ABLATE=$a python tools/kernel_time.py bf16 $R 2>&1 | grep -E "chain ms" | head -1
done
done
