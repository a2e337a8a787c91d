#!/bin/bash
# timing ablations of the chain kernel (dev tool; results of ablated runs are numerically invalid)
for R in 701 1000; do
for a in 0 1 2 4 8 16 32 64 7 15 79 127; do
  echo "== ISDFB_ABLATE=$a R=$R"
  ISDFB_ABLATE=$a python tools/kernel_time.py bf16x3 $R 2>&1 | grep "chain ms" | head -1
done
done
