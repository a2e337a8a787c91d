#!/bin/bash
R=701
for sg in 1 0; do
for a in 0 128 256 383 640 768; do
  echo "== STAGGER=$sg ABLATE=$a R=$R"
  ISDFB_STAGGER=$sg ISDFB_ABLATE=$a python tools/kernel_time.py bf16x3 $R 2>&1 | grep -E "chain ms" | head -1
  ISDFB_STAGGER=$sg ISDFB_ABLATE=$a python tools/kernel_time.py bf16 $R 2>&1 | grep -E "chain ms" | head -1
done
done
