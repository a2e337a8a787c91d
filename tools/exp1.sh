#!/bin/bash
for R in 701 1000; do
for sl in 0 1; do
  echo "== STREAM_LOADS=$sl R=$R"; ISDFB_STREAM_LOADS=$sl python tools/kernel_time.py bf16x3 $R 2>&1 | grep -E "chain ms" | head -1
done; done
ISDFB_STREAM_LOADS=1 python tools/kernel_time.py bf16 1000 2>&1 | grep -E "chain ms" | head -1
ISDFB_TEST_MODES=bf16x3 python -m pytest tests/test_gpu_engine.py -q -x 2>&1 | tail -2
