#!/bin/bash
# Round-2 measurement recipe (each block is one `gpurun` call; outputs land in gpurun_out/, the summaries that are
# judged are copied to profiles/ -- see profiles/r02_summary.md).  1 GPU unless a line says otherwise.
mkdir -p gpurun_out
# parity + sanitizer
(time python -m pytest tests -m gpu -q) > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
bash tools/sanitize.sh
# bench lines (driver-style: --steps 20 --warmup 5, and long runs)
for wl in default scannet c4 grid c5; do
  python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_$wl.json
done
python bench.py --steps 100 --warmup 5 2>/dev/null | tail -1 > gpurun_out/final_bench_default_100.json
python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/final_bench_reference.json
# launch list of the bench command (shares, not absolutes) and ncu captures of the dominant kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ISDFB_NO_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 3 -c 1 -o gpurun_out/final_chain_211 -f \
    python tools/kernel_time.py bf16x3g 1000 > /dev/null 2>&1
ISDFB_NO_OVERLAP=1 ncu --set full --clock-control none --import-source on -k regex:tc_chain -s 3 -c 1 -o gpurun_out/final_chain_148 -f \
    python tools/kernel_time.py bf16x3g 701 > /dev/null 2>&1
ISDFB_NO_OVERLAP=1 ncu --set full --clock-control none -k regex:tc_dw -s 3 -c 1 -o gpurun_out/final_dw_211 -f \
    python tools/kernel_time.py bf16x3g 1000 > /dev/null 2>&1
# here (no GPU): python tools/update_traffic.py bf16x3g default gpurun_out/final_chain_211.ncu-rep "<command>"
#                python tools/ncu_summary.py gpurun_out/final_chain_148.ncu-rep profiles/<name>.json "<note>"
# multi-GPU (gpurun --gpus N): N=2|8 bash tools/n2_check.sh
