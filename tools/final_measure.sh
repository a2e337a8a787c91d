#!/bin/bash
# round-end measurement: parity suite, bench lines, ncu launch list of the bench command, DRAM traffic of the dominant kernel
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q) > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python bench.py 2> gpurun_out/final_bench_default.err | tail -1 > gpurun_out/final_bench_default.json; cut -c1-300 gpurun_out/final_bench_default.json
python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/final_bench_c4.err | tail -1 > gpurun_out/final_bench_c4.json; cut -c1-200 gpurun_out/final_bench_c4.json
python bench.py --precision bf16 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/final_bench_bf16.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:tc_ -s 6 -c 2 --csv --log-file gpurun_out/final_traffic.csv python tools/kernel_time.py bf16x3 1000 > /dev/null 2>&1
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/final_bench_reference.json; cut -c1-200 gpurun_out/final_bench_reference.json
