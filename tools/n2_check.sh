mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $1 "${@:3}" 2>gpurun_out/$2.err | tail -1 > gpurun_out/$2.json; cut -c1-200 gpurun_out/$2.json; tail -1 gpurun_out/$2.err | cut -c1-200; }
N=${N:-2}
run $N r2_n${N}_default_20 --steps 20 --warmup 5
run $N r2_n${N}_default_200 --steps 200 --warmup 5
run $N r2_n${N}_c4_20 --workload c4 --steps 20 --warmup 5
if [ "$N" = "2" ]; then python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-220; fi
