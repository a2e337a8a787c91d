mkdir -p gpurun_out
for st in 20 200; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps $st --warmup 5 2>gpurun_out/r2h_n2_$st.err | tail -1 > gpurun_out/r2h_n2_$st.json; cut -c1-220 gpurun_out/r2h_n2_$st.json; tail -2 gpurun_out/r2h_n2_$st.err
done
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-220
python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -3
