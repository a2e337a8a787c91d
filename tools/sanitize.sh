#!/bin/bash
# compute-sanitizer passes over one small training step of the tensor-core path (40 rays x 27 samples = 9 tiles:
# chain kernel + weight-gradient kernel + the K1/K5/K6 kernels) and over the fused sampler / window kernels.
# Logs go to gpurun_out/ (copied to profiles/ after review).  Usage on the GPU box: bash tools/sanitize.sh
mkdir -p gpurun_out
export ISDFB_SMOKE_MODE=bf16x3g
SMOKE='import __graft_entry__ as g; g.smoke()'
STEP='import torch, bench, numpy as np
from isdf.modules import trainer as T
wl = dict(bench.WORKLOADS["default"]); wl["n_rays"] = 24
cfg = bench.make_config(wl, "bf16x3g", "fast"); cfg["b200"]["cuda_graph"] = 0
cfg["dataset"]["camera"].update(w=160, h=120, fx=100.0, fy=100.0, cx=79.5, cy=59.5)
tr = T.Trainer(torch.device("cuda:0"), cfg)
for k in range(7):
    tr.last_is_keyframe = True; tr.add_data(tr.get_data([k])); tr.step()
for _ in range(3):
    l, _ = tr.step()
print("sanitized steps ok, loss", float(l["total_loss"]))'
for tool in memcheck racecheck synccheck; do
  echo "== $tool: smoke (K4 + weight gradients vs oracle)"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -c "$SMOKE" > gpurun_out/sanitizer_${tool}_smoke.log 2>&1
  echo "rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_smoke.log
  echo "== $tool: Trainer.step() x10 (window, fused sampler, K4, K5, K6)"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -c "$STEP" > gpurun_out/sanitizer_${tool}_step.log 2>&1
  echo "rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_step.log
done
