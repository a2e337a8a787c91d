"""Parity figures of the tensor-core modes at the default batch shape (27 000 samples) against the fp64 oracle:
   python tools/dw_error.py bf16x3,bf16x3g"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P
DEV = torch.device("cuda:0")
cfg = O.default_cfg(noise_std=0.08)
sd = C.golden_weights(91, gain=1.3)
batch, noise = C.loss_batch(92, 1000)
ref = P.oracle_train(sd, batch, noise, cfg)
for mode in sys.argv[1].split(","):
    out = P.run_train(P.make_engine(DEV, cfg, mode, max_points=32768), sd, batch, noise, cfg, DEV)
    e = P.compare_train(out, ref)
    print(mode, "sdf %.3g g %.3g total_loss %.3g dW max rel-Fro %.3g  per tensor: %s" %
          (e["sdf"], e["g"], e["total_loss"], e["grad_max_rel_fro"], " ".join("%.2g" % x for x in e["grad_rel_fro"])))
