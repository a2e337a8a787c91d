"""Run the UNMODIFIED reference Trainer (CPU) on the synthetic sequence of trainer_case.py and store
what Trainer.step() returned.  python tests/golden/make_trainer_golden.py"""
import io
import contextlib
import json
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from oracle import ref_shim  # noqa: E402
import trainer_case as TC  # noqa: E402
import common as C  # noqa: E402

torch.set_num_threads(8)
ref = ref_shim.load()
tmp = tempfile.mkdtemp(prefix="isdf_golden_")
seq = TC.write_sequence(tmp)
cfg_path = os.path.join(tmp, "cfg.json")
json.dump(TC.config(seq), open(cfg_path, "w"))
probe = (torch.rand(256, 3, generator=C.gen(70)) - 0.5) * torch.tensor([4.0, 3.0, 6.0])
with contextlib.redirect_stdout(io.StringIO()):
    out = TC.run_schedule(ref["trainer"].Trainer, "cpu", cfg_path, probe)
torch.save(out, os.path.join(HERE, "trainer.pt"))
for i, l in enumerate(out["losses"]):
    print(i, l)
print("wrote trainer.pt")
