"""Bring-up tool for the tensor-core path (not a pytest): runs K2 / K3 / K4 on a small batch and prints,
stage by stage, the error of every intermediate (decoded from the per-tile side arrays through
isdfb_debug_buffers) against the fp64 oracle.  Usage: python tests/tc_debug.py [bf16x3|bf16] [n_rays]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import isdf_oracle as O  # noqa: E402
from tests.golden import common as C  # noqa: E402
from tests import parity as P  # noqa: E402
from isdf_b200.engine import debug_state  # noqa: E402

DEV = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    cfg = O.default_cfg(noise_std=0.05, transform=C.rigid_transform(3))
    sd = C.golden_weights(17, gain=1.5)
    batch, noise = C.loss_batch(18, R)
    n = R * 27
    nt = (n + 127) // 128
    L = 6
    ref = P.oracle_train(sd, batch, noise, cfg)
    refi = O.step_sweeps([(w.double(), b.double()) for w, b in O.layers_from_state_dict(sd, 2)],
                         {k: v.double() for k, v in batch.items()}, dict(cfg, transform=cfg["transform"].double()),
                         noise.double(), keep_intermediates=True)
    eng = P.make_engine(DEV, cfg, mode, max_points=4096)
    eng.pack_weights(P.flat_params(sd, DEV))
    torch.cuda.synchronize()
    print("== mode", mode, "rays", R, "points", n, "tiles", nt)

    x = batch["pc"].reshape(-1, 3).to(DEV)
    nz = noise.reshape(-1).to(DEV)
    sdf = eng.forward(x, noise=nz, noise_std=cfg["noise_std"])
    torch.cuda.synchronize()
    print("K2 forward       sdf rel err %.3e" % rel(sdf.reshape(R, 27), ref["sdf"]))
    sdf2, g = eng.forward(x, noise=nz, noise_std=cfg["noise_std"], want_grad=True)
    torch.cuda.synchronize()
    print("K3 forward+grad  sdf %.3e  g %.3e" % (rel(sdf2.reshape(R, 27), ref["sdf"]), rel(g.reshape(R, 27, 3), ref["g"])))

    out = P.run_train(eng, sd, batch, noise, cfg, DEV)
    errs = P.compare_train(out, ref)
    print("K4 train:", {k: ("%.3e" % v if not isinstance(v, list) else ["%.2e" % t for t in v]) for k, v in errs.items()})

    get_aux, get_dwl, get_sig = debug_state(eng)
    A_ZB2, A_PART, A_E32, A_HL = 0, L, L + 3, L + 4
    D_YH, D_YA, D_XD, D_XZ, D_V = 0, L, 2 * L, 3 * L, 4 * L

    def show(name, got, want):
        want = want[:n]
        got = got[:n, :want.shape[1]]
        print("   %-14s rel err %.3e   (ref max %.3e)" % (name, rel(got, want), float(want.abs().max())))

    print("-- intermediates (tile-decoded) vs fp64 oracle")
    show("e32", get_aux(A_E32, nt), refi["e"])
    show("Yh[0]=e", get_dwl(D_YH, nt), refi["e"])
    for l in range(L):
        show("sig[%d]" % l, get_sig(l, nt, L), refi["sig"][l])
    for l in range(1, L):
        show("Yh[%d]=h%d" % (l, l - 1), get_dwl(D_YH + l, nt), refi["inps"][l][:, :256])
    show("h_last", get_aux(A_HL, nt), refi["h_last"])
    for l in range(L - 1, -1, -1):
        show("Xd[%d]=delta" % l, get_dwl(D_XD + l, nt), refi["delta"][l])
    show("Ya[0]=abar_e", get_dwl(D_YA, nt), refi["abar_e"])
    for l in range(1, L):
        show("Ya[%d]=abar%d" % (l, l - 1), get_dwl(D_YA + l, nt), refi["abars"][l - 1])
    for l in range(L - 1):
        show("zb2[%d]" % l, get_aux(A_ZB2 + l, nt), refi["zbar2"][l])
    for l in range(L - 1, -1, -1):
        show("Xz[%d]=zbar" % l, get_dwl(D_XZ + l, nt), refi["zbars"][l])
    vref = refi["s_bar"].reshape(-1, 1) * refi["h_last"] + refi["abars"][L - 1]
    show("V", get_dwl(D_V, nt), vref)


if __name__ == "__main__":
    main()
