import os, sys, torch
os.environ["ISDFB_DEBUG_CLOCK"]="1"
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P
from isdf_b200.engine import debug_state
DEV=torch.device('cuda:0')
mode=sys.argv[1]; R=int(sys.argv[2])
cfg=O.default_cfg(noise_std=0.05)
sd=C.golden_weights(17,gain=1.5)
batch,noise=C.loss_batch(18,R)
eng=P.make_engine(DEV,cfg,mode,max_points=32768)
eng.pack_weights(P.flat_params(sd,DEV))
b={k:v.to(DEV) for k,v in batch.items()}
lc=P.loss_cfg_from(cfg,R*27)
nz=noise.to(DEV)
for i in range(4):
    eng.zero_grad(); eng.train_fwd_bwd(b['pc'],b['z_vals'],b['depth_sample'],b['dirs_C_sample'],b['T_WC_sample'],b['norm_sample'],nz,lc)
torch.cuda.synchronize()
debug_state(eng)
