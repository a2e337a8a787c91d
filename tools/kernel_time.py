import sys, torch, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P
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
for i in range(3):
    eng.zero_grad(); eng.train_fwd_bwd(b['pc'],b['z_vals'],b['depth_sample'],b['dirs_C_sample'],b['T_WC_sample'],b['norm_sample'],nz,lc)
eng.profile(True)
for i in range(10):
    eng.zero_grad(); eng.train_fwd_bwd(b['pc'],b['z_vals'],b['depth_sample'],b['dirs_C_sample'],b['T_WC_sample'],b['norm_sample'],nz,lc)
pr=eng.profile_read()
print(mode,R,'chain ms %.3f dw ms %.3f'%(pr['chain_ms']/pr['n_chain'],pr['dw_ms']/pr['n_dw']))
x=b['pc'].reshape(-1,3).contiguous()
for want in (False, True):
    for i in range(3): eng.forward(x, want_grad=want)
    eng.profile(True)
    for i in range(10): eng.forward(x, want_grad=want)
    pr=eng.profile_read()
    print(mode,R,'forward%s chain ms %.3f (%d steps)'%('+grad' if want else '', pr['chain_ms']/pr['n_chain'], 14 if want else 7))
