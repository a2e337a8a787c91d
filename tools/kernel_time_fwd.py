import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from oracle import isdf_oracle as O
from tests.golden import common as C
from tests import parity as P
DEV=torch.device('cuda:0')
mode=sys.argv[1]; n=int(sys.argv[2])
cfg=O.default_cfg()
sd=C.golden_weights(17,gain=1.5)
eng=P.make_engine(DEV,cfg,mode,max_points=32768)
eng.pack_weights(P.flat_params(sd,DEV))
x=((torch.rand(n,3,generator=C.gen(1))-0.5)*6).to(DEV)
for i in range(5): eng.forward(x)
torch.cuda.synchronize()
