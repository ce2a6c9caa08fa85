"""debug: graph replay with/without pool in fp32/tf32 (small nets)"""
import sys, os, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supervised_gan_b200 as S
from supervised_gan_b200.fcgan_model import FCGANModel
from tests.test_gpu_bench_config import fcgan_opt
prec, pool, ngf = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
S.set_precision(prec)
m = FCGANModel(); m.initialize(fcgan_opt(batchSize=2, fineSize=128, noiseSize=2, pool_size=pool, ngf=ngf, ndf=ngf, cuda_graph=True, batch_D_passes=os.environ.get('DBG_BATCHED','1')=='1'))
for t in range(6):
    m.input.copy_(torch.rand(2, 2, 128, 128, device="cuda") * 2 - 1)
    m.optimize_parameters()
    torch.cuda.synchronize()
    print(prec, pool, ngf, "step", t, float(m.loss_G), "graph" if m._graph is not None else "eager", flush=True)
