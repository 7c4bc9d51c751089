"""Two-phase tcgen05 path: per-kernel times (library events, CUDA-graph replay, L2 flushed) for the environment's
SDN_UMMA_L2KEEP_MB / SDN_UMMA_SPLIT_WEIGHTS settings.  usage: python tools/gpu_umma_l2keep.py [Q N]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200 import _native as nv
from safe_denoiser_b200.projection import NegativeBank, Projector

Q, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 3000)
nv.set_option(nv.OPT_SKIP_NEGLIGIBLE, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bank4 = orc.synthetic_bank(N, 4, 64, 64)
bank = NegativeBank(bank4.cuda(), with_planes=True)
x = orc.synthetic_queries(bank4, Q, "near").cuda()
want = orc.conditioning_fast(x.cpu().numpy(), bank4.numpy(), scale=0.33, sigma=3.15)["x_0_hat"]
proj = Projector(bank, path=nv.PATH_UMMA)
xa = x.clone()
proj.correct(xa, 3.15, 0.33, 1e-8)
torch.cuda.synchronize()
import numpy as np
err = float(np.abs(xa.cpu().numpy() - want).max() / np.abs(want).max())
ts = []
for graphed in (False, True):
    t = []
    for _ in range(30):
        xa.copy_(x)
        flush.zero_(); _ = flush.sum()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        (proj.correct_graphed if graphed else proj.correct)(xa, 3.15, 0.33, 1e-8)
        e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1) * 1e3)
    t.sort(); ts.append(t[len(t) // 2])
nv.lib().sdn_profile_enable(1)
acc = {}
for _ in range(10):
    xa.copy_(x)
    flush.zero_(); _ = flush.sum()
    proj.correct(xa, 3.15, 0.33, 1e-8)
    torch.cuda.synchronize()
    for name, ms in nv.profile_read():
        acc.setdefault(name, []).append(ms * 1e3)
print(f"Q={Q} N={N} L2KEEP_MB={os.environ.get('SDN_UMMA_L2KEEP_MB')} SPLIT_WEIGHTS={os.environ.get('SDN_UMMA_SPLIT_WEIGHTS')}: "
      f"x0 err {err:.2e} | eager {ts[0]:.1f} us graph {ts[1]:.1f} us | " + " ".join(f"{k} {sorted(v)[len(v)//2]:.1f}" for k, v in acc.items()), flush=True)
