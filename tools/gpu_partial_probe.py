"""Three-call sequence (query prepare -> partial sums -> epilogue: what N-sharded banks run per rank) on the tcgen05 path:
graph-replayed step time and per-kernel times.  usage: python tools/gpu_partial_probe.py [Q N]"""
import os, sys
import numpy as np
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
proj._few_launch[Q] = False
xa = x.clone()
proj.correct(xa, 3.15, 0.33, 1e-8, want_num=False)
torch.cuda.synchronize()
err = float(np.abs(xa.cpu().numpy() - want).max() / np.abs(want).max())
t = []
for _ in range(30):
    xa.copy_(x)
    flush.zero_(); _ = flush.sum()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    proj.correct_graphed(xa, 3.15, 0.33, 1e-8, want_num=False)
    e1.record(); torch.cuda.synchronize()
    t.append(e0.elapsed_time(e1) * 1e3)
t.sort()
nv.lib().sdn_profile_enable(1)
acc = {}
for _ in range(10):
    xa.copy_(x)
    flush.zero_(); _ = flush.sum()
    proj.correct(xa, 3.15, 0.33, 1e-8, want_num=False)
    torch.cuda.synchronize()
    for name, ms in nv.profile_read():
        acc.setdefault(name, []).append(ms * 1e3)
print(f"three-call Q={Q} N={N} UNTILE={os.environ.get('SDN_UMMA_UNTILE')}: x0 err {err:.2e} | graph {t[len(t)//2]:.1f} us | "
      + " ".join(f"{k} {sorted(v)[len(v)//2]:.1f}" for k, v in acc.items()), flush=True)
