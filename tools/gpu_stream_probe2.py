"""Stream path (Q <= 8): fused conditioning call, eager and graph replay, with per-kernel library events.
usage: python tools/gpu_stream_probe2.py   (SDN_STREAM_LEGACY_TILES=1 for the round-1 tile shapes)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200 import _native as nv
from safe_denoiser_b200.projection import NegativeBank, Projector

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for Q, N in ((1, 515), (2, 515), (1, 3000), (4, 3000), (8, 3000), (1, 20000)):
    bank4 = orc.synthetic_bank(N, 4, 64, 64)
    bank = NegativeBank(bank4.cuda(), with_planes=False)
    x4 = orc.synthetic_queries(bank4, Q, "near")
    want = orc.conditioning_fast(x4.numpy(), bank4.numpy(), scale=0.33, sigma=3.15)["x_0_hat"]
    proj = Projector(bank, path=nv.PATH_STREAM)
    x = x4.cuda(); xa = x.clone()
    proj.correct(xa, 3.15, 0.33, 1e-8); torch.cuda.synchronize()
    err = float(np.abs(xa.cpu().numpy() - want).max() / np.abs(want).max())
    res = []
    for graphed in (False, True):
        ts = []
        for _ in range(30):
            xa.copy_(x); flush.zero_(); _ = flush.sum()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); (proj.correct_graphed if graphed else proj.correct)(xa, 3.15, 0.33, 1e-8); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort(); res.append(ts[len(ts) // 2])
    nv.profile_enable(True)
    acc = {}
    for _ in range(10):
        xa.copy_(x); flush.zero_(); _ = flush.sum()
        proj.correct(xa, 3.15, 0.33, 1e-8); torch.cuda.synchronize()
        for name, ms in nv.profile_read(): acc.setdefault(name, []).append(ms * 1e3)
    nv.profile_enable(False)
    gb = (N * 16384 * 4 + N * 4 + 2 * Q * 16384 * 4) / 1e3
    print(f"Q={Q} N={N}: x0 err {err:.1e} | eager {res[0]:.1f} us, graph {res[1]:.1f} us = {gb / res[1] / 6538.6 * 100:.0f}% of peak | "
          + " ".join(f"{k} {sorted(v)[len(v)//2]:.1f}" for k, v in acc.items()), flush=True)
