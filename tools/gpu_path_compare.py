"""Two-phase tcgen05 path vs the one-pass kernel over bank sizes (CUDA events, L2 flushed, fused conditioning call)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200 import _native as nv
from safe_denoiser_b200.projection import NegativeBank, Projector

nv.set_option(nv.OPT_SKIP_NEGLIGIBLE, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for Q in (16, 64):
    for N in (128, 375, 515, 750, 1500, 3000, 6000):
        bank4 = orc.synthetic_bank(N, 4, 64, 64)
        bank = NegativeBank(bank4.cuda(), with_planes=True)
        x = orc.synthetic_queries(bank4, Q, "near").cuda()
        out = []
        for path in (nv.PATH_UMMA, nv.PATH_FLASH):
            proj = Projector(bank, path=path)
            xa = x.clone()
            for graphed in (False, True):
                ts = []
                for _ in range(25):
                    flush.zero_(); _ = flush.sum()
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                    (proj.correct_graphed if graphed else proj.correct)(xa, 3.15, 0.33, 1e-8)
                    e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                out.append(ts[len(ts) // 2])
        print(f"Q={Q:3d} N={N:5d}: two-phase {out[0]:7.1f} us (graph {out[1]:7.1f})   one-pass {out[2]:7.1f} us (graph {out[3]:7.1f})", flush=True)
