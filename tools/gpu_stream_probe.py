"""Ad-hoc probe (not a pytest file): time the kernel families on a few shapes."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200 import _native as nv
from safe_denoiser_b200.projection import NegativeBank, Projector

def bench(Q, N, C=4, H=64, W=64, paths=(1, 2), iters=30):
    bank4 = orc.synthetic_bank(N, C, H, W)
    bank = NegativeBank(bank4.cuda(), with_planes=False)
    x4 = orc.synthetic_queries(bank4, Q, "near")
    want = orc.closed_form(x4.numpy(), bank4.numpy(), sigma=3.15)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for path in paths:
        proj = Projector(bank, path=path)
        x = x4.cuda()
        try:
            s = proj.partial_sums(x, 3.15)
        except RuntimeError as e:
            print(f"Q={Q} N={N} D={C*H*W} path={path}: {e}")
            continue
        torch.cuda.synchronize()
        err_n = np.abs(s.num.cpu().numpy() - want["num"]).max() / np.abs(want["num"]).max()
        err_z = np.abs(s.z.cpu().numpy() - want["Z"]).max() / np.abs(want["Z"]).max()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); proj.partial_sums(x, 3.15); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        gb = (N * C * H * W * 4 + N * 4 + 2 * Q * C * H * W * 4) / 1e9
        print(f"Q={Q} N={N} D={C*H*W} path={path}: med {med*1e3:.1f} us min {ts[0]*1e3:.1f} us  "
              f"{gb/med*1e3:.0f} GB/s ({gb/med*1e3/6538.6*100:.1f}% of measured peak)  err num {err_n:.2e} z {err_z:.2e}")

if __name__ == "__main__":
    for Q in (1, 2, 4, 8):
        bench(Q, 515)
    for Q in (1, 4, 8):
        bench(Q, 3000)
    bench(1, 20000)
    bench(2, 515, C=16)
