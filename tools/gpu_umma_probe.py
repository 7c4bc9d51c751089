"""Ad-hoc probe: tcgen05 path vs oracle, prints error magnitudes and timings."""
import sys
import numpy as np
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200.projection import NegativeBank, Projector

def run(Q, N, C, H, W, regime="near", sigma=3.15, time_it=False):
    bank4 = orc.synthetic_bank(N, C, H, W)
    bank = NegativeBank(bank4.cuda(), with_planes=True)
    x4 = orc.synthetic_queries(bank4, Q, regime)
    want = orc.closed_form(x4.numpy(), bank4.numpy(), sigma=sigma)
    x = x4.cuda()
    out = {}
    for path in (1, 3):
        proj = Projector(bank, path=path)
        k = torch.zeros(Q, N, device="cuda")
        s = proj.partial_sums(x, sigma, k_out=k)
        torch.cuda.synchronize()
        e = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        out[path] = (e(k.cpu().numpy(), want["k"]), e(s.z.cpu().numpy(), want["Z"]), e(s.num.cpu().numpy(), want["num"]))
        msg = f"Q={Q} N={N} D={C*H*W} {regime} path={path}: err k {out[path][0]:.2e} z {out[path][1]:.2e} num {out[path][2]:.2e}"
        if time_it:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ts = []
            for _ in range(20):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); proj.partial_sums(x, sigma); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            gb = (N * C * H * W * 4 + 2 * Q * C * H * W * 4) / 1e9
            msg += f"  med {ts[len(ts)//2]*1e3:.1f} us -> {gb/ts[len(ts)//2]*1e3:.0f} GB/s"
        print(msg, flush=True)

if __name__ == "__main__":
    run(3, 37, 4, 8, 8)
    run(16, 200, 4, 16, 16)
    run(64, 384, 4, 64, 64)
    run(16, 515, 4, 64, 64, time_it=True)
    run(64, 3000, 4, 64, 64, time_it=True)
    run(64, 3000, 4, 64, 64, regime="far", sigma=1.0)
    run(64, 3000, 4, 64, 64, regime="near", sigma=1.0)
    run(64, 3000, 4, 64, 64, regime="mid", sigma=1.0)
    run(64, 3000, 4, 64, 64, regime="x0", sigma=13.15)
