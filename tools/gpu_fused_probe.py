"""Ad-hoc: per-kernel event times of the fused conditioning sequence vs the three-call sequence."""
import sys
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200 import _native as nv
from safe_denoiser_b200.projection import NegativeBank, Projector

Q, N = int(sys.argv[1]), int(sys.argv[2])
bank4 = orc.synthetic_bank(N, 4, 64, 64)
bank = NegativeBank(bank4.cuda(), with_planes=Q > 8)
x_src = orc.synthetic_queries(bank4, Q, "near").cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for fused in (True, False):
    proj = Projector(bank)
    if not fused:
        proj._few_launch[Q] = False
    x = x_src.clone()
    acc, tot = {}, []
    nv.profile_enable(True)
    for it in range(25):
        x.copy_(x_src); flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); proj.correct(x, 1.0, 0.03, 1e-8, want_num=False); e1.record()
        torch.cuda.synchronize()
        if it >= 5:
            tot.append(e0.elapsed_time(e1))
            for name, ms in nv.profile_read():
                acc.setdefault(name, []).append(ms)
    nv.profile_enable(False)
    print("fused" if fused else "three-call", "total %.1f us" % (1e3 * sum(tot) / len(tot)),
          {k: round(1e3 * sum(v) / len(v), 1) for k, v in acc.items()})
