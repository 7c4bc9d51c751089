"""Ad-hoc: run one projection shape a few times (target for ncu). usage: gpu_one_shape.py Q N path [C H W]"""
import sys
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200.projection import NegativeBank, Projector
Q, N, path = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
C, H, W = (int(v) for v in sys.argv[4:7]) if len(sys.argv) > 6 else (4, 64, 64)
bank4 = orc.synthetic_bank(N, C, H, W)
bank = NegativeBank(bank4.cuda(), with_planes=(path == 3))
x = orc.synthetic_queries(bank4, Q, "near").cuda()
proj = Projector(bank, path=path)
for _ in range(3):
    proj.partial_sums(x, 3.15)
torch.cuda.synchronize()
print("done")
