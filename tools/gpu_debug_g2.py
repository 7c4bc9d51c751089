import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import repellency_oracle as orc
from safe_denoiser_b200 import _native as nv
from safe_denoiser_b200.projection import NegativeBank, Projector
Q, N = int(sys.argv[1]), int(sys.argv[2]); regime = sys.argv[3]; sigma = float(sys.argv[4])
bank4 = orc.synthetic_bank(N, 4, 64, 64)
bank = NegativeBank(bank4.cuda(), with_planes=True)
x4 = orc.synthetic_queries(bank4, Q, regime)
want = orc.closed_form(x4.numpy(), bank4.numpy(), sigma=sigma)
for sparse in (1, 0):
    nv.set_option(nv.OPT_SKIP_NEGLIGIBLE, sparse)
    for path in (nv.PATH_UMMA, nv.PATH_FLASH, nv.PATH_GENERIC):
        proj = Projector(bank, path=path)
        k = torch.zeros(Q, N, device="cuda")
        s = proj.partial_sums(x4.cuda(), sigma, k_out=k)
        torch.cuda.synchronize()
        z = s.z.cpu().numpy().astype(np.float64); kk = k.cpu().numpy().astype(np.float64)
        ez = np.abs(z - want["Z"].reshape(-1)) / np.abs(want["Z"].reshape(-1))
        ek = np.abs(kk - want["k"]).max(axis=1) / np.abs(want["k"]).max(axis=1)
        zk = np.abs(kk.sum(axis=1) - z) / z
        w = np.argsort(-ez)[:4]
        print(f"sparse={sparse} path={path}: max rel err z {ez.max():.2e} (rows {w.tolist()}), k {ek.max():.2e}, |sum(k_out)-z|/z {zk.max():.2e}")
