"""Phase B (k_umma_accum) timeline at one shape: per-CTA %globaltimer stamps (SDN_UMMA_DBG_NOSHARED bit 10 must be set in
the environment), printed as min / median / max over the CTAs, relative to the first CTA's entry.
usage: SDN_UMMA_DBG_NOSHARED=1024 python tools/gpu_accum_trace.py [Q N]"""
import ctypes, os, sys
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
proj = Projector(bank, path=nv.PATH_UMMA)
xa = x.clone()
names = ["entry", "set-up done", "chunk 0: TMEM loads done", "chunk 0: x0 loads landed", "accumulator complete", "epilogue stores issued", "CTA end", "chunk 0: stores issued"]
rows = []
for it in range(6):
    xa.copy_(x)
    flush.zero_(); _ = flush.sum()
    proj.correct(xa, 3.15, 0.33, 1e-8)
    torch.cuda.synchronize()
    buf = np.zeros(256 * 8, dtype=np.uint64)
    nv.lib().sdn_debug_accum_trace_read(buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    t = buf.reshape(256, 8).astype(np.int64)
    t = t[t[:, 0] > 0][:, :8]
    if it >= 2:
        rows.append(t - t[:, 0].min())
t = np.stack(rows).astype(np.float64) / 1e3          # [runs][ctas][events] us
print(f"Q={Q} N={N}: {t.shape[1]} CTAs, {t.shape[0]} runs, env {os.environ.get('SDN_UMMA_DBG_NOSHARED')} lib {os.path.basename(nv.LIB_PATH)}")
for e in (0, 1, 4, 2, 3, 7, 5, 6):
    nm = names[e]
    v = t[:, :, e]
    print(f"  {nm:26s} min {v.min(axis=1).mean():6.1f}  median {np.median(v, axis=1).mean():6.1f}  max {v.max(axis=1).mean():6.1f} us")
d = t[:, :, 5] - t[:, :, 4]
print(f"  epilogue (accumulator complete -> stores issued): min {d.min():.1f} median {np.median(d):.1f} max {d.max():.1f} us")
