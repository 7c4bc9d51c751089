"""Stage latencies of the one-pass kernel from its %globaltimer trace (SDN_FLASH_TRACE=1).

    SDN_FLASH_TRACE=1 python tools/gpu_flash_trace.py [Q N]
"""
import os
import sys

os.environ.setdefault("SDN_FLASH_TRACE", "1")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import repellency_oracle as orc                      # noqa: E402
from safe_denoiser_b200 import _native as nv                     # noqa: E402
from safe_denoiser_b200.projection import NegativeBank, Projector  # noqa: E402

EV = ["tma", "A", "drain", "sent", "l1 start", "l1 stored", "w seen", "P stored", "B", "job seen", "job loaded", "job flag"]


def main(Q=64, N=3000):
    bank4 = orc.synthetic_bank(N, 4, 64, 64)
    bank = NegativeBank(bank4.cuda(), with_planes=True)
    x = orc.synthetic_queries(bank4, Q, "near").cuda()
    proj = Projector(bank, path=nv.PATH_FLASH)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        flush.zero_(); _ = flush.sum()
        proj.partial_sums(x, 3.15)
    torch.cuda.synchronize()
    buf = np.zeros(128 * 64 * 16, dtype=np.uint64)
    n = nv.lib().sdn_debug_trace_read(buf.ctypes.data, buf.nbytes)
    assert n, "no trace (SDN_FLASH_TRACE=1 must be set before the first call)"
    raw = buf.reshape(128, 64, 16).astype(np.float64)
    for e, name in ((12, "owner LL load loop (clk)"), (13, "owner LL attempts"), (14, "owner fence (clk)"), (15, "consumer gw loads (clk)")):
        v = raw[:, :, e][raw[:, :, e] > 0]
        if v.size:
            print(f"{name}: mean {v.mean():.0f}  p50 {np.median(v):.0f}  max {v.max():.0f}  n={v.size}")
    tr = raw.copy()
    ntiles = min(64, (N + 63) // 64)
    t0 = tr[:, :ntiles, 0][tr[:, :ntiles, 0] > 0].min()
    tr = np.where(tr > 0, tr - t0, np.nan)
    print(f"Q={Q} N={N}: {ntiles} tiles traced; all times in us relative to the first TMA issue")
    for cta in (0, 1, 2, 3, 127):
        print(f"-- CTA {cta}: per tile  " + " | ".join(EV[:9]))
        for t in list(range(0, min(ntiles, 6))) + list(range(20, min(ntiles, 30))) + list(range(max(30, ntiles - 2), ntiles)):
            print(f"   t={t:3d} " + " ".join(f"{tr[cta, t, e] / 1e3:8.2f}" for e in range(9)))
    # owner events live on the owner CTA: take the min / max over CTAs per tile
    print("-- per tile over all CTAs (us): tma(min) A(max) sent(max) l1stored(max) | job seen(min..max) loaded(max) flag(max) | w seen(min..max) P(max) B(min..max)")
    for t in list(range(0, min(ntiles, 14))) + list(range(max(14, ntiles - 3), ntiles)):
        g = lambda e, f: f(tr[:, t, e]) / 1e3
        print(f"   t={t:3d} {g(0, np.nanmin):7.2f} {g(1, np.nanmax):7.2f} {g(3, np.nanmax):7.2f} {g(5, np.nanmax):7.2f} | "
              f"{g(9, np.nanmin):7.2f}..{g(9, np.nanmax):7.2f} {g(10, np.nanmax):7.2f} {g(11, np.nanmax):7.2f} | "
              f"{g(6, np.nanmin):7.2f}..{g(6, np.nanmax):7.2f} {g(7, np.nanmax):7.2f} {g(8, np.nanmin):7.2f}..{g(8, np.nanmax):7.2f}")
    print("-- CTA 0 weights loop: top | loads ok | converted | pempty ok | P stored | B issued   (us)")
    for t in range(20, min(ntiles, 30)):
        print(f"   t={t:3d} " + " ".join(f"{tr[0, t, e] / 1e3:8.2f}" for e in (12, 6, 13, 14, 7, 8)))
    lo, hi = min(10, ntiles - 2), ntiles - 1
    for e, name in ((1, "A issue"), (3, "bulk sent"), (5, "l1 stored"), (6, "w seen"), (8, "B issue")):
        per = (tr[:, hi, e] - tr[:, lo, e]) / (hi - lo) / 1e3
        print(f"period of {name:10s} over tiles {lo}..{hi}: mean {np.nanmean(per):.2f} us  min {np.nanmin(per):.2f}  max {np.nanmax(per):.2f}")
    lag_cl = np.nanmax(tr[:, lo:hi, 5], axis=0) - np.nanmin(tr[:, lo:hi, 5], axis=0)
    print(f"skew of 'l1 stored' over the CTAs, per tile: mean {np.nanmean(lag_cl) / 1e3:.2f} us")
    d = lambda a_, b_: np.nanmean(tr[:, 2:ntiles, b_] - tr[:, 2:ntiles, a_]) / 1e3
    print(f"mean per-CTA deltas (us): tma->A {d(0,1):.2f}  A->drain {d(1,2):.2f}  drain->sent {d(2,3):.2f}  sent->l1 {d(3,4):.2f}  "
          f"l1->stored {d(4,5):.2f}  stored->w seen {d(5,6):.2f}  w seen->P {d(6,7):.2f}  P->B {d(7,8):.2f}  tma->B {d(0,8):.2f}")
    print(f"owner: seen->loaded {np.nanmean(tr[:, :ntiles, 10] - tr[:, :ntiles, 9]) / 1e3:.2f}  loaded->flag {np.nanmean(tr[:, :ntiles, 11] - tr[:, :ntiles, 10]) / 1e3:.2f}")
    print(f"kernel span: {np.nanmax(tr[:, :ntiles, 8]) / 1e3:.1f} us to the last B issue")


if __name__ == "__main__":
    main(*(int(v) for v in sys.argv[1:3]))
