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

# events per (CTA, 128-row tile): 0 phase-A TMA issue (chunk 0) | 1 first phase-A MMA | 2 second drain starts | 5 level-1 lines
# stored | 6 phase-B loads issued (first half) | 8 phase-B MMAs issued (first half) | 9..11 level-2 unit (worker 0 of the
# CTA that owns it): tile seen, groups loaded, counter released
EV = {0: "tma", 1: "A", 2: "drain", 5: "l1 stored", 6: "B loads", 8: "B mma", 9: "job seen", 10: "job loaded", 11: "job flag"}


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
    tr = raw.copy()
    ntiles = min(64, (N + 127) // 128)
    t0 = tr[:, :ntiles, 0][tr[:, :ntiles, 0] > 0].min()
    tr = np.where(tr > 0, tr - t0, np.nan)
    print(f"Q={Q} N={N}: {ntiles} tiles of 128 rows traced; all times in us relative to the first TMA issue")
    EV.update({3: "drain0", 4: "drain0 done", 7: "A c4", 12: "A c7", 13: "B mma h1", 14: "tma c4", 15: "tma c7"})
    cols = [0, 14, 15, 1, 3, 4, 7, 12, 2, 5, 6, 8, 13]
    for cta in (0, 1, 2, 3, 127):
        print(f"-- CTA {cta}: per tile  " + " | ".join(EV[e] for e in cols))
        for t in range(ntiles):
            print(f"   t={t:3d} " + " ".join(f"{tr[cta, t, e] / 1e3:8.2f}" for e in (cols if not np.isnan(tr[cta, t, 0]) else [6, 8, 13])))
    print("-- per tile over all CTAs (us): tma(min) A(max) l1stored(max) | job seen(min..max) loaded(max) flag(max) | B loads(min..max) B mma(min..max)")
    for t in range(ntiles):
        g = lambda e, f: f(tr[:, t, e]) / 1e3
        print(f"   t={t:3d} {g(0, np.nanmin):7.2f} {g(1, np.nanmax):7.2f} {g(5, np.nanmax):7.2f} | "
              f"{g(9, np.nanmin):7.2f}..{g(9, np.nanmax):7.2f} {g(10, np.nanmax):7.2f} {g(11, np.nanmax):7.2f} | "
              f"{g(6, np.nanmin):7.2f}..{g(6, np.nanmax):7.2f} {g(8, np.nanmin):7.2f}..{g(8, np.nanmax):7.2f}")
    lo, hi = min(4, ntiles - 2), ntiles - 1
    for e in (6, 8):
        per = (tr[:, hi, e] - tr[:, lo, e]) / (hi - lo) / 1e3
        print(f"period of {EV[e]:10s} over tiles {lo}..{hi}: mean {np.nanmean(per):.2f} us  min {np.nanmin(per):.2f}  max {np.nanmax(per):.2f}")
    d = lambda a_, b_: np.nanmean(tr[:, :ntiles, b_] - tr[:, :ntiles, a_]) / 1e3
    print(f"mean per-CTA deltas (us): tma->A {d(0,1):.2f}  A->drain {d(1,2):.2f}  drain->l1 stored {d(2,5):.2f}")
    print(f"owner: seen->loaded {d(9,10):.2f}  loaded->flag {d(10,11):.2f}")
    pub = np.nanmax(tr[:, :ntiles, 11], axis=0)
    st = np.nanmax(tr[:, :ntiles, 5], axis=0)
    print(f"per tile: last level-1 store -> last unit released: mean {np.nanmean(pub - st) / 1e3:.2f} us;  released -> first B loads {np.nanmean(np.nanmin(tr[:, :ntiles, 6], axis=0) - pub) / 1e3:.2f} us")
    print(f"kernel span: {np.nanmax(tr[:, :ntiles, 8]) / 1e3:.1f} us to the last B issue")


if __name__ == "__main__":
    main(*(int(v) for v in sys.argv[1:3]))
