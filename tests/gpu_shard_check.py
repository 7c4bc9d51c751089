"""Multi-GPU check (run under torchrun): N-sharded projection == single-GPU projection, and timing of the
fused peer-memory merge vs NCCL all-reduce + local epilogue.

    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 --master-port 29555 tests/gpu_shard_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import repellency_oracle as orc                                      # noqa: E402
from safe_denoiser_b200.projection import NegativeBank, Projector, shard_bounds  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for (Q, N) in ((64, 3000), (1, 515), (5, 777)):
        bank4 = orc.synthetic_bank(N, 4, 64, 64)
        x4 = orc.synthetic_queries(bank4, Q, "near")
        full = Projector(NegativeBank(bank4.to(dev)))
        xr = x4.to(dev).clone()
        _, sf = full.correct(xr, 3.15, 0.33, 1e-8, gate_threshold=1.0)
        torch.cuda.synchronize()
        lo, hi = shard_bounds(N, rank, world)
        for fused in (True, False):
            proj = Projector(NegativeBank(bank4[lo:hi].to(dev)), group=dist.group.WORLD)
            proj.fused_merge = fused
            x = proj.query_buffer(Q, tuple(x4.shape)) if fused else x4.to(dev).clone()   # in-place peer-mapped query
            x.copy_(x4.to(dev))
            _, s = proj.correct(x, 3.15, 0.33, 1e-8, gate_threshold=1.0)
            torch.cuda.synchronize()
            err = float((x - xr).abs().max() / xr.abs().max())
            errd = float((s.denom - sf.denom).abs().max() / sf.denom.abs().max())
            same_gate = bool((s.gate == sf.gate).all())
            mode = "fused-peer" if (fused and proj.fused_merge) else "nccl"
            if fused and not proj.fused_merge:
                print(f"[rank {rank}] fused merge unavailable: {proj.fused_merge_error}", flush=True)
            # timing
            proj.compute_mean = False
            use = proj.correct_graphed if (len(sys.argv) > 1 and sys.argv[1] == "graph") else proj.correct
            flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
            ts = []
            for _ in range(30):
                x.copy_(x4.to(dev))
                flush.zero_()
                dist.barrier()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); use(x, 3.15, 0.33, 1e-8); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            good = err <= 1e-5 and errd <= 1e-5 and same_gate
            ok = ok and good
            if rank == 0:
                print(f"Q={Q} N={N} world={world} {mode}: err x0 {err:.2e} denom {errd:.2e} gate {same_gate} "
                      f"median {ts[len(ts)//2]*1e3:.1f} us  {'OK' if good else 'FAIL'}", flush=True)
    # SPELL baseline on an N-sharded bank: the force needs every row of the bank (partial sums + one all-reduce)
    bank4 = orc.synthetic_bank(257, 4, 16, 16)
    x4 = orc.synthetic_queries(bank4, 3, "near")
    xr = x4.to(dev).clone()
    term_f, wsum_f = Projector(NegativeBank(bank4.to(dev))).sparse(xr, 13.0, 1.6, want_term=True)
    lo, hi = shard_bounds(257, rank, world)
    xs = x4.to(dev).clone()
    term_s, wsum_s = Projector(NegativeBank(bank4[lo:hi].to(dev)), group=dist.group.WORLD).sparse(xs, 13.0, 1.6, want_term=True)
    torch.cuda.synchronize()
    e1 = float((xs - xr).abs().max() / xr.abs().max())
    e2 = float((wsum_s - wsum_f).abs().max() / wsum_f.abs().max().clamp_min(1e-30))
    good = e1 <= 1e-5 and e2 <= 1e-5 and float(wsum_f.abs().max()) > 0
    ok = ok and good
    if rank == 0:
        print(f"sparse (SPELL) world={world}: err x0 {e1:.2e} wsum {e2:.2e}  {'OK' if good else 'FAIL'}", flush=True)
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(t.item()) != 1:
        sys.exit(1)


if __name__ == "__main__":
    main()
