"""Scaled-bank sweep (BASELINE configs[4]): the threshold-variant projection of Q = 128 query rows (B = 64 with
CFG) against N-sharded banks of growing size, on 1 or more GPUs of one box.

    python tools/gpu_bank_sweep.py [--rows 3000,12000,48000,192000] [--steps 30]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/gpu_bank_sweep.py

The bank is generated on the device in chunks of 375 rows seeded by the chunk index, so every world size sees the
same bank and the corrected queries can be compared across runs (`checksum`).  Per N, rank 0 prints one JSON line:
step time (CUDA events, L2 flushed between steps, max over ranks), projections/s, and the bank bytes of ONE pass
over all shards divided by the step time, as a fraction of world x the measured single-GPU HBM peak.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_denoiser_b200 import _native as nv  # noqa: E402
from safe_denoiser_b200.projection import NegativeBank, Projector, shard_bounds  # noqa: E402

CHUNK = 375
D = 4 * 64 * 64
Q = 128


def rows(lo, hi, dev):
    assert lo % CHUNK == 0 and hi % CHUNK == 0
    out = torch.empty(hi - lo, D, device=dev)
    g = torch.Generator(device=dev)
    for c in range(lo // CHUNK, hi // CHUNK):
        g.manual_seed(1234 + c)
        out[(c * CHUNK - lo):((c + 1) * CHUNK - lo)].normal_(generator=g)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="3000,12000,48000,192000")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    nv.lib()
    nv.set_option(nv.OPT_SKIP_NEGLIGIBLE, 0)          # dense accumulate: every bank row is read
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks))["hbm_gbs"]) if os.path.exists(peaks) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros((256 << 20) // 4, dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    noise = torch.randn(Q, D, device=dev, generator=g)
    x_src = rows(0, CHUNK, dev)[:Q] + 0.05 * noise      # "near" queries: each next to one negative of chunk 0
    sigma, scale, eps = 3.15, 0.33, 1e-8

    for n in [int(v) for v in args.rows.split(",")]:
        assert n % (CHUNK * world) == 0, "rows must be a multiple of 375 x world"
        lo, hi = shard_bounds(n, rank, world)
        bank = NegativeBank(rows(lo, hi, dev), with_planes=True)
        proj = Projector(bank, group=group)
        proj.compute_mean = False
        x = proj.query_buffer(Q, (Q, D)) if world > 1 else torch.empty(Q, D, device=dev)

        def step():
            proj.correct_graphed(x, sigma, scale, eps, want_num=False, gate_threshold=0.5)

        for _ in range(max(args.warmup, 3)):
            x.copy_(x_src)
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        sync_word = torch.zeros(1, device=dev)
        for _ in range(4):                  # untimed queued work: the host gets ahead of the GPU before step 0
            flush.zero_()
            flush_rd.sum()
        for i in range(args.steps):
            x.copy_(x_src)
            flush.zero_()
            flush_rd.sum()
            if world > 1:
                dist.all_reduce(sync_word)  # the flush takes a different time on every GPU: line the ranks up (as bench.py does)
            e0[i].record()
            step()
            e1[i].record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        per = sorted(a.elapsed_time(b) for a, b in zip(e0, e1))
        t = torch.tensor([sum(per), per[len(per) // 2], per[-1]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)          # every statistic: max over ranks
        ms, ms_median, ms_max = float(t[0].item()) / args.steps, float(t[1].item()), float(t[2].item())
        checksum = float(x.double().sum().item())
        absmax = float((x - x_src).abs().max().item())
        if rank == 0:
            one_pass = n * D * 4.0
            print(json.dumps({"N": n, "Q": Q, "n_gpus": world, "ms_per_step": ms, "ms_median": ms_median, "ms_slowest_step": ms_max,
                              "projections_per_s": Q / (ms * 1e-3),
                              "one_pass_GBs": one_pass / (ms * 1e-3) / 1e9,
                              "frac_of_world_x_peak": one_pass / (ms * 1e-3) / 1e9 / (world * peak),
                              "peak_GBs_per_gpu": peak, "checksum": checksum, "max_correction": absmax,
                              "steps": args.steps}), flush=True)
        del proj, bank, x
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
