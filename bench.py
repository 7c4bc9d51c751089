#!/usr/bin/env python
"""bench.py -- repellency projections / second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one conditioning() call over one batch of Q synthetic queries: query prepare ->
partial sums over the (N-sharded) negative bank -> [one NCCL all-reduce when N > 1] -> fused
correction epilogue.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (Q, N, C, H, W, module semantics, sigma, scale) -- BASELINE.json configs
    "cfg1": dict(Q=1, N=515, C=4, H=64, W=64, kind="fast", sigma=1.0, scale=0.03),
    "cfg2": dict(Q=16, N=515, C=4, H=64, W=64, kind="threshold", sigma=3.15, scale=0.33),
    "cfg3": dict(Q=64, N=3000, C=4, H=64, W=64, kind="fast", sigma=1.0, scale=0.03),
    "cfg4": dict(Q=16, N=515, C=16, H=128, W=128, kind="fast_sdv3", sigma=1.0, scale=0.03),
    "cfg5": dict(Q=128, N=30000, C=4, H=64, W=64, kind="threshold", sigma=3.15, scale=0.33),
}
WORKLOAD_TEXT = {
    "cfg1": "SD-1.4 latent projection, B=1, N=515, 4x64x64 fp32 (BASELINE configs[0])",
    "cfg2": "SD-1.4 nudity, B=8 with CFG (Q=16 rows), N=515, 4x64x64 (BASELINE configs[1], projection only)",
    "cfg3": "CoPro inappropriate bank: N=3000 negatives, B=32 with CFG (Q=64 rows), 4x64x64 fp32, N-sharded "
            "(BASELINE configs[2]; north_star target shape)",
    "cfg4": "SD3 path: 16x128x128 latents, N=515, Q=16, query channel-normalised (BASELINE configs[3])",
    "cfg5": "threshold variant, scaled bank N=30000, Q=128, 4x64x64 (BASELINE configs[4])",
}
FLUSH_BYTES = 256 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg2-loop"],
                    help="cfgN: one projection per step (BASELINE configs); cfg2-loop: the 50-step SD-1.4 sampler loop of "
                         "configs[1] with a random-init stand-in UNet, the 11 in-window steps as ONE CUDA graph")
    ap.add_argument("--path", type=int, default=0, help="kernel family: 0 auto, 1 generic, 2 stream, 3 two-phase tcgen05, 4 tcgen05 reading only the bf16 hi plane, "
                         "5 one-pass tcgen05 (one HBM read of the bank)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipe", action="store_true", help="e2e: only the one-call-at-a-time figure, no calls in flight")
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling: the bank grows with the GPU count (N x gpus rows, a fixed N-row shard per GPU) -- "
                         "the scaled-bank sweep of BASELINE configs[4]; default is strong scaling on a fixed bank")
    ap.add_argument("--sparse", action="store_true",
                    help="let the accumulate pass skip bank-row blocks whose weights are below fp32 resolution "
                         "(library default; off here so that the bench measures the dense worst case)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    return ap.parse_args()


# ----------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------- CPU baseline
def cpu_reference_arm(wl, steps, warmup, sample_q=4):
    """The reference's CPU path on all host threads.  `kind` "reference": the UNMODIFIED reference module
    (oracle/_ref, made by oracle/build_ref.py where /root/reference exists) -- kernel_fast.conditioning(); `kind`
    "port": oracle.materialised_port (the same op chain as fast.py:249-257) when the copy is absent.
    One step = conditioning of `sample_q` query rows against the full bank (the reference's [Q,N,D+1] temporaries
    do not fit at Q = 64)."""
    import tempfile
    import torch
    from oracle import repellency_oracle as orc
    from oracle import build_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    bank = orc.synthetic_bank(wl["N"], wl["C"], wl["H"], wl["W"])
    q = min(sample_q, wl["Q"])
    if wl["kind"] == "threshold":
        q = 1                      # the reference threshold module takes one query row
    x = orc.synthetic_queries(bank, q, "near")
    sdv3 = wl["kind"] == "fast_sdv3"
    sigma = wl["sigma"]
    kind, call = "port", None
    try:
        mod = build_ref.load({"fast": "fast", "fast_sdv3": "fast_sdv3", "threshold": "threshold"}[wl["kind"]])
    except Exception:
        mod = None
    if mod is not None:
        path = os.path.join(tempfile.mkdtemp(prefix="sdn_bench_ref_"), "proj_ref.pt")
        torch.save(bank.clone(), path)
        kw = dict(n_embed=16, proj_ref_path=path, cache_proj_ref=True, scale=wl["scale"])
        if wl["kind"] == "threshold":
            kw.update(sigma=sigma, beta_threshold=1.0, beta_threshold_margin=0.0)
        with torch.no_grad():
            obj = mod.get_repellency_method("kernel_fast", ref_data=torch.zeros(1, 3, 8, 8), embed_fn=None, forward_fn=None,
                                            num_timesteps=50, max_idx=1000, beta_min=0.00085, beta_max=0.012, **kw)
        kind = "reference"
        if wl["kind"] == "threshold":
            call = lambda xin: obj.conditioning(xin, beta_threshold=True)      # noqa: E731
        else:
            call = lambda xin: obj.conditioning(xin)                           # noqa: E731
    else:
        call = lambda xin: orc.conditioning_port(xin, bank, wl["scale"], sigma, 1e-8, sdv3)   # noqa: E731
    times = []
    with torch.no_grad():
        for _ in range(max(warmup, 1)):
            call(x.clone())
        for _ in range(steps):
            xin = x.clone()
            t0 = time.perf_counter()
            call(xin)
            times.append(time.perf_counter() - t0)
    total = sum(times)
    what = ("the reference module itself (oracle/_ref: repellency_methods_%s.kernel_fast.conditioning, fp32 torch on CPU)" % wl["kind"]
            if kind == "reference" else "float32 torch op chain of the reference (oracle.materialised_port)")
    return {"value": q * steps / total, "unit": "projections/s", "cores": torch.get_num_threads(),
            "kind": kind,
            "sample": f"{steps} calls x {q} query rows (of Q={wl['Q']}) against the full N={wl['N']} bank, {what}; "
                      f"best call {min(times)*1e3:.1f} ms",
            "ms_per_step": 1e3 * total / steps, "q": q}


# ----------------------------------------------------------------------------------- main
_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries print banners on stdout (NCCL: 'NCCL version ...'); the contract is ONE JSON line there.  Point fd 1
    at stderr for the whole run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def sampler_loop(args):
    """BASELINE configs[1] as the sampler runs it (SURVEY f-4; ...threshold_time.py:514-576): 50 DDPM steps, B = 8 with
    CFG (16 rows), N = 515; in the window t in [780, 1000] (11 steps) every step is  UNet -> eps->x0 -> projection ->
    gate -> re-noise -> scheduler step  through Projector.ddpm_step (device-side gate, no .item()), and the 11 steps
    are captured in ONE CUDA graph.  diffusers is absent offline, so the UNet is a random-init stand-in of three
    convolutions in stock torch and the out-of-window steps are the plain DDPM update in torch."""
    import torch
    from safe_denoiser_b200 import _native as nv
    from safe_denoiser_b200.epilogue import ddpm_coefficients, ddpm_timesteps, sd14_alphas_cumprod
    from safe_denoiser_b200.projection import NegativeBank, Projector
    from oracle import repellency_oracle as orc
    nv.lib()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    Q, N, C, H, W = 16, 515, 4, 64, 64
    sigma, scale, eps, thr = 3.15, 0.33, 1e-8, 4.54
    bank_cpu = orc.synthetic_bank(N, C, H, W)
    bank = NegativeBank(bank_cpu.to(dev), with_planes=True)
    proj = Projector(bank)
    torch.manual_seed(0)
    unet = torch.nn.Sequential(torch.nn.Conv2d(C, 64, 3, padding=1), torch.nn.SiLU(), torch.nn.Conv2d(64, 64, 3, padding=1),
                               torch.nn.SiLU(), torch.nn.Conv2d(64, C, 3, padding=1)).to(dev).eval()
    ac = sd14_alphas_cumprod()
    ts = ddpm_timesteps(50)
    window = [t for t in ts if 780 <= t <= 1000]
    rest = [t for t in ts if t < 780]
    co = {t: ddpm_coefficients(ac, t) for t in ts}
    g = torch.Generator(device=dev).manual_seed(1)
    x_init = torch.randn(Q, C, H, W, device=dev, generator=g)
    noise = torch.randn(len(ts), 2, Q, C, H, W, device=dev, generator=g)
    x = x_init.clone()

    def in_window(with_projection=True):
        nonlocal x
        for i, t in enumerate(window):
            with torch.no_grad():
                e = unet(x)
            if with_projection:
                x, _ = proj.ddpm_step(x, e, noise[i, 0], noise[i, 1], co[t], sigma, scale, eps, gate_threshold=thr)
            else:
                x = x + 0.0 * e

    def out_of_window():
        nonlocal x
        for i, t in enumerate(rest):
            c = co[t]
            with torch.no_grad():
                e = unet(x)
                x0 = (x - c["sqrt_1m_ab"] * e) / c["sqrt_ab"]
                x = c["c_x0"] * x0 + c["c_xt"] * x + c["sigma_noise"] * noise[len(window) + i, 0]

    def graph_of(fn):
        nonlocal x
        x = x_init.clone()
        fn()                                   # eager warm-up: scratch, kernel attributes, planes
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        x = x_init.clone()
        xin = x
        with torch.cuda.graph(gr):
            fn()
        return gr, xin, x

    c0 = nv.launch_count()
    x = x_init.clone()
    in_window()
    launches_window = nv.launch_count() - c0
    g_win, win_in, win_out = graph_of(in_window)
    g_unet, _, _ = graph_of(lambda: in_window(False))
    g_rest, rest_in, rest_out = graph_of(out_of_window)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    reps = max(3, args.steps // 10)

    def full_loop():
        win_in.copy_(x_init)
        g_win.replay()
        rest_in.copy_(win_out)
        g_rest.replay()

    t_win = timed(g_win.replay, reps)
    t_unet = timed(g_unet.replay, reps)
    t_loop = timed(full_loop, reps)
    finite = bool(torch.isfinite(rest_out).all().item())
    proj_ms = (t_win - t_unet) / len(window)
    line = {"metric": "repellency projections/sec", "unit": "projections/s", "n_gpus": 1, "steps": reps, "warmup": 1,
            "value": Q * len(window) / (max(t_win - t_unet, 1e-9) * 1e-3), "ms_per_step": proj_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gpu_launches": int(launches_window * reps),
            "config": {"workload": "SD-1.4 nudity sampling loop: B=8 with CFG (16 rows), N=515, 50 DDPM steps, random-init "
                                   "stand-in UNet (3 convolutions; diffusers is absent offline) -- BASELINE configs[1]",
                       "Q": Q, "N": N, "latent": [C, H, W], "in_window_steps": len(window), "window": "t in [780, 1000]",
                       "graph": "the 11 in-window steps (UNet + Projector.ddpm_step each, device-side gate, no host sync) are ONE CUDA graph; "
                                "the 39 plain DDPM steps a second one"},
            "loop": {"ms_50_steps": t_loop, "in_window_graph_ms": t_win, "same_graph_without_projection_ms": t_unet,
                     "projection_and_fused_step_ms_per_in_window_step": proj_ms,
                     "projection_share_of_in_window_step": (t_win - t_unet) / t_win,
                     "projection_share_of_the_50_step_loop": (t_win - t_unet) / t_loop,
                     "kernels_per_in_window_step": launches_window / len(window), "latents_finite": finite},
            "e2e": None, "roofline": None, "cpu_baseline": None}
    emit(line)


def main():
    args = parse()
    _quiet_stdout()
    if args.workload == "cfg2-loop":
        return sampler_loop(args)
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    Q, N, C, H, W = wl["Q"], wl["N"], wl["C"], wl["H"], wl["W"]
    if args.weak:
        N = N * max(1, args.gpus)
    D = C * H * W
    base = {"metric": "repellency projections/sec", "unit": "projections/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "bf16 bank planes (opt-in, outside the parity tolerance)" if args.path == 4 else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "Q": Q, "N": N, "latent": [C, H, W],
                       "sigma": wl["sigma"], "scale": wl["scale"], "semantics": wl["kind"],
                       "parallelism": f"N-sharded bank over {args.gpus} GPU(s), one NCCL all-reduce of [Q,D+1] fp32"
                       if args.gpus > 1 else "single GPU"}}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 10))
        r = cpu_reference_arm(wl, steps, min(args.warmup, 1))
        line = dict(base)
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"],
                     "steps": steps, "warmup": min(args.warmup, 1), "n_gpus": args.gpus,
                     "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                     "e2e": {"value": r["value"], "unit": "projections/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})
        line["config"] = dict(base["config"], note="reference arm: the reference is a flat pure-Python repo (nothing to pip "
                              "install); its own module is timed on the host cores from oracle/_ref when present (kind "
                              "'reference'), else the oracle's op-chain port (kind 'port')")
        emit(line)
        return

    import torch
    import torch.distributed as dist
    from safe_denoiser_b200 import _native as nv
    from safe_denoiser_b200.projection import NegativeBank, Projector, conditioning_host, shard_bounds
    from oracle import repellency_oracle as orc   # synthetic inputs + cpu_baseline only

    nv.lib()   # raises if the CUDA library is missing: no fallback
    nv.set_option(nv.OPT_SKIP_NEGLIGIBLE, 1 if args.sparse else 0)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD

    bank_cpu = orc.synthetic_bank(N, C, H, W)
    lo, hi = shard_bounds(N, rank, world)
    bank = NegativeBank(bank_cpu[lo:hi].to(dev), with_planes=(args.path in (0, 3, 4)))
    proj = Projector(bank, path=args.path, group=group)
    x_src = orc.synthetic_queries(bank_cpu, Q, "near").to(dev)
    x = proj.query_buffer(Q, tuple(x_src.shape)) if world > 1 else x_src.clone()
    x.copy_(x_src)
    normalize = C if wl["kind"] == "fast_sdv3" else 0
    sigma, scale, eps = wl["sigma"], wl["scale"], 1e-8
    flush = torch.empty(FLUSH_BYTES, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(FLUSH_BYTES // 4, dtype=torch.int32, device=dev)

    def flush_l2():
        # write a buffer larger than L2 (evicts the bank), then READ another one so that the dirty lines of the
        # write are written back before the timed region starts instead of during it (a 256 MiB memset leaves
        # ~126 MB of dirty lines whose write-back otherwise steals DRAM bandwidth from the first kernel timed)
        flush.zero_()
        flush_rd.sum()

    use_graph = not args.no_graph
    proj.compute_mean = False      # the logging scalar costs an extra pass on the sharded path; not part of the metric

    def step():
        fn = proj.correct_graphed if use_graph else proj.correct
        fn(x, sigma, scale, eps, normalize_channels=normalize, want_num=False,
           gate_threshold=(1.0 if wl["kind"] == "threshold" else None))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        x.copy_(x_src)
        step()
    barrier()

    # ---- timed region: K steps, each between its own CUDA events, L2 flushed in between ----
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    gate = 1.0 if wl["kind"] == "threshold" else None
    # kernels per step: counted on one eager call (graph replays launch the same kernel nodes)
    x.copy_(x_src)
    c0 = nv.launch_count()
    proj.correct(x, sigma, scale, eps, normalize_channels=normalize, gate_threshold=gate)
    launches_per_step = nv.launch_count() - c0
    barrier()
    if world > 1:
        # the ranks leave the host barrier some 100 us apart; a device-side rendezvous (one tiny all-reduce, enqueued
        # and not waited for) lines the GPU timelines up, otherwise the first timed step measures that host skew
        dist.all_reduce(torch.zeros(1, device=dev))
    sync_word = torch.zeros(1, device=dev)
    # a few untimed flushes first: ~1 ms of queued GPU work lets the host run ahead, otherwise step 0 measures the
    # host enqueueing its launch (the GPU has caught up with the host right after the barrier)
    for _ in range(4):
        flush_l2()
    for i in range(args.steps):
        x.copy_(x_src)
        flush_l2()
        if world > 1:
            # the L2 flush (512 MiB of traffic, outside the step's events) does not take the same time on every GPU:
            # without a rendezvous the rank that finishes it first starts its step early and then sits in the merge
            # kernel waiting for the others, and that wait would be billed to the step.  One tiny all-reduce, enqueued
            # and not waited for by the host, lines the GPU timelines up before the start event of every step.
            dist.all_reduce(sync_word)
        ev0[i].record()
        step()
        ev1[i].record()
    barrier()
    launches = launches_per_step * args.steps
    # a step of the job ends when its slowest rank ends: per-step max over ranks, then the sum
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in zip(ev0, ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)
    step_list = step_ms.tolist()
    slowest_step = int(max(range(len(step_list)), key=lambda j: step_list[j]))
    step_sorted = sorted(step_list)
    total_ms = float(step_ms.sum().item())
    ms_per_step = total_ms / args.steps
    value = Q * args.steps / (total_ms * 1e-3)
    pct = lambda f: step_sorted[min(len(step_sorted) - 1, int(f * len(step_sorted)))]   # noqa: E731

    # ---- parity of the timed path, in the same run: the corrected queries of one more step against the float64
    # ---- oracle on two query rows (N = 1), and against rank 0's own full-bank single-GPU projection (N > 1)
    x.copy_(x_src)
    step()
    torch.cuda.synchronize()
    got = x.detach().reshape(Q, D).clone()
    parity = None
    rows = [0, Q - 1] if Q > 1 else [0]
    if rank == 0:
        import numpy as np
        want = orc.conditioning_fast(x_src[rows].cpu().numpy(), bank_cpu.numpy(), scale=scale, sigma=sigma,
                                     sdv3=(wl["kind"] == "fast_sdv3"))["x_0_hat"].reshape(len(rows), D)
        g = got[rows].cpu().numpy().astype(np.float64)
        moved = float(np.abs(want - x_src[rows].reshape(len(rows), D).cpu().numpy()).max())
        parity = {"vs": "float64 oracle (oracle.closed_form) on query rows %s, full bank" % rows,
                  "max_rel_err_x0": float(np.abs(g - want).max() / np.abs(want).max()),
                  "correction_linf": moved, "tolerance": 1e-3}
        if world > 1:
            full = NegativeBank(bank_cpu.to(dev), with_planes=True)
            xs = x_src.clone()
            Projector(full, path=args.path).correct(xs, sigma, scale, eps, normalize_channels=normalize, gate_threshold=gate)
            torch.cuda.synchronize()
            ref1 = xs.reshape(Q, D)
            parity["vs_single_gpu"] = {"what": "all %d rows: N-sharded result vs the same projection over the full bank on rank 0" % Q,
                                       "max_rel_diff_x0": float((got - ref1).abs().max() / ref1.abs().max())}
            del full
        parity["ok"] = bool(parity["max_rel_err_x0"] <= 1e-3 and parity.get("vs_single_gpu", {}).get("max_rel_diff_x0", 0.0) <= 1e-3)

    # ---- per-kernel times of the step itself: the library's own CUDA events around every kernel it launches
    # ---- (sdn_profile_*), with the step replayed from a CUDA graph captured while profiling is on
    kernel_ms = {}
    nv.profile_enable(True)

    def prof_step():
        proj.correct(x, sigma, scale, eps, normalize_channels=normalize, want_num=False, gate_threshold=gate)

    prof_graph = None
    if use_graph and (world == 1 or proj.fused_merge):
        try:
            x.copy_(x_src)
            prof_step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                prof_step()
            g.replay()
            torch.cuda.synchronize()
            if nv.profile_read():
                prof_graph = g
        except Exception:
            prof_graph = None
            torch.cuda.synchronize()
    barrier()
    prof_steps = min(args.steps, 50)
    for i in range(prof_steps):
        x.copy_(x_src)
        flush_l2()
        if prof_graph is not None:
            prof_graph.replay()
        else:
            prof_step()
        for name, ms in nv.profile_read():
            kernel_ms.setdefault(name, []).append(ms)
    barrier()
    nv.profile_enable(False)
    kernel_avg = {k: sum(v) / len(v) for k, v in kernel_ms.items()}
    sampler.stop()

    # ---- e2e: host buffers through the public call, copies inside the timed region ----
    xh_src = x_src.cpu().contiguous()
    xh = xh_src.clone().pin_memory()
    dh = torch.empty(Q, dtype=torch.float32).pin_memory()
    e2e_steps = max(3, min(args.steps, 50))

    def e2e_step():
        if world == 1:
            conditioning_host(bank, xh, dh, sigma, scale, eps, normalize_channels=normalize, path=args.path)
        else:
            # the same graphed N-sharded step as the timed region, fed from and drained to pinned host memory
            x.copy_(xh.view_as(x), non_blocking=True)
            step()
            xh.copy_(x.view_as(xh), non_blocking=True)
            dh.copy_(proj._get(Q, False).denom, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(3):
        xh.copy_(xh_src)
        e2e_step()
    barrier()
    e2e_total = 0.0
    for _ in range(e2e_steps):
        xh.copy_(xh_src)
        flush_l2()
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        e2e_total += time.perf_counter() - t0
    te = torch.tensor([e2e_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_serial = Q * e2e_steps / float(te.item())
    e2e_value, e2e_pipe = e2e_serial, None
    # ---- e2e as a server runs it: independent requests, several host-buffer calls in flight (sdn_host_pipe_*): every
    # step still copies its query in from pinned memory and its corrected query + denominators back out, but step i+1's
    # H2D and step i-1's D2H overlap step i's kernels.  One GPU, plain query, shapes with a fused sequence.
    if world == 1 and normalize == 0 and args.path == 0 and not args.no_pipe:
        from safe_denoiser_b200.projection import HostPipe
        depth = 3
        try:
            pipe = HostPipe(bank, Q, slots=depth)
        except RuntimeError:
            pipe = None
        if pipe is not None:
            xin = xh_src.clone().pin_memory()
            outs = [torch.empty_like(xin).pin_memory() for _ in range(depth)]
            dens = [torch.empty(Q, dtype=torch.float32).pin_memory() for _ in range(depth)]
            try:
                def run(n):
                    for i in range(n):
                        sl = i % depth
                        if i >= depth:
                            pipe.wait(sl)                 # the result of step i - depth is in outs[sl] / dens[sl]
                        pipe.submit(sl, xin, outs[sl], dens[sl], sigma, scale, eps)
                    for sl in range(depth):
                        pipe.wait(sl)
                run(2 * depth)
                # same numbers as the one-at-a-time call
                xh.copy_(xh_src)
                e2e_step()
                same = all(torch.equal(o, xh) for o in outs) and all(torch.equal(d_, dh) for d_ in dens)
                pipe_steps = max(30, 3 * e2e_steps)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                run(pipe_steps)
                dt = time.perf_counter() - t0
                e2e_pipe = {"value": Q * pipe_steps / dt, "steps": pipe_steps, "in_flight": depth,
                            "us_per_step": dt / pipe_steps * 1e6, "equals_serial_call": bool(same)}
                # headline only when the bank cannot stay in the L2 between the un-flushed calls (timing rule: flush, or
                # inputs larger than the L2); a smaller bank keeps the flushed one-call-at-a-time figure as the headline
                l2_bytes = torch.cuda.get_device_properties(dev).L2_cache_size
                e2e_pipe["bank_bytes"] = int(bank.N * D * 4)
                e2e_pipe["l2_bytes"] = int(l2_bytes)
                e2e_pipe["bank_larger_than_l2"] = bool(bank.N * D * 4 > l2_bytes)
                if same and e2e_pipe["bank_larger_than_l2"]:
                    e2e_value = e2e_pipe["value"]
            except RuntimeError as ex:       # a shape without a fused sequence: the one-at-a-time figure stands
                e2e_pipe = {"unavailable": str(ex)[:120]}
            finally:
                pipe.close()

    # ---- the one-HBM-pass kernel on the same workload (one GPU, shapes it takes): timed next to the default path
    one_pass = None
    if world == 1 and args.path == 0 and normalize == 0 and nv.lib().sdn_repel_path(Q, bank.N, D, 1, nv.PATH_FLASH) == nv.PATH_FLASH \
            and D % 1024 == 0 and 8192 <= D <= 16384 and Q > 8:
        pf = Projector(bank, path=nv.PATH_FLASH)
        xo = x_src.clone()
        try:
            for _ in range(3):
                xo.copy_(x_src)
                pf.correct_graphed(xo, sigma, scale, eps, want_num=False, gate_threshold=gate)
            torch.cuda.synchronize()
            o0 = [torch.cuda.Event(enable_timing=True) for _ in range(prof_steps)]
            o1 = [torch.cuda.Event(enable_timing=True) for _ in range(prof_steps)]
            for i in range(prof_steps):
                xo.copy_(x_src)
                flush_l2()
                o0[i].record()
                pf.correct_graphed(xo, sigma, scale, eps, want_num=False, gate_threshold=gate)
                o1[i].record()
            torch.cuda.synchronize()
            oms = sorted(a.elapsed_time(b) for a, b in zip(o0, o1))
            one_pass = {"kernel": "k_flash (SDN_PATH_FLASH, opt-in)", "ms_per_step": sum(oms) / len(oms), "p50_ms": oms[len(oms) // 2],
                        "max_rel_diff_x0_vs_default_path": float((xo.reshape(Q, D) - got).abs().max() / got.abs().max())}
        except RuntimeError as e:
            one_pass = {"error": str(e)[:200]}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        n_local = hi - lo
        npad = (n_local + 127) // 128 * 128
        bank_bytes = n_local * D * (2 if args.path == 4 else 4)   # fp32 bank == bf16 hi+lo planes: 4 B per element
        step_bytes = bank_bytes + n_local * 4 + 2 * Q * D * 4    # ONE pass over the (shard of the) bank + x0 in + x0' out
        # algorithmic bytes of each kernel (what it must read + write once)
        kbytes = {
            "k_stream": step_bytes, "k_stream_reduce": 2 * Q * D * 4, "k_stream_reduce_correct": 3 * Q * D * 4,
            "k_umma_xprep": Q * D * 4 + 128 * D * 2, "k_umma_qprep": Q * D * 4 + 128 * D * 2,
            "k_umma_dots": bank_bytes + 128 * D * 2 + npad * 128 * 4,
            "k_umma_weights": npad * 128 * 4 + 128 * npad * 2,
            "k_umma_listed_accum": 0,
            "k_umma_accum": bank_bytes + 128 * npad * 2 + 2 * Q * D * 4,
            "k_umma_untile": 3 * Q * D * 4,      # tile scratch in, x0 in, x0' out
            "k_flash": step_bytes,
            "k_dots": bank_bytes + Q * D * 4 + Q * n_local * 4, "k_weights": 2 * Q * n_local * 4,
            "k_accum": bank_bytes + Q * n_local * 4 + Q * D * 4,
            "k_shard_merge_correct": (Q * D * 4 // max(1, world)) * (world + 1) + Q * D * 4,
        }
        traffic_db = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic_db = json.load(open(tpath))
        kernels = {}
        for k, ms in kernel_avg.items():
            ent = traffic_db.get(f"{args.workload}/{k}") if world == 1 else None
            b = kbytes.get(k)
            kernels[k] = {"avg_ms": round(ms, 5), "algorithmic_bytes": b,
                          "achieved_gbs": (b / (ms * 1e-3) / 1e9) if b else None,
                          "frac": (b / (ms * 1e-3) / 1e9 / peak) if b else None,
                          "traffic": ent["bytes"] if ent else None, "traffic_source": ent["source"] if ent else None}
        dom = max(kernel_avg, key=kernel_avg.get) if kernel_avg else None
        achieved = step_bytes / (ms_per_step * 1e-3) / 1e9
        traffic_sum = sum(v["traffic"] for v in kernels.values() if v["traffic"]) or None
        if one_pass and "ms_per_step" in one_pass:
            ent = traffic_db.get(f"{args.workload}/k_flash")
            one_pass.update({"frac": step_bytes / (one_pass["ms_per_step"] * 1e-3) / 1e9 / peak,
                             "traffic": ent["bytes"] if ent else None, "traffic_source": ent["source"] if ent else None})
        line = dict(base)
        line["config"] = dict(base["config"],
                              l2="flushed between steps outside the per-step events: 256 MiB memset, then a 256 MiB read so that no dirty lines are left",
                              timing="per-step CUDA-event intervals, max over ranks per step, summed"
                              + ("; at N > 1 a one-element all-reduce before every step's start event aligns the ranks "
                                 "(the L2 flush outside the events takes a different time on every GPU)" if world > 1 else ""),
                              kernel_path=args.path,
                              accumulate_pass="block-sparse (rows below fp32 resolution skipped)" if args.sparse
                              else "dense (every bank row read; the library default would skip negligible rows)",
                              launch="one CUDA graph replay per step (kernels captured from the eager call)"
                              if use_graph else "eager launches",
                              e2e_call="sdn_conditioning_host (C ABI, pinned host buffers)" if world == 1
                              else "pinned host -> graphed N-sharded step (fused NVLink merge) -> pinned host",
                              shard_merge=("fused peer-memory kernel (reduce-scatter + correction + all-gather over NVLink)"
                                           if proj.fused_merge else f"NCCL all-reduce ({proj.fused_merge_error})")
                              if world > 1 else None)
        line.update({
            "value": value, "ms_per_step": ms_per_step, "gpu_launches": int(launches),
            "step_ms": {"p50": pct(0.5), "p95": pct(0.95), "max": step_sorted[-1], "min": step_sorted[0],
                        "slowest_step_index": slowest_step},
            "e2e": {"value": e2e_value, "unit": "projections/s", "h2d_bytes_per_step": Q * D * 4,
                    "d2h_bytes_per_step": Q * D * 4 + Q * 4, "steps": e2e_pipe["steps"] if e2e_pipe and "steps" in e2e_pipe else e2e_steps,
                    "mode": (f"{e2e_pipe['in_flight']} independent host-buffer calls in flight (sdn_host_pipe_submit / _wait), wall clock "
                             "over the whole run, no L2 flush inside (the bank is larger than the L2)")
                    if e2e_pipe and e2e_value == e2e_pipe.get("value")
                    else ("one synchronous host-buffer call at a time" if world == 1
                          else "one step at a time: pinned host -> graphed N-sharded step -> pinned host"),
                    "one_call_at_a_time": {"value": e2e_serial, "steps": e2e_steps,
                                           "note": ("sdn_conditioning_host, synchronous" if world == 1 else "graphed N-sharded step fed from pinned host memory")
                                           + ", L2 flushed between calls"},
                    "pipelined": e2e_pipe},
            "parity_check": parity,
            "roofline": {"bound": "hbm",
                         "what": "the whole step: SURVEY 8(d) bytes of ONE pass over the bank (N*D*4 + N*4 + 2*Q*D*4 per GPU) / ms_per_step",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "algorithmic_bytes": step_bytes, "traffic": traffic_sum,
                         "traffic_note": "sum of the measured DRAM bytes (ncu dram__bytes_read+write) of the step's kernels that have a capture under profiles/",
                         "peak_source": peak_src, "frac_of_nominal_8TBs": achieved / 8000.0,
                         "dominant_kernel": dom, "kernels": kernels,
                         "kernel_timing": "library CUDA events around each kernel, step replayed from a CUDA graph"
                         if prof_graph is not None else "library CUDA events around each kernel, eager launches",
                         "one_pass": one_pass},
            "clocks": sampler.summary(),
        })
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_arm(wl, steps=5, warmup=1)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
