"""Generate the golden fixtures by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz (small tensors) and tests/golden/known_answers.json
(scalars at the SD-1.4 shape 515x4x64x64).  The GPU box has no
/root/reference; tests there read only the committed fixtures.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from repellency import repellency_methods_fast as ref_fast            # noqa: E402
from repellency import repellency_methods_fast_sdv3 as ref_sdv3       # noqa: E402
from repellency import repellency_methods_threshold as ref_thr        # noqa: E402
from oracle.repellency_oracle import synthetic_bank, synthetic_queries  # noqa: E402

torch.set_grad_enabled(False)
TMP = tempfile.mkdtemp(prefix="sdn_golden_")


def build(mod, name, bank, **params):
    path = os.path.join(TMP, f"bank_{bank.shape[0]}_{bank.shape[1]}_{bank.shape[2]}.pt")
    torch.save(bank.clone(), path)
    return mod.get_repellency_method(
        name, ref_data=torch.zeros(1, 3, 8, 8), embed_fn=None, forward_fn=None,
        num_timesteps=50, max_idx=1000, beta_min=0.00085, beta_max=0.012,
        n_embed=16, proj_ref_path=path, cache_proj_ref=True, **params)


def fast_cases(mod, tag, c, h, w, n, out):
    bank = synthetic_bank(n, c, h, w, seed=1234)
    out[f"{tag}/bank"] = bank.numpy()
    proc = build(mod, "kernel_fast", bank, scale=0.03, sigma=3.55)   # sigma ignored (Q4)
    for q in (1, 3):
        for regime in ("far", "x0", "near", "mid"):
            x = synthetic_queries(bank, q, regime, seed=4321 + q)
            neg, item = proc.empirical_denoiser(x_t=x.clone())
            d = proc.conditioning(x.clone(), beta_threshold=False)
            key = f"{tag}/q{q}/{regime}"
            out[key + "/x"] = x.numpy()
            out[key + "/neg"] = neg.numpy()
            out[key + "/x0"] = d["x_0_hat"].numpy()
            out[key + "/item"] = np.float64(d["mean_x_0_hat"])


def batched_cases(out):
    """Batched queries (the CFG-doubled batch): 16 and 70 rows -- the second needs two query groups of 64 in the
    tcgen05 path.  D = 2*8*8 = 128 (one d-block of that path) keeps the file small; the negative mean is not stored,
    it is (x - x0) / scale."""
    c, h, w, n = 2, 8, 8, 130
    bank = synthetic_bank(n, c, h, w, seed=2468)
    out["batched/bank"] = bank.numpy()
    proc = build(ref_fast, "kernel_fast", bank, scale=0.03, sigma=3.55)
    for q in (16, 70):
        for regime in ("near", "mid"):
            x = synthetic_queries(bank, q, regime, seed=1357 + q)
            d = proc.conditioning(x.clone(), beta_threshold=False)
            key = f"batched/q{q}/{regime}"
            out[key + "/x"] = x.numpy()
            out[key + "/x0"] = d["x_0_hat"].numpy()
            out[key + "/item"] = np.float64(d["mean_x_0_hat"])


def threshold_cases(out):
    c, h, w, n = 4, 8, 8, 37
    bank = synthetic_bank(n, c, h, w, seed=1234)
    out["thr/bank"] = bank.numpy()
    for sigma, scale, thr, margin in ((3.15, 0.33, 22.0, 1.6), (1.0, 0.9, 1e-9, 0.0), (13.15, 0.69, 35.8, 0.0)):
        proc = build(ref_thr, "kernel_fast", bank, scale=scale, sigma=sigma,
                     beta_threshold=thr, beta_threshold_margin=margin)
        for regime in ("far", "x0", "near", "mid"):
            x = synthetic_queries(bank, 1, regime, seed=99)
            key = f"thr/s{sigma}/{regime}"
            out[key + "/x"] = x.numpy()
            out[key + "/params"] = np.array([sigma, scale, thr, margin], dtype=np.float64)
            d = proc.conditioning(x.clone(), beta_threshold=True)
            out[key + "/gate/x0"] = d["x_0_hat"].numpy()
            out[key + "/gate/is_negation"] = np.bool_(d["is_negation"])
            out[key + "/gate/denominator"] = np.float64(d["mean_x_0_hat"]["denominator"])
            out[key + "/gate/nominator"] = d["mean_x_0_hat"]["nominator"].numpy()
            out[key + "/gate/item"] = np.float64(d["mean_x_0_hat"]["negative_score_item"])
            xin = x.clone()
            d = proc.conditioning(xin, beta_threshold=False)
            out[key + "/nogate/x0"] = d["x_0_hat"].numpy()          # the negative mean (Q5)
            out[key + "/nogate/x_inplace"] = xin.numpy()            # the corrected query
            out[key + "/nogate/is_negation"] = np.bool_(d["is_negation"])


class _ToyScheduler:
    """duck-type for threshold.set_noisy_proj_ref: set_timesteps/.timesteps/.add_noise"""
    def __init__(self):
        self.ab = torch.linspace(0.9999, 0.005, 1000)

    def set_timesteps(self, n, device=None):
        self.timesteps = torch.arange(n - 1, -1, -1) * (1000 // n) + 1

    def add_noise(self, x0, noise, t):
        a = self.ab[int(t)]
        return a.sqrt() * x0 + (1 - a).sqrt() * noise


def beta_cases(out):
    c, h, w, n = 4, 8, 8, 37
    bank = synthetic_bank(n, c, h, w, seed=1234)
    proc = build(ref_thr, "kernel_fast", bank, scale=0.33, sigma=3.15, beta_threshold=1.0)
    proc.proj_beta_ref_path = os.path.join(TMP, "noisy.pt")
    sched = _ToyScheduler()
    noisy = proc.set_noisy_proj_ref(sched, 5, device="cpu",
                                    generator=torch.Generator().manual_seed(42))
    proc.noisy_proj_refs = noisy
    out["beta/bank"] = bank.numpy()
    out["beta/timesteps"] = np.array(list(noisy.keys()), dtype=np.int64)
    for t, v in noisy.items():
        out[f"beta/noisy/{t}"] = v.numpy()
    for qt in (0.0, 0.25):
        res = proc.empirical_beta(sigma=3.15, quantitle=qt)
        out[f"beta/q{qt}"] = np.array([float(res[t]) for t in noisy.keys()], dtype=np.float64)


def sparse_cases(out):
    c, h, w, n = 4, 8, 8, 37
    bank = synthetic_bank(n, c, h, w, seed=1234)
    out["sparse/bank"] = bank.numpy()
    for tag, mod in (("fast", ref_fast), ("thr", ref_thr)):
        for radius in (4.0, 11.5, 13.0):
            proc = build(mod, "sparse", bank, scale=1.6, radius=radius)
            for regime in ("near", "x0", "mid"):
                x = synthetic_queries(bank, 1, regime, seed=7)
                xin = x.clone()
                d = proc.conditioning(xin, beta_threshold=False)
                key = f"sparse/{tag}/r{radius}/{regime}"
                out[key + "/x"] = x.numpy()
                out[key + "/x0"] = d["x_0_hat"].numpy()
                out[key + "/item"] = np.float64(d["mean_x_0_hat"])
                if "is_negation" in d:
                    out[key + "/is_negation"] = np.bool_(d["is_negation"])


def sparse_sdv3_cases(out):
    c, h, w, n = 16, 4, 4, 29
    bank = synthetic_bank(n, c, h, w, seed=1234)
    out["sparse_sdv3/bank"] = bank.numpy()
    for radius in (5.0, 5.7, 6.5):
        proc = build(ref_sdv3, "sparse", bank, scale=1.6, radius=radius)
        for regime in ("near", "x0", "mid"):
            x = synthetic_queries(bank, 1, regime, seed=7) * 1.7       # un-normalised query
            xin = x.clone()
            d = proc.conditioning(xin, beta_threshold=False)
            key = f"sparse_sdv3/r{radius}/{regime}"
            out[key + "/x"] = x.numpy()
            out[key + "/x0"] = d["x_0_hat"].numpy()
            out[key + "/item"] = np.float64(d["mean_x_0_hat"])


def flow_step_cases(out):
    """The SD3 caller epilogue: the reference pipeline cannot be imported (diffusers is absent), so the
    lines of its in-window branch are text-extracted from models/sdv3/safe_denoiser_pipeline.py and EXECUTED here,
    unmodified, around the reference's own fast_sdv3 processor -- fp16 latents as in the pipeline, including the
    cast back to fp16 (:1161)."""
    import math
    src = open(os.path.join(REF, "models", "sdv3", "safe_denoiser_pipeline.py")).read().splitlines()
    # :1141 "# current time step" ... :1161 "latents = latents.to(latents_dtype)"
    first = next(i for i, l in enumerate(src) if l.strip() == "# current time step")
    last = next(i for i, l in enumerate(src) if i > first and l.strip() == "latents = latents.to(latents_dtype)")
    assert (first + 1, last + 1) == (1141, 1161), (first + 1, last + 1)
    body = [l for l in src[first:last + 1]]
    # the block spans two indentation levels (inside / after `if repellency_processor is not None:`) and holds only
    # simple statements: left-strip every line (the continuation line of the conditioning(...) call sits in parentheses)
    code = "\n".join(l.strip() for l in body)
    c, h, w, n, q = 16, 8, 8, 40, 3
    bank = synthetic_bank(n, c, h, w, seed=1234)
    out["flow/bank"] = bank.numpy()
    proc = build(ref_sdv3, "kernel_fast", bank, scale=0.03, sigma=3.55)
    sig_list = [1.0, 0.98, 0.8, 0.78, 0.02]
    for case, (i, steps) in enumerate(((0, 5), (2, 5), (4, 5))):       # the last one: sigma_next = 0.0
        g = torch.Generator().manual_seed(100 + case)
        idx = torch.randint(0, n, (q,), generator=g)
        lat = (1.7 * bank[idx] + 0.3 * torch.randn(q, c, h, w, generator=g)).half()
        v = (0.5 * torch.randn(q, c, h, w, generator=g)).half()
        ns = {"math": math, "torch": torch, "latents": lat.clone(), "noise_pred": v.clone(),
              "sigmas": torch.tensor(sig_list), "i": i, "num_inference_steps": steps,
              "repellency_processor": proc, "latents_dtype": torch.float16}
        torch.manual_seed(7 + case)
        exec(compile(code, "safe_denoiser_pipeline.py:1141-1161", "exec"), ns)
        torch.manual_seed(7 + case)
        z = torch.randn_like(lat)                      # the draw the executed block made
        key = f"flow/case{case}"
        out[key + "/latents"] = lat.numpy()
        out[key + "/v"] = v.numpy()
        out[key + "/z"] = z.numpy()
        out[key + "/sigma"] = np.float64(float(ns["sigma"]))
        out[key + "/sigma_next"] = np.float64(float(ns["sigma_next"]))
        out[key + "/x0_corrected"] = ns["latents_pred_0_repellenced"].float().numpy()
        out[key + "/latents_next"] = ns["latents"].numpy()      # fp16


def known_answers():
    """Scalars at the real SD-1.4 shape (SURVEY 8c recipe)."""
    g = torch.Generator().manual_seed(1234)
    bank = torch.randn(515, 4, 64, 64, generator=g)
    bank /= bank.norm(dim=1, keepdim=True)
    g2 = torch.Generator().manual_seed(4321)
    xf = torch.randn(1, 4, 64, 64, generator=g2)
    xn = bank[7:8] + 0.05 * torch.randn(1, 4, 64, 64, generator=g2)
    ka = {"recipe": "bank seed 1234 randn(515,4,64,64)/channel-norm; g2 seed 4321: xf=randn, xn=bank[7:8]+0.05*randn"}
    thr = build(ref_thr, "kernel_fast", bank, scale=0.33, sigma=3.15, beta_threshold=1e-9)
    for name, x in (("xf", xf), ("half_xf", 0.5 * xf), ("xn", xn)):
        xin = x.clone()
        d = thr.conditioning(xin, beta_threshold=True)
        ka[f"threshold/{name}"] = {
            "denominator": d["mean_x_0_hat"]["denominator"],
            "max_abs_delta": float((xin - x).abs().max()),
            "sum_x0": float(xin.double().sum()),
            "is_negation": bool(d["is_negation"]),
            "x0_first8": xin.reshape(-1)[:8].tolist(),
        }
    fast = build(ref_fast, "kernel_fast", bank, scale=0.03, sigma=3.55)
    for name, x in (("xf", xf), ("xn", xn)):
        xin = x.clone()
        d = fast.conditioning(xin, beta_threshold=False)
        ka[f"fast/{name}"] = {
            "mean_x_0_hat": d["mean_x_0_hat"],
            "max_abs_delta": float((xin - x).abs().max()),
            "sum_x0": float(xin.double().sum()),
        }
    return ka


def main():
    if "--flow-only" in sys.argv:           # added in round 2: leaves the other fixture files untouched
        fx = {}
        flow_step_cases(fx)
        np.savez_compressed(os.path.join(HERE, "flow_cases.npz"), **fx)
        print("flow_cases.npz written to", HERE)
        return
    if "--batched-only" in sys.argv:        # added later: leaves the other fixture files untouched
        fx = {}
        batched_cases(fx)
        np.savez_compressed(os.path.join(HERE, "batched_cases.npz"), **fx)
        print("batched_cases.npz written to", HERE)
        return
    fx = {}
    batched_cases(fx)
    np.savez_compressed(os.path.join(HERE, "batched_cases.npz"), **fx)
    fx = {}
    fast_cases(ref_fast, "fast", 4, 8, 8, 37, fx)
    fast_cases(ref_sdv3, "sdv3", 16, 4, 4, 29, fx)
    np.savez_compressed(os.path.join(HERE, "fast_cases.npz"), **fx)
    fx = {}
    threshold_cases(fx)
    np.savez_compressed(os.path.join(HERE, "threshold_cases.npz"), **fx)
    fx = {}
    beta_cases(fx)
    np.savez_compressed(os.path.join(HERE, "beta_cases.npz"), **fx)
    fx = {}
    sparse_cases(fx)
    sparse_sdv3_cases(fx)
    np.savez_compressed(os.path.join(HERE, "sparse_cases.npz"), **fx)
    fx = {}
    flow_step_cases(fx)
    np.savez_compressed(os.path.join(HERE, "flow_cases.npz"), **fx)
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(known_answers(), f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
