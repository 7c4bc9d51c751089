"""Pin the oracle: every restatement in oracle/ vs fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import repellency_oracle as orc

G = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def fast_fx():
    return np.load(os.path.join(G, "fast_cases.npz"))


@pytest.mark.parametrize("tag,sdv3", [("fast", False), ("sdv3", True)])
@pytest.mark.parametrize("q", [1, 3])
@pytest.mark.parametrize("regime", ["far", "x0", "near", "mid"])
def test_fast_closed_form_matches_reference(fast_fx, tag, sdv3, q, regime):
    bank = fast_fx[f"{tag}/bank"]
    key = f"{tag}/q{q}/{regime}"
    x = fast_fx[key + "/x"]
    r = orc.conditioning_fast(x, bank, scale=0.03, sigma=1.0, sdv3=sdv3)
    # reference is fp32: agree to fp32 rounding of the expansion (2-5e-7 typical)
    assert rel(r["neg"], fast_fx[key + "/neg"]) < 5e-5
    assert rel(r["x_0_hat"], fast_fx[key + "/x0"]) < 1e-6
    assert abs(r["mean_x_0_hat"] - float(fast_fx[key + "/item"])) < 1e-6 * max(1.0, abs(r["mean_x_0_hat"]))


@pytest.mark.parametrize("q", [16, 70])
@pytest.mark.parametrize("regime", ["near", "mid"])
def test_batched_queries_match_reference(q, regime):
    """CFG-doubled batches (16 and 70 query rows) through the reference's fast module."""
    fx = np.load(os.path.join(G, "batched_cases.npz"))
    key = f"batched/q{q}/{regime}"
    x = fx[key + "/x"]
    r = orc.conditioning_fast(x, fx["batched/bank"], scale=0.03, sigma=1.0)
    assert rel(r["x_0_hat"], fx[key + "/x0"]) < 1e-6
    assert rel(r["neg"], (x - fx[key + "/x0"]) / 0.03) < 5e-4      # (x - x0) / scale loses digits to cancellation
    assert abs(r["mean_x_0_hat"] - float(fx[key + "/item"])) < 1e-6 * max(1.0, abs(r["mean_x_0_hat"]))


@pytest.mark.parametrize("q", [1, 3])
@pytest.mark.parametrize("regime", ["far", "x0", "near", "mid"])
def test_materialised_port_matches_reference(fast_fx, q, regime):
    for tag, sdv3 in (("fast", False), ("sdv3", True)):
        bank = torch.from_numpy(fast_fx[f"{tag}/bank"])
        key = f"{tag}/q{q}/{regime}"
        x = torch.from_numpy(fast_fx[key + "/x"]).clone()
        neg, item, _, _ = orc.materialised_port(x.clone(), bank, 1.0, 1e-8, normalise_query=sdv3)
        np.testing.assert_allclose(neg.numpy(), fast_fx[key + "/neg"], rtol=1e-5, atol=1e-7)
        x0, _, _ = orc.conditioning_port(x, bank, 0.03, normalise_query=sdv3)
        np.testing.assert_allclose(x0.numpy(), fast_fx[key + "/x0"], rtol=1e-6, atol=1e-7)


def test_threshold_cases_match_reference():
    fx = np.load(os.path.join(G, "threshold_cases.npz"))
    bank = fx["thr/bank"]
    seen_gate = set()
    for sigma in (3.15, 1.0, 13.15):
        for regime in ("far", "x0", "near", "mid"):
            key = f"thr/s{sigma}/{regime}"
            sg, scale, thr, margin = fx[key + "/params"]
            x = fx[key + "/x"]
            r = orc.conditioning_threshold(x, bank, sg, scale, 1e-8, thr, margin, use_gate=True)
            assert rel(r["x_0_hat"], fx[key + "/gate/x0"]) < 1e-5
            assert rel(r["denominator"], fx[key + "/gate/denominator"]) < 2e-5
            assert rel(r["nominator"], fx[key + "/gate/nominator"]) < 5e-5
            assert bool(r["is_negation"][0]) == bool(fx[key + "/gate/is_negation"])
            seen_gate.add(bool(r["is_negation"][0]))
            r = orc.conditioning_threshold(x, bank, sg, scale, 1e-8, thr, margin, use_gate=False)
            assert rel(r["x_0_hat"], fx[key + "/nogate/x0"]) < 5e-5          # negative mean (Q5)
            assert rel(r["x_0_hat_inplace"], fx[key + "/nogate/x_inplace"]) < 1e-5
            assert bool(fx[key + "/nogate/is_negation"]) is True
    assert seen_gate == {True, False}, "fixtures must exercise both gate outcomes"


def test_empirical_beta_matches_reference():
    fx = np.load(os.path.join(G, "beta_cases.npz"))
    bank = fx["beta/bank"]
    ts = [int(t) for t in fx["beta/timesteps"]]
    noisy = {t: fx[f"beta/noisy/{t}"] for t in ts}
    for qt in (0.0, 0.25):
        got = orc.empirical_beta(noisy, bank, sigma=3.15, quantile=qt)
        np.testing.assert_allclose([got[t] for t in ts], fx[f"beta/q{qt}"], rtol=2e-5)


def test_sparse_matches_reference():
    fx = np.load(os.path.join(G, "sparse_cases.npz"))
    bank = fx["sparse/bank"]
    outcomes = set()
    for tag in ("fast", "thr"):
        for radius in (4.0, 11.5, 13.0):
            for regime in ("near", "x0", "mid"):
                key = f"sparse/{tag}/r{radius}/{regime}"
                r = orc.sparse_repellency(fx[key + "/x"], bank, radius, scale=1.6)
                assert rel(r["x_0_hat"], fx[key + "/x0"]) < 2e-5, key
                assert abs(r["force_norm"] - float(fx[key + "/item"])) <= 2e-4 * max(1.0, r["force_norm"])
                if tag == "thr":
                    assert r["is_negation"] == bool(fx[key + "/is_negation"])
                    outcomes.add(r["is_negation"])
    assert outcomes == {True, False}
    # SD3 variant: force on the channel-normalised query, update on the original (fast_sdv3.py:332)
    bank3 = fx["sparse_sdv3/bank"]
    for radius in (5.0, 5.7, 6.5):
        for regime in ("near", "x0", "mid"):
            key = f"sparse_sdv3/r{radius}/{regime}"
            r = orc.sparse_repellency(fx[key + "/x"], bank3, radius, scale=1.6, normalise_query=True)
            assert rel(r["x_0_hat"], fx[key + "/x0"]) < 2e-5, key
            assert abs(r["force_norm"] - float(fx[key + "/item"])) <= 2e-4 * max(1.0, r["force_norm"])


def test_known_answers_sd14_shape():
    """SURVEY 8c anchors at 515x4x64x64, regenerated from seeds."""
    ka = json.load(open(os.path.join(G, "known_answers.json")))
    g = torch.Generator().manual_seed(1234)
    bank = torch.randn(515, 4, 64, 64, generator=g)
    bank /= bank.norm(dim=1, keepdim=True)
    g2 = torch.Generator().manual_seed(4321)
    xf = torch.randn(1, 4, 64, 64, generator=g2)
    xn = bank[7:8] + 0.05 * torch.randn(1, 4, 64, 64, generator=g2)
    b = bank.numpy()
    for name, x in (("xf", xf), ("half_xf", 0.5 * xf), ("xn", xn)):
        r = orc.conditioning_threshold(x.numpy(), b, 3.15, 0.33, 1e-8, 1e-9, 0.0, True)
        want = ka[f"threshold/{name}"]
        assert abs(r["denominator"][0] / want["denominator"] - 1) < 2e-5
        assert abs(r["x_0_hat"].sum() - want["sum_x0"]) < 1e-3
        np.testing.assert_allclose(r["x_0_hat"].reshape(-1)[:8], want["x0_first8"], rtol=1e-5)
        assert abs(np.abs(r["x_0_hat"] - x.numpy()).max() - want["max_abs_delta"]) < 1e-6
    for name, x in (("xf", xf), ("xn", xn)):
        r = orc.conditioning_fast(x.numpy(), b, scale=0.03, sigma=1.0)
        want = ka[f"fast/{name}"]
        assert abs(np.abs(r["x_0_hat"] - x.numpy()).max() - want["max_abs_delta"]) < 1e-6
        assert abs(r["mean_x_0_hat"] - want["mean_x_0_hat"]) < 2e-3 * abs(want["mean_x_0_hat"]) + 1e-30


def test_flow_step_oracle_against_reference_executed_lines():
    """SURVEY A8 pin: flow_cases.npz was produced by EXECUTING models/sdv3/safe_denoiser_pipeline.py:1141-1161 (text
    extracted by make_golden.py) around the reference's fast_sdv3 processor, fp16 latents and the cast back to fp16
    included.  The float64 restatement must agree to fp16 rounding (the reference forms x0 = latents - sigma v in
    fp16: eps 9.8e-4)."""
    from oracle import scheduler_oracle as so
    d = np.load(os.path.join(G, "flow_cases.npz"))
    for c in range(3):
        k = f"flow/case{c}"
        x, v, z = (d[k + s].astype(np.float64) for s in ("/latents", "/v", "/z"))
        sg, sn = float(d[k + "/sigma"]), float(d[k + "/sigma_next"])
        ref = orc.closed_form(x - sg * v, d["flow/bank"], sigma=1.0, normalise_query=True)
        w = so.flow_fused_step(x, v, ref["neg"].reshape(x.shape), 0.03, sg, sn, z)
        want_x0c = d[k + "/x0_corrected"].astype(np.float64)
        want_next = d[k + "/latents_next"].astype(np.float64)
        assert d[k + "/latents_next"].dtype == np.float16
        assert np.abs(w["x0_corrected"] - want_x0c).max() / np.abs(want_x0c).max() <= 1.5e-3
        assert np.abs(w["next"] - want_next).max() / np.abs(want_next).max() <= 1.5e-3
    assert float(d["flow/case2/sigma_next"]) == 0.0        # the last step: sigmas[i+1] does not exist


def test_chunked_closed_form_equals_closed_form():
    bank = orc.synthetic_bank(300, 4, 8, 8)
    x = orc.synthetic_queries(bank, 5, "near")
    a = orc.closed_form(x.numpy(), bank.numpy(), sigma=3.15)
    b = orc.closed_form_chunked(x.numpy(), bank.numpy(), sigma=3.15, chunk=64, weight_rows=(0, 4))
    assert np.allclose(a["Z"], b["Z"], rtol=1e-12) and np.allclose(a["num"], b["num"], rtol=1e-10, atol=1e-14)
    assert np.allclose(a["k"][[0, 4]], b["k"], rtol=1e-12)
