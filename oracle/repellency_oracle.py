"""CPU restatement of Safe_Denoiser's per-step repellency projection.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product path
(``safe_denoiser_b200``) never imports this file.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the pin is the reference itself: the modules under
``/root/reference/repellency`` were imported unchanged in the build container
and their outputs on seeded synthetic inputs were frozen into
``tests/golden/*.npz`` / ``tests/golden/known_answers.json`` by
``tests/golden/make_golden.py``.  ``tests/test_oracle_golden.py`` checks every
function below against those fixtures.

Two restatements are kept side by side:

* ``closed_form`` -- float64 numpy, one line of maths per quantity.  This is
  the checker the CUDA kernels are compared with.
* ``materialised_port`` -- float32 torch, following the reference's op
  sequence (cdist -> broadcast to [Q,N,D+1] -> exp -> multiply by the bank
  augmented with a ones column -> sum over N).  It is what ``bench.py`` times
  as the CPU baseline because it moves the same bytes the reference moves.

Reference lines restated (all under /root/reference/repellency/):
  repellency_methods_fast.py:223-262       kernel_fast.empirical_denoiser
  repellency_methods_fast.py:120-132       conditioning / conditioning_1
  repellency_methods_fast_sdv3.py:229-271  same + query channel-normalisation (:239)
  repellency_methods_threshold.py:171-193  conditioning_threshold / conditioning_1
  repellency_methods_threshold.py:309-349  kernel_fast.empirical_denoiser (dict return)
  repellency_methods_threshold.py:351-384  empirical_beta
  repellency_methods_fast.py:306-340       SparseRepellency (SPELL baseline)
  repellency_methods_threshold.py:415-454  SparseRepellency with is_negation
"""
from __future__ import annotations

import numpy as np

try:  # torch is only needed by the float32 port and the quantile helper
    import torch
except Exception:  # pragma: no cover
    torch = None


# ---------------------------------------------------------------------------
# float64 closed form
# ---------------------------------------------------------------------------

def _flat64(a):
    a = np.asarray(a, dtype=np.float64)
    return a.reshape(a.shape[0], -1)


def channel_normalise(x4):
    """x / ||x||_2 over dim=1 (fast_sdv3.py:239; also ``project``, fast.py:55-56)."""
    x4 = np.asarray(x4, dtype=np.float64)
    nrm = np.sqrt((x4 * x4).sum(axis=1, keepdims=True))
    return x4 / nrm


def closed_form(x, bank, sigma=1.0, eps=1e-8, dist_power=1, bank_alpha=1.0,
                normalise_query=False):
    """k_i = exp(-dist(x, a*n_i) / (2 sigma^2)); neg = sum k_i n_i / (sum k_i + eps).

    fast.py:249-257.  ``dist_power`` 1 is the reference (un-squared L2,
    SURVEY Q1); 2 is the squared form of the paper / dead ``lsh`` variant
    (fast.py:413).  ``bank_alpha`` scales the bank inside the distance only.
    ``normalise_query`` applies the SD3 per-pixel channel normalisation to the
    query before the distance (x must then be 4-D).

    Returns dict of float64 arrays: dist [Q,N], k [Q,N], Z [Q], denom [Q]
    (= Z + eps), num [Q,D], neg [Q,D], weights [Q,N] (= k / denom).
    """
    xq = channel_normalise(x) if normalise_query else np.asarray(x, dtype=np.float64)
    xf = _flat64(xq)
    bf = _flat64(bank)
    d2 = ((xf * xf).sum(1)[:, None] + (bank_alpha ** 2) * (bf * bf).sum(1)[None, :]
          - 2.0 * bank_alpha * (xf @ bf.T))
    d2 = np.maximum(d2, 0.0)
    dist = np.sqrt(d2) if dist_power == 1 else d2
    k = np.exp(-dist / (2.0 * float(sigma) ** 2))
    Z = k.sum(1)
    num = k @ bf
    denom = Z + eps
    return {
        "dist": dist, "k": k, "Z": Z, "denom": denom, "num": num,
        "neg": num / denom[:, None], "weights": k / denom[:, None],
    }


def closed_form_chunked(x, bank, sigma=1.0, eps=1e-8, chunk=2048, weight_rows=()):
    """closed_form for banks too large for float64 copies (BASELINE configs[4]: N = 30 000 .. 200 000): the bank rows
    are taken ``chunk`` at a time (float64 temporaries [chunk, D] and [Q, chunk]); same arithmetic, dist_power 1.
    Returns float64 Z [Q], denom [Q], num [Q,D], neg [Q,D] and k [len(weight_rows), N] for the listed query rows."""
    xf = _flat64(x)
    bank = np.asarray(bank)
    n = bank.shape[0]
    q, d = xf.shape
    xs = (xf * xf).sum(1)
    Z = np.zeros(q)
    num = np.zeros((q, d))
    rows = list(weight_rows)
    k_rows = np.zeros((len(rows), n))
    for lo in range(0, n, chunk):
        bf = bank[lo:lo + chunk].reshape(min(chunk, n - lo), -1).astype(np.float64)
        d2 = np.maximum(xs[:, None] + (bf * bf).sum(1)[None, :] - 2.0 * (xf @ bf.T), 0.0)
        k = np.exp(-np.sqrt(d2) / (2.0 * float(sigma) ** 2))
        Z += k.sum(1)
        num += k @ bf
        if rows:
            k_rows[:, lo:lo + bf.shape[0]] = k[rows]
    denom = Z + eps
    return {"Z": Z, "denom": denom, "num": num, "neg": num / denom[:, None], "k": k_rows}


def conditioning_fast(x4, bank4, scale=1.0, eps=1e-8, sigma=1.0, sdv3=False):
    """fast.py:120-132 (sdv3: fast_sdv3.py:120-132).

    sigma is 1.0 whatever the YAML says.  Returns the corrected
    x0 (the reference mutates in place) and the logging scalar.
    """
    r = closed_form(x4, bank4, sigma=sigma, eps=eps, normalise_query=sdv3)
    neg = r["neg"].reshape(np.shape(x4))
    x0 = np.asarray(x4, dtype=np.float64) - scale * neg
    return {"x_0_hat": x0, "mean_x_0_hat": float(np.clip(r["neg"], -1e10, 1e10).mean()),
            "neg": neg, "denom": r["denom"], "weights": r["weights"]}


def conditioning_threshold(x4, bank4, sigma, scale, eps, beta_threshold, margin=0.0,
                           use_gate=True):
    """threshold.py:171-193.

    use_gate=True  -> conditioning_threshold: returns corrected x0 and
                      is_negation = denominator > beta_threshold - margin.
    use_gate=False -> conditioning_1: returns the NEGATIVE MEAN under
                      "x_0_hat"  and is_negation=True; the query is
                      still corrected in place (reported as "x_0_hat_inplace").
    The reference supports Q == 1 only (threshold.py:348); here ``denominator``
    and ``is_negation`` are per-row arrays and collapse to scalars when Q == 1.
    """
    r = closed_form(x4, bank4, sigma=sigma, eps=eps)
    neg = r["neg"].reshape(np.shape(x4))
    corrected = np.asarray(x4, dtype=np.float64) - scale * neg
    denom = r["denom"]
    if use_gate:
        gate = denom > (beta_threshold - margin)
        ret = corrected
    else:
        gate = np.ones_like(denom, dtype=bool)
        ret = neg
    return {"x_0_hat": ret, "x_0_hat_inplace": corrected, "is_negation": gate,
            "denominator": denom, "nominator": r["num"],
            "negative_score_item": float(np.clip(r["neg"], -1e10, 1e10).mean()),
            "weights": r["weights"]}


def empirical_beta(noisy_by_t, bank4, sigma, eps=1e-8, quantile=0.0):
    """threshold.py:351-384: beta_j = sum_i exp(-||noisy_j - n_i|| / 2 sigma^2) + eps,
    then torch.quantile over j (linear interpolation), per timestep key."""
    out = {}
    for t, noisy in noisy_by_t.items():
        r = closed_form(noisy, bank4, sigma=sigma, eps=eps)
        out[t] = float(np.quantile(r["denom"], quantile, method="linear"))
    return out


def sparse_repellency(x4, bank4, radius, scale=1.0, normalise_query=False):
    """SPELL baseline, fast.py:306-340 / threshold.py:415-454.

    Neighbours are selected with the FIRST query row broadcast against the
    bank when Q == 1 (``x_0_hat - ref`` broadcasts, fast.py:312-318); for
    Q > 1 the reference's broadcast is ill-formed unless Q == N, so this
    restatement defines the per-row generalisation: row q uses the negatives
    within ``radius`` of row q.  term_q = sum_i relu(radius/d_qi - 1) (x_q - n_i);
    x0' = x0 + scale * term.
    """
    x_orig = _flat64(x4)
    # fast_sdv3.py:332: the force is taken on the channel-normalised query, the update lands on the original x0
    xf = _flat64(channel_normalise(x4)) if normalise_query else x_orig
    bf = _flat64(bank4)
    # ||x - n|| computed directly (no expansion), as the reference does
    d = np.stack([np.sqrt(((xq[None, :] - bf) ** 2).sum(1)) for xq in xf], 0)
    inside = d < radius
    with np.errstate(divide="ignore"):
        w = np.where(inside, np.maximum(radius / d - 1.0, 0.0), 0.0)
    term = w.sum(1)[:, None] * xf - w @ bf
    x0 = x_orig + scale * term
    return {"x_0_hat": x0.reshape(np.shape(x4)), "term": term.reshape(np.shape(x4)),
            "trunc_weight": w, "force_norm": float(np.sqrt((term ** 2).sum())),
            "is_negation": bool(w.sum() != 0.0)}


# ---------------------------------------------------------------------------
# float32 port with the reference's memory behaviour (the timed CPU baseline)
# ---------------------------------------------------------------------------

def materialised_port(x4, bank4, sigma=1.0, eps=1e-8, normalise_query=False):
    """float32 torch restatement that materialises the [Q,N,D+1] broadcast the
    way fast.py:249-250 does, so that timing it measures the reference's CPU
    cost (same ops, same temporaries).  Returns (neg [Q,C,H,W], mean_item,
    denominator [Q,1], nominator [Q,D])."""
    if normalise_query:
        x4 = x4 / torch.norm(x4, dim=1, keepdim=True)
    q = x4.shape[0]
    n, c, h, w = bank4.shape
    d = c * h * w
    bank2 = bank4.reshape(n, d)
    dist = torch.cdist(x4.reshape(1, q, d), bank2.reshape(1, n, d))[0]
    logits = -(dist.reshape(q, n, 1).repeat(1, 1, d + 1)) / (2.0 * sigma ** 2)
    aug = torch.cat((bank2, torch.ones(n, 1, dtype=bank2.dtype, device=bank2.device)), dim=1)
    summed = (logits.exp() * aug.reshape(1, n, d + 1)).sum(dim=1)
    denominator = summed[:, -1].reshape(-1, 1) + eps
    nominator = summed[:, :-1]
    neg = nominator / denominator
    item = neg.clamp(min=-1e10, max=1e10).mean().item()
    return neg.reshape(-1, c, h, w), item, denominator, nominator


def conditioning_port(x4, bank4, scale, sigma=1.0, eps=1e-8, normalise_query=False):
    """One full reference call: denoiser + in-place correction (fast.py:129-132)."""
    neg, item, denominator, _ = materialised_port(x4, bank4, sigma, eps, normalise_query)
    x4 -= scale * neg
    return x4, item, denominator


# ---------------------------------------------------------------------------
# synthetic inputs shared by the tests, the goldens and bench.py
# ---------------------------------------------------------------------------

def synthetic_bank(n, c, h, w, seed=1234):
    g = torch.Generator().manual_seed(seed)
    b = torch.randn(n, c, h, w, generator=g)
    return b / b.norm(dim=1, keepdim=True)


def synthetic_queries(bank4, q, regime, seed=4321):
    """far: randn; x0: 0.5*randn; near: bank[idx] + 0.05*randn."""
    g = torch.Generator().manual_seed(seed)
    shape = (q,) + tuple(bank4.shape[1:])
    if regime == "far":
        return torch.randn(shape, generator=g)
    if regime == "x0":
        return 0.5 * torch.randn(shape, generator=g)
    if regime == "near":
        idx = torch.randint(0, bank4.shape[0], (q,), generator=g)
        return bank4[idx] + 0.05 * torch.randn(shape, generator=g)
    if regime == "mid":  # midpoint of two negatives: two (near-)equal weights
        idx = torch.randint(0, bank4.shape[0], (q, 2), generator=g)
        return 0.5 * (bank4[idx[:, 0]] + bank4[idx[:, 1]])
    raise ValueError(regime)
