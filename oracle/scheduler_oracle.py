"""CPU restatement of the caller-side scheduler arithmetic around the projection.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Parity status: UNPINNED.  The arithmetic lives in a third-party dependency that
is absent from /root/reference and from this image: ``diffusers==0.29.0``
(/root/reference/requirements.txt:3) -- ``DDPMScheduler.step`` /
``.add_noise`` / ``.set_timesteps`` and the flow-matching re-noise the SD3
pipeline writes by hand.  The formulas below restate the published algorithm of
that version; the *structure* (what is called, in which order, on which
tensors) is anchored on the reference's own call sites:

  models/textuals_visual/modified_safree_diffusion_pipeline_threshold_time.py:550-576
      x0 = step(eps, t, x_t).pred_original_sample ; conditioning(x0, beta_threshold=True)
      if is_negation: x_t = add_noise(x0', randn, t) ; x_prev = step(eps, t, x_t).prev_sample
  models/sdv3/safe_denoiser_pipeline.py:1142-1161
      sigma = t/1000 ; x0 = x - sigma v ; x1 = x + (1-sigma) v ; x0' = conditioning(x0)
      noise = sqrt(sigma') x1 + sqrt(1-sigma') z ; x_next = x0' + sigma' (noise - x0')

No reference test pins results at that boundary, hence "unpinned".
"""
from __future__ import annotations

import numpy as np


def sd14_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """scaled-linear betas of the SD-1.4 scheduler config, in float32 like
    diffusers (torch.linspace(sqrt(b0), sqrt(b1), T, float32) ** 2 -> cumprod)."""
    betas = np.linspace(np.float32(beta_start) ** 0.5, np.float32(beta_end) ** 0.5,
                        num_train_timesteps, dtype=np.float32) ** 2
    return np.cumprod((1.0 - betas).astype(np.float32)).astype(np.float64)


def ddpm_timesteps(num_inference_steps=50, num_train_timesteps=1000, steps_offset=1):
    """'leading' spacing: t = round(arange(n) * (T // n))[::-1] + offset -> 981 ... 1."""
    ratio = num_train_timesteps // num_inference_steps
    return (np.arange(num_inference_steps) * ratio).round()[::-1].astype(np.int64) + steps_offset


def ddpm_coefficients(alphas_cumprod, t, num_inference_steps=50, final_alpha_cumprod=None):
    """All scalars one in-window DDPM step needs.  At the last step DDPMScheduler takes alpha_bar_prev = 1,
    DDIMScheduler its final_alpha_cumprod (= alphas_cumprod[0] for the SD-1.4 config, set_alpha_to_one = False):
    the two *_prev entries used by the DDIM step follow the latter."""
    T = len(alphas_cumprod)
    t_prev = t - T // num_inference_steps
    ab_t = float(alphas_cumprod[t])
    ab_prev = float(alphas_cumprod[t_prev]) if t_prev >= 0 else 1.0
    fin = float(alphas_cumprod[0]) if final_alpha_cumprod is None else float(final_alpha_cumprod)
    ab_prev_ddim = float(alphas_cumprod[t_prev]) if t_prev >= 0 else fin
    a_t = ab_t / ab_prev
    b_t = 1.0 - a_t
    var = max((1.0 - ab_prev) / (1.0 - ab_t) * b_t, 1e-20)
    return {
        "sqrt_ab": ab_t ** 0.5, "sqrt_1m_ab": (1.0 - ab_t) ** 0.5,
        "c_x0": (ab_prev ** 0.5) * b_t / (1.0 - ab_t),
        "c_xt": (a_t ** 0.5) * (1.0 - ab_prev) / (1.0 - ab_t),
        "sigma_noise": (var ** 0.5) if t > 0 else 0.0,
        "sqrt_ab_prev": ab_prev_ddim ** 0.5, "sqrt_1m_ab_prev": (1.0 - ab_prev_ddim) ** 0.5,
    }


def eps_to_x0(x_t, eps, co):
    return (np.asarray(x_t, np.float64) - co["sqrt_1m_ab"] * np.asarray(eps, np.float64)) / co["sqrt_ab"]


def ddpm_fused_step(x_t, eps, neg, gate, scale, co, z1, z2, return_neg_as_x0=False):
    """One in-window SD-1.4 step after the projection produced ``neg`` and ``gate``
    (per-row bool).  Mirrors ...threshold_time.py:554-576.  Rows whose gate is
    False keep x_t (no re-noise) but still take the ancestral step."""
    x_t = np.asarray(x_t, np.float64)
    eps = np.asarray(eps, np.float64)
    x0 = eps_to_x0(x_t, eps, co)
    x0c = x0 - scale * np.asarray(neg, np.float64)
    src = np.asarray(neg, np.float64) if return_neg_as_x0 else x0c
    g = np.asarray(gate, bool).reshape((-1,) + (1,) * (x_t.ndim - 1))
    x_t2 = np.where(g, co["sqrt_ab"] * src + co["sqrt_1m_ab"] * np.asarray(z1, np.float64), x_t)
    x0b = eps_to_x0(x_t2, eps, co)
    prev = co["c_x0"] * x0b + co["c_xt"] * x_t2 + co["sigma_noise"] * np.asarray(z2, np.float64)
    return {"x0": x0, "x0_corrected": x0c, "x_t_renoised": x_t2, "prev": prev}


def ddim_fused_step(x_t, eps, neg, gate, scale, co, z1, return_neg_as_x0=False):
    """DDIM (eta = 0) variant named by north_star: x_prev = sqrt(ab_prev) x0'' + sqrt(1-ab_prev) eps."""
    x_t = np.asarray(x_t, np.float64)
    eps = np.asarray(eps, np.float64)
    x0 = eps_to_x0(x_t, eps, co)
    x0c = x0 - scale * np.asarray(neg, np.float64)
    src = np.asarray(neg, np.float64) if return_neg_as_x0 else x0c
    g = np.asarray(gate, bool).reshape((-1,) + (1,) * (x_t.ndim - 1))
    x_t2 = np.where(g, co["sqrt_ab"] * src + co["sqrt_1m_ab"] * np.asarray(z1, np.float64), x_t)
    x0b = eps_to_x0(x_t2, eps, co)
    prev = co["sqrt_ab_prev"] * x0b + co["sqrt_1m_ab_prev"] * eps
    return {"x0": x0, "x0_corrected": x0c, "x_t_renoised": x_t2, "prev": prev}


def flow_fused_step(x, v, neg, scale, sigma, sigma_next, z):
    """SD3 flow-matching step, safe_denoiser_pipeline.py:1142-1161."""
    x = np.asarray(x, np.float64)
    v = np.asarray(v, np.float64)
    x0 = x - sigma * v
    x1 = x + (1.0 - sigma) * v
    x0c = x0 - scale * np.asarray(neg, np.float64)
    noise = (sigma_next ** 0.5) * x1 + ((1.0 - sigma_next) ** 0.5) * np.asarray(z, np.float64)
    nxt = x0c + sigma_next * (noise - x0c)
    return {"x0": x0, "x1": x1, "x0_corrected": x0c, "next": nxt}
