"""Scalar coefficients of the caller-side scheduler arithmetic that the fused epilogue kernels take.

The arithmetic itself is third-party (diffusers==0.29.0 DDPMScheduler, /root/reference/
requirements.txt:3); the call structure it serves is the reference's in-window step
(models/textuals_visual/modified_safree_diffusion_pipeline_threshold_time.py:550-576).
Everything here is host-side float maths on a handful of scalars.
"""
from __future__ import annotations


def alphas_cumprod_of(scheduler):
    """alphas_cumprod of a diffusers-style scheduler as a list of python floats."""
    ac = scheduler.alphas_cumprod
    return [float(v) for v in (ac.tolist() if hasattr(ac, "tolist") else ac)]


def sd14_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """scaled-linear schedule of the SD-1.4 scheduler config (float32 like diffusers)."""
    import torch
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).double().tolist()


def ddpm_coefficients(alphas_cumprod, t: int, num_inference_steps: int = 50, final_alpha_cumprod=None) -> dict:
    """Scalars of one DDPM (ancestral, fixed_small) / DDIM (eta 0) step at timestep t.

    At the last step (t_prev < 0) DDPMScheduler uses alpha_bar_prev = 1; DDIMScheduler uses its
    ``final_alpha_cumprod`` -- alphas_cumprod[0] with the SD-1.4 config (set_alpha_to_one = False).  The DDIM
    coefficients (sqrt_ab_prev, sqrt_1m_ab_prev) follow that; pass ``final_alpha_cumprod=1.0`` for set_alpha_to_one."""
    T = len(alphas_cumprod)
    t = int(t)
    t_prev = t - T // int(num_inference_steps)
    ab_t = float(alphas_cumprod[t])
    ab_prev = float(alphas_cumprod[t_prev]) if t_prev >= 0 else 1.0
    if final_alpha_cumprod is None:
        final_alpha_cumprod = float(alphas_cumprod[0])
    ab_prev_ddim = float(alphas_cumprod[t_prev]) if t_prev >= 0 else float(final_alpha_cumprod)
    a_t = ab_t / ab_prev
    b_t = 1.0 - a_t
    var = max((1.0 - ab_prev) / (1.0 - ab_t) * b_t, 1e-20)
    return {
        "sqrt_ab": ab_t ** 0.5,
        "sqrt_1m_ab": (1.0 - ab_t) ** 0.5,
        "c_x0": (ab_prev ** 0.5) * b_t / (1.0 - ab_t),
        "c_xt": (a_t ** 0.5) * (1.0 - ab_prev) / (1.0 - ab_t),
        "sigma_noise": (var ** 0.5) if t > 0 else 0.0,
        "sqrt_ab_prev": ab_prev_ddim ** 0.5,
        "sqrt_1m_ab_prev": (1.0 - ab_prev_ddim) ** 0.5,
    }


def ddpm_timesteps(num_inference_steps=50, num_train_timesteps=1000, steps_offset=1):
    """'leading' spacing with offset: 981, 961, ..., 1 for the SD-1.4 defaults."""
    ratio = num_train_timesteps // num_inference_steps
    return [int(round(i * ratio)) + steps_offset for i in range(num_inference_steps - 1, -1, -1)]
