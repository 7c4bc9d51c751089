"""Method classes shared by repellency_methods_fast and repellency_methods_fast_sdv3.

Reference: /root/reference/repellency/repellency_methods_fast.py:120-340 (and the sdv3 twin, whose
only difference is the query normalisation at fast_sdv3.py:152,:194,:239,:332).
"""
from __future__ import annotations

import torch

from ._base import LazyScalar, RepellencyBase


class FastRepellencyMethod(RepellencyBase):
    """conditioning / conditioning_1 / conditioning_2 of fast.py:120-137."""

    def sigma(self, cont_time, **kwargs):     # fast.py:82-83: a method, which is why YAML sigma is ignored (Q4)
        pass

    def empirical_denoiser(self, x_t, sigma=1.0, **kwargs):
        raise NotImplementedError

    def conditioning(self, x_0_hat, **kwargs):
        if x_0_hat.dtype != self.ref_data.dtype:
            x_0_hat = x_0_hat.to(self.ref_data.dtype)        # a copy: the caller's tensor is untouched (Q8)
        gs = kwargs.get("guidance_scale", None)
        if gs is not None and gs > 0.0:
            return self.conditioning_2(x_0_hat, **kwargs)
        return self.conditioning_1(x_0_hat, **kwargs)

    def _corrected(self, x_0_hat, scale, want_neg, **kwargs):
        """Shared body: one fused projection + in-place correction on the caller's tensor."""
        q, copied = self._as_query(x_0_hat)
        neg, s = self.projector().correct(q, kwargs.get("sigma", 1.0), scale, self.epsilon,
                                          normalize_channels=self._channels(), want_neg=want_neg)
        if copied:
            if x_0_hat.dtype == torch.float32:
                x_0_hat.copy_(q)                             # keep the in-place contract for strided input
            else:
                x_0_hat = q.to(x_0_hat.dtype)                # the reference hands back ref_data.dtype (fast.py:121-122)
        return x_0_hat, neg, s

    def conditioning_1(self, x_0_hat, **kwargs):
        x_0_hat, _, s = self._corrected(x_0_hat, self.scale, False, **kwargs)
        return {"x_0_hat": x_0_hat, "mean_x_0_hat": LazyScalar(s.mean)}

    def conditioning_2(self, x_0_hat, **kwargs):
        x_0_hat, neg, s = self._corrected(x_0_hat, 1.0, True, **kwargs)
        return {"x_0_hat": neg.view_as(x_0_hat), "mean_x_0_hat": LazyScalar(s.mean)}


class KernelFast(FastRepellencyMethod):
    """'kernel_fast' (fast.py:217-262): RBF-style empirical denoiser on the un-squared distance."""

    def __init__(self, ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs):
        super().__init__(ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs)
        self.scale = kwargs.get('scale', 1.0)

    def empirical_denoiser(self, x_t, sigma=1.0, **kwargs):
        """-> (negative mean [Q,C,H,W], its clamped mean).  x_t is not modified."""
        q, _ = self._as_query(x_t)
        neg, s = self.projector().correct(q, sigma, 0.0, self.epsilon, normalize_channels=self._channels(),
                                          want_neg=True, apply=False)
        return neg.view((-1,) + tuple(self.proj_refs.shape[1:])), LazyScalar(s.mean)


class RandomNoise(FastRepellencyMethod):
    """'random_noise' control (fast.py:264-297): the 'negative mean' is one Gaussian draw."""

    def __init__(self, ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs):
        super().__init__(ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs)
        self.scale = kwargs.get('scale', 1.0)

    def empirical_denoiser(self, x_t, sigma=1.0, **kwargs):
        bank = self.get_proj_ref()
        draw = torch.randn(size=(1, bank[0].numel())).to(bank.device)
        item = draw.clamp(min=-1e10, max=1e10).mean().item()
        return draw.reshape((-1,) + tuple(bank.shape[1:])), item

    def conditioning_1(self, x_0_hat, **kwargs):
        neg, item = self.empirical_denoiser(x_t=x_0_hat, **kwargs)
        x_0_hat -= self.scale * neg
        return {"x_0_hat": x_0_hat, "mean_x_0_hat": item}


class Sparse(FastRepellencyMethod):
    """'sparse' SPELL baseline (fast.py:299-340)."""

    def __init__(self, ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs):
        super().__init__(ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs)
        self.radius = kwargs.get('radius', 1.0)
        self.scale = kwargs.get('scale', 1.0)

    def repellency_force(self, x_0_hat, **kwargs):
        """-> (sum_i relu(radius/d_i - 1)(x - n_i), its L2 norm); x_0_hat is not modified."""
        q, _ = self._as_query(x_0_hat)
        term, _ = self.projector().sparse(q.clone(), self.radius, 0.0, want_term=True,
                                          normalize_channels=self._channels())
        return term.view_as(x_0_hat), LazyScalar(term.norm(p=2))

    def empirical_denoiser(self, x_0_hat, **kwargs):
        return self.repellency_force(x_0_hat, **kwargs)

    def conditioning_1(self, x_0_hat, **kwargs):
        q, copied = self._as_query(x_0_hat)
        term, _ = self.projector().sparse(q, self.radius, self.scale, want_term=True,
                                          normalize_channels=self._channels())
        if copied:
            x_0_hat.copy_(q)
        return {"x_0_hat": x_0_hat, "mean_x_0_hat": LazyScalar(term.norm(p=2))}


def _unconstructable(label):
    """'euclidean', 'kernel', 'lsh' are registered by the reference but take six positionals
    (fast.py:142,:181,:344) while the factory passes seven: construction raises TypeError there,
    and here."""
    class Dead(FastRepellencyMethod):
        def __init__(self, ref_data, embed_fn, forward_fn, max_idx, beta_min, beta_max, **kwargs):
            raise TypeError(f"{label} repellency is not constructible through get_repellency_method")
    Dead.__name__ = f"{label.capitalize()}Repellency"
    return Dead
