"""Drop-in for /root/reference/repellency/repellency_methods_threshold.py (run_nudity.py:54,
run_coco30k.py:55, run_ann_graham.py:46, run_munch.py:48).

kernel_fast + beta-threshold gate (``is_negation``), the noisy-bank builder and the empirical-beta
calibration.  Generalisation over the reference: Q > 1 is accepted (the reference calls
``denominator.item()``, threshold.py:348, so it is Q == 1 only); the gate is then evaluated per row,
``is_negation`` is true if ANY row passes, and the per-row tensor is returned as
``is_negation_rows``.
"""
from __future__ import annotations

import torch

from ._base import LazyScalar, RepellencyBase, make_registry

__CONDITIONING_METHOD__, register_conditioning_method, get_repellency_method = make_registry()


class LazyFlag:
    """bool(any(gate rows)) evaluated when the caller branches on it (one host sync, as in the
    reference's ``if denominator > beta_threshold``, threshold.py:183-185)."""
    __slots__ = ("_t", "_v")

    def __init__(self, gate_rows: torch.Tensor):
        self._t = gate_rows
        self._v = None

    def __bool__(self):
        if self._v is None:
            self._v = bool(self._t.any().item())
            self._t = None
        return self._v

    def __eq__(self, other):
        return bool(self) == bool(other)

    def __repr__(self):
        return repr(bool(self))

    __hash__ = None


class RepellencyMethod(RepellencyBase):
    float_after_project = False      # threshold.py:54-72 has no .float()

    def _init_extra(self, kwargs):   # threshold.py:35-47
        self.sigma = kwargs.get('sigma', 1.0)
        self.quantile = kwargs.get('quantile', 0.0)
        self.beta_threshold = kwargs.get('beta_threshold', False)
        self.beta_threshold_margin = kwargs.get('beta_threshold_margin', 0.0)
        self.proj_beta_ref_path = kwargs.get('proj_noisy_ref_path_for_beta', None)
        self.cache_proj_beta_ref = kwargs.get('cache_noisy_ref_path_for_beta', False)
        # not a reference kwarg: calibrate beta without materialising the {t: noisy bank} dictionary (50 N D floats,
        # 9.8 GB at N = 3000) and without writing its cache file -- see empirical_beta_streaming
        self.stream_beta_calibration = kwargs.get('stream_beta_calibration', False)

    # ---- noisy bank for the calibration (threshold.py:108-155) ---------------------------------
    def set_noisy_proj_ref(self, scheduler, num_timesteps=None, **kwargs):
        """{int t: add_noise(bank, randn(seed 42), t)} for every inference timestep, saved with
        torch.save in scheduler order (last key = smallest t)."""
        steps = num_timesteps if num_timesteps is not None else 50
        device = kwargs.get("device", "cuda")
        generator = kwargs.get("generator", None)
        if generator is None:
            generator = torch.Generator(device=device).manual_seed(42)
        bank = self.proj_refs
        scheduler.set_timesteps(steps, device=device)
        out = {}
        with torch.no_grad():
            for t in scheduler.timesteps:
                chunks = []
                for lo in range(0, len(bank), self.n_embed):
                    clean = bank[lo:lo + self.n_embed]
                    noise = torch.randn(clean.shape, generator=generator, device=device, dtype=torch.float32)
                    chunks.append(scheduler.add_noise(clean, noise, t))
                out[t.item()] = torch.cat(chunks, 0)
        print("[Proj_Ref] Save the cached proj_beta_ref")
        self.mkdir_cache(self.proj_beta_ref_path)
        torch.save(out, self.proj_beta_ref_path)
        return out

    def get_noisy_proj_refs(self):
        return self.noisy_proj_refs

    # ---- conditioning (threshold.py:171-193) ---------------------------------------------------
    def empirical_denoiser(self, x_t, sigma=1.0, **kwargs):
        raise NotImplementedError

    def conditioning(self, x_0_hat, **kwargs):
        if kwargs.get("beta_threshold", False):
            return self.conditioning_threshold(x_0_hat, **kwargs)
        return self.conditioning_1(x_0_hat, **kwargs)

    def _project(self, x_0_hat, gate_threshold):
        q, copied = self._as_query(x_0_hat)
        neg, s = self.projector().correct(q, self.sigma, self.scale, self.epsilon, want_neg=True,
                                          gate_threshold=gate_threshold)
        if copied:
            if x_0_hat.dtype == torch.float32:
                x_0_hat.copy_(q)
            else:
                x_0_hat = q
        rows = s.gate.clone()
        item = {"negative_score_item": LazyScalar(s.mean),
                "denominator": LazyScalar(s.denom) if s.denom.numel() == 1 else s.denom.clone(),
                "nominator": s.num.clone()}
        return x_0_hat, neg.view_as(x_0_hat), item, rows

    def conditioning_threshold(self, x_0_hat, **kwargs):
        thr = self.beta_threshold - self.beta_threshold_margin
        x_0_hat, _, item, rows = self._project(x_0_hat, float(thr))
        return {"x_0_hat": x_0_hat, "mean_x_0_hat": item, "is_negation": LazyFlag(rows),
                "is_negation_rows": rows}

    def conditioning_1(self, x_0_hat, **kwargs):
        # the query is corrected in place, but what is RETURNED under "x_0_hat" is the negative
        # mean (threshold.py:190-193)
        _, neg, item, _ = self._project(x_0_hat, None)
        return {"x_0_hat": neg, "mean_x_0_hat": item, "is_negation": True}


@register_conditioning_method(name='euclidean')
class EuclideanRepellency(RepellencyMethod):
    def __init__(self, ref_data, embed_fn, forward_fn, max_idx, beta_min, beta_max, **kwargs):
        raise TypeError("euclidean repellency is not constructible through get_repellency_method")


@register_conditioning_method(name='kernel')
class RBFKernelRepellencyLegacy(RepellencyMethod):
    def __init__(self, ref_data, embed_fn, forward_fn, max_idx, beta_min, beta_max, **kwargs):
        raise TypeError("kernel repellency is not constructible through get_repellency_method")


class _Calibrated(RepellencyMethod):
    """Shared by kernel_fast (beta) and sparse (radius): load or build the noisy bank."""

    def _noisy_refs(self, kwargs):
        if self.cache_proj_beta_ref:
            return self.import_proj_ref(self.proj_beta_ref_path)
        scheduler = kwargs.get("scheduler", None)
        assert scheduler != None, "We need scheduler for computing \\beta reference"  # noqa: E711
        return self.set_noisy_proj_ref(scheduler, self.num_timesteps,
                                       **{k: kwargs[k] for k in ("device", "generator") if k in kwargs})


@register_conditioning_method(name='kernel_fast')
class RBFKernelRepellency(_Calibrated):
    def __init__(self, ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs):
        super().__init__(ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs)
        self.scale = kwargs.get('scale', 1.0)
        self.beta_threshold = kwargs.get('beta_threshold', -1.0)
        if self.beta_threshold <= 0 and self.stream_beta_calibration and not self.cache_proj_beta_ref:
            scheduler = kwargs.get("scheduler", None)
            assert scheduler != None, "We need scheduler for computing \\beta reference"  # noqa: E711
            betas = self.empirical_beta_streaming(scheduler, sigma=self.sigma, quantitle=self.quantile,
                                                  **{k: kwargs[k] for k in ("device", "generator") if k in kwargs})
            self.beta_threshold = float(betas[list(betas.keys())[-1]])
        if self.beta_threshold <= 0:                      # threshold.py:291-306
            self.noisy_proj_refs = self._noisy_refs(kwargs)
            self.noisy_refs_beta_quantitle = self.empirical_beta(sigma=self.sigma, quantitle=self.quantile)
            last_t = list(self.noisy_refs_beta_quantitle.keys())[-1]
            # a python float once, here: the 0-d CUDA tensor torch.quantile returns would cost one device-to-host sync
            # per sampling step in conditioning_threshold (the calibration prints .item() per level anyway)
            self.beta_threshold = float(self.noisy_refs_beta_quantitle[last_t])
            del self.noisy_proj_refs, self.noisy_refs_beta_quantitle
            torch.cuda.empty_cache()

    def empirical_denoiser(self, x_t, sigma=1.0, **kwargs):
        """-> (negative mean, {"negative_score_item", "denominator", "nominator"}) (threshold.py:309-349)."""
        q, _ = self._as_query(x_t)
        neg, s = self.projector().correct(q, sigma, 0.0, self.epsilon, want_neg=True, apply=False)
        item = {"negative_score_item": LazyScalar(s.mean),
                "denominator": LazyScalar(s.denom) if s.denom.numel() == 1 else s.denom.clone(),
                "nominator": s.num.clone()}
        return neg.view((-1,) + tuple(self.proj_refs.shape[1:])), item

    def empirical_beta(self, sigma=1.0, quantitle=0.25, rows_per_call=128, **kwargs):
        """beta_j = sum_i exp(-||noisy_j - n_i|| / 2 sigma^2) + eps per noise level, then the
        ``quantitle`` quantile over j (threshold.py:351-384).  The noisy rows are the QUERIES of the
        same projection kernel (only z is needed)."""
        print("*" * 10, "Set Beta Thresholds", "*" * 10)
        proj = self.projector()
        out = {}
        for t, latents in self.get_noisy_proj_refs().items():
            print(f"[Empirical Betas] Computing empirical beta for {t}-th step")
            lat, _ = self._as_query(latents.to(proj.bank.device))
            betas = []
            for lo in range(0, lat.shape[0], rows_per_call):
                s = proj.partial_sums(lat[lo:lo + rows_per_call], sigma, z_only=True)
                betas.append(s.z.clone())
            beta = torch.cat(betas) + self.epsilon
            q1 = torch.quantile(beta, quantitle)
            print(f"Top {100*(1-quantitle):.1f} % of radius at t={t}: {q1.item():.3f}")
            out[t] = q1
        return out


    def empirical_beta_streaming(self, scheduler, sigma=1.0, quantitle=0.25, num_timesteps=None, **kwargs):
        """set_noisy_proj_ref + empirical_beta (threshold.py:108-155, :351-384) in one sweep that never holds more than
        one chunk of noisy rows: per timestep, per chunk of n_embed bank rows -- the SAME order in which the reference
        draws from its generator, so every noise value, every beta_j and every quantile is the one the two-step path
        produces -- add_noise, then the z-only projection of the chunk (distance kernel + weights, no phase B), and the
        chunk is dropped.  Memory: n_embed x D instead of 50 x N x D; no cache file is written."""
        steps = num_timesteps if num_timesteps is not None else (self.num_timesteps or 50)
        device = kwargs.get("device", "cuda")
        generator = kwargs.get("generator", None)
        if generator is None:
            generator = torch.Generator(device=device).manual_seed(42)
        bank = self.proj_refs
        proj = self.projector()
        scheduler.set_timesteps(steps, device=device)
        print("*" * 10, "Set Beta Thresholds", "*" * 10)
        out = {}
        with torch.no_grad():
            for t in scheduler.timesteps:
                print(f"[Empirical Betas] Computing empirical beta for {t.item()}-th step")
                betas = []
                for lo in range(0, len(bank), self.n_embed):
                    clean = bank[lo:lo + self.n_embed]
                    noise = torch.randn(clean.shape, generator=generator, device=device, dtype=torch.float32)
                    lat, _ = self._as_query(scheduler.add_noise(clean, noise, t).to(proj.bank.device))
                    for r0 in range(0, lat.shape[0], 128):
                        betas.append(proj.partial_sums(lat[r0:r0 + 128], sigma, z_only=True).z.clone())
                beta = torch.cat(betas) + self.epsilon
                q1 = torch.quantile(beta, quantitle)
                print(f"Top {100*(1-quantitle):.1f} % of radius at t={t.item()}: {q1.item():.3f}")
                out[t.item()] = q1
        return out


@register_conditioning_method(name='sparse')
class SparseRepellency(_Calibrated):
    """SPELL baseline with is_negation (threshold.py:386-490)."""

    def __init__(self, ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs):
        super().__init__(ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max, **kwargs)
        self.radius = kwargs.get('radius', -1.0)
        self.scale = kwargs.get('scale', 1.0)
        if self.radius <= 0:
            self.noisy_proj_refs = self._noisy_refs(kwargs)
            self.noisy_refs_beta_quantitle = self.empirical_radius(quantitle=self.quantile)
            last_t = list(self.noisy_refs_beta_quantitle.keys())[-1]
            self.radius = self.noisy_refs_beta_quantitle[last_t]
            del self.noisy_proj_refs, self.noisy_refs_beta_quantitle
            torch.cuda.empty_cache()

    def repellency_force(self, x_0_hat, **kwargs):
        q, _ = self._as_query(x_0_hat)
        term, wsum = self.projector().sparse(q.clone(), float(self.radius), 0.0, want_term=True)
        return term.view_as(x_0_hat), {"repellency_force": LazyScalar(term.norm(p=2)), "trunc_weight": wsum}

    def empirical_denoiser(self, x_0_hat, **kwargs):
        return self.repellency_force(x_0_hat, **kwargs)

    def conditioning_1(self, x_0_hat, **kwargs):
        q, copied = self._as_query(x_0_hat)
        term, wsum = self.projector().sparse(q, float(self.radius), self.scale, want_term=True)
        if copied:
            x_0_hat.copy_(q)
        return {"x_0_hat": x_0_hat, "mean_x_0_hat": LazyScalar(term.norm(p=2)),
                "is_negation": LazyFlag(wsum != 0.0)}

    def conditioning_threshold(self, x_0_hat, **kwargs):
        return self.conditioning_1(x_0_hat, **kwargs)

    def empirical_radius(self, quantitle=0.25, **kwargs):
        """quantile of all pairwise ||noisy_j - n_i|| per noise level (threshold.py:461-490)."""
        print("*" * 10, "Set Radius Thresholds", "*" * 10)
        bank = self.get_proj_ref()
        flat = bank.reshape(bank.shape[0], -1)
        out = {}
        with torch.no_grad():
            for t, latents in self.get_noisy_proj_refs().items():
                print(f"[Empirical Betas] Computing empirical radius for {t}-th step")
                d = torch.cdist(latents.reshape(latents.shape[0], -1).to(flat.device), flat,
                                compute_mode="donot_use_mm_for_euclid_dist")
                q1 = torch.quantile(d.reshape(-1), quantitle)
                print(f"Top {100*(1-quantitle):.1f} % of beta at t={t}: {q1.item():.3f}")
                out[t] = q1
        return out
