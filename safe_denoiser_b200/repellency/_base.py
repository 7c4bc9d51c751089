"""Shared machinery of the three drop-in modules.

Mirrors the interface of /root/reference/repellency/repellency_methods_{fast,fast_sdv3,threshold}.py
(registry :9-22, RepellencyMethod :24-137) with the torch op chain replaced by the CUDA projection.
Each public module builds its own registry with ``make_registry()`` so that, as in the reference,
the three modules do not share registered names.
"""
from __future__ import annotations

import os

import torch

from .. import _native as nv
from ..projection import NegativeBank, Projector


def make_registry():
    """(registry dict, register_conditioning_method, get_repellency_method) -- fast.py:9-22."""
    registry = {}

    def register_conditioning_method(name: str):
        def wrapper(cls):
            if registry.get(name, None):
                raise NameError(f"Name {name} is already registered!")
            registry[name] = cls
            return cls
        return wrapper

    def get_repellency_method(name: str, ref_data, embed_fn, forward_fn, num_timesteps, max_idx,
                              beta_min, beta_max, **kwargs):
        if registry.get(name, None) is None:
            raise NameError(f"Name {name} is not defined!")
        return registry[name](ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min,
                              beta_max, **kwargs)

    return registry, register_conditioning_method, get_repellency_method


class LazyScalar:
    """A device scalar that turns into a python float only when somebody looks at it.

    The reference returns ``tensor.item()`` from every call (fast.py:258, threshold.py:344,:348),
    i.e. a host sync per step.  The drop-in keeps the dict keys but defers the sync.
    """
    __slots__ = ("_t", "_v")

    def __init__(self, t: torch.Tensor):
        self._t = t.clone()       # scratch buffers are reused by the next call
        self._v = None

    def item(self) -> float:
        if self._v is None:
            self._v = float(self._t.item())
            self._t = None
        return self._v

    __float__ = item

    def __repr__(self):
        return repr(self.item())

    def __format__(self, spec):
        return format(self.item(), spec)

    def __bool__(self):
        return bool(self.item())

    def _cmp(self, other, op):
        return op(self.item(), float(other))

    def __lt__(self, o): return self._cmp(o, float.__lt__)
    def __le__(self, o): return self._cmp(o, float.__le__)
    def __gt__(self, o): return self._cmp(o, float.__gt__)
    def __ge__(self, o): return self._cmp(o, float.__ge__)
    def __eq__(self, o): return self._cmp(o, float.__eq__)
    def __add__(self, o): return self.item() + o
    def __radd__(self, o): return o + self.item()
    def __sub__(self, o): return self.item() - o
    def __rsub__(self, o): return o - self.item()
    def __mul__(self, o): return self.item() * o
    def __rmul__(self, o): return o * self.item()
    def __hash__(self): return hash(self.item())


class RepellencyBase:
    """Attributes and cache I/O of RepellencyMethod (fast.py:24-114 / threshold.py:24-166)."""

    # set by subclasses / modules
    normalize_query = False      # fast_sdv3.py:239
    float_after_project = True   # fast.py:58-59 casts, threshold.py:54-72 does not

    def __init__(self, ref_data, embed_fn, forward_fn, num_timesteps, max_idx, beta_min, beta_max,
                 n_embed, **kwargs):
        self.ref_data = ref_data
        self.embed_fn = embed_fn
        self.forward_fn = forward_fn
        self.num_timesteps = num_timesteps
        self.max_idx = max_idx
        self.beta_min = beta_min
        self.beta_max = beta_max
        self.n_embed = n_embed
        self.scale = kwargs.get('scale', 1.0)
        self.epsilon = kwargs.get('epsilon', 1e-8)
        self.proj_ref_path = kwargs.get('proj_ref_path', None)
        self.cache_proj_ref = kwargs.get('cache_proj_ref', False)
        # extensions (ignored by the reference, which swallows unknown kwargs)
        self.kernel_path = kwargs.get('kernel_path', nv.PATH_AUTO)
        self.process_group = kwargs.get('process_group', None)
        self._projector = None
        self._init_extra(kwargs)
        if self.cache_proj_ref:
            self.proj_refs = self.import_proj_ref(self.proj_ref_path)
        else:
            self.proj_refs = self.set_proj_ref()

    def _init_extra(self, kwargs):
        pass

    # ---- bank construction and cache (format: torch.save of one fp32 CPU tensor [N,C,H,W]) ----
    @torch.no_grad()
    def project(self, data, **kwargs):
        """VAE-encode in chunks of n_embed, then divide by the per-pixel channel norm (fast.py:45-70)."""
        if len(data) > self.n_embed:
            parts = [self.embed_fn(data[i:i + self.n_embed]) for i in range(0, len(data), self.n_embed)]
            latents = torch.cat(parts, 0)
        else:
            latents = self.embed_fn(data)
        if latents.is_cuda and latents.dim() == 4 and latents.dtype == torch.float32:
            return NegativeBank.from_latents(latents).tensor          # fused normalise (+ ||n||^2) kernel
        latents = latents / torch.norm(latents, dim=1, keepdim=True)
        return latents.float() if self.float_after_project else latents

    def discrete_to_continous_time(self, idx, **kwargs):
        return 0.001 if idx == 0 else idx / self.max_idx

    def sigma_edm(self, cont_time, **kwargs):
        return torch.sqrt(torch.exp(0.5 * self.beta_max * cont_time ** 2 + self.beta_min * cont_time) - 1.)

    def mkdir_cache(self, path=None):
        target = self.proj_ref_path if path is None else path
        os.makedirs(os.path.split(target)[0], exist_ok=True)

    def set_proj_ref(self):
        with torch.no_grad():
            bank_cpu = self.project(self.ref_data).cpu()
        print("[Proj_Ref] Save the cached proj_ref")
        self.mkdir_cache(self.proj_ref_path)
        torch.save(bank_cpu, self.proj_ref_path)
        return bank_cpu.to('cuda')

    def import_proj_ref(self, proj_ref_path):
        return torch.load(proj_ref_path, map_location=self.ref_data.device)

    def get_proj_ref(self):
        return self.proj_refs

    # ---- device state -------------------------------------------------------------------------
    def projector(self) -> Projector:
        """The prepared bank (built on first use; rebuilt if ``proj_refs`` was replaced).  With a
        ``process_group`` each rank keeps only its contiguous N-shard on the device."""
        src = self.proj_refs
        if self._projector is None or self._projector_src is not src:
            if not src.is_cuda:
                raise RuntimeError(
                    "the repellency projection runs on CUDA only (bank is on %s); load the proj_ref "
                    "cache with ref_data on a cuda device" % src.device)
            if self.process_group is not None:
                import torch.distributed as dist
                from ..projection import shard_bounds
                lo, hi = shard_bounds(src.shape[0], dist.get_rank(self.process_group),
                                      dist.get_world_size(self.process_group))
                rows = src[lo:hi]
            else:
                rows = src
            want_planes = self.kernel_path == nv.PATH_UMMA
            self._projector = Projector(NegativeBank(rows, with_planes=want_planes),
                                        path=self.kernel_path, group=self.process_group)
            self._projector_src = src
        return self._projector

    def _channels(self):
        return int(self.proj_refs.shape[1]) if self.normalize_query else 0

    @staticmethod
    def _as_query(x):
        """fp32 contiguous CUDA view of the query; second value tells whether it is a copy."""
        q = x
        if q.dtype != torch.float32:
            q = q.float()
        if not q.is_contiguous():
            q = q.contiguous()
        return q, (q is not x)
