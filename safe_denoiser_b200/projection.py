"""Host-side driver of the repellency projection: the prepared negative bank, the
per-call scratch, and the three steps of one projection (query prepare ->
partial sums over the bank -> fused epilogue), all through the C ABI.

Reference behaviour replaced: RBFKernelRepellency.empirical_denoiser and
RepellencyMethod.conditioning_1 (/root/reference/repellency/
repellency_methods_fast.py:223-262, :129-132) and their sdv3 / threshold
siblings.  The UNet / scheduler objects stay the caller's.
"""
from __future__ import annotations

import ctypes
import math

import torch

from . import _native as nv


class NegativeBank:
    """The device-resident ``proj_ref`` tensor plus its derived data.

    ``proj_refs`` is the [N,C,H,W] fp32 tensor of the reference's cache file
    (fast.py:99-111).  Derived once here instead of on every call: ||n_i||^2
    (torch.cdist recomputes it, fast.py:249) and, for the tcgen05 path, the
    bf16 hi/lo planes.
    """

    def __init__(self, proj_refs: torch.Tensor, with_planes: bool = False):
        if proj_refs.dim() < 2:
            raise ValueError("bank must be [N, ...]")
        if not proj_refs.is_cuda:
            raise RuntimeError("NegativeBank needs a CUDA tensor; there is no CPU fallback")
        t = proj_refs
        if t.dtype != torch.float32:
            t = t.float()
        self.tensor = t.contiguous()
        self.N = int(self.tensor.shape[0])
        self.D = int(self.tensor[0].numel())
        self.item_shape = tuple(self.tensor.shape[1:])
        self.flat = self.tensor.view(self.N, self.D)
        self.sqnorm = torch.empty(self.N, dtype=torch.float32, device=t.device)
        self.planes = (torch.empty(2, self.N, self.D, dtype=torch.bfloat16, device=t.device)
                       if with_planes else None)
        nv.check(nv.lib().sdn_bank_prepare(nv.ptr(self.flat), self.N, self.D, nv.ptr(self.sqnorm),
                                           nv.ptr(self.planes), nv.current_stream()))

    @classmethod
    def from_latents(cls, latents: torch.Tensor, with_planes: bool = False) -> "NegativeBank":
        """Build the bank from raw VAE latents [N,C,H,W]: per-pixel channel normalisation (``project``,
        fast.py:55-56), ||n_i||^2 and the planes in one fused pass.  ``.tensor`` is what the proj_ref cache stores."""
        if not latents.is_cuda or latents.dim() != 4:
            raise RuntimeError("from_latents needs a CUDA tensor [N,C,H,W]")
        lat = latents.float().contiguous()
        n, c, h, w = lat.shape
        self = cls.__new__(cls)
        self.tensor = torch.empty_like(lat)
        self.N, self.D, self.item_shape = n, c * h * w, (c, h, w)
        self.flat = self.tensor.view(n, self.D)
        self.sqnorm = torch.empty(n, dtype=torch.float32, device=lat.device)
        self.planes = torch.empty(2, n, self.D, dtype=torch.bfloat16, device=lat.device) if with_planes else None
        nv.check(nv.lib().sdn_bank_build(nv.ptr(lat), n, c, h * w, nv.ptr(self.flat), nv.ptr(self.sqnorm),
                                         nv.ptr(self.planes), nv.current_stream()))
        return self

    @property
    def device(self):
        return self.tensor.device

    def ensure_planes(self):
        """bf16 hi/lo planes for the tcgen05 path, made on first batched use."""
        if self.planes is None:
            self.planes = torch.empty(2, self.N, self.D, dtype=torch.bfloat16, device=self.device)
            nv.check(nv.lib().sdn_bank_prepare(nv.ptr(self.flat), self.N, self.D, nv.ptr(self.sqnorm),
                                               nv.ptr(self.planes), nv.current_stream()))
        return self.planes

    def shard(self, rank: int, world: int) -> "NegativeBank":
        """Contiguous N-shard [lo, hi) of this bank."""
        lo, hi = shard_bounds(self.N, rank, world)
        return NegativeBank(self.tensor[lo:hi], with_planes=self.planes is not None)


def shard_bounds(n: int, rank: int, world: int):
    """Rows [lo, hi) of rank ``rank`` when n rows are split as evenly as possible."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_partials(packed: torch.Tensor, group) -> torch.Tensor:
    """The one exchange step of the N-sharded projection: every rank holds the
    un-normalised sums over its bank shard packed as [Q*D floats of num | Q floats of z]; one
    all-reduce(SUM) leaves the full-bank sums on every rank, which then all run the same epilogue
    (latents stay replicated, no broadcast).  ``group`` None means single GPU: nothing to do."""
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


class ShardLink:
    """Peer-mapped buffers of one N-sharded projection shape (torch symmetric memory = CUDA VMM/IPC handles
    exchanged over the process group).  ``packed`` receives this rank's partial sums, ``out`` the merged,
    corrected query; ``sig`` holds the epoch flags of sdn_shard_merge_correct."""

    def __init__(self, Q: int, D: int, group, device):
        import ctypes as C
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.packed = symm.empty(Q * D + Q, dtype=torch.float32, device=device)
        self.out = symm.empty(Q * D, dtype=torch.float32, device=device)
        self.sig = symm.empty(64, dtype=torch.int32, device=device)
        self.sig.zero_()
        handles = [symm.rendezvous(t, group) for t in (self.packed, self.out, self.sig)]
        arr = C.c_void_p * self.world
        self.p_packed, self.p_out, self.p_sig = (arr(*[int(p) for p in h.buffer_ptrs]) for h in handles)
        self._handles = handles
        self.counter = torch.zeros(4, dtype=torch.int32, device=device)
        self.epoch = torch.ones(4, dtype=torch.int32, device=device)    # advanced by the kernel itself
        torch.cuda.synchronize(device)
        dist.barrier(group)          # every rank's flag words are zero before the first merge


def _nvtx(name):
    """NVTX range around an entry point when SDN_NVTX=1 (nsys / ncu --nvtx timelines)."""
    import contextlib
    import os
    if os.environ.get("SDN_NVTX", "0") != "1":
        return contextlib.nullcontext()
    return torch.cuda.nvtx.range(name)


class _Scratch:
    __slots__ = ("num", "z", "xsq", "xq", "denom", "gate", "mean", "ws", "ws_bytes", "packed", "link")


class Projector:
    """One bank + cached scratch; every method enqueues on the current stream and never syncs."""

    def __init__(self, bank: NegativeBank, path: int = nv.PATH_AUTO, group=None):
        self.bank = bank
        self.path = path
        self.group = group          # torch.distributed group for N-sharded banks (None = single GPU)
        self.fused_merge = group is not None    # merge + correction in one kernel over NVLink peer memory
        self.fused_merge_error = None
        self.compute_mean = True                # the reference's logging scalar (extra pass when sharded)
        self._scratch = {}
        self._graphs = {}
        self._few_launch = {}
        self._needs_xsq = {}

    # -- scratch -----------------------------------------------------------------------------
    def _get(self, Q: int, need_xq: bool) -> _Scratch:
        s = self._scratch.get(Q)
        if s is None:
            dev, D = self.bank.device, self.bank.D
            s = _Scratch()
            s.link = None
            if self.group is not None and self.fused_merge:
                import torch.distributed as dist
                try:
                    s.link = ShardLink(Q, D, self.group, dev)
                except Exception as e:           # no peer mapping on this box: NCCL all-reduce instead
                    self.fused_merge_error = repr(e)
                # every rank must take the same exchange (a rank in the peer kernel and a rank in all_reduce
                # would wait for each other forever): agree on the outcome of the rendezvous
                ok = torch.tensor([1 if s.link is not None else 0], dtype=torch.int32, device=dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
                if int(ok.item()) == 0:
                    s.link = None
                    self.fused_merge = False
                    self.fused_merge_error = self.fused_merge_error or "a peer rank could not map the buffers"
            # num and z live in one packed [Q, D+1]-sized buffer so the N-shard merge is ONE exchange
            s.packed = s.link.packed if s.link is not None else torch.empty(Q * D + Q, dtype=torch.float32, device=dev)
            s.num = s.packed[: Q * D].view(Q, D)
            s.z = s.packed[Q * D:]
            s.xsq = torch.empty(Q, dtype=torch.float32, device=dev)
            s.denom = torch.empty(Q, dtype=torch.float32, device=dev)
            s.gate = torch.empty(Q, dtype=torch.int32, device=dev)
            s.mean = torch.zeros(1, dtype=torch.float32, device=dev)
            s.xq = None
            s.ws_bytes = int(nv.lib().sdn_repel_workspace_bytes(Q, self.bank.N, D, self.path))
            s.ws = torch.empty(max(s.ws_bytes, 16), dtype=torch.uint8, device=dev)
            self._scratch[Q] = s
        if need_xq and s.xq is None:
            s.xq = torch.empty(Q, self.bank.D, dtype=torch.float32, device=self.bank.device)
        return s

    def _batched(self, Q: int) -> bool:
        """Same rule as the library's AUTO dispatch: tensor-core path for Q > 8, and from Q = 5 on large banks."""
        return Q > 8 or (Q > 4 and self.bank.N >= 1024)

    def _flat_query(self, x: torch.Tensor):
        if x.dtype != torch.float32 or not x.is_cuda:
            raise RuntimeError("query must be a CUDA fp32 tensor")
        if not x.is_contiguous():
            raise RuntimeError("query must be contiguous")
        Q = int(x.shape[0])
        if x[0].numel() != self.bank.D:
            raise RuntimeError(
                f"query rows have {x[0].numel()} elements but the bank rows have {self.bank.D}")
        return Q, x.view(Q, self.bank.D)

    # -- step 1+2: partial sums ----------------------------------------------------------------
    def partial_sums(self, x0: torch.Tensor, sigma: float, *, normalize_channels: int = 0,
                     model_out: torch.Tensor | None = None, c_x: float = 1.0, c_m: float = 0.0,
                     x0_out: torch.Tensor | None = None, dist_power: int = 1, bank_alpha: float = 1.0,
                     k_out: torch.Tensor | None = None, z_only: bool = False, merge: bool = True) -> _Scratch:
        """num[Q,D] = sum_i k_qi n_i and z[Q] = sum_i k_qi over this (shard of the) bank, merged over
        the process group when the bank is N-sharded.  ``x0`` contiguous fp32 [Q,...]."""
        L = nv.lib()
        st = nv.current_stream()
        Q, xf = self._flat_query(x0)
        if ((self._batched(Q) and self.path == nv.PATH_AUTO)
                or self.path in (nv.PATH_UMMA, nv.PATH_UMMA_BF16, nv.PATH_FLASH)) \
                and self.bank.D % 128 == 0:
            self.bank.ensure_planes()        # batched calls go to the tcgen05 kernels
        s = self._get(Q, normalize_channels > 0)
        mo = None
        if model_out is not None:
            mo = self._flat_query(model_out)[1]
        xo = None if x0_out is None else self._flat_query(x0_out)[1]
        plain = mo is None and normalize_channels == 0 and xo is None and c_x == 1.0
        key = (Q, z_only, self.bank.planes is not None)
        if plain and key not in self._needs_xsq:
            chosen = L.sdn_repel_path(Q, self.bank.N, self.bank.D, 1 if self.bank.planes is not None else 0, self.path)
            # the generic kernels (also used for z-only calls the one-pass kernel would otherwise take) need ||x||^2
            self._needs_xsq[key] = chosen == nv.PATH_GENERIC or (z_only and chosen == nv.PATH_STREAM)
        have_xsq = not (plain and not self._needs_xsq.get(key, True))
        if have_xsq:
            nv.check(L.sdn_query_prepare(nv.ptr(xf), nv.ptr(mo), c_x, c_m, Q, self.bank.D, int(normalize_channels),
                                         nv.ptr(xo), nv.ptr(s.xq) if normalize_channels > 0 else None,
                                         nv.ptr(s.xsq), st))
        if normalize_channels > 0:
            query = s.xq
        elif xo is not None:
            query = xo
        elif not plain:
            # a combined query (c_x * x + c_m * model_out) with no caller buffer for it: it must exist somewhere,
            # the distances are taken on it, not on the raw input
            if s.xq is None:
                s.xq = torch.empty(Q, self.bank.D, dtype=torch.float32, device=self.bank.device)
            nv.check(L.sdn_query_prepare(nv.ptr(xf), nv.ptr(mo), c_x, c_m, Q, self.bank.D, 0, nv.ptr(s.xq), None,
                                         nv.ptr(s.xsq), st))
            query = s.xq
        else:
            query = xf
        b = self.bank
        nv.check(L.sdn_repel_partial(nv.ptr(b.flat), nv.ptr(b.sqnorm), nv.ptr(b.planes), b.N, b.D,
                                     nv.ptr(query), nv.ptr(s.xsq) if have_xsq else None, Q,
                                     1.0 / (2.0 * float(sigma) ** 2), int(dist_power), float(bank_alpha),
                                     None if z_only else nv.ptr(s.num), nv.ptr(s.z), nv.ptr(k_out),
                                     nv.ptr(s.ws), s.ws_bytes, self.path, st))
        if merge:
            merge_partials(s.z if z_only else s.packed, self.group)
        return s

    # -- step 3: epilogues ---------------------------------------------------------------------
    def correct(self, x0: torch.Tensor, sigma: float, scale: float, eps: float, *,
                normalize_channels: int = 0, gate_threshold: float | None = None,
                want_neg: bool = False, apply: bool = True, dist_power: int = 1,
                bank_alpha: float = 1.0, k_out: torch.Tensor | None = None, want_num: bool = True):
        """conditioning(): x0 <- x0 - scale * neg in place.  Returns (neg or None, scratch); scratch
        holds denom [Q], gate [Q] (int32), mean [1] and num [Q,D] as device tensors."""
        with _nvtx("sdn.correct"):
            return self._correct(x0, sigma, scale, eps, normalize_channels=normalize_channels,
                                 gate_threshold=gate_threshold, want_neg=want_neg, apply=apply, dist_power=dist_power,
                                 bank_alpha=bank_alpha, k_out=k_out, want_num=want_num)

    def _correct(self, x0, sigma, scale, eps, *, normalize_channels, gate_threshold, want_neg, apply, dist_power,
                 bank_alpha, k_out, want_num):
        Q, xf = self._flat_query(x0)
        if (self.group is None and apply and normalize_channels == 0
                and self.path in (nv.PATH_AUTO, nv.PATH_STREAM, nv.PATH_UMMA, nv.PATH_FLASH)
                and self._few_launch.get(Q, True)):
            # one GPU, plain query: the fused sequences (1 launch one-pass tcgen05, 2 for Q <= 8, 5 two-phase tcgen05)
            b = self.bank
            use_planes = ((self._batched(Q) and self.path == nv.PATH_AUTO)
                          or self.path in (nv.PATH_UMMA, nv.PATH_FLASH)) and b.D % 128 == 0
            if use_planes:
                b.ensure_planes()
            s = self._get(Q, False)
            neg = torch.empty_like(xf) if want_neg else None
            flags = nv.EPI_GATE if gate_threshold is not None else 0
            rc = nv.lib().sdn_conditioning_fused(
                nv.ptr(b.flat), nv.ptr(b.sqnorm), nv.ptr(b.planes) if use_planes else None, b.N, b.D, nv.ptr(xf), Q,
                1.0 / (2.0 * float(sigma) ** 2), int(dist_power), float(bank_alpha), float(eps), float(scale),
                float(gate_threshold if gate_threshold is not None else 0.0), flags,
                nv.ptr(s.num) if want_num else None, nv.ptr(s.z), nv.ptr(neg), nv.ptr(s.denom), nv.ptr(s.gate),
                nv.ptr(s.mean), nv.ptr(k_out), nv.ptr(s.ws), s.ws_bytes, self.path, nv.current_stream())
            if rc == 0:
                return neg, s
            if rc != -6:
                nv.check(rc)
            self._few_launch[Q] = False           # shape not taken by a fused sequence: three-call sequence
        fused = (self.group is not None and self.fused_merge and apply and not want_neg
                 and self._get(Q, normalize_channels > 0).link is not None)
        s = self.partial_sums(x0, sigma, normalize_channels=normalize_channels,
                              dist_power=dist_power, bank_alpha=bank_alpha, k_out=k_out, merge=not fused)
        if fused:
            return None, self._merge_correct(s, xf, Q, scale, eps, gate_threshold)
        neg = torch.empty_like(xf) if want_neg else None
        s.mean.zero_()
        flags = nv.EPI_GATE if gate_threshold is not None else 0
        nv.check(nv.lib().sdn_epilogue_correct(
            nv.ptr(s.num), nv.ptr(s.z), Q, self.bank.D, float(eps), float(scale),
            float(gate_threshold if gate_threshold is not None else 0.0), flags,
            nv.ptr(xf) if apply else None, nv.ptr(neg), nv.ptr(s.denom), nv.ptr(s.gate), nv.ptr(s.mean),
            nv.current_stream()))
        return neg, s

    def _merge_correct(self, s, xf, Q, scale, eps, gate_threshold):
        """N-sharded: reduce-scatter + correction + all-gather in one kernel over peer memory."""
        link = s.link
        flags = nv.EPI_GATE if gate_threshold is not None else 0
        nv.check(nv.lib().sdn_shard_merge_correct(
            link.p_packed, link.p_out, link.p_sig, link.rank, link.world, nv.ptr(link.epoch), Q, self.bank.D,
            float(eps), float(scale), float(gate_threshold if gate_threshold is not None else 0.0), flags,
            nv.ptr(xf), nv.ptr(s.denom), nv.ptr(s.gate), nv.ptr(link.counter), nv.current_stream()))
        merged = link.out.view(Q, self.bank.D)
        if xf.data_ptr() == merged.data_ptr():
            return s          # the query lives in the peer-mapped buffer (query_buffer): peers wrote it in place
        if self.compute_mean:
            s.mean.copy_(((xf - merged) / (scale if scale != 0 else 1.0)).clamp_(-1e10, 1e10).mean().reshape(1))
        xf.copy_(merged)
        return s

    def shard_error(self) -> bool:
        """True when a merge kernel of this projector gave up waiting for a peer (sticky; synchronises)."""
        return any(sc.link is not None and int(sc.link.counter[1].item()) != 0 for sc in self._scratch.values())

    def query_buffer(self, Q: int, shape=None) -> torch.Tensor:
        """N-sharded banks: a [Q, D] query tensor that lives in this rank's peer-mapped result buffer.  A query
        placed there is corrected in place by the merge kernel itself (rank r rewrites D-slice r on every rank), which
        saves the copy back into a private tensor.  Single GPU: a plain tensor."""
        D = self.bank.D
        if self.group is None or not self.fused_merge:
            t = torch.empty(Q, D, dtype=torch.float32, device=self.bank.device)
        else:
            s = self._get(Q, False)
            t = s.link.out.view(Q, D) if s.link is not None else torch.empty(Q, D, dtype=torch.float32,
                                                                            device=self.bank.device)
        return t.view(shape) if shape is not None else t

    def correct_graphed(self, x0: torch.Tensor, sigma: float, scale: float, eps: float, **kw):
        """``correct`` replayed from a CUDA graph: the 6-8 small launches of one projection are launch bound
        for the shapes the samplers use (N = 515..3000), so the whole call is captured once per
        (query buffer, arguments) and replayed with one ``cudaGraphLaunch``.  The first call with a given key
        runs eagerly (it also performs the one-time kernel attribute setup), the second captures, later calls
        replay.  Not used with a process group (the all-reduce stays outside graphs) nor with ``want_neg`` /
        ``k_out`` outputs, which callers own."""
        if (self.group is not None and not self.fused_merge) or kw.get("want_neg") or kw.get("k_out") is not None:
            return self.correct(x0, sigma, scale, eps, **kw)       # an NCCL all-reduce stays outside graphs
        key = (x0.data_ptr(), tuple(x0.shape), float(sigma), float(scale), float(eps),
               tuple(sorted((k, v) for k, v in kw.items())))
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "seen"
            return self.correct(x0, sigma, scale, eps, **kw)
        if entry == "seen":
            if len(self._graphs) > 64:                      # bound the cache; pointers churn in long runs
                self._graphs = {key: "seen"}
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.correct(x0, sigma, scale, eps, **kw)
            entry = (g, out)
            self._graphs[key] = entry
        entry[0].replay()
        return entry[1]

    def ddpm_step(self, x_t, eps_pred, z1, z2, coeffs: dict, sigma: float, scale: float, eps: float, *,
                  gate_threshold: float | None = None, return_neg: bool = False, ddim: bool = False,
                  x0c_out: torch.Tensor | None = None):
        """One fused in-window SD-1.4 step (...threshold_time.py:554-576): eps -> x0, projection,
        correction, gate, re-noise and the scheduler update; no host sync.  ``coeffs`` as produced by
        ``epilogue.ddpm_coefficients``.  Returns (latents, scratch)."""
        sa, s1 = coeffs["sqrt_ab"], coeffs["sqrt_1m_ab"]
        Q, xt = self._flat_query(x_t)
        s0 = self._get(Q, True)
        # x0 is only needed as the query: written to the xq scratch, never handed back
        s = self.partial_sums(x_t, sigma, model_out=eps_pred, c_x=1.0 / sa, c_m=-s1 / sa, x0_out=s0.xq)
        out = torch.empty_like(xt)
        s.mean.zero_()
        flags = (nv.EPI_GATE if gate_threshold is not None else 0) | (nv.EPI_RETURN_NEG if return_neg else 0)
        thr = float(gate_threshold if gate_threshold is not None else 0.0)
        L, st = nv.lib(), nv.current_stream()
        e = self._flat_query(eps_pred)[1]
        n1 = self._flat_query(z1)[1]
        xc = None if x0c_out is None else self._flat_query(x0c_out)[1]
        if ddim:
            nv.check(L.sdn_epilogue_ddim(nv.ptr(s.num), nv.ptr(s.z), Q, self.bank.D, float(eps), float(scale),
                                         thr, flags, nv.ptr(xt), nv.ptr(e), nv.ptr(n1), sa, s1,
                                         coeffs["sqrt_ab_prev"], coeffs["sqrt_1m_ab_prev"],
                                         nv.ptr(out), nv.ptr(xc), nv.ptr(s.denom), nv.ptr(s.gate),
                                         nv.ptr(s.mean), st))
        else:
            n2 = None if z2 is None else self._flat_query(z2)[1]
            nv.check(L.sdn_epilogue_ddpm(nv.ptr(s.num), nv.ptr(s.z), Q, self.bank.D, float(eps), float(scale),
                                         thr, flags, nv.ptr(xt), nv.ptr(e), nv.ptr(n1), nv.ptr(n2), sa, s1,
                                         coeffs["c_x0"], coeffs["c_xt"], coeffs["sigma_noise"],
                                         nv.ptr(out), nv.ptr(xc), nv.ptr(s.denom), nv.ptr(s.gate),
                                         nv.ptr(s.mean), st))
        return out.view_as(x_t), s

    def flow_step(self, x, v, zn, sigma_t: float, sigma_next: float, kernel_sigma: float, scale: float,
                  eps: float, *, normalize_channels: int, x0c_out: torch.Tensor | None = None,
                  out_dtype: torch.dtype | None = None):
        """One fused SD3 flow-matching step (safe_denoiser_pipeline.py:1141-1161).  fp32 in; the pipeline casts the
        new latents back to its fp16 working dtype (:1161): pass ``out_dtype=torch.float16`` for that."""
        Q, xf = self._flat_query(x)
        s0 = self._get(Q, True)
        x0_tmp = None if normalize_channels > 0 else s0.xq   # un-normalised x0 is only the query
        s = self.partial_sums(x, kernel_sigma, normalize_channels=normalize_channels, model_out=v,
                              c_x=1.0, c_m=-float(sigma_t), x0_out=x0_tmp)
        out = torch.empty_like(xf)
        s.mean.zero_()
        xc = None if x0c_out is None else self._flat_query(x0c_out)[1]
        nv.check(nv.lib().sdn_epilogue_flow(nv.ptr(s.num), nv.ptr(s.z), Q, self.bank.D, float(eps), float(scale),
                                            nv.ptr(xf), nv.ptr(self._flat_query(v)[1]),
                                            nv.ptr(self._flat_query(zn)[1]), float(sigma_t), float(sigma_next),
                                            nv.ptr(out), nv.ptr(xc), nv.ptr(s.denom), nv.ptr(s.mean),
                                            nv.current_stream()))
        out = out.view_as(x)
        return (out.to(out_dtype) if out_dtype is not None and out_dtype != out.dtype else out), s

    def sparse(self, x0: torch.Tensor, radius: float, scale: float, want_term: bool = False,
               normalize_channels: int = 0):
        """SPELL baseline (fast.py:306-340): x0 += scale * sum_i relu(radius/d_i - 1)(xq - n_i) in place, where xq is
        x0, or x0 channel-normalised (fast_sdv3.py:332).  Returns (term or None, wsum [Q])."""
        L, st = nv.lib(), nv.current_stream()
        Q, xf = self._flat_query(x0)
        s = self._get(Q, normalize_channels > 0)
        b = self.bank
        xq = s.xq if normalize_channels > 0 else None
        nv.check(L.sdn_query_prepare(nv.ptr(xf), None, 1.0, 0.0, Q, b.D, int(normalize_channels), None, nv.ptr(xq),
                                     nv.ptr(s.xsq), st))
        # generic kernels: [Q,N] scratch + [Q,D]; the one-pass (Q <= 8) and tcgen05 (batched) kernels: their workspace + [Q,D]
        need = max(Q * b.N * 4, int(L.sdn_repel_workspace_bytes(Q, b.N, b.D, nv.PATH_AUTO))) + Q * b.D * 4 + 1024
        ws = torch.empty(need, dtype=torch.uint8, device=b.device)
        term = torch.empty_like(xf) if want_term else None
        query = xq if xq is not None else xf
        batched = Q > 8 and b.D % 128 == 0
        if batched:
            b.ensure_planes()
            # SPELL on the tcgen05 two-phase kernels (weights step with the relu(radius/d - 1) functor)
            packed = torch.empty(Q * b.D + Q, dtype=torch.float32, device=b.device)
            num, wsum = packed[: Q * b.D].view(Q, b.D), packed[Q * b.D:]
            nv.check(L.sdn_sparse_partial_planes(nv.ptr(b.planes), nv.ptr(b.sqnorm), b.N, b.D, nv.ptr(query), nv.ptr(s.xsq), Q,
                                                 float(radius), nv.ptr(num), nv.ptr(wsum), nv.ptr(ws), need, st))
            if self.group is not None:
                merge_partials(packed, self.group)
            nv.check(L.sdn_sparse_apply(nv.ptr(num), nv.ptr(wsum), Q, b.D, float(scale), nv.ptr(query), nv.ptr(xf),
                                        nv.ptr(term), st))
            return term, wsum
        if self.group is None:
            wsum = torch.empty(Q, dtype=torch.float32, device=b.device)
            nv.check(L.sdn_sparse_repel(nv.ptr(b.flat), nv.ptr(b.sqnorm), b.N, b.D, nv.ptr(xf), nv.ptr(xq), nv.ptr(s.xsq), Q,
                                        float(radius), float(scale), nv.ptr(term), nv.ptr(wsum),
                                        nv.ptr(ws), need, st))
            return term, wsum
        # N-sharded bank: every rank sums its rows (sum_i w_i n_i and sum_i w_i), ONE all-reduce of the packed
        # [Q*D | Q] buffer, then the same apply on every rank -- the force needs all rows of the bank
        packed = torch.empty(Q * b.D + Q, dtype=torch.float32, device=b.device)
        num, wsum = packed[: Q * b.D].view(Q, b.D), packed[Q * b.D:]
        nv.check(L.sdn_sparse_partial(nv.ptr(b.flat), nv.ptr(b.sqnorm), b.N, b.D, nv.ptr(xq if xq is not None else xf),
                                      nv.ptr(s.xsq), Q, float(radius), nv.ptr(num), nv.ptr(wsum), nv.ptr(ws), need, st))
        merge_partials(packed, self.group)
        nv.check(L.sdn_sparse_apply(nv.ptr(num), nv.ptr(wsum), Q, b.D, float(scale), nv.ptr(xq if xq is not None else xf),
                                    nv.ptr(xf), nv.ptr(term), st))
        return term, wsum


def conditioning_host(bank: NegativeBank, x0_host: torch.Tensor, denom_host: torch.Tensor, sigma: float,
                      scale: float, eps: float = 1e-8, normalize_channels: int = 0, path: int = nv.PATH_AUTO):
    """The e2e call: HOST query in, corrected HOST query out (in place), through the C ABI with the
    host<->device copies inside the call."""
    if x0_host.is_cuda or x0_host.dtype != torch.float32 or not x0_host.is_contiguous():
        raise RuntimeError("x0_host must be a contiguous CPU fp32 tensor")
    Q = int(x0_host.shape[0])
    nv.check(nv.lib().sdn_conditioning_host(
        nv.ptr(bank.flat), nv.ptr(bank.sqnorm), nv.ptr(bank.planes), bank.N, bank.D,
        x0_host.data_ptr(), Q, int(normalize_channels), 1.0 / (2.0 * float(sigma) ** 2), 1, 1.0,
        float(eps), float(scale), denom_host.data_ptr(), path, nv.current_stream()))
    return x0_host


class HostPipe:
    """Several host-buffer conditioning() calls in flight (sdn_host_pipe_*): independent requests overlap their
    PCIe copies with each other's kernels.  `submit(slot, x_in, x_out, denom)` returns at once, `wait(slot)` blocks
    until x_out / denom hold that slot's result.  Host tensors: contiguous CPU fp32, pinned for real overlap."""

    def __init__(self, bank: NegativeBank, Q: int, slots: int = 3):
        self.bank, self.Q, self.slots = bank, int(Q), int(slots)
        h = ctypes.c_void_p()
        with torch.cuda.device(bank.device):
            nv.check(nv.lib().sdn_host_pipe_create(self.Q, bank.N, bank.D, self.slots, ctypes.byref(h)))
        self._h = h

    def submit(self, slot: int, x_in: torch.Tensor, x_out: torch.Tensor, denom: torch.Tensor, sigma: float,
               scale: float, eps: float = 1e-8):
        for t in (x_in, x_out, denom):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("host tensors must be contiguous CPU fp32")
        if x_in.numel() != self.Q * self.bank.D or x_out.numel() != x_in.numel() or denom.numel() != self.Q:
            raise RuntimeError("host tensors do not match the pipe's (Q, D)")
        b = self.bank
        nv.check(nv.lib().sdn_host_pipe_submit(self._h, int(slot), nv.ptr(b.flat), nv.ptr(b.sqnorm), nv.ptr(b.planes),
                                               x_in.data_ptr(), x_out.data_ptr(), denom.data_ptr(),
                                               1.0 / (2.0 * float(sigma) ** 2), 1, 1.0, float(eps), float(scale)))

    def wait(self, slot: int):
        nv.check(nv.lib().sdn_host_pipe_wait(self._h, int(slot)))

    def close(self):
        if self._h is not None and self._h.value:
            nv.lib().sdn_host_pipe_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def inv_two_sigma_sq(sigma: float) -> float:
    return 1.0 / (2.0 * math.pow(float(sigma), 2))
