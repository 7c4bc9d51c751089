"""CPU-only checks: the C-ABI library loads and exports what include/sdn_repel.h declares, the host-side
mirror of the reference interface behaves like the reference (registry, errors, cache format), and the
N-shard merge works over gloo with world_size 2.  No kernel is launched here."""
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nv():
    from safe_denoiser_b200.build import build_native
    build_native()
    from safe_denoiser_b200 import _native
    _native.lib()
    return _native


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdn_repel.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdn_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(nv):
    import ctypes
    handle = ctypes.CDLL(nv.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(handle, name), f"{name} declared in sdn_repel.h but not exported"
        assert name in nv.SIGNATURES, f"{name} has no ctypes signature"
    assert set(nv.SIGNATURES) == set(names)


def test_host_only_entry_points(nv):
    L = nv.lib()
    assert L.sdn_abi_version() == 1
    assert L.sdn_error_string(0) == b"ok"
    assert b"NULL" in L.sdn_error_string(-1)
    assert L.sdn_repel_workspace_bytes(4, 100, 16384, nv.PATH_GENERIC) >= 4 * 100 * 4
    assert L.sdn_repel_workspace_bytes(0, 100, 16384, nv.PATH_AUTO) == 0
    # argument validation happens before any CUDA call
    assert L.sdn_repel_partial(None, None, None, 1, 4, None, None, 1, 1.0, 1, 1.0, None, None, None, None, 0, 0, None) == -1
    assert L.sdn_epilogue_correct(None, None, 1, 4, 0.0, 0.0, 0.0, 0, None, None, None, None, None, None) == -1
    assert L.sdn_launch_count() == 0


def test_host_pipe_argument_checks(nv):
    """sdn_host_pipe_*: every argument error is reported before any CUDA call (this box has no GPU)."""
    import ctypes
    L = nv.lib()
    h = ctypes.c_void_p()
    assert L.sdn_host_pipe_create(4, 100, 16384, 3, None) == -1            # no place for the handle
    assert L.sdn_host_pipe_create(0, 100, 16384, 3, ctypes.byref(h)) == -2  # Q
    assert L.sdn_host_pipe_create(4, 100, 16384, 0, ctypes.byref(h)) == -2  # slots: 1..8
    assert L.sdn_host_pipe_create(4, 100, 16384, 9, ctypes.byref(h)) == -2
    assert not h.value
    assert L.sdn_host_pipe_submit(None, 0, None, None, None, None, None, None, 1.0, 1, 1.0, 1e-8, 0.3) == -1
    assert L.sdn_host_pipe_wait(None, 0) == -1
    L.sdn_host_pipe_destroy(None)                                           # a null handle is ignored


def test_missing_library_is_loud(monkeypatch):
    from safe_denoiser_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libsdn_repel.so")
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()


MODULES = ["fast", "fast_sdv3", "threshold"]


def _mod(kind):
    import importlib
    return importlib.import_module(f"safe_denoiser_b200.repellency.repellency_methods_{kind}")


def _make(kind, name, tmp_path, bank=None, **params):
    if bank is None:
        bank = torch.randn(6, 4, 8, 8)
    path = str(tmp_path / "sub" / "dir" / f"{kind}_{name}.pt")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(bank, path)
    return _mod(kind).get_repellency_method(
        name, ref_data=torch.zeros(1, 3, 8, 8), embed_fn=None, forward_fn=None, num_timesteps=50,
        max_idx=1000, beta_min=0.00085, beta_max=0.012, n_embed=4, proj_ref_path=path,
        cache_proj_ref=True, **params), bank


@pytest.mark.parametrize("kind", MODULES)
def test_registry_matches_reference_surface(kind, tmp_path):
    m = _mod(kind)
    live = {"fast": {"kernel_fast", "sparse", "random_noise"}, "fast_sdv3": {"kernel_fast", "sparse", "random_noise"},
            "threshold": {"kernel_fast", "sparse"}}[kind]
    assert live <= set(m.__CONDITIONING_METHOD__)
    with pytest.raises(NameError):
        m.get_repellency_method("nope", None, None, None, 50, 1000, 0.1, 0.2, n_embed=1)
    with pytest.raises(NameError):
        m.register_conditioning_method("kernel_fast")(object)
    for dead in ("euclidean", "kernel"):           # fast.py:142,:181 take six positionals -> TypeError
        with pytest.raises(TypeError):
            _make(kind, dead, tmp_path)
    # registries are per module, as in the reference
    others = [o for o in MODULES if o != kind]
    assert all(_mod(o).__CONDITIONING_METHOD__ is not m.__CONDITIONING_METHOD__ for o in others)


@pytest.mark.parametrize("kind", MODULES)
def test_constructor_attributes_and_cache_format(kind, tmp_path):
    proc, bank = _make(kind, "kernel_fast", tmp_path, scale=0.33, sigma=3.15, epsilon=1e-6,
                       beta_threshold=2.5, beta_threshold_margin=1.6, unknown_kwarg=123)
    assert proc.scale == 0.33 and proc.epsilon == 1e-6 and proc.n_embed == 4
    assert proc.num_timesteps == 50 and proc.max_idx == 1000 and proc.cache_proj_ref is True
    assert torch.equal(proc.get_proj_ref(), bank) and proc.get_proj_ref().dtype == torch.float32
    if kind == "threshold":
        assert proc.sigma == 3.15 and proc.beta_threshold == 2.5 and proc.beta_threshold_margin == 1.6
        assert proc.quantile == 0.0
    else:
        assert callable(proc.sigma)       # the YAML sigma is ignored by fast / fast_sdv3
    for name in ("project", "set_proj_ref", "import_proj_ref", "get_proj_ref", "mkdir_cache",
                 "empirical_denoiser", "conditioning", "conditioning_1", "discrete_to_continous_time", "sigma_edm"):
        assert callable(getattr(proc, name))
    assert proc.discrete_to_continous_time(0) == 0.001 and proc.discrete_to_continous_time(500) == 0.5


def test_project_normalises_channels_and_chunks(tmp_path):
    proc, _ = _make("fast", "kernel_fast", tmp_path)
    calls = []

    def embed(x):
        calls.append(len(x))
        return x[:, :2].double() * 3.0
    proc.embed_fn = embed
    out = proc.project(torch.randn(10, 3, 4, 4))
    assert calls == [4, 4, 2] and out.dtype == torch.float32
    np.testing.assert_allclose(out.norm(dim=1).numpy(), 1.0, rtol=1e-6)
    thr, _ = _make("threshold", "kernel_fast", tmp_path, beta_threshold=1.0)
    thr.embed_fn = embed
    assert thr.project(torch.randn(3, 3, 4, 4)).dtype == torch.float64    # threshold.py has no .float()


@pytest.mark.parametrize("kind", MODULES)
def test_no_cpu_fallback(kind, tmp_path):
    proc, _ = _make(kind, "kernel_fast", tmp_path, beta_threshold=1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        proc.conditioning(torch.randn(1, 4, 8, 8), beta_threshold=True)


def test_auto_beta_needs_scheduler(tmp_path):
    with pytest.raises(AssertionError):
        _make("threshold", "kernel_fast", tmp_path, sigma=3.15)       # threshold.py:297


def test_shard_bounds_partition():
    from safe_denoiser_b200.projection import shard_bounds
    for n in (1, 7, 515, 3000, 200000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_epilogue_coefficients_match_scheduler_oracle():
    from oracle import scheduler_oracle as so
    from safe_denoiser_b200 import epilogue as ep
    ac = ep.sd14_alphas_cumprod()
    np.testing.assert_allclose(ac, so.sd14_alphas_cumprod(), rtol=5e-5)   # float32 linspace: torch vs numpy rounding
    assert ep.ddpm_timesteps() == list(so.ddpm_timesteps())
    assert ep.ddpm_timesteps()[0] == 981 and ep.ddpm_timesteps()[-1] == 1
    for t in (981, 781, 1):
        a, b = ep.ddpm_coefficients(ac, t), so.ddpm_coefficients(np.asarray(ac), t)
        for k in b:
            assert abs(a[k] - b[k]) <= 1e-12 * max(1.0, abs(b[k])), (t, k)
    last = ep.ddpm_coefficients(ac, 1)
    assert last["c_xt"] == 0.0                                       # DDPM, t_prev < 0 -> abar_prev := 1
    assert abs(last["sqrt_ab_prev"] - ac[0] ** 0.5) < 1e-15          # DDIM: final_alpha_cumprod = alphas_cumprod[0]
    assert ep.ddpm_coefficients(ac, 1, final_alpha_cumprod=1.0)["sqrt_ab_prev"] == 1.0     # set_alpha_to_one


def test_lazy_scalar_behaves_like_a_float():
    from safe_denoiser_b200.repellency._base import LazyScalar
    v = LazyScalar(torch.tensor([2.5]))
    assert float(v) == 2.5 and v > 1.0 and v < 3 and f"{v:.1f}" == "2.5" and v + 1 == 3.5 and bool(v)


# ---------------------------------------------------------------- N-shard merge over gloo, world_size 2
def _gloo_worker(rank, world, port, tmpdir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import repellency_oracle as orc
    from safe_denoiser_b200.projection import merge_partials, shard_bounds
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    bank = orc.synthetic_bank(37, 4, 8, 8)
    x = orc.synthetic_queries(bank, 3, "near")
    Q, D = 3, 256
    lo, hi = shard_bounds(37, rank, world)
    part = orc.closed_form(x.numpy(), bank[lo:hi].numpy(), sigma=3.15)
    packed = torch.zeros(Q * D + Q, dtype=torch.float32)
    packed[:Q * D] = torch.from_numpy(part["num"].astype(np.float32)).reshape(-1)
    packed[Q * D:] = torch.from_numpy(part["Z"].astype(np.float32))
    merge_partials(packed, dist.group.WORLD)
    full = orc.closed_form(x.numpy(), bank.numpy(), sigma=3.15)
    num = packed[:Q * D].view(Q, D).numpy()
    z = packed[Q * D:].numpy()
    ok = (np.abs(num - full["num"]).max() <= 1e-5 * np.abs(full["num"]).max()
          and np.abs(z - full["Z"]).max() <= 1e-5 * np.abs(full["Z"]).max())
    neg = num / (z[:, None] + 1e-8)
    ok = ok and np.abs(neg - full["neg"]).max() <= 1e-5 * np.abs(full["neg"]).max()
    open(os.path.join(tmpdir, f"rank{rank}.ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_nshard_merge_gloo_world2(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.ok").read() == "1"


def test_new_entry_points_validate_arguments(nv):
    import ctypes as C
    L = nv.lib()
    assert L.sdn_set_option(nv.OPT_SKIP_NEGLIGIBLE, 1) == 0
    assert L.sdn_set_option(12345, 1) == -4
    assert L.sdn_repel_path(1, 515, 16384, 0, nv.PATH_AUTO) == nv.PATH_STREAM
    assert L.sdn_repel_path(64, 3000, 16384, 1, nv.PATH_AUTO) == nv.PATH_UMMA        # the one-pass kernel is opt-in
    assert L.sdn_repel_path(64, 3000, 16384, 1, nv.PATH_FLASH) == nv.PATH_FLASH
    assert L.sdn_repel_path(16, 515, 65536, 1, nv.PATH_AUTO) == nv.PATH_UMMA
    assert L.sdn_repel_path(8, 3000, 16384, 1, nv.PATH_AUTO) == nv.PATH_UMMA         # FMA-bound for the cluster kernel
    assert L.sdn_repel_path(8, 515, 16384, 1, nv.PATH_AUTO) == nv.PATH_STREAM
    assert L.sdn_repel_path(4, 3000, 16384, 1, nv.PATH_AUTO) == nv.PATH_STREAM
    assert L.sdn_repel_path(64, 3000, 16384, 0, nv.PATH_AUTO) == nv.PATH_GENERIC     # no planes: CUDA cores
    assert L.sdn_repel_path(3, 37, 256, 0, nv.PATH_AUTO) == nv.PATH_GENERIC          # D too small for the cluster kernel
    assert L.sdn_conditioning_fused(None, None, None, 1, 4, None, 1, 1.0, 1, 1.0, 0.0, 0.0, 0.0, 0,
                                    None, None, None, None, None, None, None, None, 0, 0, None) == -1
    arr = (C.c_void_p * 2)(None, None)
    assert L.sdn_shard_merge_correct(arr, arr, arr, 0, 9, None, 1, 4, 0.0, 0.0, 0.0, 0, None, None, None, None, None) == -1
    assert L.sdn_bank_build(None, 1, 4, 16, None, None, None, None) == -1
