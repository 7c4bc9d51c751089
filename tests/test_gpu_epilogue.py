"""Fused epilogue kernels (eps->x0, correction, beta gate, re-noise, scheduler update) against the
scheduler oracle chained with the projection oracle.  Tolerance 1e-3 max-rel as for the projection."""
import numpy as np
import pytest
import torch

from oracle import repellency_oracle as orc
from oracle import scheduler_oracle as so

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(a, b):
    a = np.asarray(a.detach().cpu().numpy() if torch.is_tensor(a) else a, np.float64).reshape(np.shape(b))
    return np.abs(a - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30)


def _setup(Q, N=64, C=4, H=32, W=32, seed=0):
    from safe_denoiser_b200.projection import NegativeBank, Projector
    bank4 = orc.synthetic_bank(N, C, H, W)
    g = torch.Generator().manual_seed(seed)
    shape = (Q, C, H, W)
    idx = torch.randint(0, N, (Q,), generator=g)
    # latents whose x0-prediction lands near a negative, so that the correction is not trivially small
    ac = so.sd14_alphas_cumprod()
    return bank4, Projector(NegativeBank(bank4.cuda())), g, shape, idx, ac


@pytest.mark.parametrize("t", [981, 781, 1])
@pytest.mark.parametrize("mode", ["ddpm", "ddim"])
@pytest.mark.parametrize("return_neg", [False, True])
def test_fused_sd14_step(t, mode, return_neg):
    from safe_denoiser_b200.epilogue import ddpm_coefficients, sd14_alphas_cumprod
    Q = 5
    bank4, proj, g, shape, idx, ac = _setup(Q)
    co = ddpm_coefficients(sd14_alphas_cumprod(), t)
    eps_pred = torch.randn(shape, generator=g)
    x0_target = bank4[idx] + 0.05 * torch.randn(shape, generator=g)
    x_t = co["sqrt_ab"] * x0_target + co["sqrt_1m_ab"] * eps_pred
    z1 = torch.randn(shape, generator=g)
    z2 = torch.randn(shape, generator=g)
    sigma, scale, eps = 3.15, 0.33, 1e-8
    x0 = so.eps_to_x0(x_t.numpy(), eps_pred.numpy(), co)
    ref = orc.closed_form(x0, bank4.numpy(), sigma=sigma, eps=eps)
    srt = np.sort(ref["denom"])
    thr = float(0.5 * (srt[1] + srt[2]))                      # some rows pass the gate, some do not
    gate = ref["denom"] > thr
    assert gate.any() and not gate.all()
    neg = ref["neg"].reshape(shape)
    if mode == "ddpm":
        want = so.ddpm_fused_step(x_t.numpy(), eps_pred.numpy(), neg, gate, scale, co, z1.numpy(), z2.numpy(),
                                  return_neg_as_x0=return_neg)
    else:
        want = so.ddim_fused_step(x_t.numpy(), eps_pred.numpy(), neg, gate, scale, co, z1.numpy(),
                                  return_neg_as_x0=return_neg)
    x0c = torch.empty(shape, device="cuda")
    out, s = proj.ddpm_step(x_t.cuda(), eps_pred.cuda(), z1.cuda(), z2.cuda(), co, sigma, scale, eps,
                            gate_threshold=thr, return_neg=return_neg, ddim=(mode == "ddim"), x0c_out=x0c)
    torch.cuda.synchronize()
    assert (s.gate.cpu().numpy().astype(bool) == gate).all()
    assert rel(s.denom, ref["denom"]) <= TOL
    assert rel(x0c, want["x0_corrected"]) <= TOL
    assert rel(out, want["prev"]) <= TOL


def test_fused_sd14_step_without_gate_renoises_every_row():
    from safe_denoiser_b200.epilogue import ddpm_coefficients, sd14_alphas_cumprod
    Q = 3
    bank4, proj, g, shape, idx, ac = _setup(Q, seed=3)
    co = ddpm_coefficients(sd14_alphas_cumprod(), 881)
    eps_pred, z1, z2 = (torch.randn(shape, generator=g) for _ in range(3))
    x_t = co["sqrt_ab"] * bank4[idx] + co["sqrt_1m_ab"] * eps_pred
    x0 = so.eps_to_x0(x_t.numpy(), eps_pred.numpy(), co)
    ref = orc.closed_form(x0, bank4.numpy(), sigma=1.0)
    want = so.ddpm_fused_step(x_t.numpy(), eps_pred.numpy(), ref["neg"].reshape(shape), np.ones(Q, bool), 0.03, co,
                              z1.numpy(), z2.numpy())
    out, s = proj.ddpm_step(x_t.cuda(), eps_pred.cuda(), z1.cuda(), z2.cuda(), co, 1.0, 0.03, 1e-8)
    torch.cuda.synchronize()
    assert s.gate.cpu().numpy().all()
    assert rel(out, want["prev"]) <= TOL


@pytest.mark.parametrize("sigma_t,sigma_next", [(1.0, 0.98), (0.8, 0.78), (0.02, 0.0)])
def test_fused_flow_step_sd3(sigma_t, sigma_next):
    from safe_denoiser_b200.projection import NegativeBank, Projector
    Q, N, C, H, W = 3, 48, 16, 16, 16
    bank4 = orc.synthetic_bank(N, C, H, W)
    proj = Projector(NegativeBank(bank4.cuda()))
    g = torch.Generator().manual_seed(11)
    shape = (Q, C, H, W)
    x = torch.randn(shape, generator=g)
    v = torch.randn(shape, generator=g)
    zn = torch.randn(shape, generator=g)
    x0 = x.numpy().astype(np.float64) - sigma_t * v.numpy()
    ref = orc.closed_form(x0, bank4.numpy(), sigma=1.0, normalise_query=True)
    want = so.flow_fused_step(x.numpy(), v.numpy(), ref["neg"].reshape(shape), 0.03, sigma_t, sigma_next, zn.numpy())
    x0c = torch.empty(shape, device="cuda")
    out, s = proj.flow_step(x.cuda(), v.cuda(), zn.cuda(), sigma_t, sigma_next, 1.0, 0.03, 1e-8,
                            normalize_channels=C, x0c_out=x0c)
    torch.cuda.synchronize()
    assert rel(s.denom, ref["denom"]) <= TOL
    assert rel(x0c, want["x0_corrected"]) <= TOL
    assert rel(out, want["next"]) <= TOL


def test_in_window_steps_replay_from_one_cuda_graph():
    """SURVEY f-4: several consecutive in-window SD-1.4 steps (device-side gate, no host sync) captured in ONE CUDA
    graph give the same latents as the same steps launched eagerly."""
    from safe_denoiser_b200.epilogue import ddpm_coefficients, ddpm_timesteps, sd14_alphas_cumprod
    Q = 16
    bank4, proj, g, shape, idx, ac = _setup(Q, N=200, C=4, H=64, W=64, seed=4)
    acp = sd14_alphas_cumprod()
    steps = [t for t in ddpm_timesteps(50) if t >= 900]            # 981, 961, 941, 921, 901
    cos = [ddpm_coefficients(acp, t) for t in steps]
    noise = torch.randn((len(steps), 3) + shape, generator=g).cuda()
    x_init = (cos[0]["sqrt_ab"] * bank4[idx] + cos[0]["sqrt_1m_ab"] * torch.randn(shape, generator=g)).cuda()

    def run(x):
        for i, co in enumerate(cos):
            eps_pred = 0.1 * x + noise[i, 0]                        # stand-in for the UNet
            x, s = proj.ddpm_step(x, eps_pred, noise[i, 1], noise[i, 2], co, 3.15, 0.33, 1e-8, gate_threshold=2.0)
        return x

    want = run(x_init.clone())
    torch.cuda.synchronize()
    xin = x_init.clone()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = run(xin)
    for _ in range(2):
        xin.copy_(x_init)
        gr.replay()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert torch.equal(out, want)


def test_fused_flow_step_against_reference_executed_golden():
    """A8 pinned: the fused SD3 step vs the golden made by executing safe_denoiser_pipeline.py:1141-1161 (fp16 in/out)."""
    import os
    from safe_denoiser_b200.projection import NegativeBank, Projector
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "flow_cases.npz"))
    bank4 = torch.from_numpy(d["flow/bank"])
    proj = Projector(NegativeBank(bank4.cuda()))
    for c in range(3):
        k = f"flow/case{c}"
        x, v, z = (torch.from_numpy(d[k + s]).cuda() for s in ("/latents", "/v", "/z"))      # fp16, as the pipeline holds them
        sg, sn = float(d[k + "/sigma"]), float(d[k + "/sigma_next"])
        x0c = torch.empty(x.shape, device="cuda")
        out, _ = proj.flow_step(x.float(), v.float(), z.float(), sg, sn, 1.0, 0.03, 1e-8,
                                normalize_channels=bank4.shape[1], x0c_out=x0c, out_dtype=torch.float16)
        torch.cuda.synchronize()
        assert out.dtype == torch.float16
        assert rel(x0c, d[k + "/x0_corrected"]) <= 1.5e-3
        assert rel(out.float(), d[k + "/latents_next"].astype(np.float32)) <= 1.5e-3


def test_query_prepare_eps_to_x0_and_channel_norm():
    from safe_denoiser_b200 import _native as nv
    Q, C, HW = 3, 16, 64
    D = C * HW
    g = torch.Generator().manual_seed(5)
    x = torch.randn(Q, D, generator=g).cuda()
    m = torch.randn(Q, D, generator=g).cuda()
    x0 = torch.empty_like(x)
    xq = torch.empty_like(x)
    xsq = torch.empty(Q, device="cuda")
    nv.check(nv.lib().sdn_query_prepare(x.data_ptr(), m.data_ptr(), 1.25, -0.5, Q, D, C, x0.data_ptr(),
                                        xq.data_ptr(), xsq.data_ptr(), nv.current_stream()))
    torch.cuda.synchronize()
    want0 = 1.25 * x.double() - 0.5 * m.double()
    wantq = orc.channel_normalise(want0.cpu().numpy().reshape(Q, C, HW)).reshape(Q, D)
    assert rel(x0, want0.cpu().numpy()) <= 1e-6
    assert rel(xq, wantq) <= 1e-5
    assert rel(xsq, (wantq ** 2).sum(1)) <= 1e-5


def test_bank_build_from_latents_matches_project():
    """sdn_bank_build == RepellencyMethod.project's normalisation (fast.py:55-56) + sdn_bank_prepare."""
    from safe_denoiser_b200.projection import NegativeBank
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(33, 4, 16, 16, generator=g) * 3.0
    want = (lat / lat.norm(dim=1, keepdim=True)).double()
    bank = NegativeBank.from_latents(lat.cuda(), with_planes=True)
    torch.cuda.synchronize()
    assert rel(bank.tensor, want.numpy()) <= 1e-6
    assert rel(bank.sqnorm, (want.reshape(33, -1) ** 2).sum(1).numpy()) <= 1e-5
    ref = NegativeBank(bank.tensor.clone(), with_planes=True)
    assert torch.equal(ref.planes, bank.planes)
    recon = bank.planes[0].float() + bank.planes[1].float()
    assert rel(recon, bank.flat.cpu().numpy()) <= 2e-5
