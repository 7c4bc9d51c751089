"""Where does the host-buffer (e2e) call spend its time?  Times, on cfg3's shape, the H2D and D2H copies alone,
the device-only projection, and sdn_conditioning_host; prints one line each.  Run on a GPU box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from safe_denoiser_b200.projection import NegativeBank, Projector, conditioning_host  # noqa: E402


def wall(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    t.sort()
    return t[len(t) // 2] * 1e6, t[0] * 1e6


def main():
    Q, N, D = 64, 3000, 16384
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(0)
    bank_t = torch.randn(N, D, generator=g).to(dev)
    x_src = (bank_t[torch.arange(Q) % N] + 0.05 * torch.randn(Q, D, generator=g).to(dev)).contiguous()
    bank = NegativeBank(bank_t)
    bank.ensure_planes()
    proj = Projector(bank)
    x = x_src.clone()
    xh = x_src.cpu().pin_memory()
    dh = torch.empty(Q).pin_memory()

    def h2d():
        x.copy_(xh, non_blocking=True)
        torch.cuda.synchronize()

    def d2h():
        xh.copy_(x, non_blocking=True)
        torch.cuda.synchronize()

    def dev_only():
        proj.correct(x, 1.0, 0.03, 1e-8)
        torch.cuda.synchronize()

    def graphed():
        proj.correct_graphed(x, 1.0, 0.03, 1e-8)
        torch.cuda.synchronize()

    def host():
        conditioning_host(bank, xh, dh, 1.0, 0.03, 1e-8)

    for name, fn in (("h2d 4MiB + sync", h2d), ("d2h 4MiB + sync", d2h), ("device correct + sync", dev_only),
                     ("graphed correct + sync", graphed), ("conditioning_host", host)):
        med, best = wall(fn)
        print(f"{name:28s} median {med:8.1f} us   best {best:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
