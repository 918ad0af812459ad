"""``e3nn.math.normalize2mom`` restated.  TEST INFRASTRUCTURE ONLY (oracle)."""
import functools

import torch


@functools.lru_cache(maxsize=None)
def _second_moment_constant(key: str) -> float:
    f = {"silu": torch.nn.functional.silu, "tanh": torch.tanh}[key]
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    return f(z).pow(2).mean().pow(-0.5).item()


def _moment_of(f) -> float:
    if isinstance(f, torch.nn.SiLU) or f is torch.nn.functional.silu:
        return _second_moment_constant("silu")
    if f is torch.tanh or isinstance(f, torch.nn.Tanh):
        return _second_moment_constant("tanh")
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    return f(z).pow(2).mean().pow(-0.5).item()


class normalize2mom(torch.nn.Module):
    def __init__(self, f, dtype=None, device=None) -> None:
        super().__init__()
        self.f = f
        self.cst = _moment_of(f)
        self._is_id = abs(self.cst - 1.0) < 1e-4

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.f(x) if self._is_id else self.f(x).mul(self.cst)
