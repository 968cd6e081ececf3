"""CPU oracle (TEST INFRASTRUCTURE, never imported by the package) for PhiSNet's Clebsch-Gordan mixing layers -- SURVEY.md section 8 f4:

  PairMixing       nablaDFT/phisnet/nn/modules/pair_mixing.py:47-69      y_L = sum_{l1,l2} (rbf . W_{l1 l2 L}) * CG^{l1 l2 L} : (x1_{l1} (x) x2_{l2})
  SelfMixing       nablaDFT/phisnet/nn/modules/self_mixing.py:55-83      y_L = keep_L x_L + sum_{l1<l2} mix_{l1 l2 L} * CG^{l1 l2 L} : (x_{l1} (x) x_{l2})
  SphericalLinear  nablaDFT/phisnet/nn/modules/spherical_linear.py:50-59 SelfMixing (optional) followed by one Linear per order (bias on L = 0)

Features are lists over the order L of tensors [..., 2L+1, F]; every feature channel is mixed independently.
The real Clebsch-Gordan tensors are the reference's vendored table (`clebsch_gordan_coefficients_L10.npz`, module clebsch_gordan.py:15-28)
restricted to l <= 4: tests/golden/phisnet_cg_L4.npz, written by tests/golden/make_golden_phisnet.py.
PINNED: tests/test_oracle_phisnet.py compares these classes with outputs of the reference's own modules (tests/golden/phisnet_mixing.npz).
Parameter names equal the reference's, so state dicts are interchangeable.
"""
import os
from typing import List

import numpy as np
import torch
from torch import nn

_CG_FILE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "phisnet_cg_L4.npz")


class ClebschGordan(nn.Module):
    """(l1, l2, l3) -> real CG tensor [2 l1 + 1, 2 l2 + 1, 2 l3 + 1]  (clebsch_gordan.py:27-28)."""

    def __init__(self, path: str = _CG_FILE):
        super().__init__()
        self._t = {k: torch.from_numpy(v) for k, v in np.load(path).items()}

    def forward(self, l1: int, l2: int, l3: int) -> torch.Tensor:
        return self._t[f"{l1}_{l2}_{l3}"]


def paths(o1: int, o2: int, oo: int, strict_upper: bool = False):
    """Loop order of the reference (pair_mixing.py:28-36 / self_mixing.py:18-25): (l1, l2, L)."""
    return [(l1, l2, L) for l1 in range(o1 + 1) for l2 in range((l1 + 1) if strict_upper else 0, o2 + 1)
            for L in range(abs(l1 - l2), min(l1 + l2, oo) + 1)]


def _couple(cg: torch.Tensor, x1: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    # sum_{m1 m2} CG[m1, m2, m3] x1[..., m1, f] x2[..., m2, f]
    return torch.einsum("abc,...af,...bf->...cf", cg.to(x1.dtype), x1, x2)


class PairMixing(nn.Module):
    def __init__(self, order_in1, order_in2, order_out, num_basis_functions, num_features, clebsch_gordan):
        super().__init__()
        self.order_in1, self.order_in2, self.order_out = order_in1, order_in2, order_out
        self.num_basis_functions, self.num_features, self.clebsch_gordan = num_basis_functions, num_features, clebsch_gordan
        for l1, l2, L in paths(order_in1, order_in2, order_out):
            self.add_module(f"coeff_{l1}_{l2}_{L}", nn.Linear(num_basis_functions, num_features, bias=False))

    def forward(self, x1s: List[torch.Tensor], x2s: List[torch.Tensor], rbf: torch.Tensor) -> List[torch.Tensor]:
        ys = [x1s[0].new_zeros(*x1s[0].shape[:-2], 2 * L + 1, x1s[0].shape[-1]) for L in range(self.order_out + 1)]
        for l1, l2, L in paths(self.order_in1, self.order_in2, self.order_out):
            ys[L] = ys[L] + getattr(self, f"coeff_{l1}_{l2}_{L}")(rbf) * _couple(self.clebsch_gordan(l1, l2, L), x1s[l1], x2s[l2])
        return ys


class SelfMixing(nn.Module):
    def __init__(self, order_in, order_out, num_features, clebsch_gordan):
        super().__init__()
        self.order_in, self.order_out, self.num_features, self.clebsch_gordan = order_in, order_out, num_features, clebsch_gordan
        for l1, l2, L in paths(order_in, order_in, order_out, strict_upper=True):
            self.register_parameter(f"mixcoeff_{l1}_{l2}_{L}", nn.Parameter(torch.zeros(num_features)))
        for L in range(min(order_in, order_out) + 1):
            self.register_parameter(f"keepcoeff_{L}", nn.Parameter(torch.ones(num_features)))

    def forward(self, xs: List[torch.Tensor]) -> List[torch.Tensor]:
        ys = [getattr(self, f"keepcoeff_{L}") * xs[L] if L <= self.order_in else xs[0].new_zeros(*xs[0].shape[:-2], 2 * L + 1, xs[0].shape[-1])
              for L in range(self.order_out + 1)]
        for l1, l2, L in paths(self.order_in, self.order_in, self.order_out, strict_upper=True):
            ys[L] = ys[L] + getattr(self, f"mixcoeff_{l1}_{l2}_{L}") * _couple(self.clebsch_gordan(l1, l2, L), xs[l1], xs[l2])
        return ys


class SphericalLinear(nn.Module):
    def __init__(self, order_in, num_in, order_out, num_out, clebsch_gordan=None, mix_orders=True, bias=True, zero_init=False):
        super().__init__()
        self.order_in, self.num_in, self.order_out, self.num_out, self.mix_orders = order_in, num_in, order_out, num_out, mix_orders
        if mix_orders:
            self.mixing = SelfMixing(order_in, order_out, num_in, clebsch_gordan)
        else:
            assert order_in == order_out
        self.linear = nn.ModuleList([nn.Linear(num_in, num_out, bias=(bias and L == 0)) for L in range(order_out + 1)])

    def forward(self, xs: List[torch.Tensor]) -> List[torch.Tensor]:
        ys = self.mixing(xs) if self.mix_orders else xs
        return [lin(y) for lin, y in zip(self.linear, ys)]
