"""GPU parity of the PhiSNet mixing layers (csrc/phisnet.cu behind nabladft_b200.phisnet) against the outputs of the REFERENCE'S OWN modules
(tests/golden/phisnet_mixing.npz) and, at a pair count of config size, against the CPU oracle.  Inputs are O(1) random features, so the
outputs are O(1)..O(10): the tolerance is the Hamiltonian-block one of north_star scaled to that magnitude (1e-6 relative to the largest entry)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN
from test_oracle_phisnet import G, feats, load_named

pytestmark = pytest.mark.gpu
REL = 2e-6


def dev():
    return torch.device("cuda:0")


def check(ys, tag, order_out):
    for L in range(order_out + 1):
        ref = G[f"{tag}/y/{L}"]
        got = ys[L].double().cpu().numpy()
        assert got.shape == ref.shape
        err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
        assert err < REL, (tag, L, err)


@pytest.mark.parametrize("tag,orders", [("pair444", (4, 4, 4)), ("pair214", (2, 1, 4))])
def test_pair_mixing_matches_reference_golden(tag, orders):
    from nabladft_b200 import phisnet as ph

    o1, o2, oo = orders
    pm = load_named(ph.PairMixing(o1, o2, oo, 128, 128, ph.ClebschGordan()), tag, torch.float32).to(dev()).eval()
    ys = pm([t.float().to(dev()) for t in feats(tag, "x1", o1)], [t.float().to(dev()) for t in feats(tag, "x2", o2)],
            torch.from_numpy(G[f"{tag}/rbf"]).float().to(dev()))
    check(ys, tag, oo)


@pytest.mark.parametrize("tag,orders", [("self44", (4, 4)), ("self42", (4, 2)), ("self24", (2, 4))])
def test_self_mixing_matches_reference_golden(tag, orders):
    from nabladft_b200 import phisnet as ph

    oi, oo = orders
    sm = load_named(ph.SelfMixing(oi, oo, 128, ph.ClebschGordan()), tag, torch.float32).to(dev()).eval()
    check(sm([t.float().to(dev()) for t in feats(tag, "x", oi)]), tag, oo)


@pytest.mark.parametrize("tag,cfg", [("lin44", (4, 128, 4, 128, True)), ("lin40", (4, 128, 0, 64, True)), ("lin22n", (2, 128, 2, 64, False))])
def test_spherical_linear_matches_reference_golden(tag, cfg):
    from nabladft_b200 import phisnet as ph

    oi, fi, oo, fo, mix = cfg
    sl = load_named(ph.SphericalLinear(oi, fi, oo, fo, ph.ClebschGordan(), mix_orders=mix, bias=True), tag, torch.float32).to(dev()).eval()
    check(sl([t.float().to(dev()) for t in feats(tag, "x", oi)]), tag, oo)


def test_pair_mixing_config_size_vs_oracle_and_state_dict_names():
    """2 000 atom pairs, order 4, 128 features, 128 basis functions (phisnet/configs/args_*.txt): strict state-dict exchange with the oracle
    (= the reference's parameter names), values against a float64 oracle pass on a 64-pair slice, determinism."""
    from nabladft_b200 import phisnet as ph
    from oracle import phisnet as op

    g = torch.Generator().manual_seed(3)
    P = 2000
    ours = ph.PairMixing(4, 4, 4, 128, 128).to(dev()).eval()
    ref = op.PairMixing(4, 4, 4, 128, 128, op.ClebschGordan()).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in ours.state_dict().items()}, strict=True)
    x1 = [torch.randn(1, P, 2 * L + 1, 128, generator=g) for L in range(5)]
    x2 = [torch.randn(1, P, 2 * L + 1, 128, generator=g) for L in range(5)]
    rbf = torch.rand(1, P, 1, 128, generator=g)
    ys = ours([t.to(dev()) for t in x1], [t.to(dev()) for t in x2], rbf.to(dev()))
    ys2 = ours([t.to(dev()) for t in x1], [t.to(dev()) for t in x2], rbf.to(dev()))
    yr = ref([t[:, :64].double() for t in x1], [t[:, :64].double() for t in x2], rbf[:, :64].double())
    for L in range(5):
        assert torch.equal(ys[L], ys2[L]) and ys[L].shape == (1, P, 2 * L + 1, 128)
        err = (ys[L][:, :64].double().cpu() - yr[L]).abs().max().item() / max(1.0, yr[L].abs().max().item())
        assert err < REL, (L, err)
