"""Per-tensor comparison of the engine's parameter gradients with the oracle's autograd (energy seed, force-only loss = double backward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from helpers import load_fixture
from test_gpu_painn import _Data, _oc_model, dev
from oracle.painn_oc import PaiNNOC

which = sys.argv[1] if len(sys.argv) > 1 else "force"
for L in (1, 3):
    kw = dict(hidden_channels=128, num_layers=L, num_rbf=100, cutoff=5.0, max_neighbors=100, num_elements=100)
    net = _oc_model(L)
    ref = PaiNNOC(**kw).double()
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()}, strict=True)
    z, pos, batch = load_fixture([0, 4, 7])
    c = torch.tensor([0.7, -1.3, 0.4], dtype=torch.float64)
    g = torch.Generator().manual_seed(3)
    f_t = 0.05 * torch.randn(pos.shape, generator=g, dtype=torch.float64)
    loss = (lambda e, f, dt: (c.to(dt).to(e.device) * e).sum()) if which == "energy" else (lambda e, f, dt: ((f - f_t.to(dt).to(f.device)) ** 2).mean())
    e_ref, f_ref = ref(z, pos.clone(), batch, create_graph=True)
    loss(e_ref, f_ref, torch.float64).backward()
    net = net.to(dev()).train()
    e, f = net(_Data(z.to(dev()), pos.float().to(dev()), batch.to(dev())))
    loss(e, f, torch.float32).backward()
    print("L", L, which)
    for (k, p), (k2, p2) in zip(net.named_parameters(), ref.named_parameters()):
        g2 = p2.grad if p2.grad is not None else torch.zeros_like(p2)
        g1 = p.grad.double().cpu() if p.grad is not None else torch.zeros_like(p2)
        if float(g2.abs().max()) == 0 and float(g1.abs().max()) == 0:
            continue
        rel = float((g1 - g2).abs().max() / (g2.abs().max() + 1e-30))
        print(f"{k:45s} rel {rel:9.2e}  |ref| {float(g2.abs().max()):9.2e}  |ours| {float(g1.abs().max()):9.2e}")
