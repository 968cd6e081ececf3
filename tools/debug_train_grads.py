import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch
from helpers import load_fixture
from test_gpu_painn import _Data, _oc_model, dev
from oracle.painn_oc import PaiNNOC

for L in (1, 3):
    kw = dict(hidden_channels=128, num_layers=L, num_rbf=100, cutoff=5.0, max_neighbors=100, num_elements=100)
    net = _oc_model(L)
    ref = PaiNNOC(**kw).double()
    ref.load_state_dict({k: v.double() for k, v in net.state_dict().items()}, strict=True)
    z, pos, batch = load_fixture([0, 4, 7])
    c = torch.tensor([0.7, -1.3, 0.4], dtype=torch.float64)
    e_ref, f_ref = ref(z, pos.clone(), batch, create_graph=True)
    (c * e_ref).sum().backward()
    net = net.to(dev()).train()
    e, f = net(_Data(z.to(dev()), pos.float().to(dev()), batch.to(dev())))
    (c.float().to(dev()) * e).sum().backward()
    print("L", L, "dE", float((e.detach().cpu().double() - e_ref.detach()).abs().max()))
    for (k, p), (k2, p2) in zip(net.named_parameters(), ref.named_parameters()):
        g, g2 = p.grad, p2.grad
        if g2 is None:
            print(f"{k:45s} ref none; ours {None if g is None else float(g.abs().max())}")
            continue
        rel = float((g.double().cpu() - g2).abs().max() / (g2.abs().max() + 1e-30))
        print(f"{k:45s} rel {rel:9.2e}  |ref| {float(g2.abs().max()):9.2e}  |ours| {float(g.abs().max()):9.2e}")
