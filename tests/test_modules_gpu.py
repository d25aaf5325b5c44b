"""GPU tests of the smaller drop-in modules through the nn.Module API (CUDA kernels behind the C ABI, no emulation)."""
import pytest
import torch

from helpers import load_gold
from oracle import functional as Fn
from oracle import synth

pytestmark = pytest.mark.gpu


def test_text_shifting_n_layer_on_gpu_matches_reference_golden():
    """mmtr.py:249-273 TextShiftingNLayer (the `hybrid=True` head): N = 5 inputs, golden from the reference module"""
    import bpmult_b200.modules as M
    g = load_gold("modules.pt")["gmu"]
    D, rows = g["dims"]
    n_in = 5
    m = M.TextShiftingNLayer([D] * n_in, D)
    shp = {"hiddens.%d.weight" % i: (D, D) for i in range(n_in)}
    shp.update({"x_gates.%d.weight" % i: (D, n_in * D) for i in range(n_in)})
    m.load_state_dict(synth.make_state_dict(shp, g["seed"] + 20))
    m.cuda()
    xs = [synth.randn((rows, D), g["seed"] + i).cuda().requires_grad_() for i in range(n_in)]
    o, z = m(*xs)
    (o * synth.randn(o.shape, g["seed"] + 9).cuda()).sum().backward()
    torch.cuda.synchronize()
    r = g["tsN"]
    assert Fn.max_rel(o.detach().cpu(), r["out"]) < 2e-5 and Fn.max_rel(z.cpu(), r["z"]) < 2e-5
    for a, b in zip(xs, r["dx"]):
        assert Fn.max_rel(a.grad.cpu(), b) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad.cpu(), r["pgrads"][n]) < 5e-5, n
    with pytest.raises(AssertionError):
        m(*xs[:3])


@pytest.mark.parametrize("n_in", [3, 4])
def test_text_shifting_3_4_on_gpu_match_reference_golden(n_in):
    import bpmult_b200.modules as M
    g = load_gold("modules.pt")["gmu"]
    D, rows = g["dims"]
    cls = M.TextShifting3Layer if n_in == 3 else M.TextShifting4Layer
    m = cls(D, D, D, D) if n_in == 3 else cls(D, D, D, D, D)
    shp = {}
    for i in range(n_in):
        shp["hidden%d.weight" % (i + 1)] = (D, D)
    for i in range(n_in):
        shp["x%d_gate.weight" % (i + 1)] = (D, n_in * D)
    m.load_state_dict(synth.make_state_dict(shp, g["seed"] + n_in))
    m.cuda()
    xs = [synth.randn((rows, D), g["seed"] + i).cuda().requires_grad_() for i in range(n_in)]
    o, z = m(xs)
    (o * synth.randn(o.shape, g["seed"] + 9).cuda()).sum().backward()
    r = g["ts%d" % n_in]
    assert Fn.max_rel(o.detach().cpu(), r["out"]) < 2e-5 and Fn.max_rel(z.cpu(), r["z"]) < 2e-5
    for a, b in zip(xs, r["dx"]):
        assert Fn.max_rel(a.grad.cpu(), b) < 5e-5
    for n, p in m.named_parameters():
        assert Fn.rel_l2(p.grad.cpu(), r["pgrads"][n]) < 5e-5, n
