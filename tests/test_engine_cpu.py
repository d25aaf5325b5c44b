"""Host-logic tests (no GPU): the product engines (kernel sequencing, padded layout, hand-derived backward) driven by
the pure-torch op emulation in tests/emu_ops.py must reproduce the golden vectors generated from the reference."""
import pytest
import torch

from emu_ops import EmuOps
from helpers import load_gold, run_encoder_engine, run_model_engine
from oracle import functional as Fn
from oracle import synth

ENC = load_gold("encoder.pt")


@pytest.mark.parametrize("rec", ENC, ids=[r["case"][0] for r in ENC])
def test_encoder_engine_fp32_matches_reference_golden(rec):
    name, T, S, B, D, H, L, bi, mask, self_only, zt = rec["case"]
    sd = synth.make_state_dict(synth.encoder_shapes(D, L, bi), rec["seed"])
    x = synth.randn((T, B, D), rec["seed"] + 100)
    k = synth.randn((S, B, D), rec["seed"] + 101)
    g = synth.randn((T, B, D), rec["seed"] + 102)
    if zt:
        x[T - zt:] = 0
        k[S - zt:] = 0
    out, dx, dk, grads, _ = run_encoder_engine(EmuOps(), sd, x, None if self_only else k, g, H, L, mask, bi, self_only)
    assert Fn.max_rel(out, rec["out"]) < 2e-5
    assert Fn.max_rel(dx, rec["dx"]) < 5e-5
    if not self_only:
        assert Fn.max_rel(dk, rec["dk"]) < 5e-5
    if "pgrads" in rec:
        for n, ref in rec["pgrads"].items():
            assert Fn.rel_l2(grads[n], ref) < 5e-5, n
    else:
        for n, s in rec["pgrad_summ"].items():
            f = grads[n].reshape(-1).double()
            assert abs(f.norm().item() - s["norm"]) <= 5e-5 * s["norm"] + 1e-9, n
            assert torch.allclose(f[s["idx"]].float(), s["val"], rtol=2e-3, atol=2e-5 * max(1e-6, s["norm"])), n


def test_mmtrvat_engine_fp32_matches_reference_golden():
    rec = load_gold("mmtrvat_tiny.pt")
    logits, z, loss, dtxt, grads, eng = run_model_engine(EmuOps(), rec)
    assert Fn.max_rel(logits, rec["logits"]) < 2e-5
    assert Fn.max_rel(z, rec["z"]) < 2e-5
    assert abs(loss.item() - rec["loss"].item()) < 1e-5
    assert Fn.max_rel(dtxt, rec["dtxt"]) < 1e-4
    assert sorted(eng.unused_params()) == sorted(rec["nograd"])
    worst = 0.0
    for n, ref in rec["pgrads"].items():
        e = Fn.rel_l2(grads[n], ref)
        worst = max(worst, e)
        assert e < 2e-4, (n, e)
    print("worst param-grad rel-l2", worst)
