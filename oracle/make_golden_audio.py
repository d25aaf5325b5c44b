"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the AudioEncoder (reference models/mmtr.py:93-108, used at :307,452), generated from the
UNMODIFIED reference class (oracle/ref_shim.py keeps it as `AudioEncoderReal`; the shim's feature bypass is swapped out again here):

  audio_encoder   the module alone: (2, 96, 900) spectrogram -> (2, 96, 200), output + parameter-gradient fingerprints
  mmtrvapt_audio  the 4-modality model with its real AudioEncoder upstream of the trunk, raw audio (2, 96, 700)

Asserts restatement (oracle/functional.py) == reference and writes tests/golden/audio_encoder.pt.
Run:  python oracle/make_golden_audio.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import functional as Fn  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_audio_encoder(ref, seed):
    sd = synth.make_state_dict(synth.audio_encoder_shapes(96), seed)
    m = ref.AudioEncoderReal(None)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    x = synth.randn((2, 96, 900), seed + 1)
    g = synth.randn((2, 96, 200), seed + 2)
    y = m(x)
    (y * g).sum().backward()
    rgr = {n: p.grad.clone() for n, p in m.named_parameters()}
    sdo = {k: v.clone().requires_grad_() for k, v in sd.items()}
    y2 = Fn.audio_encoder(sdo, "", x)
    (y2 * g).sum().backward()
    assert Fn.max_rel(y2, y) < 1e-6
    for n in rgr:
        assert Fn.rel_l2(sdo[n].grad, rgr[n]) < 1e-5, n
    print("AudioEncoder ok: out", tuple(y.shape))
    return dict(seed=seed, out=y.detach(), pgrad_fp={n: synth.summarize(v) for n, v in rgr.items()})


def run_model(ref, seed):
    cfg = synth.tiny_cfg(layers=1, n_classes=13, orig_d_p=48, orig_d_a=96, audio_encoder=True)
    shapes = synth.mmtrvapt_shapes(cfg)
    shapes.update(synth.audio_encoder_shapes(96, "audio_enc."))
    sd = synth.make_state_dict(shapes, seed)
    m = ref.mmtr.MultiprojectionMMTransformerGMUClf(cfg)
    m.audio_enc = ref.AudioEncoderReal(cfg)                              # undo the shim's bypass: the model as the reference ships it
    ref_keys = {k for k in m.state_dict().keys() if not (k.endswith(".version") or k.endswith("_float_tensor"))}
    assert ref_keys == set(shapes.keys()), sorted(ref_keys ^ set(shapes.keys()))[:10]
    m.load_state_dict(sd, strict=False)
    m.train()
    B = 2
    txt, img, _, poster, tgt = synth.mmtrvapt_inputs(cfg, B, 20, 30, 25)
    audio = synth.randn((B, 96, 700), seed + 5)
    pw = torch.linspace(0.5, 2.0, cfg.n_classes)
    txt_r = txt.clone().requires_grad_()
    logits, z = m(txt_r, None, None, img, audio, poster, output_gate=True)
    loss = torch.nn.BCEWithLogitsLoss(pos_weight=pw)(logits, tgt)
    loss.backward()
    rgr = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    sdo = {k: v.clone().requires_grad_() for k, v in sd.items()}
    l2, z2 = Fn.mmtrvapt_forward(sdo, cfg, txt, img, audio, poster)
    Fn.bce_with_logits(l2, tgt, pw).backward()
    assert Fn.max_rel(l2, logits) < 2e-5 and Fn.max_rel(z2, z) < 2e-5
    worst = max(Fn.rel_l2(sdo[n].grad, rgr[n]) for n in rgr)
    assert worst < 1e-4, worst
    print("mmtrvapt + AudioEncoder ok: worst grad rel-l2 %.2e, %d gradient tensors" % (worst, len(rgr)))
    return dict(cfg=vars(cfg), dims=(B, 20, 700, 25), seed=seed, pos_weight=pw, logits=logits.detach(), z=z.detach(), loss=loss.detach(),
                dtxt=txt_r.grad, pgrad_fp={n: synth.summarize(g) for n, g in rgr.items()})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = load_reference()
    assert ref is not None, "reference tree not found"
    ref.mmtr.AudioEncoder = ref.AudioEncoderReal                          # (the class body looks itself up by its module-level name)
    torch.save(dict(audio_encoder=run_audio_encoder(ref, 91), mmtrvapt_audio=run_model(ref, 92)), os.path.join(OUT, "audio_encoder.pt"))
    print("audio_encoder.pt", os.path.getsize(os.path.join(OUT, "audio_encoder.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
