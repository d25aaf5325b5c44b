"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference (`/root/reference`, or `$BPMULT_REF`, or `baseline/_ref`)
with the five monkey-patch shims it needs to run on torch 2.11 / CPU:

  shim 1  missing packages `detectron2*`, `pytorch_pretrained_bert*`      (models/image.py:15-20, mmtr.py:6)
  shim 2  non-contiguous `.view(-1)` in the positional embedding, B > 1   (position_embedding.py:27,76)
  shim 3  in-place `q *= scaling` on a chunk() view in the self-attn path (multihead_attention.py:72,86,138)
  shim 4  hard-coded `.cuda()` in transfm_2dim                            (mmtr.py:434-438,725-729)
  shim 5  TextShifting3Layer called with 4 ctor args instead of 5         (mmtr.py:199 vs :663)
  shim 6  hybrid branch of mmtrvat: `gmu_early(a, b, c)` on a forward(xs: list) and `gmu([a, b, c, d])` on a forward(*xs)
          (mmtr.py:775 vs :211, :855 vs :258): both gates accept either calling convention; nothing else changes
  bypass  BertEncoder -> identity on float features                       (mmtr.py:144-158)

Only `tests/`, `oracle/make_golden.py` and `bench.py --impl reference` may use this.
"""
import os
import sys
from argparse import Namespace
from unittest.mock import MagicMock

import torch

_LOADED = None


def find_reference():
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("BPMULT_REF"), "/root/reference",
                 os.path.join(here, "..", "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "bpmult", "models")):
            return os.path.abspath(cand)
    return None


def load_reference(patch_cuda=True):
    """Returns a namespace with the shimmed reference modules, or None when no reference tree exists."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    root = find_reference()
    if root is None:
        return None
    for n in ["detectron2", "detectron2.model_zoo", "detectron2.modeling", "detectron2.config",
              "detectron2.checkpoint", "detectron2.structures", "detectron2.structures.image_list",
              "pytorch_pretrained_bert", "pytorch_pretrained_bert.modeling"]:
        sys.modules.setdefault(n, MagicMock())                                  # shim 1
    if root not in sys.path:
        sys.path.insert(0, root)
    import bpmult.models.position_embedding as pe
    import bpmult.models.multihead_attention as mha
    import bpmult.models.mmtr as mmtr
    import bpmult.models.transformer as tr

    _pe_fwd = pe.SinusoidalPositionalEmbedding.forward                           # shim 2
    pe.SinusoidalPositionalEmbedding.forward = lambda self, inp: _pe_fwd(self, inp.contiguous())

    _qkv = mha.MultiheadAttention.in_proj_qkv                                    # shim 3
    mha.MultiheadAttention.in_proj_qkv = lambda self, q: tuple(t.clone() for t in _qkv(self, q))

    if patch_cuda and not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self                           # shim 4 (CPU runs only)

    _ts3_init = mmtr.TextShifting3Layer.__init__                                 # shim 5

    def _ts3_fixed(self, a, b, c, d, e=None):
        if e is None:
            d, e = 0, d
        _ts3_init(self, a, b, c, d, e)
    mmtr.TextShifting3Layer.__init__ = _ts3_fixed

    _ts3_fwd, _tsn_fwd = mmtr.TextShifting3Layer.forward, mmtr.TextShiftingNLayer.forward     # shim 6
    mmtr.TextShifting3Layer.forward = lambda self, *xs: _ts3_fwd(self, list(xs[0]) if len(xs) == 1 else list(xs))
    mmtr.TextShiftingNLayer.forward = lambda self, *xs: _tsn_fwd(self, *(xs[0] if len(xs) == 1 and isinstance(xs[0], (list, tuple)) else xs))

    class FeatEnc(torch.nn.Module):                                              # BERT bypass
        def __init__(self, args):
            super().__init__()

        def forward(self, txt, mask, segment):
            return txt
    mmtr.BertEncoder = FeatEnc

    class AudioFeat(torch.nn.Module):                                            # AudioEncoder bypass (mmtr.py:452-453): audio arrives as
        def __init__(self, args):                                               # post-encoder features (B, T_a, orig_d_a)
            super().__init__()

        def forward(self, audio):
            return audio.transpose(1, 2)
    audio_real = mmtr.AudioEncoder                                               # the unmodified class, for the AudioEncoder goldens
    mmtr.AudioEncoder = AudioFeat

    _LOADED = Namespace(root=root, pe=pe, mha=mha, mmtr=mmtr, tr=tr, AudioEncoderReal=audio_real)
    return _LOADED


def mmtrvat_args(**kw):
    """The flat namespace the reference constructors read (mmtr.py:594-613); README MOSEI defaults."""
    d = dict(orig_d_l=300, orig_d_v=35, orig_d_a=74, orig_d_p=4096, hidden_sz=300, num_heads=12, layers=8,
             vonly=True, lonly=True, aonly=True, attn_mask=True, hybrid=False, n_classes=6,
             attn_dropout=0.1, attn_dropout_v=0.0, attn_dropout_a=0.0, relu_dropout=0.1, res_dropout=0.1,
             out_dropout=0.0, embed_dropout=0.25, bert_model="bert-base-uncased")
    d.update(kw)
    return Namespace(**d)


def zero_dropout(args):
    for k in ("attn_dropout", "attn_dropout_v", "attn_dropout_a", "relu_dropout", "res_dropout",
              "out_dropout", "embed_dropout"):
        setattr(args, k, 0.0)
    return args
