"""N > 1 host logic on CPU: two gloo ranks, each with half of the batch, must end a Trainer step with identical parameters that
equal the single-process full-batch step (equal shards: mean of local means = global mean; gradients all-reduced per bucket in
backward-completion order and divided by the world size inside the Adam launch).  Device ops are the test-only emulation."""
import os
import tempfile
from argparse import Namespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth


def _build(cfg):
    import bpmult_b200.modules as M
    from emu_ops import EmuOps
    o = EmuOps()
    M._ops_for = lambda device: o
    m = M.MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision="fp32")
    m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5), strict=False)
    return m.train()


def _worker(rank, world, initfile, out, hybrid=False):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", init_method="file://" + initfile, rank=rank, world_size=world)
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1, hybrid=hybrid)
    m = _build(cfg)
    tr = Trainer(m, lr=1e-2, use_graph=False)
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, 4, 8, 12, 10)
    sl = slice(rank * 2, rank * 2 + 2)
    losses = [tr.step(txt[sl], img[sl], audio[sl], tgt[sl]) for _ in range(2)]
    torch.save(dict(p=tr.flat_p.clone(), losses=losses, buckets=tr.buckets), out % rank)
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("hybrid", [False, True], ids=["plain", "hybrid"])
def test_two_gloo_ranks_match_full_batch_step(hybrid):
    with tempfile.TemporaryDirectory() as d:
        initfile, out = os.path.join(d, "init"), os.path.join(d, "r%d.pt")
        mp.spawn(_worker, args=(2, initfile, out, hybrid), nprocs=2, join=True)
        r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    assert torch.equal(r0["p"], r1["p"])                          # replicas stay bit-identical
    # single process, full batch
    from bpmult_b200.trainer import Trainer
    cfg = synth.tiny_cfg(layers=1, hybrid=hybrid)
    tr = Trainer(_build(cfg), lr=1e-2, use_graph=False)
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, 4, 8, 12, 10)
    losses = [tr.step(txt, img, audio, tgt) for _ in range(2)]
    assert abs(0.5 * (r0["losses"][0] + r1["losses"][0]) - losses[0]) < 1e-6
    rel = ((tr.flat_p - r0["p"]).double().norm() / tr.flat_p.double().norm()).item()
    assert rel < 2e-4, rel
    # bucket order = backward completion order: wave-2 encoders first, misc last, contiguous cover of the flat buffer
    names = [b[0] for b in r0["buckets"]]
    if hybrid:                                                    # the early-fusion stacks finish their backward first
        assert names[:3] == ["a_early", "v_early", "l_early"]
        names = names[3:]
    assert names[0] in ("a_with_l2v", "a_with_v2l") and names[-1] == "misc" and len(names) == 13
    nb = len(r0["buckets"])
    assert all(r0["buckets"][i][2] == r0["buckets"][i + 1][1] for i in range(nb - 1))
