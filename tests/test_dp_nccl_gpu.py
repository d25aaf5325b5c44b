"""N > 1 on real hardware: two NCCL ranks (one process per GPU), each with half of the batch, must end a Trainer step with
bit-identical parameters that equal the single-GPU full-batch step (replaces nn.DataParallel of train.py:354-356).
Skipped when the box has fewer than 2 GPUs (`gpurun --gpus 2` runs it); tests/test_dp_gloo_cpu.py covers the same host logic on CPU."""
import os
import tempfile
from argparse import Namespace

import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _model(cfg, precision):
    from bpmult_b200 import MultiprojectionMMTransformer3DGMUClf
    m = MultiprojectionMMTransformer3DGMUClf(Namespace(**vars(cfg)), precision=precision)
    m.load_state_dict(synth.make_state_dict(synth.mmtrvat_shapes(cfg), 5), strict=False)
    return m.train()


def _worker(rank, world, port, out, use_graph):
    import sys
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from bpmult_b200 import Trainer
    cfg = synth.tiny_cfg(layers=1)
    m = _model(cfg, "fp32").cuda()
    if rank == 1:                                     # replicas that start different must be synchronised by the Trainer
        with torch.no_grad():
            for p in m.parameters():
                p.add_(0.5)
    tr = Trainer(m, lr=1e-2, use_graph=use_graph)
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, 4, 8, 12, 10)
    sl = slice(rank * 2, rank * 2 + 2)
    losses = [tr.step(txt[sl], img[sl], audio[sl], tgt[sl]) for _ in range(4)]
    torch.cuda.synchronize()
    torch.save(dict(p=tr.flat_p.cpu(), losses=losses), out % rank)
    tr.close()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("use_graph", [False, True], ids=["eager", "graph"])
def test_two_nccl_ranks_match_full_batch_step(use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "r%d.pt")
        mp.spawn(_worker, args=(2, 29611 + int(use_graph), out, use_graph), nprocs=2, join=True)
        r0, r1 = torch.load(out % 0, weights_only=False), torch.load(out % 1, weights_only=False)
    assert torch.equal(r0["p"], r1["p"])                          # replicas stay bit-identical (rank 1 started from other weights)
    from bpmult_b200 import Trainer
    cfg = synth.tiny_cfg(layers=1)
    tr = Trainer(_model(cfg, "fp32").cuda(), lr=1e-2, use_graph=False)
    txt, img, audio, tgt = synth.mmtrvat_inputs(cfg, 4, 8, 12, 10)
    losses = [tr.step(txt, img, audio, tgt) for _ in range(4)]
    assert abs(0.5 * (r0["losses"][0] + r1["losses"][0]) - losses[0]) < 1e-5
    rel = ((tr.flat_p.cpu() - r0["p"]).double().norm() / tr.flat_p.double().norm().cpu()).item()
    assert rel < 5e-4, rel
