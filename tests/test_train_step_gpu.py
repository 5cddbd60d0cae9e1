"""The training step that hosts the path (main.py:845-871 recipe) — CUDA-graph replay and data-parallel equivalence.

* a step replayed from the per-gate-combination CUDA graphs equals the same step launched eagerly (same seeds);
* SURVEY.md §4 "distributed": a 2-rank step (one flat-buffer all-reduce; gloo over CUDA tensors so that both ranks can
  share the test box's single GPU) equals the single-process computation on the same per-rank shards — every MRFP / BN
  statistic is local to the rank's shard (parity is defined per shard), only the gradients are averaged.
"""
import os
import random
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 128
NB = 2               # per-rank batch (>= 2: NP+ statistics need two samples, deepv3.py:272)
LR = 1e-2


def _data(rank):
    g = torch.Generator().manual_seed(1000 + rank)
    img = (torch.rand(NB, 3, H, W, generator=g) * 255.0).cuda()
    lab = torch.randint(0, 19, (NB, H, W), generator=g).cuda()
    return img, lab


def _build(rank, math_mode, init_state=None):
    """Model + optimiser of one rank, RNG streams as bench.py's train leg: torch per rank, gates shared."""
    from mrfp_b200 import dist as D
    from mrfp_b200.model import MRFPPlus
    D.seed_rank_streams(3, rank)
    model = MRFPPlus(19, criterion=torch.nn.CrossEntropyLoss(ignore_index=255), math_mode=math_mode).cuda()
    if init_state is not None:
        model.load_state_dict(init_state)
    random.seed(100)
    opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=LR, momentum=0.9, weight_decay=5e-4)
    return model, opt


def _trainable(model):
    return torch.cat([p.detach().flatten() for p in model.parameters() if p.requires_grad])


@pytest.mark.parametrize("math_mode", [0, 2])
def test_graph_replay_equals_eager_steps(math_mode):
    from mrfp_b200.train_step import GraphedTrainStep
    torch.backends.cudnn.benchmark = False
    img, lab = _data(0)
    results = []
    for graphs in (False, True):
        model, opt = _build(0, math_mode)
        step = GraphedTrainStep(model, opt, img, lab, eager_steps=1, use_graphs=graphs)
        losses = []
        for i in range(14):                     # 8 combinations x (1 eager + replays): most steps of the second run replay
            losses.append(step(img, lab))
        torch.cuda.synchronize()
        results.append((torch.stack([l.detach() for l in losses]).cpu(), _trainable(model).cpu(), len(step.captured())))
    (l_e, p_e, n_e), (l_g, p_g, n_g) = results
    assert n_e == 0 and n_g >= 3, (n_e, n_g)
    assert torch.isfinite(l_g).all()
    # same gates, same draws (the CUDA generator advances identically under replay), same data: the two runs differ by
    # the order of the atomics in the BN statistics only
    tol = 2e-3 if math_mode == 0 else 5e-2
    assert torch.allclose(l_e, l_g, rtol=tol, atol=tol), (l_e, l_g)
    rel = float((p_e - p_g).norm() / p_e.norm())
    assert rel <= (1e-4 if math_mode == 0 else 2e-3), rel


def _worker(rank, world, port, init_path, out_path, use_graphs):
    import torch.distributed as dist
    from mrfp_b200.train_step import GraphedTrainStep
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.backends.cudnn.benchmark = False
    model, opt = _build(rank, 0, torch.load(init_path))
    img, lab = _data(rank)
    step = GraphedTrainStep(model, opt, img, lab, eager_steps=1, use_graphs=use_graphs)
    n_steps = 3 if use_graphs else 1
    for _ in range(n_steps):
        loss = step(img, lab, ) if not use_graphs else step.step_with_gates(img, lab, (0.25, 0.25, 0.25))
    torch.cuda.synchronize()
    torch.save({"params": _trainable(model).cpu(), "loss": float(loss)}, out_path + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_graphs", [False, True])
def test_two_rank_step_equals_single_process_on_the_same_shards(use_graphs):
    import torch.multiprocessing as mp
    from mrfp_b200.train_step import GraphedTrainStep
    torch.backends.cudnn.benchmark = False
    world = 2
    tmp = tempfile.mkdtemp()
    init_path, out_path = os.path.join(tmp, "init.pt"), os.path.join(tmp, "out.pt")
    model0, _ = _build(0, 0)
    torch.save(model0.state_dict(), init_path)
    port = 29500 + (os.getpid() % 1000)
    mp.spawn(_worker, args=(world, port, init_path, out_path, use_graphs), nprocs=world, join=True)
    got = [torch.load(out_path + f".{r}") for r in range(world)]
    # both ranks hold the same parameters after the step(s)
    assert torch.allclose(got[0]["params"], got[1]["params"], rtol=0, atol=1e-6)
    if use_graphs:
        return      # (multi-step replay: equality across ranks is the property; the one-step arithmetic is checked below)
    # single process: per-shard gradients with each rank's own streams, averaged, one SGD step
    grads, p0 = [], None
    for r in range(world):
        model, opt = _build(r, 0, torch.load(init_path))
        img, lab = _data(r)
        gates = (random.random(), random.random(), random.random())        # seed 100: what every rank drew
        loss = model(img, lab, training=True, gates=gates)
        loss.backward()
        grads.append(torch.cat([p.grad.detach().flatten() for p in model.parameters() if p.requires_grad]).cpu())
        p0 = _trainable(model).cpu()
        assert abs(float(loss) - got[r]["loss"]) <= 1e-3 * abs(float(loss)) + 1e-5      # the rank's loss is its shard's loss
    g = sum(grads) / world
    expect = p0 - LR * (g + 5e-4 * p0)                                         # first SGD step: momentum buffer = gradient
    rel = float((got[0]["params"] - expect).norm() / (expect - p0).norm())
    assert rel <= 2e-3, rel
