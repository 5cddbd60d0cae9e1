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


def _pin_host_arithmetic():
    """The host trunk on deterministic fp32 cuDNN algorithms: at this tiny size (4x4 maps at layer4, BN over 32 values)
    the trunk amplifies TF32 / algorithm-choice noise to 1e-3 of the loss, which would mask what these tests compare."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


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
@pytest.mark.parametrize("gates", [(0.75, 0.75, 0.25), (0.75, 0.75, 0.75)])
def test_graph_replay_equals_the_eager_step(gates, math_mode):
    """Gate combinations that consume no random numbers inside the step (p >= .5: no HRFP re-initialisation, p2 >= .5: no
    NP+ draws; HRFP+ on / off), learning rate 0: the eager step, the first replay and a later replay of the captured
    graph all evaluate the same function of the same parameters — same loss, same flat gradient up to the order of the
    floating-point atomics (cuDNN pinned to deterministic fp32 algorithms so that the host trunk cannot differ)."""
    from mrfp_b200.train_step import GraphedTrainStep
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    try:
        img, lab = _data(0)
        model, _ = _build(0, math_mode)
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=0.0)
        step = GraphedTrainStep(model, opt, img, lab, eager_steps=1, use_graphs=True)
        runs = []
        for i in range(4):
            loss = step.step_with_gates(img, lab, gates)
            torch.cuda.synchronize()
            runs.append((float(loss), step.flat_grad.clone()))
        assert len(step.captured()) == 1                       # step 0 eager, step 1 capture + replay, steps 2-3 replays
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old
    l0, g0 = runs[0]
    assert g0.abs().sum() > 0
    for i, (l, g) in enumerate(runs[1:], 1):
        rel = float((g - g0).norm() / g0.norm())
        tol = 1e-4 if math_mode == 0 else 5e-2                 # bf16 chain: statistics atomics flip roundings (and ReLU masks)
        assert abs(l - l0) <= 1e-4 * abs(l0) + (0 if math_mode == 0 else 1e-3), (i, l, l0)
        assert rel <= tol, (i, rel)


def test_replayed_graph_advances_the_random_streams():
    """All three gates on: every replay re-draws the HRFP weights (deepv3.py:290-306) and the NP+ factors through torch's
    CUDA generator — consecutive replays of the SAME graph must see different draws (Philox offset advancing)."""
    from mrfp_b200.train_step import GraphedTrainStep
    torch.backends.cudnn.benchmark = False
    img, lab = _data(0)
    model, opt = _build(0, 2)
    step = GraphedTrainStep(model, opt, img, lab, eager_steps=1, use_graphs=True)
    gammas, losses = [], []
    for _ in range(5):
        losses.append(float(step.step_with_gates(img, lab, (0.25, 0.25, 0.25))))
        gammas.append(model.OC1_bn.weight.detach().clone())
    assert step.captured() == [(True, True, True)]
    assert all(l == l and abs(l) < 1e3 for l in losses), losses
    for a, b in zip(gammas[1:-1], gammas[2:]):          # steps 2.. are replays
        assert not torch.equal(a, b)
    assert abs(float(torch.stack(gammas).std()) - 0.5) < 0.1          # gamma ~ N(0, 0.5), mynn.py:73


def _worker(rank, world, port, init_path, out_path, use_graphs):
    import torch.distributed as dist
    from mrfp_b200.train_step import GraphedTrainStep
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    _pin_host_arithmetic()
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
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    _pin_host_arithmetic()
    try:
        _two_rank_body(use_graphs)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = old


def _two_rank_body(use_graphs):
    import torch.multiprocessing as mp
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
