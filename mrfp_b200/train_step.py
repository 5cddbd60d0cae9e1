"""One optimisation step of the drop-in MRFP+ model as CUDA graphs (host-side mirror of the reference's training loop,
/root/reference/main.py:845-871: zero_grad -> loss = net(images, gts) -> loss.backward() -> optimizer.step()).

Why: at the per-GPU batches of BASELINE config[2] (global batch 16 over 8 GPUs = 2 images per GPU) a step is ~890 kernel
launches of 5-20 us each; eager PyTorch needs ~15 ms of host time to enqueue them while the GPU needs ~19 ms to run
them, and with 8 ranks sharing the host's cores the step becomes launch-bound (round-1 measurement: 34.5 ms per step at
8 GPUs against 19.7 ms for the same per-GPU work on one GPU).  A replayed graph needs ~0.1 ms of host time per step.

What is captured: forward + backward of the local replica for ONE combination of the three MRFP gates
(p < .5, p2 < .5, p3 < .5 — deepv3.py:281-283).  The gates are drawn on the host with `random.random()` exactly as the
reference's forward does, then the graph of that combination is replayed: up to 8 graphs, captured lazily and sharing
one memory pool.  The first `eager_steps` occurrences of a combination run eagerly (they ARE training steps: cuDNN picks
its algorithms, the library sizes its scratch buffers); capture itself executes nothing, so the sequence of parameter
updates is the same as in eager mode.  Random draws inside the step (HRFP re-initialisation, NP+ factors) go through
torch's CUDA generator, which CUDA graphs replay with advancing Philox offsets.

Data parallelism: gradients are written into ONE flat buffer (every `param.grad` is a view of it), so the
only collective of a step is a single all-reduce (AVG) of that buffer over NCCL between the graph replay and the
optimiser step — the same result as DistributedDataParallel's bucketed all-reduce; BatchNorm statistics and the MRFP
statistics stay per replica, as under the reference's nn.DataParallel scatter (main.py:824).
"""
import random
from typing import Optional

import torch
import torch.distributed as dist


class GraphedTrainStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, example_images: torch.Tensor,
                 example_labels: torch.Tensor, eager_steps: int = 2, process_group: Optional[dist.ProcessGroup] = None,
                 use_graphs: bool = True):
        self.model, self.opt = model, optimizer
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = example_images.device
        self.images = torch.empty_like(example_images)
        self.labels = torch.empty_like(example_labels)
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:                       # every .grad is a window of the flat buffer: one collective per step
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.eager_steps = max(1, eager_steps)
        self.use_graphs = use_graphs
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.stream = torch.cuda.Stream(device=dev)   # eager warm-up and capture share it (stream-private scratch buffers)
        self.pool = torch.cuda.graph_pool_handle() if use_graphs else None
        self.graphs = {}          # combo -> (CUDAGraph, static loss)
        self.seen = {}            # combo -> number of eager steps so far
        self.loss = None

    # ---- one forward + backward of the local replica with fixed gates, gradients accumulated into flat_grad ----
    def _fwd_bwd(self, gates):
        self.flat_grad.zero_()
        loss = self.model(self.images, self.labels, training=True, gates=gates)
        loss.backward()
        return loss

    def __call__(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Runs one training step on (images, labels); returns the loss tensor (device, no sync)."""
        gates = (random.random(), random.random(), random.random())             # deepv3.py:281-283, same order
        return self.step_with_gates(images, labels, gates)

    def warm_all(self, images: torch.Tensor, labels: torch.Tensor):
        """Runs every gate combination until its graph exists (eager_steps eager steps + the capturing one each).  These
        are real optimisation steps; a benchmark calls this in its untimed warm-up so that the timed steps replay."""
        for mask in range(8):
            combo = (bool(mask & 1), bool(mask & 2), bool(mask & 4))
            gates = tuple(0.25 if c else 0.75 for c in combo)
            for _ in range(self.eager_steps + (1 if self.use_graphs else 0)):
                self.step_with_gates(images, labels, gates)

    def step_with_gates(self, images: torch.Tensor, labels: torch.Tensor, gates) -> torch.Tensor:
        combo = tuple(g < 0.5 for g in gates)
        fixed = tuple(0.25 if c else 0.75 for c in combo)                        # any value on the same side of 0.5
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.images.copy_(images, non_blocking=True)
            self.labels.copy_(labels, non_blocking=True)
            entry = self.graphs.get(combo)
            if entry is None and self.use_graphs and self.seen.get(combo, 0) >= self.eager_steps:
                g = torch.cuda.CUDAGraph()
                # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
                with torch.cuda.graph(g, pool=self.pool, stream=self.stream, capture_error_mode="thread_local"):
                    static_loss = self._fwd_bwd(fixed)
                entry = self.graphs[combo] = (g, static_loss)
            if entry is not None:
                entry[0].replay()
                # the graphs share one memory pool: a later replay of ANOTHER graph may use this graph's static output
                # as scratch, so the loss leaves the pool right behind the replay (stream-ordered)
                loss = entry[1].clone()
            else:
                loss = self._fwd_bwd(fixed)
                self.seen[combo] = self.seen.get(combo, 0) + 1
            if self.world > 1:
                if dist.get_backend(self.pg) == "nccl":
                    dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG, group=self.pg)
                else:                                   # gloo (CPU-side tests): no AVG
                    dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
                    self.flat_grad.div_(self.world)
            self.opt.step()
        cur.wait_stream(self.stream)
        self.loss = loss
        return loss

    def captured(self):
        return sorted(self.graphs)
