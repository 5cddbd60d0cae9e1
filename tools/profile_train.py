"""Where a DeepLabV3+/MRFP+ training step spends its time at a given per-GPU batch (torch profiler / CUPTI).

    python tools/profile_train.py [--batch 2 16] [--steps 6] [--out gpurun_out/train_profile.json]

For each per-GPU batch it reports, for the natural Bernoulli gate mix of bench.py's `train` leg:
  * step time by CUDA events, and the host time needed to ENQUEUE a step (no sync inside) — when the two are equal the
    step is launch-bound, i.e. the GPU waits for the Python / driver side;
  * number of GPU kernels per step, sum of their durations (GPU-busy time) and busy fraction of the step;
  * the top kernels by time, and the share of this repo's kernels vs cuDNN/ATen ones.
Under torchrun (world > 1) it wraps the model in DDP exactly as bench.py does and also reports the NCCL kernel time.
"""
import argparse
import collections
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[2, 16])
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--trunk", default="resnet-50")
    ap.add_argument("--graphs", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "train_profile.json"))
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from torch.profiler import profile, ProfilerActivity
    from mrfp_b200.model import MRFPPlus
    from mrfp_b200 import dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    report = {"world": world, "trunk": args.trunk, "rows": []}
    for nb in args.batch:
        D.seed_rank_streams(3, rank)
        random.seed(100 + rank)
        model = MRFPPlus(19, trunk=args.trunk, criterion=torch.nn.CrossEntropyLoss(ignore_index=255)).to(dev)
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], broadcast_buffers=False)
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-2, momentum=0.9, weight_decay=5e-4)
        img = torch.rand(nb, 3, 768, 768, device=dev) * 255.0
        lab = torch.randint(0, 19, (nb, 768, 768), device=dev)
        lab[torch.rand(nb, 768, 768, device=dev) < 0.05] = 255

        def one():
            opt.zero_grad(set_to_none=True)
            loss = model(img, lab, training=True)
            loss.backward()
            opt.step()
            return loss

        for _ in range(args.warmup):
            one()
        torch.cuda.synchronize()
        # event-timed step and host enqueue time of the same steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            one()
        e1.record()
        t_enq = (time.perf_counter() - t0) / args.steps * 1e3
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(args.steps):
                one()
            torch.cuda.synchronize()
        kern = collections.OrderedDict()
        nk, busy, launches = 0, 0.0, 0
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                nm = ev.name
                dur = ev.time_range.end - ev.time_range.start
                short = nm.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:70]
                a = kern.setdefault(short, [0, 0.0])
                a[0] += 1; a[1] += dur
                nk += 1; busy += dur
            elif ev.name in ("cudaLaunchKernel", "cudaLaunchKernelExC", "cuLaunchKernel", "cuLaunchKernelEx",
                             "cudaLaunchCooperativeKernel", "cudaMemsetAsync", "cudaMemcpyAsync"):
                launches += 1
        ours = sum(v[1] for k, v in kern.items() if "mrfp::" in k)
        nccl = sum(v[1] for k, v in kern.items() if "nccl" in k.lower())
        top = sorted(kern.items(), key=lambda kv: -kv[1][1])[:25]
        row = {"per_gpu_batch": nb, "ms_per_step_events": ms, "ms_host_enqueue_per_step": t_enq,
               "img_per_s": nb * world / (ms * 1e-3), "gpu_kernels_per_step": nk / args.steps,
               "host_launch_calls_per_step": launches / args.steps,
               "gpu_busy_ms_per_step": busy / args.steps / 1e3, "gpu_busy_frac": busy / args.steps / 1e3 / ms,
               "mrfp_kernels_ms_per_step": ours / args.steps / 1e3, "nccl_ms_per_step": nccl / args.steps / 1e3,
               "top_kernels": [{"name": k, "count_per_step": v[0] / args.steps, "us_per_step": v[1] / args.steps} for k, v in top]}
        report["rows"].append(row)
        if rank == 0:
            print(json.dumps({k: v for k, v in row.items() if k != "top_kernels"}))
            for t in row["top_kernels"][:12]:
                print("   %-72s %6.1f x %9.1f us" % (t["name"], t["count_per_step"], t["us_per_step"]))
        del model, opt, img, lab
        torch.cuda.empty_cache()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
