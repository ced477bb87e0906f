#!/usr/bin/env python
"""bench.py -- B-cos ResNet-50 forward + explanation throughput (img/s @224) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU arithmetic (oracle port)

One "step" = forward + explanation maps of `--batch` (default 256) synthetic images per GPU.  The batch is sharded
across ranks, there is no collective on this path: scaling = "weak".
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _claim_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries print there too (NCCL's "NCCL version ..." banner at
    communicator creation when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the
    result line is written to a private duplicate of the original stdout."""
    sys.stdout.flush()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return out


def _host_threads() -> int:
    """Threads for the CPU arm: the physical cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to every rank;
    the reference arm runs on rank 0 alone and takes all the cores."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        import psutil
        phys = psutil.cpu_count(logical=False)
        if phys:
            n = min(n, phys)
    except Exception:
        pass
    return max(1, n)


METRIC = "B-cos RN50 fwd+explain img/s @224 at 1/2/4/8 B200; BcosConv tensor-pipe % peak"
WORKLOAD = "B-cosified ResNet-50 forward + explanation maps, batch 256 per GPU bf16, 1/2/4/8 B200"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arch", default="resnet50")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--planes", type=int, default=1, help="precision planes (1 = 16-bit throughput mode)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"],
                    help="operand format of the tensor-core launches (BASELINE names bf16; fp16 runs at the same rate)")
    ap.add_argument("--cpu-batch", type=int, default=32, help="images per CPU-baseline / reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-table", default=None, help="write the per-launch timing table (JSON) here")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"], bf16_tflops_sustained=p.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, windows):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(args, steps, warmup):
    """The reference's own arithmetic on the host cores: the oracle port (the reference is pure Python on ATen and
    cannot travel to the GPU box; oracle/bcos_oracle.py is pinned bit-exactly against it by oracle/make_golden.py)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bcos_oracle as OR
    from bcos_b200.models import resnet_state_shapes
    from bcos_b200.utils import synth
    if torch.get_num_threads() < _host_threads():
        torch.set_num_threads(_host_threads())
    sd = synth.synthetic_checkpoint(args.arch, resnet_state_shapes(args.arch))
    x6 = synth.to_bcos_input(synth.synth_images_u8(args.cpu_batch, 224, 7))
    model = OR.OracleResNet(args.arch, sd)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        OR.explain_batched(model.forward, x6)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=args.cpu_batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=torch.get_num_threads(),
                sample=f"{len(times)} steps x {args.cpu_batch} images (forward + batched explanation), fp32, torch CPU")


def main():
    args = parse_args()
    real_stdout = _claim_stdout()
    from bcos_b200.utils import dist as D
    rank, local_rank, world = D.env_rank()

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(args, max(1, args.steps), max(1, args.warmup))       # exactly K timed steps after W warm-ups
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arch": args.arch, "image": 224, "batch_per_step": args.cpu_batch},
            "cpu_baseline": {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }), file=real_stdout, flush=True)
        return

    import torch
    from bcos_b200 import build as bbuild
    from bcos_b200.engine import ops as O
    from bcos_b200.models import synthetic_resnet_plan
    from bcos_b200.utils import synth

    torch.cuda.set_device(local_rank)
    D.init("nccl")
    bbuild.build()

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    plan = synthetic_resnet_plan(args.arch, B, planes=args.planes, dtype=args.dtype, device=f"cuda:{local_rank}", input_u8=True,
                                seed_scale=4096.0 if args.dtype == "fp16" else 1.0)
    imgs = torch.from_numpy(synth.synth_images_u8(min(B, 64), 224, 1000 + rank))
    imgs = imgs.repeat((B + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1)[:B].contiguous()
    h_in = imgs.pin_memory()
    h_logits = torch.empty(B, plan.ncls, dtype=torch.float32).pin_memory()
    h_cmap = torch.empty(B, 224, 224, dtype=torch.float32).pin_memory()
    plan.load_input(h_in)
    plan.capture()

    barrier, max_over_ranks = D.barrier, D.max_over_ranks

    sampler = ClockSampler(local_rank)
    sampler.start()
    windows = []

    # ---------------- device-resident throughput: inputs already in HBM, K graph replays ----------------
    for _ in range(W):
        plan.replay_all()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(K):
        plan.replay_all()
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    ms_dev = max_over_ranks(e0.elapsed_time(e1))

    # ---------------- end to end through the public API: pinned host images in, logits + maps out ----------------
    # PipelinedExplainer: every step copies its own batch host->device and its own results device->host (pinned memory);
    # copies of step i+1 / i-1 overlap the compute of step i.  The timed region ends when the last result is on the host.
    from bcos_b200.engine import PipelinedExplainer
    pipe = PipelinedExplainer(plan)
    h_ins = [h_in, h_in.clone().pin_memory()]
    for i in range(W):
        pipe.submit(h_ins[i % 2])
    pipe.drain()
    barrier()
    w0 = time.time()
    e0.record()
    last = 0
    for i in range(K):
        last = pipe.submit(h_ins[i % 2])
    res_last = pipe.result(last)
    pipe.drain()
    t_host_done = time.time()
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    ms_e2e = max_over_ranks((t_host_done - w0) * 1e3)        # host clock: includes the final device->host copies
    h_logits, h_cmap = res_last["logits"], res_last["contribution_map"]
    sampler.stop()

    # ---------------- per-launch timing (eager launches, CUDA events on the launching stream) ----------------
    all_ops = plan.fwd_ops + plan.bwd_ops
    per_op = [0.0] * len(all_ops)
    reps = 3
    for rep in range(reps + 1):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(all_ops) + 1)]
        evs[0].record()
        for i, o in enumerate(all_ops):
            o.run()
            evs[i + 1].record()
        torch.cuda.synchronize()
        if rep > 0:
            for i in range(len(all_ops)):
                per_op[i] += evs[i].elapsed_time(evs[i + 1]) / reps
    ig = [(o, t) for o, t in zip(all_ops, per_op) if isinstance(o, O.IgemmOp)]
    ig_ms = sum(t for _, t in ig)
    ig_bytes = sum(o.algo_bytes() for o, _ in ig)
    ig_flops = sum(o.algo_flops for o, _ in ig)
    step_ms_eager = sum(per_op)

    if rank != 0:
        D.shutdown()
        return

    pk = peaks()
    # measured DRAM bytes of the same launches from the committed ncu capture (dram__bytes_read + dram__bytes_write over
    # one step of this workload); only meaningful for the configuration it was captured on
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_final_dram_traffic.json")
    if os.path.exists(tpath) and args.arch == "resnet50" and B == 256 and args.planes == 1:
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic, traffic_src = float(tj["igemm_dram_bytes_per_step"]), "profiles/r01_final_dram_traffic.json (ncu, bytes per step)"
    hbm_ach = ig_bytes / (ig_ms * 1e-3) / 1e9
    tc_ach = ig_flops / (ig_ms * 1e-3) / 1e12
    tpk = pk["bf16_tflops_sustained"] or pk["bf16_tflops"]
    imgs_total = world * B * K
    res = {
        "metric": METRIC, "value": imgs_total / (ms_dev * 1e-3), "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "arch": args.arch, "image": 224, "batch_per_gpu": B, "precision_planes": args.planes,
                   "parallelism": f"batch-sharded x{world}, no collective", "weights": "random-init synthetic checkpoint, BN calibrated",
                   "l2": "per-layer activations (>=100 MB at batch 256) exceed the 126 MB L2; no explicit flush",
                   "cuda_graph": True},
        "e2e": {"value": imgs_total / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": int(h_in.numel() * h_in.element_size()),
                "d2h_bytes_per_step": int(h_logits.numel() * 4 + h_cmap.numel() * 4),
                "api": "PipelinedExplainer.submit(pinned uint8 images) / .result(): per-step H2D + CUDA-graph replay + D2H of logits and "
                       "contribution maps, copies overlapped with the neighbouring steps' compute"},
        "gpu_launches": plan.num_launches() * K,
        "launches_per_step": plan.num_launches(),
        "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "bcosk_igemm_* (all conv / dgrad launches of one step)",
                     "peak_source": pk["source"], "kernel_ms_per_step": ig_ms, "kernel_share_of_step": ig_ms / step_ms_eager,
                     "algorithmic_bytes_per_step": ig_bytes},
        "roofline_tensor": {"bound": "tensor", "achieved": tc_ach, "peak": tpk, "unit": "TFLOP/s", "frac": tc_ach / tpk,
                            "algorithmic_flops_per_step": ig_flops, "peak_source": pk["source"] + " sustained"},
        "clocks": sampler.summary(windows),
    }
    if args.layer_table:
        rows = []
        for o, t in zip(all_ops, per_op):
            row = {"name": o.name, "kind": type(o).__name__, "ms": t}
            if isinstance(o, O.IgemmOp):
                row.update(M=o.M, N=o.n, K=o.ktot, block_n=o.resolved_block_n(), gflop=o.algo_flops / 1e9,
                           mbytes=o.algo_bytes() / 1e6, tflops=o.algo_flops / (t * 1e-3) / 1e12,
                           gbs=o.algo_bytes() / (t * 1e-3) / 1e9)
            rows.append(row)
        os.makedirs(os.path.dirname(os.path.abspath(args.layer_table)), exist_ok=True)
        with open(args.layer_table, "w") as fh:
            json.dump({"batch": B, "arch": args.arch, "planes": args.planes, "step_ms_eager": step_ms_eager, "rows": rows}, fh, indent=1)
    if world == 1 and not args.no_cpu_baseline:
        c = cpu_reference_arm(args, 16, 1)     # ~10 s of host work: 16 steps of 32 images
        res["cpu_baseline"] = {"value": c["value"], "unit": "img/s", "cores": c["cores"], "kind": "port", "sample": c["sample"]}
    print(json.dumps(res), file=real_stdout, flush=True)
    D.shutdown()


if __name__ == "__main__":
    main()
