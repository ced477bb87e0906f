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
    ap.add_argument("--mode", default="parity", choices=["parity", "parity_full", "throughput", "throughput_fp16"],
                    help="operand format (engine/resnet.py PRECISION_MODES).  parity (default): two fp16 planes + fp32-faithful "
                         "accumulation forward, one fp16 plane in the explanation pass - meets the parity contract.  throughput: one "
                         "bf16 plane, the format BASELINE.json names - faster, does not meet the map tolerances")
    ap.add_argument("--no-throughput-record", action="store_true", help="skip the extra bf16 x1 measurement")
    ap.add_argument("--no-train-record", action="store_true", help="skip the fine-tuning-step measurement (BASELINE config 5)")
    ap.add_argument("--no-densenet-record", action="store_true", help="skip the DenseNet-121 measurement")
    ap.add_argument("--no-vit-record", action="store_true", help="skip the SimpleViT measurements (BASELINE config 3)")
    ap.add_argument("--vit-batch", type=int, default=256, help="images per GPU of the SimpleViT records")
    ap.add_argument("--no-clip-record", action="store_true", help="skip the CLIP RN50 measurement (BASELINE config 4)")
    ap.add_argument("--clip-batch", type=int, default=512, help="images per GPU of the CLIP RN50 record (config 4: 512)")
    ap.add_argument("--train-batch", type=int, default=64, help="images per GPU of the fine-tuning step (reference recipe: 64)")
    ap.add_argument("--train-steps", type=int, default=8)
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per CPU-baseline / reference-arm step (8 = the CPU's best)")
    ap.add_argument("--cpu-steps", type=int, default=64, help="steps of the in-run CPU baseline (~10-20 s of host work)")
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
    """The reference's own CPU implementation of the path on the host cores.  shrebox/B-cosification is pure Python on ATen;
    its unmodified hot-path modules travel to the GPU box as the build artefact oracle/_ref/bcos_reference.zip (recipe:
    oracle/stage_ref.py) and are run through the reference's own BcosifyNetwork / explanation_mode (kind "reference").
    Without the archive (and without /root/reference) the oracle port runs instead (kind "port"; pinned bit-exactly against
    the reference by oracle/make_golden.py)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bcos_oracle as OR
    import refload
    from bcos_b200.models import resnet_state_shapes
    from bcos_b200.utils import synth
    if torch.get_num_threads() < _host_threads():
        torch.set_num_threads(_host_threads())
    sd = synth.synthetic_checkpoint(args.arch, resnet_state_shapes(args.arch))
    x6 = synth.to_bcos_input(synth.synth_images_u8(args.cpu_batch, 224, 7))
    kind = "port"
    if refload.available() and not os.environ.get("BCOS_BENCH_CPU_PORT"):
        import make_golden as G
        model = G.build_reference_resnet(args.arch)                  # the reference's bcosify.BcosifyNetwork over its ResNetBcos
        model.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        model.eval()
        step = lambda: G.reference_explain_batched(model, x6)        # noqa: E731
        kind = "reference"
        what = f"the reference's own modules ({refload.source()})"
    else:
        oracle_model = OR.OracleResNet(args.arch, sd)
        step = lambda: OR.explain_batched(oracle_model.forward, x6)  # noqa: E731
        what = "oracle port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=args.cpu_batch * len(times) / total, ms_per_step=1e3 * total / len(times), cores=torch.get_num_threads(), kind=kind,
                sample=f"{len(times)} steps x {args.cpu_batch} images (forward + batched explanation), fp32, torch CPU, {what}")


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
            "config": {"workload": WORKLOAD, "arch": args.arch, "image": 224, "batch_per_step": args.cpu_batch,
                       "note": "bounded sample of the workload: the CPU's best step size (8 images; 32-image steps are ~1.5x slower per image)"},
            "cpu_baseline": {"value": r["value"], "unit": "img/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }), file=real_stdout, flush=True)
        return

    import torch
    from bcos_b200 import build as bbuild
    from bcos_b200.engine import ops as O
    from bcos_b200.engine import PipelinedExplainer
    from bcos_b200.models import synthetic_resnet_plan
    from bcos_b200.utils import synth

    torch.cuda.set_device(local_rank)
    D.init("nccl")
    bbuild.build()
    barrier, max_over_ranks = D.barrier, D.max_over_ranks

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    dev = f"cuda:{local_rank}"
    imgs = torch.from_numpy(synth.synth_images_u8(min(B, 64), 224, 1000 + rank))
    imgs = imgs.repeat((B + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1)[:B].contiguous()
    h_in = imgs.pin_memory()

    def device_resident_ms(pl):
        """W warm-ups, then K CUDA-graph replays of forward + explanation with the inputs resident in HBM; device time, max over ranks"""
        for _ in range(W):
            pl.replay_all()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(K):
            pl.replay_all()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), (w0, time.time())

    sampler = ClockSampler(local_rank)
    windows = []

    def guarded(what, fn):
        """The extra records must never cost the headline line: a failure in one of them is recorded (single-rank run) instead of
        ending the run.  With several ranks the exception propagates - a rank that silently skipped a barrier would hang the others."""
        try:
            return fn()
        except Exception as e:  # noqa: BLE001
            if world > 1:
                raise
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            print(f"[bench] extra record `{what}` failed: {type(e).__name__}: {e}", file=sys.stderr)
            return {"error": f"{type(e).__name__}: {e}"}

    # ---------------- optional extra record: the opt-in bf16 x1 throughput mode (BASELINE's named format), device resident -------
    thr = None
    if args.mode != "throughput" and not args.no_throughput_record:
        tplan = synthetic_resnet_plan(args.arch, B, mode="throughput", device=dev, input_u8=True)
        tplan.load_input(h_in)
        tplan.capture()
        ms_t, _ = device_resident_ms(tplan)
        thr = {"mode": "throughput", "operands": "bf16 x1", "ms_per_step": ms_t / K, "value": world * B * K / (ms_t * 1e-3), "unit": "img/s",
               "launches_per_step": tplan.num_launches(),
               "note": "device-resident, same timing method; does NOT meet the map tolerances (see `parity_check.throughput`)"}
        del tplan
        torch.cuda.empty_cache()

    # ---------------- extra record: the fine-tuning step (config 5), batch sharded, gradients all-reduced over NCCL ----------------
    train = None
    if not args.no_train_record:
        train = guarded("train_step", lambda: measure_train_step(args, dev, rank, world, barrier, max_over_ranks))

    # ---------------- extra record: BASELINE config 4, the B-cos CLIP RN50 image encoder (embedding + explanation, batch 512) ----
    clip = None
    if not args.no_clip_record:
        clip = guarded("clip_rn50", lambda: measure_clip_rn50(args, dev, world, barrier, max_over_ranks))

    clip_vit = None
    if not args.no_clip_record:
        clip_vit = guarded("clip_vit_b32", lambda: measure_clip_vit(args, dev, world, barrier, max_over_ranks))

    # ---------------- extra record: BASELINE config 3, B-cosified SimpleViT-Ti/16 and ViT-B/16 forward + explanation at 224^2 ------
    vit = None
    if not args.no_vit_record:
        vit = [guarded("vit", lambda a=arch: measure_vit(args, a, dev, world, barrier, max_over_ranks))
               for arch in ("simple_vit_ti_patch16_224", "simple_vit_b_patch16_224")]

    # ---------------- extra record: B-cosified DenseNet-121 (the other network of BASELINE config 5) forward + explanation -----------
    dense = None
    if not args.no_densenet_record:
        dense = guarded("densenet121", lambda: measure_densenet(args, dev, world, barrier, max_over_ranks))

    plan = synthetic_resnet_plan(args.arch, B, mode=args.mode, device=dev, input_u8=True)
    prec = plan.precision
    plan.load_input(h_in)
    plan.capture()
    sampler.start()

    # ---------------- device-resident throughput: inputs already in HBM, K graph replays ----------------
    ms_dev, win = device_resident_ms(plan)
    windows.append(win)

    # ---------------- end to end through the public API: pinned host images in, logits + maps out ----------------
    # PipelinedExplainer: every step copies its own batch host->device and its own results device->host (pinned memory);
    # copies of step i+1 / i-1 overlap the compute of step i.  The timed region ends when the last result is on the host.
    pipe = PipelinedExplainer(plan)
    h_ins = [h_in, h_in.clone().pin_memory()]
    for i in range(W):
        pipe.submit(h_ins[i % 2])
    pipe.drain()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
    ig_design_bytes = sum(o.algo_bytes() for o, _ in ig)
    ig_flops = sum(o.algo_flops for o, _ in ig)
    ig_exec_flops = sum(o.flops() for o, _ in ig)
    step_ms_eager = sum(per_op)
    # tensor-class launches: the ones SURVEY 8(d) marks tensor bound (3x3 at <= 28^2, K >= 1024 1x1, fc)
    tc = [(o, t) for o, t in ig if (len(o.taps) > 1 and o.op <= 28 and o.n >= 128) or (len(o.taps) == 1 and o.ktot // len(o.seg_a_choff) >= 1024)]
    tc_ms, tc_flops = sum(t for _, t in tc), sum(o.algo_flops for o, _ in tc)
    tc_exec_flops, tc_fill = sum(o.flops() for o, _ in tc), sum(o.smem_fill_bytes() for o, _ in tc)

    if rank != 0:
        D.shutdown()
        return

    # ---------------- parity of the benchmarked mode against the reference golden (4 images, same code path) ----------------
    parity_check = run_parity_check(args, prec, dev)

    pk = peaks()
    # 133 MB per image forward + explanation (bf16: x, y, gain, g_out, g_in, weights once each) - the yardstick
    # for every operand mode.  The design's own traffic (every tensor the launches touch once; includes second planes, masks,
    # shortcut streams) and the measured DRAM bytes of the same launches (ncu) are reported beside it.
    SURVEY_BYTES_PER_IMG = 133e6
    alg_bytes = SURVEY_BYTES_PER_IMG * B if args.arch == "resnet50" else ig_design_bytes
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", f"r02_dram_traffic_{args.mode}.json")
    if os.path.exists(tpath) and args.arch == "resnet50" and B == 256:
        with open(tpath) as fh:
            tj = json.load(fh)
        traffic, traffic_src = float(tj["igemm_dram_bytes_per_step"]), f"profiles/r02_dram_traffic_{args.mode}.json (ncu, bytes per step)"
    hbm_ach = alg_bytes / (ig_ms * 1e-3) / 1e9
    tc_ach = ig_flops / (ig_ms * 1e-3) / 1e12
    imgs_total = world * B * K
    res = {
        "metric": METRIC, "value": imgs_total / (ms_dev * 1e-3), "unit": "img/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": prec["dtype"],
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "arch": args.arch, "image": 224, "batch_per_gpu": B, "mode": args.mode,
                   "precision": {"forward_planes": prec["planes"], "explain_planes": prec["explain_planes"] or prec["planes"],
                                 "operand_dtype": prec["dtype"], "accumulate": "fp32" + (" (fresh TMEM accumulator per K chunk, RN adds)" if prec["planes"] > 1 else ""),
                                 "seed_scale": prec["seed_scale"]},
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
        "parity_check": parity_check,
        "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / pk["hbm_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": "bcosk_igemm_* (all conv / dgrad launches of one step)",
                     "peak_source": pk["source"], "kernel_ms_per_step": ig_ms, "kernel_share_of_step": ig_ms / step_ms_eager,
                     "algorithmic_bytes_per_step": alg_bytes, "algorithmic_bytes_source": "SURVEY.md 8(d): 133 MB / image x batch",
                     "design_traffic_bytes_per_step": ig_design_bytes,
                     "design_traffic_frac_of_peak": ig_design_bytes / (ig_ms * 1e-3) / 1e9 / pk["hbm_gbs"]},
        "roofline_tensor": {"bound": "tensor", "achieved": tc_ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tc_ach / pk["bf16_tflops"],
                            "algorithmic_flops_per_step": ig_flops, "executed_flops_per_step": ig_exec_flops,
                            "executed_tflops": ig_exec_flops / (ig_ms * 1e-3) / 1e12, "peak_source": pk["source"] + " burst",
                            "tensor_class_launches": {"count": len(tc), "ms": tc_ms, "algorithmic_tflops": tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                                                      "frac_of_burst_peak": tc_flops / (tc_ms * 1e-3) / 1e12 / pk["bf16_tflops"] if tc_ms else None,
                                                      "executed_tflops": tc_exec_flops / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                                                      "executed_frac_of_burst_peak": tc_exec_flops / (tc_ms * 1e-3) / 1e12 / pk["bf16_tflops"] if tc_ms else None,
                                                      "what": "3x3 at <= 28^2 and K >= 1024 1x1 launches (SURVEY 8d tensor-bound rows)"}},
        "clocks": sampler.summary(windows),
    }
    # operand fill of the same tensor-class launches: bytes their TMA loads pull from L2 into shared memory (one 128 x 64 tile of the
    # contract mode fetches 48 KB per 3.1 MFLOP executed, a 128 x 128 one 64 KB per 6.3).  Reference points: the peak ncu reports for
    # l1tex__m_xbar2l1tex_read_bytes on this B200 (profiles/r02_parity_mode.md: 8.49 TB/s = 25.97 % -> 32.7 TB/s), the guide's measured
    # chip-wide L2 throughput cap (B300_MICROARCH.md: ~6300 B / cycle) and the best our own launches reach (one-plane 128-wide dgrad: 15.0 TB/s)
    sm_mhz = res["clocks"].get("sm_mhz") or res["clocks"].get("sm_max_mhz") or 1965.0
    ncu_peak = 32700.0
    res["roofline_l2"] = {"bound": "l2_to_smem_fill", "achieved": tc_fill / (tc_ms * 1e-3) / 1e9 if tc_ms else None, "peak": ncu_peak, "unit": "GB/s",
                          "frac": tc_fill / (tc_ms * 1e-3) / 1e9 / ncu_peak if tc_ms else None, "fill_bytes_per_step": tc_fill,
                          "kernel": "tensor-class launches (roofline_tensor.tensor_class_launches)",
                          "peak_source": "ncu l1tex__m_xbar2l1tex_read_bytes peak_sustained on this B200 (profiles/r02_parity_mode.md)",
                          "guide_l2_cap_gbs": 6300.0 * sm_mhz * 1e6 / 1e9, "best_observed_gbs": 15017.0,
                          "best_observed_source": "ncu, layer3.1.conv2.dgrad (one-plane 128-wide tiles, three CTAs per SM)"}
    if thr is not None:
        res["throughput_mode"] = thr
    if train is not None:
        res["train_step"] = train
    if clip is not None:
        res["clip_rn50"] = clip
    if clip_vit is not None:
        res["clip_vit_b32"] = clip_vit
    if vit is not None:
        res["vit"] = vit
    if dense is not None:
        res["densenet121"] = dense
    if args.layer_table:
        rows = []
        for o, t in zip(all_ops, per_op):
            row = {"name": o.name, "kind": type(o).__name__, "ms": t}
            if isinstance(o, O.IgemmOp):
                row.update(M=o.M, N=o.n, K=o.ktot, block_n=o.resolved_block_n(), hp=bool(o.hp_accum), gflop=o.algo_flops / 1e9,
                           mbytes=o.algo_bytes() / 1e6, tflops=o.algo_flops / (t * 1e-3) / 1e12, exec_tflops=o.flops() / (t * 1e-3) / 1e12,
                           gbs=o.algo_bytes() / (t * 1e-3) / 1e9)
            rows.append(row)
        os.makedirs(os.path.dirname(os.path.abspath(args.layer_table)), exist_ok=True)
        with open(args.layer_table, "w") as fh:
            json.dump({"batch": B, "arch": args.arch, "mode": args.mode, "step_ms_eager": step_ms_eager, "rows": rows}, fh, indent=1)
    if world == 1 and not args.no_cpu_baseline:
        c = cpu_reference_arm(args, args.cpu_steps, 2)
        res["cpu_baseline"] = {"value": c["value"], "unit": "img/s", "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]}
    print(json.dumps(res), file=real_stdout, flush=True)
    D.shutdown()


def measure_densenet(args, dev, world, barrier, max_over_ranks):
    """B-cosified DenseNet-121 forward + explanation at 224^2, batch 256, through the fused plan (engine/densenet.py: block feature
    tensors written slice by slice, per-consumer BN + ReLU kernels, fp32 feature-gradient accumulation), same operand format as the
    main record.  Device-resident uint8 inputs, CUDA events, max over ranks."""
    import torch
    from bcos_b200.models import synthetic_densenet_plan
    from bcos_b200.utils import synth
    Bd, reps = 256, 5
    plan = synthetic_densenet_plan("densenet121", Bd, mode=args.mode, device=dev, input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(32, 224, 9)).repeat(Bd // 32, 1, 1, 1).to(dev)
    plan.load_input(x)
    plan.capture()
    for _ in range(3):
        plan.replay_all()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        plan.replay_forward()
    ev[1].record()
    for _ in range(reps):
        plan.replay_all()
    ev[2].record()
    barrier()
    ms_f = max_over_ranks(ev[0].elapsed_time(ev[1])) / reps
    ms_x = max_over_ranks(ev[1].elapsed_time(ev[2])) / reps
    ok = bool(torch.isfinite(plan.logits).all() and torch.isfinite(plan.cmap).all())
    rec = {"workload": "B-cosified DenseNet-121 forward + explanation maps at 224^2", "batch_per_gpu": Bd, "n_gpus": world, "mode": args.mode,
           "forward_ms_per_step": ms_f, "forward_value": world * Bd / (ms_f * 1e-3), "ms_per_step": ms_x, "value": world * Bd / (ms_x * 1e-3),
           "unit": "img/s", "launches_per_step": plan.num_launches(), "finite": ok,
           "parity": "tests/test_densenet_gpu.py: argmax equal, logits 2e-7 rel, map cosine 0.9999999, max-abs 2.7e-4 of range vs the reference golden (contract mode)"}
    del plan
    torch.cuda.empty_cache()
    return rec


def measure_vit(args, arch, dev, world, barrier, max_over_ranks):
    """BASELINE config 3: B-cosified SimpleViT (bcosify_vit simple_vit, BcosLinear attention / MLP) forward + explanation at 224^2
    through the fused plan (engine/vit.py: 1x1 tcgen05 launches, tensor-core attention, CUDA graph), same operand format as the
    main record.  Device-resident uint8 inputs, CUDA events, max over ranks."""
    import torch
    from bcos_b200.models import synthetic_vit_plan
    from bcos_b200.utils import synth
    Bv, reps = args.vit_batch, 5
    plan = synthetic_vit_plan(arch, Bv, mode=args.mode, device=dev, input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(32, 224, 7)).repeat((Bv + 31) // 32, 1, 1, 1)[:Bv].to(dev)
    plan.load_input(x)
    plan.capture()
    for _ in range(3):
        plan.replay_all()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        plan.replay_forward()
    ev[1].record()
    for _ in range(reps):
        plan.replay_all()
    ev[2].record()
    barrier()
    ms_f = max_over_ranks(ev[0].elapsed_time(ev[1])) / reps
    ms_x = max_over_ranks(ev[1].elapsed_time(ev[2])) / reps
    ok = bool(torch.isfinite(plan.logits).all() and torch.isfinite(plan.cmap).all())
    rec = {"workload": f"B-cosified {arch} forward + explain at 224^2 (BASELINE config 3)", "arch": arch, "batch_per_gpu": Bv, "n_gpus": world,
           "mode": args.mode, "forward_ms_per_step": ms_f, "forward_value": world * Bv / (ms_f * 1e-3), "ms_per_step": ms_x,
           "value": world * Bv / (ms_x * 1e-3), "unit": "img/s", "launches_per_step": plan.num_launches(),
           "executed_gemm_tflops_per_gpu": plan.gemm_flops() / (ms_x * 1e-3) / 1e12, "finite": ok,
           "operands": f"residual stream {plan.sp} plane(s), branch operands {plan.bp} plane(s), explanation pass 1 plane ({plan.precision['dtype']})",
           "parity": "tests/test_vit_gpu.py: argmax equal, logits <= 1.5e-4 rel, map cosine 0.9999997, max-abs <= 5.0e-4 of range vs the reference goldens (contract mode)"}
    del plan
    torch.cuda.empty_cache()
    return rec


def measure_clip_vit(args, dev, world, barrier, max_over_ranks):
    """The other CLIP image encoder north_star names: B-cos CLIP ViT-B/32 (CLIP/clip/model.py:206-241 through bcosify.py with clip_kd),
    embedding + explanation of cos(embedding, text direction), batch 512, through the fused plan (engine/clip_vit.py)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from bcos_b200.engine import CLIPViTPlan
    from bcos_b200.models import clip_vit_state_shapes
    from bcos_b200.utils import synth
    Bc, reps = args.clip_batch, 4
    plan = CLIPViTPlan(synth.synth_state_dict(clip_vit_state_shapes(), 0), Bc, mode=args.mode, device=dev, input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(32, 224, 5)).repeat((Bc + 31) // 32, 1, 1, 1)[:Bc].to(dev)
    plan.load_input(x)
    plan.capture()
    g = torch.Generator().manual_seed(0)
    t = torch.nn.functional.normalize(torch.randn(plan.out_dim, generator=g), dim=0)
    for _ in range(2):
        plan.explain_direction(None, t)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        plan.embed(None)
    ev[1].record()
    for _ in range(reps):
        out = plan.explain_direction(None, t)
    ev[2].record()
    barrier()
    ms_e = max_over_ranks(ev[0].elapsed_time(ev[1])) / reps
    ms_x = max_over_ranks(ev[1].elapsed_time(ev[2])) / reps
    ok = bool(torch.isfinite(out["embedding"]).all() and torch.isfinite(out["contribution_map"]).all())
    rec = {"workload": "B-cos CLIP ViT-B/32 image encoder embedding + explanation on synthetic images", "batch_per_gpu": Bc, "n_gpus": world,
           "mode": args.mode, "embed_ms_per_step": ms_e, "embed_value": world * Bc / (ms_e * 1e-3), "ms_per_step": ms_x,
           "value": world * Bc / (ms_x * 1e-3), "unit": "img/s", "launches_per_step": plan.num_launches(), "finite": ok,
           "parity": "tests/test_vit_gpu.py: embedding 4.0e-4 rel, map cosine 0.9999987, max-abs 7.7e-4 of range vs the reference golden (contract mode)"}
    del plan
    torch.cuda.empty_cache()
    return rec


def measure_clip_rn50(args, dev, world, barrier, max_over_ranks):
    """BASELINE config 4: B-cos CLIP RN50 image encoder, embedding + explanation of cos(embedding, text direction) on synthetic
    images through the fused plan (engine/clip_rn.py; trunk and attention-pool head as CUDA graphs),
    same operand format as the main record.  Device-resident inputs, CUDA events, max over ranks."""
    import torch
    from bcos_b200.models import synthetic_clip_rn50_plan
    from bcos_b200.utils import synth
    Bc, reps = args.clip_batch, 4
    plan = synthetic_clip_rn50_plan(Bc, mode=args.mode, device=dev, input_u8=True)
    x = torch.from_numpy(synth.synth_images_u8(32, 224, 5)).repeat((Bc + 31) // 32, 1, 1, 1)[:Bc].to(dev)
    plan.load_input(x)
    plan.capture()
    g = torch.Generator().manual_seed(0)
    t = torch.nn.functional.normalize(torch.randn(1024, generator=g), dim=0)
    for _ in range(2):
        plan.explain_direction(None, t)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(reps):
        plan.embed(None)
    ev[1].record()
    for _ in range(reps):
        out = plan.explain_direction(None, t)
    ev[2].record()
    barrier()
    ms_e = max_over_ranks(ev[0].elapsed_time(ev[1])) / reps
    ms_x = max_over_ranks(ev[1].elapsed_time(ev[2])) / reps
    ok = bool(torch.isfinite(out["embedding"]).all() and torch.isfinite(out["contribution_map"]).all())
    rec = {"workload": "B-cos CLIP RN50 image encoder (resnet_50_clip_b2_noBias) embedding + explanation throughput on synthetic images, batch 512 (BASELINE config 4)",
           "batch_per_gpu": Bc, "n_gpus": world, "mode": args.mode, "embed_ms_per_step": ms_e, "embed_value": world * Bc / (ms_e * 1e-3),
           "ms_per_step": ms_x, "value": world * Bc / (ms_x * 1e-3), "unit": "img/s",
           "launches_per_step": plan.num_launches(), "finite": ok,
           "parity": "tests/test_clip_gpu.py: embedding 1e-4 rel, map cosine 0.9999994, max-abs 8.8e-4 .. 9.7e-4 of range vs the reference golden (contract mode)"}
    del plan
    torch.cuda.empty_cache()
    return rec


def measure_train_step(args, dev, rank, world, barrier, max_over_ranks):
    """BASELINE config 5: one fine-tuning step (forward in train mode, UniformOffLabelsBCE loss, full backward incl. the
    tcgen05 weight-gradient kernel, bucketed NCCL all-reduce of the gradients on a side stream, AGC + AdamW) per call."""
    import torch
    from bcos_b200.engine import ResNetTrainPlan
    from bcos_b200.models import resnet_state_shapes
    from bcos_b200.utils import synth
    Bt, Kt = args.train_batch, args.train_steps
    sd = synth.synthetic_checkpoint(args.arch, resnet_state_shapes(args.arch))
    plan = ResNetTrainPlan(args.arch, sd, Bt, dtype="bf16", device=dev, world_size=world)
    imgs = torch.from_numpy(synth.synth_images_u8(Bt, 224, 2000 + rank)).to(dev)
    labels = (torch.arange(Bt) * 37 + rank) % 1000
    plan.load_batch(imgs, labels.to(dev))
    captured = plan.capture()            # the whole step (incl. the bucketed all-reduce on its side stream) as one CUDA graph
    losses = []
    for _ in range(3):
        losses.append(float(plan.train_step()))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(Kt):
        plan.train_step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / Kt
    losses.append(float(plan.loss))
    rec = {"workload": "B-cosification fine-tuning step of ResNet-50, batch-sharded backward, NCCL gradient all-reduce (BASELINE config 5)",
           "batch_per_gpu": Bt, "n_gpus": world, "steps": Kt, "ms_per_step": ms, "value": world * Bt / (ms * 1e-3), "unit": "img/s",
           "dtype": "bf16 operands, fp32 accumulate / master weights / optimizer state",
           "tflops_per_gpu": plan.train_flops() / (ms * 1e-3) / 1e12, "launches_per_step": plan.num_train_launches(), "cuda_graph": bool(captured),
           "collective": None if world == 1 else f"NCCL all_reduce(sum) of {len(plan.buckets)} fp32 gradient buckets on a side stream, overlapped with the backward pass",
           "allreduce_bytes_per_step": 0 if world == 1 else int(plan.g_flat.numel() * 4),
           "loss_first_steps": losses[:3], "loss_last": losses[-1], "loss_decreases_on_fixed_batch": bool(losses[-1] < losses[0])}
    del plan
    torch.cuda.empty_cache()
    return rec


def _metrics(logits, maps, gold):
    """BASELINE.json's criteria against the committed reference run: argmax equality, logit relative error, per-image map
    cosine (min) and max-abs / range (max).  The reference's fp32 run and its exact (fp64) evaluation differ by
    `reference_fp32_vs_fp64` (chaotic random-init net: one ReLU decision differs on RN50 image 1); per image the map error is
    the distance to the nearer of the two, both distances are reported."""
    import torch
    logits, maps = logits.detach().double().cpu(), maps.detach().double().cpu().flatten(1)
    ref_logits = torch.from_numpy(gold["logits"]).double()
    r32 = torch.from_numpy(gold["contribution_map"]).double().flatten(1)
    r64 = torch.from_numpy(gold["contribution_map_fp64"]).double().flatten(1)
    cos = torch.nn.functional.cosine_similarity(maps, r32, dim=1)
    rng = r32.max(1).values - r32.min(1).values
    e32 = (maps - r32).abs().max(1).values / rng
    e64 = (maps - r64).abs().max(1).values / rng
    return {"argmax_equal": bool((logits.argmax(1) == ref_logits.argmax(1)).all()),
            "logit_rel_err": float((logits - ref_logits).abs().max() / ref_logits.abs().max()),
            "map_cos_min": float(cos.min()), "map_maxabs_over_range": float(torch.minimum(e32, e64).max()),
            "map_maxabs_vs_fp32_ref": float(e32.max()), "map_maxabs_vs_fp64_eval": float(e64.max()),
            "reference_fp32_vs_fp64": float(gold["fp32_noise_floor_maxabs_over_range"])}


def run_parity_check(args, prec, dev):
    """The benchmarked operand mode (and the bf16 x1 throughput mode) on the committed reference golden of this network
    (tests/golden/<arch>_b<n>.npz: weights seed, calibrated BN variances, images, reference fp32 logits and maps)."""
    import numpy as np
    import torch
    from bcos_b200.engine import ResNetPlan
    from bcos_b200.models import resnet_state_shapes
    from bcos_b200.utils import synth
    nb = {"resnet50": 4, "resnet18": 8}.get(args.arch)
    path = os.path.join(ROOT, "tests", "golden", f"{args.arch}_b{nb}.npz")
    if nb is None or not os.path.exists(path):
        return None
    gold = np.load(path)
    sd = synth.synth_state_dict(resnet_state_shapes(args.arch), int(gold["seed"]))
    off = 0
    for k, n in zip(gold["bn_keys"].tolist(), gold["bn_sizes"].tolist()):
        sd[k] = torch.from_numpy(gold["bn_var"][off:off + n].copy())
        off += n
    imgs = torch.from_numpy(gold["images_u8"])
    out = {"golden": f"tests/golden/{args.arch}_b{nb}.npz (reference fp32 run, {nb} images)",
           "contract": {"argmax_equal": True, "logit_rel_err": 2e-3, "map_cos_min": 0.999, "map_maxabs_over_range": 1e-3}}
    for name in dict.fromkeys([args.mode, "throughput"]):
        plan = ResNetPlan(args.arch, sd, nb, mode=name, input_u8=True, device=dev)
        o = plan.explain(imgs)
        torch.cuda.synchronize()
        m = _metrics(o["logits"], o["contribution_map"], gold)
        m["meets_contract"] = bool(m["argmax_equal"] and m["logit_rel_err"] <= 2e-3 and m["map_cos_min"] >= 0.999 and m["map_maxabs_over_range"] <= 1e-3)
        out[name] = m
        del plan
    return out


if __name__ == "__main__":
    main()
