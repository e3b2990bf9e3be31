#!/usr/bin/env python
"""Headline benchmark: stereo pairs/s at 540x960, K=3, D=24 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp32|tc]

One "step" = one pass of the hot path (s8 tensor -> s32 disparity) over one batch of synthetic
stereo pairs.  `value` is device-timed with inputs resident in HBM; `e2e` goes through the
reference-facing C-ABI call (snb_infer) with pinned HOST buffers, H2D/D2H inside the timed region.
Prints ONE JSON line on rank 0.  Only the cpu_baseline / --impl reference legs touch oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, K, D, BATCH = 540, 960, 3, 24, 1
COALESCE_MAX = 4          # max_batch of the e2e context: queued one-pair calls may share a pass (never more than task_num = 4)
WORKLOAD = "SceneFlow-shape 540x960, 1/8-res cost volume D=24, 3x refinement, batch=1 per GPU (BASELINE.json configs[1])"
METRIC = "stereo pairs/sec at 540x960 D=24"
SEED = 1234
L2_BYTES = 126 * 1024 * 1024
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
# `ncu --set full` capture (profiles/): filled in by hand after each capture, None when not captured.
TRAFFIC = {
    # profiles/r01_final_ncu_full_resblock.csv: one full-resolution refinement block, 78.8 + 27.2 MB (algorithmic 134 MB:
    # part of the output is still in L2 when the kernel ends)
    "k_resblock_tc": 106.0e6,
    # profiles/r01_final_ncu_full_stream.csv: layer2.1.conv_a / conv_b, the shape of 30 of the 58 launches: 4.6 / 9.0 MB read,
    # 0 written (algorithmic 8.4 MB in + 8.4 MB out (+ 8.4 MB residual): the 68x120 maps live in L2 between launches);
    # head.filter.1 (profiles/r01_final_ncu_full_stream3d.csv): 26.9 MB read for 25 MB in + 25 MB out
    "k_conv_stream": 6.8e6,
}


def synth_inputs(n: int) -> np.ndarray:
    """n distinct synthetic s8 tensors [n,6,H,W] (seeded; oracle/synth.py is test data generation,
    re-stated here with numpy only so the product bench does not import oracle/)."""
    rng = np.random.default_rng(SEED)
    base = rng.integers(-128, 128, (6, H // 4 + 1, W // 4 + 1), dtype=np.int16)
    out = np.empty((n, 6, H, W), np.int8)
    up = np.repeat(np.repeat(base, 4, axis=1), 4, axis=2)[:, :H, :W]
    for i in range(n):
        noise = rng.integers(-24, 25, (6, H, W), dtype=np.int16)
        shifted = np.roll(up, i * 3, axis=2)
        out[i] = np.clip(shifted // 2 + noise, -128, 127).astype(np.int8)
    return out


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev: int):
        self.dev, self.rows, self.p = dev, [], None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None
        return self

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.p:
            time.sleep(0.15)
            self.p.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def _oracle():
    import torch
    from hobot_stereonet_b200 import capi
    from oracle import arch, weights
    from oracle.stereonet_ref import Oracle
    torch.set_num_threads(os.cpu_count() or 1)
    return Oracle(arch.Config(H, W, K, D), weights.from_blob(capi.synthesize_weights(K, SEED))[1])


def cpu_baseline(max_pairs: int = 2):
    """The fp32 oracle (port of the reference float model) on this box's host cores."""
    import torch
    o = _oracle()
    x = synth_inputs(1)
    o.forward_s32(x)                                   # warm-up (oneDNN primitive creation)
    t0 = time.perf_counter()
    n = 0
    while n < max_pairs and (n == 0 or time.perf_counter() - t0 < 20):
        o.forward_s32(x); n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} pair(s) of the same 540x960 K=3 D=24 workload through oracle/stereonet_ref.py (fp32 PyTorch CPU), 1 warm-up"}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU-side float inference (oracle port; the reference itself cannot
    be built here: ROS 2 + closed hobot_dnn + BPU binary) on all host threads."""
    if rank != 0:
        return
    import torch
    o = _oracle()
    xs = synth_inputs(2)
    for _ in range(min(args.warmup, 1)):
        o.forward_s32(xs[:1])
    t0 = time.perf_counter()
    for i in range(args.steps):
        o.forward_s32(xs[i % 2:i % 2 + 1])
    dt = time.perf_counter() - t0
    v = args.steps * BATCH / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "H": H, "W": W, "K": K, "D": D, "batch": BATCH},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} step(s) x 1 pair through oracle/stereonet_ref.py"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SNB_PRECISION", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs under ncu only: device-resident loop, no e2e legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        args.steps = args.steps or 3
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args, rank)
    args.steps = args.steps or 50
    args.warmup = max(3, 5 if args.warmup is None else args.warmup)

    import torch
    import torch.distributed as dist
    from hobot_stereonet_b200 import Model, capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- init (untimed): rank 0 builds the weight blob, one NCCL broadcast installs it everywhere ----
    from hobot_stereonet_b200.shard import broadcast_blob
    blob = capi.synthesize_weights(K, SEED) if rank == 0 else None      # a deployment passes model_file instead
    blob = broadcast_blob(blob, src=0, device=dev)                        # the single collective of this workload
    prec = capi.PREC_TC_F16X2 if args.precision == "tc" else capi.PREC_FP32
    m = Model(H, W, K, D, max_batch=BATCH, device=local_rank, task_num=4, precision=prec, weights=blob)
    # the e2e leg's context: same network, same one-pair calls, but room for the library to merge queued snb_infer_async
    # calls into passes of up to COALESCE_MAX pairs (capi.cu worker_main)
    mc = None if args.no_e2e else Model(H, W, K, D, max_batch=COALESCE_MAX, device=local_rank, task_num=4, precision=prec, weights=blob)
    del blob

    # ---- inputs: a rotating pool larger than L2, so no step finds its input cached ----
    in_bytes = 6 * H * W * BATCH
    pool_n = L2_BYTES // in_bytes + 8
    host_in = torch.from_numpy(synth_inputs(min(pool_n, 8))).repeat((pool_n + 7) // 8, 1, 1, 1)[:pool_n].contiguous()
    for i in range(pool_n):                          # make every pool entry distinct
        host_in[i, :, 0, 0] = i % 127
    host_in = host_in.pin_memory()
    d_in = host_in.to(dev)
    d_out = torch.empty((pool_n, 1, H, W), dtype=torch.int32, device=dev)
    host_out = torch.empty((pool_n, 1, H, W), dtype=torch.int32).pin_memory()
    stream = torch.cuda.Stream(dev)                  # non-default: the library launches on this handle

    def step_device(i):
        j = i % pool_n
        m.infer_device(d_in[j], d_out[j], BATCH, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        step_device(i)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        ev0.record(stream)
        for i in range(args.steps):
            step_device(args.warmup + i)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1)
        barrier()
        e2e_s = e2e_1_s = e2e_sync_s = float('nan'); e2e_passes = 0
        if not args.no_e2e:
            # ---- e2e: the reference-facing call with HOST buffers.  The node calls DnnNode::Run(is_sync=false) with
            # task_num = 4 calls in flight (stereonet_node.cpp:144,812): snb_infer_async, pinned buffers, copies timed.
            for i in range(3):
                m.infer(host_in[i:i + 1].numpy(), host_out[i:i + 1].numpy())
            for b in range(1, COALESCE_MAX + 1):             # warm the e2e context: the CUDA graph of every pass size it may use
                mc.infer(host_in[0:b].numpy(), host_out[0:b].numpy())
            for i in range(12):
                mc.infer_async(host_in[i:i + 1].numpy(), host_out[i:i + 1].numpy())
            mc.wait_all()
            barrier()
            passes0 = mc.pass_count()
            t0 = time.perf_counter()
            for i in range(args.steps):
                j = (3 + i) % pool_n
                mc.infer_async(host_in[j:j + 1].numpy(), host_out[j:j + 1].numpy())
            mc.wait_all()
            torch.cuda.synchronize(dev)
            e2e_s = time.perf_counter() - t0
            e2e_passes = mc.pass_count() - passes0
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):                      # same calls, one pass per call (max_batch = 1 context)
                j = (3 + i) % pool_n
                m.infer_async(host_in[j:j + 1].numpy(), host_out[j:j + 1].numpy())
            m.wait_all()
            torch.cuda.synchronize(dev)
            e2e_1_s = time.perf_counter() - t0
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):                      # same through the synchronous call, one pair in flight
                j = (3 + i) % pool_n
                m.infer(host_in[j:j + 1].numpy(), host_out[j:j + 1].numpy())
            torch.cuda.synchronize(dev)
            e2e_sync_s = time.perf_counter() - t0
    clocks = clk.summary()
    launches_per_step = m.rt_stat().kernel_launches

    t = torch.tensor([ms, e2e_s * 1e3, e2e_sync_s * 1e3, e2e_1_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_sync_ms, e2e_1_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    result = None
    if rank == 0:
        # ---- roofline of the dominant kernel, measured live with CUDA events per launch ----
        prof = {}
        reps = 5
        for _ in range(reps):
            for name, kms, fl, by in m.profile_pass(BATCH):
                a = prof.setdefault(name, [0.0, fl, by])
                a[0] += kms / reps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)       # kernel timed inside a long step
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        fam_of = lambda n: ("k_resblock_tc" if "[tc-block]" in n else "k_conv_stream" if "[tc-stream]" in n else
                            "k_conv_tc" if "[tc]" in n else "k_costvol" if n == "costvol" else "cuda_core_and_hbm")
        fam = {}
        for n, v in prof.items():
            a = fam.setdefault(fam_of(n), [0.0, 0.0, 0.0, 0])
            a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3] += 1
        total_ms = sum(v[0] for v in prof.values())
        tc_fams = {k: v for k, v in fam.items() if k in ("k_resblock_tc", "k_conv_stream", "k_conv_tc") or args.precision != "tc"}
        dom = max(tc_fams, key=lambda k: tc_fams[k][0])
        d = fam[dom]
        cv = fam.get("k_costvol", [1e-9, 0, 0, 0])
        tf = lambda v: v[1] / (v[0] * 1e-3) / 1e12
        roofline = {"kernel": dom + (" (fused residual block: conv+ReLU+conv+residual+ReLU, tcgen05, split-fp16 operands = 3 fp16 MMAs per algorithmic MAC)"
                                     if dom == "k_resblock_tc" else ""),
                    "bound": "tensor", "achieved": tf(d), "peak": tf_peak, "unit": "TFLOP/s", "frac": tf(d) / tf_peak,
                    # split-fp16 operands: 3 fp16 MMAs are issued per algorithmic MAC, so the tensor pipe works at 3 x frac
                    "fp16_mma_issued_frac": 3 * tf(d) / tf_peak,
                    "traffic": TRAFFIC.get(dom), "peak_source": peak_src, "launches": d[3], "share_of_step": d[0] / total_ms,
                    "algorithmic_flops_per_launch": d[1] / max(d[3], 1), "avg_launch_ms": d[0] / max(d[3], 1),
                    "families": {k: {"ms": v[0], "share": v[0] / total_ms, "launches": v[3],
                                     "tflops": tf(v) if v[1] else None, "gbs": v[2] / (v[0] * 1e-3) / 1e9} for k, v in fam.items()},
                    "all_tensor_kernels": {"achieved": sum(v[1] for v in tc_fams.values()) / (sum(v[0] for v in tc_fams.values()) * 1e-3) / 1e12,
                                           "unit": "TFLOP/s"},
                    "hbm_kernels": {"costvol": {"bound": "hbm", "achieved": cv[2] / (cv[0] * 1e-3) / 1e9, "peak": hbm_peak,
                                                "unit": "GB/s", "frac": cv[2] / (cv[0] * 1e-3) / 1e9 / hbm_peak}}}
        result = {
            "metric": METRIC, "value": world * args.steps * BATCH / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16x2->f32" if args.precision == "tc" else "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "H": H, "W": W, "K": K, "D": D, "batch_per_gpu": BATCH, "precision": args.precision,
                       "l2": f"inputs rotate through a pool of {pool_n} tensors = {pool_n * in_bytes / 2**20:.0f} MiB > 126 MiB L2",
                       "parallelism": f"replicas x{world}, batch-sharded, one NCCL weight broadcast at init"},
            "e2e": {"value": world * args.steps * BATCH / (e2e_ms * 1e-3), "unit": "pairs/s",
                    "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4 * H * W * BATCH,
                    "api": "snb_infer_async, one pair per call, 4 calls in flight as the reference node (task_num = 4), pinned host "
                           f"buffers; the library merges queued calls into passes of <= {COALESCE_MAX} pairs "
                           f"(rank 0: {e2e_passes} passes for {args.steps} calls)",
                    "one_pass_per_call_value": world * args.steps * BATCH / (e2e_1_ms * 1e-3),
                    "one_pass_per_call_api": "snb_infer_async on a max_batch = 1 context (no merging), 4 calls in flight",
                    "sync_value": world * args.steps * BATCH / (e2e_sync_ms * 1e-3), "sync_api": "snb_infer (one call in flight)"},
            # device-resident loop + one-pass-per-call async loop + sync loop, and the merged passes of the e2e loop (rank 0)
            "gpu_launches": launches_per_step * (args.steps * (1 if args.no_e2e else 3) + e2e_passes),
            "clocks": clocks, "roofline": roofline,
        }
        if args.no_e2e:
            result["e2e"] = None                             # profiling run: no e2e legs were executed
        if not args.no_cpu_baseline and world == 1:
            result["cpu_baseline"] = cpu_baseline()
    m.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
