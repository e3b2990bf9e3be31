#!/usr/bin/env python
"""Benchmark of the B200 StereoNet path: stereo pairs/s on BASELINE.json's configurations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1|2|3|4|5|deployed]
                  [--precision fp32|tc]

Default = configs[1] (540x960, K=3, D=24, batch 1 per GPU), the configuration BASELINE.json's metric is quoted on.
One "step" = one pass of the hot path (s8 tensor -> s32 disparity) over one batch of synthetic stereo pairs:
  configs 2, 3, deployed   the batch lives on every GPU ("weak" scaling: per-GPU work fixed),
  configs 4, 5             ONE global batch (32 / 64 pairs) is cut into contiguous shards (hobot_stereonet_b200.shard
                           .shard_range) over the ranks ("strong" scaling: total work fixed, time = max over ranks),
  config 1                 plumbing on the host only (pre-process -> float model on the CPU -> packing), no GPU.
`value` is device-timed with inputs resident in HBM; `e2e` goes through the reference-facing C-ABI call
(snb_infer_async, one pair per call as DnnNode::Run, 4 calls in flight) with pinned HOST buffers, H2D/D2H inside the
timed region.  Prints ONE JSON line on rank 0.  Only the cpu_baseline / --impl reference / config-1 legs touch oracle/.
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[0..4] + the reference's deployed instance
CONFIGS = {
    "1": dict(H=540, W=960, K=3, D=24, batch=1, scaling="weak", max_batch=1, cpu_only=True,
              workload="single 540x960 synthetic stereo pair through stereonet_infer pre/post-process + float model on host CPU "
                       "(BASELINE.json configs[0], plumbing, no GPU)"),
    "2": dict(H=540, W=960, K=3, D=24, batch=1, scaling="weak", max_batch=1,
              workload="SceneFlow-shape 540x960, 1/8-res cost volume D=24, 3x refinement, batch=1 per GPU (BASELINE.json configs[1])"),
    "3": dict(H=720, W=1280, K=3, D=48, batch=8, scaling="weak", max_batch=8,
              workload="ZED-2i 1280x720, D=48, K=3 (max disparity 384), batch=8 per GPU (BASELINE.json configs[2])"),
    "4": dict(H=540, W=960, K=3, D=192, batch=32, scaling="strong", max_batch=4,
              workload="SceneFlow-shape 540x960, D=192 high-disparity (max disparity 1536), ONE batch of 32 pairs sharded across the GPUs "
                       "(BASELINE.json configs[3])"),
    "5": dict(H=375, W=1242, K=3, D=192, batch=64, scaling="strong", max_batch=8,
              workload="KITTI-shape 1242x375, D=192, ONE batch of 64 pairs sharded across the GPUs (BASELINE.json configs[4])"),
    "deployed": dict(H=720, W=1280, K=4, D=12, batch=1, scaling="weak", max_batch=1,
                     workload="the reference's deployed instance 720x1280, K=4, D=12 (hbm tensor table), batch=1 per GPU"),
}
TASK_NUM = 4              # calls in flight in the e2e legs: the reference's task_num (stereonet_node.cpp:144)
COALESCE_MAX = 4          # max_batch of the e2e context: queued one-pair calls may share a pass (never more than task_num)
SEED = 1234
L2_BYTES = 126 * 1024 * 1024


def metric_name(c):
    return f"stereo pairs/sec at {c['H']}x{c['W']} D={c['D']}"


def config_dict(key, c, world):
    """The workload, identical in the product and the reference arm."""
    per_gpu = c["batch"] if c["scaling"] == "weak" else None
    return {"workload": c["workload"], "config": key, "H": c["H"], "W": c["W"], "K": c["K"], "D": c["D"],
            "batch": c["batch"] * (world if c["scaling"] == "weak" else 1), "batch_per_gpu": per_gpu,
            "sharding": "replicas, batch per GPU" if c["scaling"] == "weak" else "one global batch, contiguous shards (shard_range)"}


def synth_inputs(n: int, H: int, W: int) -> np.ndarray:
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
    """SM clock and throttle reasons DURING the device-timed region (B200_PROFILING.md recipe), rank 0 only.  Sampled
    in-process through NVML every 5 ms (the timed loop of the default run is tens of milliseconds: an `nvidia-smi -lms 100`
    child would see it once at best); falls back to that child process when NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, dev: int, enabled: bool = True, uuid: str = ""):
        self.dev, self.rows, self.p, self.enabled, self.uuid = dev, [], None, enabled, uuid
        self.sm, self.mx, self.reasons, self.stop, self.how = [], [], set(), False, None

    def _nvml_loop(self, nv, h):
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByUUID(self.uuid.encode() if hasattr(self.uuid, "encode") else self.uuid) if self.uuid else nv.nvmlDeviceGetHandleByIndex(self.dev)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
            self.how = "nvml, 5 ms"
            return self
        except Exception:
            pass
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            self.how = "nvidia-smi -lms 100"
        except OSError:
            self.p = None
        return self

    def _read(self):
        for line in self.p.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[1])); self.mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def __exit__(self, *a):
        self.stop = True
        if self.p:
            time.sleep(0.15)
            self.p.terminate()
        if getattr(self, "t", None):
            self.t.join(timeout=2)

    def count(self):
        return len(self.sm)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "how": self.how}


def ncu_traffic(kernel: str, d192: bool = False):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, averaged over the launches captured in the newest
    round's committed `ncu --set full` raw pages under profiles/ (captures of the D = 192 configuration are used for the D = 192
    configurations only).  None when no capture names the kernel."""
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*ncu_full*.csv")), reverse=True)
    files = [f for f in files if ("_d192" in os.path.basename(f)) == d192]
    newest = os.path.basename(files[0]).split("_")[0] if files else ""
    vals, used = [], []
    for path in files:
        if not os.path.basename(path).startswith(newest + "_"):
            continue
        try:
            rows = list(csv.reader(open(path, newline="")))
            hdr = rows[0]
            ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except (OSError, ValueError, IndexError):
            continue
        ur, uw = unit.get(rows[1][ir], 1.0), unit.get(rows[1][iw], 1.0)
        v = [float(r[ir]) * ur + float(r[iw]) * uw for r in rows[2:] if len(r) > max(ir, iw) and kernel in r[ik]]
        if v:
            vals += v
            used.append(os.path.relpath(path, ROOT))
    if not vals:
        return None
    return {"bytes_per_launch": float(np.mean(vals)), "launches_captured": len(vals), "source": used}


# ---- CPU legs (the only code that may touch oracle/) ------------------------------------------------------------------
def _oracle(c):
    import torch
    from oracle import arch, weights
    from oracle.stereonet_ref import Oracle
    torch.set_num_threads(os.cpu_count() or 1)
    # oracle/weights.py builds the blob: the reference arm does not load the product library
    return Oracle(arch.Config(c["H"], c["W"], c["K"], c["D"]), weights.generate(c["K"], SEED))


def cpu_prepost_ms(c, reps: int = 5):
    """The reference's host-side pre/post-process for one frame, single thread as in the reference: the
    C ABI host functions (restating stereonet_node.cpp:702-738, preprocess.cpp:913-1059, stereonet_node.cpp:1033-1049,
    parser.cpp:79-87) timed through ctypes."""
    from hobot_stereonet_b200 import capi
    H, W = c["H"] + c["H"] % 2, c["W"] + c["W"] % 2
    rng = np.random.default_rng(1)
    frame = rng.integers(0, 256, H * 3 // 2 * 2 * W, dtype=np.uint8)
    q = rng.integers(0, 400000, (1, 1, H, W), dtype=np.int64).astype(np.int32)
    t_pre = t_post = t_jpeg = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        left, right = capi.pre_split_nv12(frame, H, 2 * W)
        capi.pre_cvt_nv12_to_tensor(left, right, W, H)
        t1 = time.perf_counter()
        jpg = capi.jpeg_encode_nv12(left, W, H)
        t2 = time.perf_counter()
        capi.post_pack(q, jpg)
        capi.post_parse_depth(q)
        t3 = time.perf_counter()
        t_pre, t_jpeg, t_post = min(t_pre, (t1 - t0) * 1e3), min(t_jpeg, (t2 - t1) * 1e3), min(t_post, (t3 - t2) * 1e3)
    return {"preprocess_ms": t_pre, "jpeg_ms": t_jpeg, "postprocess_ms": t_post, "threads": 1,
            "what": "C ABI host functions: NV12 split + CvtNV12Data2Tensors | JPEG of the left view | payload pack + ParseTensor depth"}


def cpu_baseline(c, max_pairs: int = 2, budget_s: float = 20.0):
    """The fp32 oracle (port of the reference float model) on this box's host cores, bounded sample."""
    import torch
    o = _oracle(c)
    x = synth_inputs(1, c["H"], c["W"])
    o.forward_s32(x)                                   # warm-up (oneDNN primitive creation)
    t0 = time.perf_counter()
    n = 0
    while n < max_pairs and (n == 0 or time.perf_counter() - t0 < budget_s):
        o.forward_s32(x); n += 1
    dt = time.perf_counter() - t0
    out = {"value": n / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"{n} pair(s) of the same {c['H']}x{c['W']} K={c['K']} D={c['D']} workload through oracle/stereonet_ref.py "
                     "(fp32 PyTorch CPU), 1 warm-up"}
    out["host_prepost"] = cpu_prepost_ms(c)
    return out


def run_reference(args, key, c, rank: int, world: int):
    """--impl reference: the reference's CPU-side float inference (oracle port; the reference itself cannot be built
    here: ROS 2 + closed hobot_dnn + BPU binary) on all host threads.  A step = one pair of the config."""
    if rank != 0:
        return
    import torch
    o = _oracle(c)
    xs = synth_inputs(2, c["H"], c["W"])
    for i in range(args.warmup):
        o.forward_s32(xs[i % 2:i % 2 + 1])
    t0 = time.perf_counter()
    for i in range(args.steps):
        o.forward_s32(xs[i % 2:i % 2 + 1])
    dt = time.perf_counter() - t0
    v = args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": metric_name(c), "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": c["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(key, c, world),
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} step(s) x 1 pair of the config's shape through oracle/stereonet_ref.py (fp32 PyTorch CPU)"},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_config1(args, key, c):
    """BASELINE.json configs[0]: one pair through the host plumbing, no GPU: NV12 frame -> C++ pre-process -> float model on
    the CPU (oracle) -> s32 quantisation -> payload pack -> render-tool decode; asserts the known-answer relations."""
    import torch
    from hobot_stereonet_b200 import capi
    from oracle import arch, prepost_ref as pp, synth
    cfg = arch.Config(c["H"], c["W"], c["K"], c["D"])
    o = _oracle(c)
    frame = synth.frame(cfg.H, cfg.W, cfg.max_disp, seed=SEED)
    t = {}
    t0 = time.perf_counter()
    left, right = capi.pre_split_nv12(frame, cfg.H, 2 * cfg.W)
    s8 = capi.pre_cvt_nv12_to_tensor(left, right, cfg.W, cfg.H)
    t["preprocess_ms"] = (time.perf_counter() - t0) * 1e3
    assert (s8 == pp.cvt_nv12_to_tensor_fast(left, right, cfg.W, cfg.H)).all()
    o.forward_s32(s8)                                  # warm-up
    t0 = time.perf_counter()
    reps = max(1, args.steps)
    for _ in range(reps):
        q = o.forward_s32(s8)
    t["model_ms"] = (time.perf_counter() - t0) * 1e3 / reps
    t0 = time.perf_counter()
    jpg = capi.jpeg_encode_nv12(left, cfg.W, cfg.H)
    t["jpeg_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    payload = capi.post_pack(q, jpg)
    depth = capi.post_parse_depth(q)
    t["postprocess_ms"] = (time.perf_counter() - t0) * 1e3
    q2, jpg2 = pp.unpack_output(payload, cfg.H, cfg.W)
    assert (q2.reshape(q.shape) == q).all() and bytes(jpg2) == jpg and np.isfinite(depth[q > 0]).all()
    total = t["preprocess_ms"] + t["model_ms"] + t["jpeg_ms"] + t["postprocess_ms"]
    v = 1e3 / total
    print(json.dumps({
        "metric": metric_name(c), "value": v, "unit": "pairs/s", "n_gpus": 0, "steps": reps, "warmup": 1, "ms_per_step": total,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(key, c, 1), "stages_ms": t, "gpu_launches": 0,
        "cpu_baseline": {"value": 1e3 / t["model_ms"], "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{reps} pair(s), float model only", "host_prepost": cpu_prepost_ms(c)},
        "note": "plumbing configuration: host pre/post-process (C ABI, single thread) around the float model on the CPU; no GPU kernel runs",
    }))


# ---- the product arm ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=os.environ.get("SNB_BENCH_CONFIG", "2"), choices=sorted(CONFIGS))
    ap.add_argument("--precision", default=os.environ.get("SNB_PRECISION", "tc"), choices=["fp32", "tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs under ncu only: device-resident loop, no e2e legs")
    args = ap.parse_args()
    key, c = args.config, CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    heavy = c["D"] >= 96 or c["batch"] > 1
    if args.impl == "reference":
        args.steps = args.steps or 3
        args.warmup = 1 if args.warmup is None else args.warmup
        return run_reference(args, key, c, rank, world)
    if c.get("cpu_only"):
        args.steps = args.steps or 2
        return run_config1(args, key, c) if rank == 0 else None
    args.steps = args.steps or (10 if heavy else 50)
    args.warmup = max(3, (3 if heavy else 5) if args.warmup is None else args.warmup)

    import torch
    import torch.distributed as dist
    from hobot_stereonet_b200 import Model, capi
    from hobot_stereonet_b200.shard import broadcast_blob, shard_range

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    H, W, K, D = c["H"], c["W"], c["K"], c["D"]
    # ---- this rank's share of a step ----
    if c["scaling"] == "weak":
        lo, hi = 0, c["batch"]                        # every GPU runs the configured batch
        global_batch = c["batch"] * world
    else:
        lo, hi = shard_range(c["batch"], world, rank)  # one global batch, contiguous shards
        global_batch = c["batch"]
    nloc = hi - lo
    if nloc < 1:
        raise SystemExit(f"config {key}: {c['batch']} pairs cannot be sharded over {world} ranks")
    max_batch = max(1, min(c["max_batch"], nloc))

    # ---- init (untimed): rank 0 builds the weight blob, one NCCL broadcast installs it everywhere ----
    blob = capi.synthesize_weights(K, SEED) if rank == 0 else None      # a deployment passes model_file instead
    blob = broadcast_blob(blob, src=0, device=dev)                        # the single collective of this workload
    prec = capi.PREC_TC_F16X2 if args.precision == "tc" else capi.PREC_FP32
    m = Model(H, W, K, D, max_batch=max_batch, device=local_rank, task_num=TASK_NUM, precision=prec, weights=blob)
    # the e2e context: same network, one-pair calls, room for the library to merge queued snb_infer_async calls into
    # passes of up to COALESCE_MAX pairs (capi.cu worker_main); and a max_batch = 1 context for the unmerged leg
    mc = m1 = None
    if not args.no_e2e:
        mc = m if max_batch == COALESCE_MAX else Model(H, W, K, D, max_batch=COALESCE_MAX, device=local_rank, task_num=TASK_NUM, precision=prec, weights=blob)
        m1 = m if max_batch == 1 else Model(H, W, K, D, max_batch=1, device=local_rank, task_num=TASK_NUM, precision=prec, weights=blob)
    del blob

    # ---- inputs: a rotating pool larger than L2, so no step finds its input cached ----
    in_pair, out_pair = 6 * H * W, 4 * H * W
    nstep = max(nloc, 1)
    pool_steps = max(2, -(-L2_BYTES // (in_pair * nstep)) + 1)
    pool_n = pool_steps * nstep
    base = torch.from_numpy(synth_inputs(min(pool_n, 8), H, W))
    host_in = base.repeat((pool_n + base.shape[0] - 1) // base.shape[0], 1, 1, 1)[:pool_n].contiguous()
    for i in range(pool_n):                          # make every pool entry distinct
        host_in[i, :, 0, 0] = i % 127
    host_in = host_in.pin_memory()
    d_in = host_in.to(dev)
    d_out = torch.empty((pool_n, 1, H, W), dtype=torch.int32, device=dev)
    host_out = torch.empty((pool_n, 1, H, W), dtype=torch.int32).pin_memory()
    stream = torch.cuda.Stream(dev)                  # non-default: the library launches on this handle
    nv12_ok = H % 2 == 0 and W % 2 == 0
    host_frames = None
    if nv12_ok and not args.no_e2e:
        rng = np.random.default_rng(SEED + rank)
        host_frames = torch.from_numpy(rng.integers(0, 256, (min(pool_n, 16), H * 3 // 2, 2 * W), dtype=np.uint8)).pin_memory()

    def step_device(i):
        if nloc == 0:
            return
        j = 0 if os.environ.get("SNB_BENCH_FIXED_IO") else (i % pool_steps) * nstep      # diagnostics: the same buffers every step
        m.infer_device(d_in[j:j + nloc], d_out[j:j + nloc], nloc, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def e2e_leg(model, nv12=False):
        """steps x nloc one-pair calls through the asynchronous C-ABI entry, TASK_NUM in flight, host buffers."""
        t0 = time.perf_counter()
        for i in range(args.steps * nloc):
            j = (3 + i) % pool_n
            if nv12:
                model.infer_nv12_async(host_frames[j % host_frames.shape[0]:j % host_frames.shape[0] + 1].numpy(), host_out[j:j + 1].numpy())
            else:
                model.infer_async(host_in[j:j + 1].numpy(), host_out[j:j + 1].numpy())
        model.wait_all()
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = ""
    # one sampler for the job (rank 0), running over warm-up + the device-timed loop and never during the e2e legs
    with ClockSampler(local_rank, enabled=rank == 0, uuid=uuid) as clk:
        for i in range(args.warmup):
            step_device(i)
        barrier()
        n0 = clk.count()
        ev0.record(stream)
        for i in range(args.steps):
            step_device(args.warmup + i)
        ev1.record(stream)
        torch.cuda.synchronize(dev)
        ms = ev0.elapsed_time(ev1)
        n_timed = clk.count() - n0
        if rank == 0 and clk.count() < 8:                          # a very short run: keep the same load up (untimed) until the clocks are characterised
            t_end = time.perf_counter() + 0.5
            i = 0
            while time.perf_counter() < t_end:
                step_device(i); i += 1
            torch.cuda.synchronize(dev)
    clocks = clk.summary()
    clocks["samples_in_timed_loop"] = n_timed
    barrier()
    nan = float("nan")
    e2e_s = e2e_1_s = e2e_sync_s = e2e_nv12_s = nan
    e2e_passes = 0
    e2e_nv12_passes = 0
    if not args.no_e2e:
        # ---- e2e: the reference-facing call with HOST buffers.  The node calls DnnNode::Run(is_sync=false) with
        # task_num = 4 calls in flight (stereonet_node.cpp:144,812): snb_infer_async, pinned buffers, copies timed.
        for mm in {id(x): x for x in (m1, mc)}.values():         # warm every context: the CUDA graph of every pass size it may use
            for b in range(1, mm.max_batch + 1):
                mm.infer(host_in[0:b].numpy(), host_out[0:b].numpy())
        for i in range(8):
            mc.infer_async(host_in[i:i + 1].numpy(), host_out[i:i + 1].numpy())
        mc.wait_all()
        barrier()
        passes0 = mc.pass_count()
        e2e_s = e2e_leg(mc)
        e2e_passes = mc.pass_count() - passes0
        barrier()
        e2e_1_s = e2e_leg(m1)                                    # same calls, one pass per call (max_batch = 1 context)
        barrier()
        if nv12_ok:
            mc.infer_nv12(host_frames[0:1].numpy(), host_out[0:1].numpy())
            barrier()
            passes0 = mc.pass_count()
            e2e_nv12_s = e2e_leg(mc, nv12=True)                  # raw camera frames in, pre-process on the GPU (half the H2D bytes)
            e2e_nv12_passes = mc.pass_count() - passes0
            barrier()
        t0 = time.perf_counter()
        for i in range(args.steps * nloc):                       # same through the synchronous call, one pair in flight
            j = (3 + i) % pool_n
            m1.infer(host_in[j:j + 1].numpy(), host_out[j:j + 1].numpy())
        torch.cuda.synchronize(dev)
        e2e_sync_s = time.perf_counter() - t0
    launches_per_pass = m.rt_stat().kernel_launches if nloc > 0 else 0

    t = torch.tensor([ms, e2e_s * 1e3, e2e_sync_s * 1e3, e2e_1_s * 1e3, e2e_nv12_s * 1e3], device=dev, dtype=torch.float64)
    t = torch.nan_to_num(t, nan=0.0)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_sync_ms, e2e_1_ms, e2e_nv12_ms = [float(v) for v in t]
    pairs_total = args.steps * global_batch                       # pairs processed by all ranks in the timed region
    rate = lambda tot_ms: pairs_total / (tot_ms * 1e-3) if tot_ms > 0 else None

    result = None
    if rank == 0:
        # ---- roofline of the dominant kernel.  Per-launch device times come from an eager pass with one CUDA event per
        # launch (snb_profile_pass, median of `reps`); inside the replayed CUDA graph launches overlap their neighbours'
        # tails (programmatic dependent launch), so the eager times are scaled by (graph step / eager sum) to make them add
        # up to the measured step: `achieved` and `frac` are quoted on that basis, the raw eager figure is kept beside it.
        pb = max(1, min(max_batch, nloc))
        passes_per_step = -(-nloc // max_batch) if nloc else 0
        prof = {}
        reps = 5
        for _ in range(reps):
            for name, kms, fl, by in m.profile_pass(pb):
                a = prof.setdefault(name, [[], fl, by])
                a[0].append(kms)
        prof = {n: [float(np.median(v[0])), v[1], v[2]] for n, v in prof.items()}
        eager_pass_ms = sum(v[0] for v in prof.values())
        graph_pass_ms = ms / args.steps * pb / max(nloc, 1)       # time of a pass of pb pairs inside the timed loop
        scale = graph_pass_ms / eager_pass_ms if eager_pass_ms > 0 else 1.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)       # kernel timed inside a long step
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        fam_of = lambda n: ("k_resblock_tc" if "[tc-block]" in n else "k_conv_stream" if "[tc-stream" in n else "k_cost3d" if "[tc-cost3d]" in n else
                            "k_conv_tc" if "[tc]" in n else "k_costvol" if n == "costvol" else
                            "k_refine_head" if ".head [" in n else "cuda_core_and_hbm")
        fam = {}
        for n, v in prof.items():
            a = fam.setdefault(fam_of(n), [0.0, 0.0, 0.0, 0])
            a[0] += v[0] * scale; a[1] += v[1]; a[2] += v[2]; a[3] += 1
        total_ms = sum(v[0] for v in fam.values())
        tc_fams = {k: v for k, v in fam.items() if k in ("k_resblock_tc", "k_conv_stream", "k_conv_tc", "k_cost3d") or args.precision != "tc"}
        dom = max(tc_fams, key=lambda k: tc_fams[k][0])
        d = fam[dom]
        tf = lambda v: v[1] / (v[0] * 1e-3) / 1e12 if v[0] > 0 else 0.0
        gbs = lambda v: v[2] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0
        traffic = ncu_traffic(dom, d192=D >= 96)
        hbm = {}
        for k in ("k_costvol", "k_refine_head"):      # the two HBM-bound kernels north_star names: cost-volume build, soft-argmin + upsample + refinement input
            if k in fam:
                hbm[k] = {"bound": "hbm", "achieved": gbs(fam[k]), "peak": hbm_peak, "unit": "GB/s", "frac": gbs(fam[k]) / hbm_peak,
                          "algorithmic_bytes_per_launch": fam[k][2] / max(fam[k][3], 1), "traffic": (ncu_traffic(k, d192=D >= 96) or {}).get("bytes_per_launch")}
        roofline = {"kernel": dom, "bound": "tensor", "achieved": tf(d), "peak": tf_peak, "unit": "TFLOP/s", "frac": tf(d) / tf_peak,
                    # split-fp16 operands: 3 fp16 MMAs are issued per algorithmic MAC, so the tensor pipe works at 3 x frac
                    "fp16_mma_issued_frac": 3 * tf(d) / tf_peak,
                    "traffic": traffic["bytes_per_launch"] if traffic else None, "traffic_source": traffic,
                    "peak_source": peak_src, "launches": d[3], "share_of_step": d[0] / total_ms,
                    "algorithmic_flops_per_launch": d[1] / max(d[3], 1), "avg_launch_ms": d[0] / max(d[3], 1),
                    "timing": {"pairs_per_pass": pb, "graph_pass_ms": graph_pass_ms, "eager_pass_ms_sum_of_launches": eager_pass_ms,
                               "scale_applied_to_eager_launch_times": scale},
                    "families": {k: {"ms": v[0], "share": v[0] / total_ms, "launches": v[3],
                                     "tflops": tf(v) if v[1] else None, "gbs": gbs(v)} for k, v in fam.items()},
                    "all_tensor_kernels": {"achieved": sum(v[1] for v in tc_fams.values()) / (sum(v[0] for v in tc_fams.values()) * 1e-3) / 1e12,
                                           "unit": "TFLOP/s"},
                    "whole_step": {"achieved": sum(v[1] for v in fam.values()) / (graph_pass_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                   "frac": sum(v[1] for v in fam.values()) / (graph_pass_ms * 1e-3) / 1e12 / tf_peak},
                    "hbm_kernels": hbm}
        n_calls = args.steps * nloc
        e2e_passes_all = 0 if args.no_e2e else 2 * n_calls + e2e_passes + e2e_nv12_passes     # per-call async leg + sync leg (one pass per call) + the merged legs
        result = {
            "metric": metric_name(c), "value": rate(ms), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": c["scaling"], "vs_baseline": None, "dtype": "f16x2->f32" if args.precision == "tc" else "f32",
            "data": "synthetic", "config": config_dict(key, c, world),
            "run": {"precision": args.precision, "pairs_per_step_all_gpus": global_batch, "pairs_per_step_rank0": nloc, "max_batch_per_pass": max_batch,
                    "l2": f"inputs rotate through a pool of {pool_n} tensors = {pool_n * in_pair / 2**20:.0f} MiB > 126 MiB L2",
                    "parallelism": f"{world} rank(s), one process per GPU, one NCCL weight broadcast at init, no per-frame collective"},
            "e2e": None if args.no_e2e else {
                "value": rate(e2e_ms), "unit": "pairs/s", "h2d_bytes_per_step": in_pair * global_batch, "d2h_bytes_per_step": out_pair * global_batch,
                "api": f"snb_infer_async, one pair per call, {TASK_NUM} calls in flight as the reference node (task_num = 4), pinned host "
                       f"buffers; the library merges queued calls into passes of <= {COALESCE_MAX} pairs "
                       f"(rank 0: {e2e_passes} passes for {n_calls} calls)",
                "one_pass_per_call_value": rate(e2e_1_ms),
                "one_pass_per_call_api": f"snb_infer_async on a max_batch = 1 context (no merging), {TASK_NUM} calls in flight",
                "nv12_value": rate(e2e_nv12_ms) if nv12_ok else None,
                "nv12_api": "snb_infer_nv12_async: raw side-by-side NV12 frames in (3 B per pixel pair instead of 6), pre-process on the GPU, merging on",
                "nv12_h2d_bytes_per_step": 3 * H * W * global_batch if nv12_ok else None,
                "sync_value": rate(e2e_sync_ms), "sync_api": "snb_infer (one call in flight)"},
            # kernels of this library launched by rank 0 inside the timed regions: device-resident loop + the e2e legs
            "gpu_launches": launches_per_pass * (args.steps * passes_per_step + e2e_passes_all),
            "clocks": clocks, "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:
            result["cpu_baseline"] = cpu_baseline(c, max_pairs=2 if not heavy else 1)
    for mm in {id(x): x for x in (m, mc, m1) if x is not None}.values():
        mm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


if __name__ == "__main__":
    main()
