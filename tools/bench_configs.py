"""Device-resident throughput of the other BASELINE.json configurations on ONE GPU (bench.py measures configs[1], the
headline; these are the per-GPU shards of configs[2..4] and the deployed shape).  CUDA events around `steps` graph replays
of the whole pass at the configuration's per-GPU batch, inputs resident in HBM.  Prints one JSON line per configuration."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hobot_stereonet_b200 import Model, capi

CONFIGS = [
    ("configs[1] SceneFlow 540x960 D=24 batch 1", 540, 960, 3, 24, 1),
    ("configs[2] ZED-2i 720x1280 D=48 batch 8", 720, 1280, 3, 48, 8),
    ("configs[3] SceneFlow 540x960 D=192, 4 pairs per GPU (batch 32 over 8 GPUs)", 540, 960, 3, 192, 4),
    ("configs[4] KITTI 375x1242 D=192, 8 pairs per GPU (batch 64 over 8 GPUs)", 375, 1242, 3, 192, 8),
    ("deployed shape 720x1280 K=4 D=12 batch 1", 720, 1280, 4, 12, 1),
]

def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    for name, H, W, K, D, B in CONFIGS:
        blob = capi.synthesize_weights(K, 1234)
        m = Model(H, W, K, D, max_batch=B, device=0, task_num=4, precision=capi.PREC_TC_F16X2, weights=blob)
        pool = 4
        d_in = torch.randint(-128, 127, (pool, B, 6, H, W), dtype=torch.int8, device=dev)
        d_out = torch.empty((pool, B, 1, H, W), dtype=torch.int32, device=dev)
        st = torch.cuda.Stream(dev)
        for i in range(3):
            m.infer_device(d_in[i % pool], d_out[i % pool], B, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            m.infer_device(d_in[i % pool], d_out[i % pool], B, st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        flops = sum(f for _, _, f, _ in m.profile_pass(B))
        print(json.dumps({"config": name, "H": H, "W": W, "K": K, "D": D, "batch_per_gpu": B, "ms_per_batch": round(ms, 3),
                          "pairs_per_s": round(B / ms * 1e3, 1), "algorithmic_tflops": round(flops / ms * 1e-9, 1) if flops else None,
                          "hbm_gb_allocated": round(torch.cuda.memory_allocated() / 2**30, 2)}), flush=True)
        m.close()
        del d_in, d_out
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
