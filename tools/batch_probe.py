"""Device-resident pairs/s of configs[1] (540x960 K=3 D=24) as a function of the batch per pass (CUDA graph replays)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hobot_stereonet_b200 import Model, capi
H, W, K, D = 540, 960, 3, 24
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
blob = capi.synthesize_weights(K, 1234)
for B in (1, 2, 3, 4, 8):
    m = Model(H, W, K, D, max_batch=B, device=0, task_num=4, precision=capi.PREC_TC_F16X2, weights=blob)
    pool = 8
    d_in = torch.randint(-128, 127, (pool, B, 6, H, W), dtype=torch.int8, device=dev)
    d_out = torch.empty((pool, B, 1, H, W), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)
    for i in range(5): m.infer_device(d_in[i % pool], d_out[i % pool], B, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 60
    e0.record(st)
    for i in range(steps): m.infer_device(d_in[i % pool], d_out[i % pool], B, st.cuda_stream)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"batch {B}: {ms:.3f} ms/pass, {B / ms * 1e3:.1f} pairs/s", flush=True)
    m.close(); del d_in, d_out; torch.cuda.empty_cache()
