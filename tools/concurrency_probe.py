"""How much does a second (third) stereo pair in flight on the SAME GPU add?  N contexts, N streams, alternating launches.
Diagnostic for the frame-level concurrency of snb_infer_async (the reference keeps task_num = 4 calls in flight)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hobot_stereonet_b200 import Model, capi

H, W, K, D = 540, 960, 3, 24
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
blob = capi.synthesize_weights(K, 1234)
for nctx in (1, 2, 3, 4):
    ms_ = [Model(H, W, K, D, max_batch=1, device=0, task_num=4, precision=capi.PREC_TC_F16X2, weights=blob) for _ in range(nctx)]
    streams = [torch.cuda.Stream(dev) for _ in range(nctx)]
    pool = 64
    d_in = torch.randint(-128, 127, (pool, 6, H, W), dtype=torch.int8, device=dev)
    d_out = torch.empty((pool, 1, H, W), dtype=torch.int32, device=dev)
    steps = 120
    for i in range(8):
        ms_[i % nctx].infer_device(d_in[i % pool], d_out[i % pool], 1, streams[i % nctx].cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        ms_[i % nctx].infer_device(d_in[i % pool], d_out[i % pool], 1, streams[i % nctx].cuda_stream)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"contexts {nctx}: {steps / dt:.1f} pairs/s ({dt / steps * 1e3:.3f} ms/pair)", flush=True)
    for m in ms_:
        m.close()
