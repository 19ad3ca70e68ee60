"""Ad-hoc interleaved A/B of the p2p launch shapes through the library (20 back-to-back steps per sample)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from moptimizer_0_b200 import capi

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
n = 100_000_000
ctx = capi.Context(0)
st = capi.Store(ctx, capi.MODEL_POINT2POINT, n, capi.F32)
st.generate(seed=2, gt=[0.5, -0.3, 0.2, 0.10, -0.05, 0.08], noise_sigma=0.01, outlier_fraction=0.05, outlier_range=1.0)
prob = capi.make_problem(capi.MODEL_POINT2POINT, capi.JAC_ANALYTICAL, capi.F32, loss=capi.LOSS_HUBER, loss_param=0.05)
x0 = np.zeros(6)
stream = torch.cuda.ExternalStream(ctx.stream())
res = {0: [], 256: []}
for rep in range(12):
    for threads in (0, 256):
        ctx.set_launch(0, threads)
        for _ in range(3):
            ctx.linearize_async(st, prob, x0)
        ctx.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(20):
            ctx.linearize_async(st, prob, x0)
        b.record(stream)
        ctx.synchronize()
        res[threads].append(a.elapsed_time(b) / 20 * 1e3)
    print(rep, f"1024: {res[0][-1]:.1f} us   256: {res[256][-1]:.1f} us   sm {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz mem {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)} MHz "
          f"T {pynvml.nvmlDeviceGetTemperature(h, 0)} C P {pynvml.nvmlDeviceGetPowerUsage(h)/1000:.0f} W", flush=True)
for k, v in res.items():
    print(k, "median", np.median(v), "min", np.min(v), "max", np.max(v))
