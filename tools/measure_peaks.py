"""Measures the TF32 / FP32-FMA / fp16 matmul peaks of this B200 the same way MEASURED_PEAKS.json measured bf16
(torch.matmul 8192^3, best of 10, CUDA events) — SURVEY.md §8(d) asks for them next to every roofline fraction.
Library GEMMs are used for the DENOMINATOR only.  Writes profiles/r02_measured_peaks_tf32_fp32.json."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = 8192
dev = torch.device("cuda:0")
out = {"how": "torch.matmul 8192^3 (2 n^3 flops), best of 10, CUDA events", "gpu": torch.cuda.get_device_name(0)}
for name, dtype, tf32 in (("fp32_fma_tflops", torch.float32, False), ("tf32_tflops", torch.float32, True),
                          ("fp16_tflops", torch.float16, False), ("bf16_tflops", torch.bfloat16, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dtype); b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
torch.backends.cuda.matmul.allow_tf32 = False
print(json.dumps(out))
with open(os.path.join(ROOT, "gpurun_out", "r02_measured_peaks_tf32_fp32.json"), "w") as f:
    json.dump(out, f, indent=1)
