"""Measure the pipe peaks MEASURED_PEAKS.json does not hold: cuBLAS TF32 and FP32 (FFMA) GEMM."""
import json, torch, time
def bench(fn, flops, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return flops / best / 1e9
n = 8192
A = torch.randn(n, n, device="cuda"); B = torch.randn(n, n, device="cuda")
out = {}
torch.backends.cuda.matmul.allow_tf32 = True
out["tf32_tflops"] = bench(lambda: A @ B, 2 * n ** 3)
torch.backends.cuda.matmul.allow_tf32 = False
out["fp32_tflops"] = bench(lambda: A @ B, 2 * n ** 3)
Ah, Bh = A.bfloat16(), B.bfloat16()
out["bf16_tflops"] = bench(lambda: Ah @ Bh, 2 * n ** 3)
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda"); y = torch.empty_like(x)
out["copy_gbs"] = bench(lambda: y.copy_(x), 2 * (1 << 30)) * 1e3 / 1e3
print(json.dumps(out))
