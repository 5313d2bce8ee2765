"""Measured TF32 tensor-core peak on this box, by the recipe MEASURED_PEAKS.json uses for bf16: the library GEMM (cuBLAS
through torch.matmul, allow_tf32) on 8192^3, CUDA-event timed after warm-up — burst (one GEMM timed alone x 20) and
sustained (200 back to back).  Printed as JSON; the denominator for every tf32 'fraction of peak' in DESIGN.md."""
import json, torch
torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda"); b = torch.randn(n, n, device="cuda")
for _ in range(10): a @ b
torch.cuda.synchronize()
def run(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): a @ b
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
burst = min(run(1) for _ in range(20))
sustained = run(200)
fl = 2 * n ** 3
ab, bb = a.bfloat16(), b.bfloat16()
for _ in range(10): ab @ bb
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ab @ bb
e1.record(); torch.cuda.synchronize()
print(json.dumps({"tf32_tflops_burst": round(fl / (burst * 1e-3) / 1e12, 1), "tf32_tflops_sustained": round(fl / (sustained * 1e-3) / 1e12, 1),
                  "bf16_tflops_same_box": round(fl / (e0.elapsed_time(e1) / 50 * 1e-3) / 1e12, 1),
                  "how": "torch.matmul (cuBLAS, allow_tf32) 8192^3 f32, CUDA events; burst = best of 20 single launches, sustained = 200 back to back"}))
