"""Measure the dtype-specific GEMM peaks MEASURED_PEAKS.json does not hold (fp64 / fp32 / tf32
torch.matmul 8192^3, best of 5, CUDA events) — the roofline denominators of SURVEY.md §8(d)."""
import json
import torch


def peak(dtype, n=8192, tf32=False, reps=5):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n**3 / (best * 1e-3) / 1e12


if __name__ == "__main__":
    out = {
        "gpu": torch.cuda.get_device_name(0),
        "fp64_tflops": peak(torch.float64),
        "fp32_tflops": peak(torch.float32),
        "tf32_tflops": peak(torch.float32, tf32=True),
        "how": "torch.matmul 8192^3, best of 5, CUDA events",
    }
    print(json.dumps(out))
