"""cuBLAS FP64 GEMM ceiling on this box (the DGEMM figure SURVEY.md §8(d) asks for).

Shapes: one large square GEMM (library ceiling) and the two skinny CAVI shapes
(K=100 haplotypes) so the roofline of the contraction kernels can be quoted against
what the vendor library reaches on the same shape.
"""
import json
import torch


def bench(m, n, k, reps=10):
    a = torch.randn(m, k, dtype=torch.float64, device="cuda")
    b = torch.randn(k, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return {"m": m, "n": n, "k": k, "ms": best, "tflops": 2.0 * m * n * k / best * 1e-9}


if __name__ == "__main__":
    out = [bench(8192, 8192, 8192), bench(4096, 4096, 4096),
           bench(100000, 100, 5000),   # X = E . h^T  (N x LB) (LB x K), config-5 shape
           bench(100, 5000, 100000),   # C = wz^T . E (K x N) (N x LB)
           bench(5000, 100, 1005), bench(100, 1005, 5000)]
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "dgemm": out}))
