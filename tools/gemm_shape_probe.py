"""What can a library GEMM do on the fine net's layer shape on this box?  (calibration, not product)

Times torch.matmul (cuBLAS, fp16) and the engine's dense kernels on C[M,1024] = A[M,1024]·W[1024,1024]^T in a
27-launch ping-pong loop (the fine net's structure: every launch reads the previous launch's output), M = 262144
(2048 rays x 128 samples), for several seconds so the clocks settle under the power cap."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from mofanerf_b200 import get_engine  # noqa: E402

dev = "cuda:0"
M, N, K = 262144, 1024, 1024
g = torch.Generator().manual_seed(0)
A = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
B = (torch.randn(M, K, generator=g) * 0.5).half().to(dev)
W = (torch.randn(N, K, generator=g) * (2.0 / K) ** 0.5).half().to(dev)   # He init: ReLU activations keep their scale
bias = torch.zeros(N, device=dev)
eng = get_engine(dev)
flops = 2.0 * M * N * K


def loop(fn, secs=4.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    n = 0
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(27):
            fn()
        n += 27
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return flops * n / (e0.elapsed_time(e1) / 1e3) / 1e12


state = {"a": A, "b": B}


def cublas():
    torch.matmul(state["a"], W.t(), out=state["b"])
    torch.relu_(state["b"])          # keep the data alive (all-zero operands draw less power and clock higher)
    state["a"], state["b"] = state["b"], state["a"]


def make(mode):
    def f():
        out = eng.dense(state["a"], W, bias, relu=True, mode=mode)
        state["a"] = out
    return f


res = {"shape": [M, N, K], "note": "cuBLAS figure includes a separate relu_ kernel per GEMM (~0.17 ms of ~0.45)",
       "cublas_fp16_plus_relu_tflops": loop(cublas)}


def cublas_only():
    torch.matmul(A, W.t(), out=B)


res["cublas_fp16_gemm_only_tflops"] = loop(cublas_only)
state["a"] = A
res["engine_pair_tflops"] = loop(make("default"))
state["a"] = A
res["engine_1cta_tflops"] = loop(make("1cta"))
print(json.dumps(res))
