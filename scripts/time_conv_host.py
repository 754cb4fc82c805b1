"""Host time per tcgen05 convolution launch, with and without the launch context (pmu_ctx: cached TMA descriptors, kernel
attributes, device properties).  The C entry point is called directly through ctypes with prebuilt arguments, 500 calls
back to back (below the launch-queue depth, so the host never waits for the GPU); also the same through the Python op."""
import os, sys, time
from ctypes import c_void_p
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmu_b200
from pmu_b200 import ops, _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
st = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
N = 500


def p(t):
    return None if t is None else c_void_p(t.data_ptr())


cases = []
# (name, x0, x1, Cout, ntaps): a 3x3 layer, a two-source (skip concat) layer, a transposed convolution (4 output maps)
x = torch.randn(1, 16, 16, 64, device=dev).half()
cases.append(("conv3x3 64->64 @16x16", x, None, 64, 9))
cases.append(("conv3x3 64+64->64 @16x16 (two sources)", x, x.clone(), 64, 9))
xt = torch.randn(1, 8, 8, 128, device=dev).half()
cases.append(("convT 128->64 @8x8 (4 output phases)", xt, None, 64, 4))
for name, x0, x1, Cout, ntaps in cases:
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[3]
    K = (9 if ntaps == 9 else 1) * (C0 + C1)
    Ntot = (4 if ntaps == 4 else 1) * Cout
    w = (torch.randn(Ntot, K, device=dev) * 0.05).half()
    bias = torch.zeros(Cout, device=dev)
    oh, ow = (2 * H, 2 * W) if ntaps == 4 else (H, W)
    y = torch.empty(B, oh, ow, Cout, device=dev, dtype=torch.float16)
    args = (p(x0), C0, p(x1), C1, p(w), p(bias), p(y), B, H, W, Cout, ntaps, 0, 1, 0, 0, st)
    res = {}
    for use_ctx in (False, True):
        _lib.USE_CTX = use_ctx
        _lib.bind_device(lib, 0)
        for _ in range(20):
            assert lib.pmu_conv_gemm_bf16(*args) == 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(N):
            lib.pmu_conv_gemm_bf16(*args)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        for _ in range(20):
            ops.conv_gemm_bf16(x0, w, bias, Cout, ntaps, False, x1=x1)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for _ in range(N):
            ops.conv_gemm_bf16(x0, w, bias, Cout, ntaps, False, x1=x1)
        t3 = time.perf_counter()
        torch.cuda.synchronize()
        res[use_ctx] = ((t1 - t0) / N * 1e6, (t3 - t2) / N * 1e6)
    print(f"{name:48s} C call: {res[False][0]:6.2f} us -> {res[True][0]:6.2f} us with pmu_ctx | Python op: {res[False][1]:6.2f} -> {res[True][1]:6.2f} us")
print("context stats (tensor maps, hits, misses):", _lib.ctx_stats(0))
