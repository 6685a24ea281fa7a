"""Stage-0 ConvNeXt MLP half-block in split-bf16 mode: fused kernel (gdrn_mlp_fused_x3) vs the two-GEMM path, us per block.
Usage (under gpurun): [GDRN_MLP_TRACE=1] python tools/mlp_fused_x3_bench.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import _lib as L  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = L.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    M, C = 262144, 128
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, 2 * C, generator=g).to(dev).bfloat16()
    W1 = (torch.randn(4 * C, 2 * C, generator=g) / np.sqrt(C)).to(dev).bfloat16()
    W2 = (torch.randn(C, 8 * C, generator=g) / np.sqrt(4 * C)).to(dev).bfloat16()
    b1, b2, gam = torch.randn(4 * C).to(dev), torch.randn(C).to(dev), torch.rand(C).to(dev)
    X = torch.zeros(M, C, device=dev)
    Hb = torch.zeros(M, 8 * C, dtype=torch.bfloat16, device=dev)
    st = L.current_stream()

    def unfused():
        L.check(lib.gdrn_gemm_x3(L.ptr(A), L.ptr(W1), L.ptr(b1), None, None, L.ptr(Hb), M, 4 * C, C, 1, 256, st), "fc1")
        L.check(lib.gdrn_gemm_x3(L.ptr(Hb), L.ptr(W2), L.ptr(b2), L.ptr(gam), L.ptr(X), L.ptr(X), M, C, 4 * C, 2, 128, st), "fc2")

    def fused():
        L.check(lib.gdrn_mlp_fused_x3(L.ptr(A), L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), L.ptr(gam), L.ptr(X), M, C, st), "fused")

    for name, fn in (("unfused", unfused), ("fused", fused)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot, n = 0.0, 10
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            tot += e0.elapsed_time(e1)
        print(name, "us per block: %.1f" % (tot / n * 1e3), flush=True)


if __name__ == "__main__":
    main()
