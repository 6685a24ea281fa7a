"""fc1 -> fc2 of one ConvNeXt block in split-bf16 mode, with the rows processed in `chunks` pieces through ONE reused hidden
buffer (so that the 4C-wide hidden activation of a piece stays in L2 between the two GEMMs).  us per block for each chunk count.
Usage (under gpurun): python tools/mlp_chunk_bench.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import _lib as L  # noqa: E402

STAGES = [("s0", 262144, 128), ("s1", 65536, 256), ("s2", 16384, 512)]


def main():
    dev = torch.device("cuda:0")
    lib = L.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    for name, M, C in STAGES:
        g = torch.Generator().manual_seed(M)
        A = torch.randn(M, 2 * C, generator=g).to(dev).bfloat16()
        W1 = (torch.randn(4 * C, 2 * C, generator=g) / np.sqrt(C)).to(dev).bfloat16()
        W2 = (torch.randn(C, 8 * C, generator=g) / np.sqrt(4 * C)).to(dev).bfloat16()
        b1, b2, gam = torch.randn(4 * C).to(dev), torch.randn(C).to(dev), torch.rand(C).to(dev)
        X = torch.zeros(M, C, device=dev)
        bn2 = 256 if C >= 256 else 128
        for chunks in (1, 2, 4, 8, 16):
            Mc = M // chunks
            if Mc % 256:
                continue
            Hb = torch.zeros(Mc, 8 * C, dtype=torch.bfloat16, device=dev)

            def block():
                for c in range(chunks):
                    a = A[c * Mc:(c + 1) * Mc]
                    x = X[c * Mc:(c + 1) * Mc]
                    L.check(lib.gdrn_gemm_x3(L.ptr(a), L.ptr(W1), L.ptr(b1), None, None, L.ptr(Hb), Mc, 4 * C, C, 1, 256,
                                             L.current_stream()), "fc1")
                    L.check(lib.gdrn_gemm_x3(L.ptr(Hb), L.ptr(W2), L.ptr(b2), L.ptr(gam), L.ptr(x), L.ptr(x), Mc, C, 4 * C, 2, bn2,
                                             L.current_stream()), "fc2")
            for _ in range(2):
                block()
            torch.cuda.synchronize()
            # capture the block in a CUDA graph (as the model step is) so that host launch cost does not count
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                block()
            tot, n = 0.0, 8
            for _ in range(n):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gr.replay(); e1.record(); e1.synchronize()
                tot += e0.elapsed_time(e1)
            res.append({"stage": name, "chunks": chunks, "us_per_block": round(tot / n * 1e3, 1)})
            print(res[-1], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
