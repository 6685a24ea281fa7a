"""Scratch: the ConvNeXt MLP GEMM shapes of one step (B=64) through gdrn_gemm_bf16 vs torch.matmul (cuBLAS)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
# (name, M, N, K, epi, block_n)
SHAPES = [
    ("s0_fc1", 262144, 512, 128, 1, 256), ("s0_fc2", 262144, 128, 512, 2, 128),
    ("s1_fc1", 65536, 1024, 256, 1, 256), ("s1_fc2", 65536, 256, 1024, 2, 256),
    ("s2_fc1", 16384, 2048, 512, 1, 256), ("s2_fc2", 16384, 512, 2048, 2, 256),
    ("s3_fc1", 4096, 4096, 1024, 1, 256), ("s3_fc2", 4096, 1024, 4096, 2, 256),
]
res = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, M, N, K, epi, bn in SHAPES:
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev)
    gamma = torch.rand(N, device=dev)
    resid = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, dtype=torch.float32 if epi == 2 else torch.bfloat16, device=dev)

    def run():
        rc = L.gdrn_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(resid if epi == 2 else out),
                              _lib.ptr(resid if epi == 2 else out), M, N, K, epi, 0, bn, _lib.current_stream())
        assert rc == 0, _lib.last_error()

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]
    ms = timeit(run)
    ms_cublas = timeit(lambda: torch.matmul(A, W.t()))
    fl = 2.0 * M * N * K
    res[name] = {"us": ms * 1e3, "tflops": fl / ms / 1e9, "cublas_us": ms_cublas * 1e3, "cublas_tflops": fl / ms_cublas / 1e9}
    print(name, json.dumps(res[name]), flush=True)
    del A, W, resid, out
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = os.environ.get("TAG", "base")
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_shapes_%s.json" % tag), "w"), indent=1)
