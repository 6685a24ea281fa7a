"""Scratch: per-role wait accounting of the GEMM kernel on the MLP shapes (GDRN_GEMM_TRACE=1)."""
import os, sys
os.environ["GDRN_GEMM_TRACE"] = "1"
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import _lib  # noqa: E402
dev = torch.device("cuda:0")
L = _lib.lib()
SHAPES = [("s0_fc1", 262144, 512, 128, 1, 256), ("s0_fc2", 262144, 128, 512, 2, 128),
          ("s1_fc1", 65536, 1024, 256, 1, 256), ("s1_fc2", 65536, 256, 1024, 2, 256),
          ("s2_fc1", 16384, 2048, 512, 1, 256), ("s2_fc2", 16384, 512, 2048, 2, 256),
          ("s2_fc1_store", 16384, 2048, 512, 0, 256),
          ("s3_fc1", 4096, 4096, 1024, 1, 256), ("s3_fc2", 4096, 1024, 4096, 2, 256)]
for name, M, N, K, epi, bn in SHAPES:
    A = torch.randn(M, K, device=dev).bfloat16()
    W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=dev); gamma = torch.rand(N, device=dev)
    out = torch.zeros(M, N, dtype=torch.float32 if epi == 2 else torch.bfloat16, device=dev)
    for i in range(3):
        sys.stderr.write(name + " ")
        sys.stderr.flush()
        rc = L.gdrn_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(out), _lib.ptr(out),
                              M, N, K, epi, 0, bn, _lib.current_stream())
        assert rc == 0, _lib.last_error()
    del A, W, out
