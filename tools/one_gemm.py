"""Scratch: run one MLP GEMM shape a few times (for ncu captures). usage: one_gemm.py M N K epi bn"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gdrnpp_bop2022_b200 import _lib  # noqa: E402
M, N, K, epi, bn = [int(a) for a in sys.argv[1:6]]
dev = torch.device("cuda:0")
L = _lib.lib()
A = torch.randn(M, K, device=dev).bfloat16()
W = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
bias = torch.randn(N, device=dev); gamma = torch.rand(N, device=dev)
out = torch.zeros(M, N, dtype=torch.float32 if epi == 2 else torch.bfloat16, device=dev)
for i in range(4):
    rc = L.gdrn_gemm_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(gamma), _lib.ptr(out), _lib.ptr(out),
                          M, N, K, epi, 0, bn, _lib.current_stream())
    assert rc == 0, _lib.last_error()
torch.cuda.synchronize()
