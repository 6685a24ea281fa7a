"""Accuracy of the branch-free GELU used by the split-bf16 epilogues (csrc/common.cuh gelu_erf: erfc by Abramowitz & Stegun
7.1.26 with one rcp and one ex2), emulated in float32 numpy against the float64 erf form."""
import math

import numpy as np
from scipy.special import erf as erf64

f = np.float32
A = [0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429]


def gelu_as(x):
    x = x.astype(f)
    z = (np.abs(x) * f(0.70710678118654752440)).astype(f)
    t = (f(1) / (f(0.3275911) * z + f(1)).astype(f)).astype(f)
    p = f(0.5 * A[4])
    for c in (A[3], A[2], A[1], A[0]):
        p = (p * t + f(0.5 * c)).astype(f)
    p = (p * t).astype(f)
    e = np.exp2((z * (z * f(-1.4426950408889634))).astype(f)).astype(f)
    h = (p * e).astype(f)
    return (x * np.where(x >= 0, (f(1) - h).astype(f), h)).astype(f)


def gelu_as_packed(x):
    """common.cuh gelu_erf2 (the packed FFMA2 form the GEMM epilogue runs): z keeps its sign, exp argument (z*z)*c,
    Phi = 0.5 + copysign(0.5 - h, x)."""
    x = x.astype(f)
    z = (x * f(0.70710678118654752440)).astype(f)
    t = (f(1) / (np.abs(z) * f(0.3275911) + f(1)).astype(f)).astype(f)
    p = f(0.5 * A[4])
    for c in (A[3], A[2], A[1], A[0]):
        p = (p * t + f(0.5 * c)).astype(f)
    p = (p * t).astype(f)
    e = np.exp2(((z * z).astype(f) * f(-1.4426950408889634)).astype(f)).astype(f)
    a = (f(0.5) - (p * e).astype(f)).astype(f)
    s = (a.view(np.uint32) | (x.view(np.uint32) & np.uint32(0x80000000))).view(f)
    return (x * (s + f(0.5)).astype(f)).astype(f)


if __name__ == "__main__":
    import torch

    x = np.linspace(-12, 12, 4000001).astype(f)
    gref = 0.5 * x.astype(np.float64) * (1 + erf64(x.astype(np.float64) / math.sqrt(2)))
    print("gelu_erf (A&S 7.1.26) max abs err %.3e" % np.abs(gelu_as(x) - gref).max())
    print("gelu_erf2 (packed form)  max abs err %.3e" % np.abs(gelu_as_packed(x) - gref).max())
    print("torch fp32 gelu       max abs err %.3e" % np.abs(torch.nn.functional.gelu(torch.from_numpy(x)).numpy() - gref).max())
