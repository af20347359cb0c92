#!/usr/bin/env python
"""Numerics experiment: how does the tcgen05 engine's error grow with K, and does adding the K-chunks
outside the tensor core (FP32 round-to-nearest, via the Cin path of ust_test_cgemm) remove the growth?

Prints, for an (M x K) x (K x N) complex64 product, the normwise relative error against a complex128
reference and the signed bias  Re<C - Cref, Cref> / |Cref|^2  for
  simt            : FP32 FMA engine
  tc (one shot)   : all K accumulated in TMEM
  tc sliced by S  : C += A[:, k:k+S] B[k:k+S, :] one launch per slice, slices added in the epilogue
  tc2             : TMA-fed engine, leading products drained to FP32 registers every 16 k (with / without the
                    first-order truncation-bias correction UST_TC2_BIAS_FIX)
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveforminversionust_b200 import _lib  # noqa: E402

L = _lib.lib()


def gemm(engine, A, B, Cin, k0, k1):
    M, N = A.shape[0], B.shape[1]
    out = torch.empty((M, N), dtype=torch.complex64, device="cuda")
    Ak = A[:, k0:k1]
    Bk = B[k0:k1, :]
    rc = L.ust_test_cgemm(engine, 0, M, N, k1 - k0, C.c_void_p(Ak.data_ptr()), A.shape[1], C.c_void_p(Bk.data_ptr()), N,
                          C.c_void_p(Cin.data_ptr()) if Cin is not None else None, N, C.c_void_p(out.data_ptr()), N,
                          C.c_float(1.0), 0, 0, 0, 0, None)
    _lib.check(rc, "ust_test_cgemm")
    return out


def stats(got, ref):
    d = got.to(torch.complex128) - ref
    err = float(torch.linalg.norm(d) / torch.linalg.norm(ref))
    bias = float((d * ref.conj()).real.sum() / (ref.abs() ** 2).sum())
    return err, bias


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    for kind in ("randn", "positive"):
        for K in (64, 256, 512, 2048):
            M, N = 256, 256
            if kind == "randn":
                mk = lambda *s: torch.complex(torch.randn(*s, generator=g, device="cuda"), torch.randn(*s, generator=g, device="cuda"))
            else:  # coherent sums: the accumulator grows monotonically, truncation bias is visible
                mk = lambda *s: torch.complex(torch.rand(*s, generator=g, device="cuda") + 0.5, 0.1 * torch.randn(*s, generator=g, device="cuda"))
            A, B = mk(M, K), mk(K, N)
            ref = A.to(torch.complex128) @ B.to(torch.complex128)
            row = [f"{kind:8s} K={K:5d}"]
            e, b = stats(gemm(1, A, B, None, 0, K), ref)
            row.append(f"simt {e:.2e}/{b:+.1e}")
            e, b = stats(gemm(2, A, B, None, 0, K), ref)
            row.append(f"tc {e:.2e}/{b:+.1e}")
            for fix in ("0", "2.5e-8"):
                os.environ["UST_TC2_BIAS_FIX"] = fix
                e, b = stats(gemm(3, A, B, None, 0, K), ref)
                row.append(f"tc2[fix={fix}] {e:.2e}/{b:+.1e}")
            os.environ["UST_TC2_BIAS_FIX"] = "0"
            for S in (16, 64):
                if S >= K:
                    continue
                acc = None
                for k0 in range(0, K, S):
                    acc = gemm(2, A, B, acc, k0, k0 + S)
                e, b = stats(acc, ref)
                row.append(f"tc/{S} {e:.2e}/{b:+.1e}")
            print("  ".join(row), flush=True)


if __name__ == "__main__":
    main()
