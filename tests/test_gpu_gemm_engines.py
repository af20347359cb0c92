"""Block-GEMM engines in isolation (ust_test_cgemm) against a float64 torch reference of the same op:
the SIMT fp32 engine and the tcgen05 engine (BF16x3 split, FP32-accurate)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ENG = {"simt": 1, "tc2": 3, "tc2h": 4}  # tc2h: the 128 x 64-tile, two-CTAs-per-SM form used by the Gauss-Jordan kernels


def _run(torch, engine, ta, M, N, K, with_cin=True, mask=(0, 0), skip=(0, 0), sgn=-1.0, seed=0, pad=0):
    from waveforminversionust_b200 import _lib
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    cplx = lambda *s: torch.complex(torch.randn(*s, generator=g, device="cuda"), torch.randn(*s, generator=g, device="cuda"))
    lda = (M if ta else K) + pad
    A = cplx(K if ta else M, lda)
    ldb = N + pad
    B = cplx(K, ldb)
    Cin = cplx(M, N + pad) if with_cin else None
    Cout = torch.full((M, N + pad), 7.0 + 7.0j, dtype=torch.complex64, device="cuda")
    rc = L.ust_test_cgemm(ENG[engine], int(ta), M, N, K, C.c_void_p(A.data_ptr()), lda, C.c_void_p(B.data_ptr()), ldb,
                          C.c_void_p(Cin.data_ptr()) if with_cin else None, N + pad, C.c_void_p(Cout.data_ptr()), N + pad,
                          C.c_float(sgn), mask[0], mask[1], skip[0], skip[1], None)
    _lib.check(rc, "ust_test_cgemm")
    torch.cuda.synchronize()
    A64 = A.to(torch.complex128)
    opA = A64[:, :M].conj().T if ta else A64[:, :K]
    ref = sgn * (opA @ B.to(torch.complex128)[:, :N])
    if with_cin:
        c = Cin.to(torch.complex128)[:, :N].clone()
        c[:, mask[0]:mask[1]] = 0
        ref = ref + c
    got = Cout[:, :N].to(torch.complex128)
    keep = torch.ones(M, dtype=torch.bool, device="cuda")
    keep[skip[0]:skip[1]] = False
    err = float(torch.linalg.norm((got - ref)[keep]) / torch.linalg.norm(ref[keep]))
    untouched = bool((Cout[~keep][:, :N] == (7.0 + 7.0j)).all()) and bool((Cout[:, N:] == (7.0 + 7.0j)).all())
    return err, untouched


@pytest.mark.parametrize("engine", ["simt", "tc2", "tc2h"])
@pytest.mark.parametrize("ta", [False, True])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (192, 40, 190), (510, 256, 510), (64, 6, 30)])
def test_cgemm_engine_matches_float64_reference(engine, ta, M, N, K):
    import torch
    if engine == "tc2h" and ta:
        pytest.skip("the 128 x 64 form is forward only (Gauss-Jordan panels and updates)")
    err, untouched = _run(torch, engine, ta, M, N, K)
    print(f"{engine} ta={ta} {M}x{N}x{K}: rel err {err:.3e}")
    assert err < 3e-6 and untouched
    err2, _ = _run(torch, engine, ta, M, N, K, with_cin=False, sgn=1.0, seed=3)
    assert err2 < 3e-6


@pytest.mark.parametrize("engine", ["simt", "tc2", "tc2h"])
def test_cgemm_engine_mask_and_unaligned(engine):
    import torch
    err, untouched = _run(torch, engine, False, 320, 320, 64, mask=(128, 192), pad=0)
    assert err < 3e-6 and untouched
    err, untouched = _run(torch, engine, False, 130, 67, 77, pad=1)  # odd leading dimensions: scalar paths
    assert err < 3e-6 and untouched
    if engine == "tc2h":
        return
    err, untouched = _run(torch, engine, True, 130, 67, 77, pad=1)
    assert err < 3e-6 and untouched


@pytest.mark.parametrize("engine", ["tc2", "tc2h"])
def test_tc_engine_row_skip(engine):
    import torch
    err, untouched = _run(torch, engine, False, 512, 512, 64, mask=(64, 128), skip=(64, 128))
    assert err < 3e-6 and untouched
