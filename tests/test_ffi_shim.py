"""csrc/xla_ffi_shim.cc: compiled against the test double of the XLA FFI header (tests/xla_ffi_stub/, jaxlib is not
installable here) and, on the GPU, driven through hand-filled call frames against the plain C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

from common import observed_data, rel, small_case

F32, S32, C64 = 11, 4, 15  # xla::ffi::DataType values of the test double


def _lib():
    from waveforminversionust_b200 import jax_frontend
    L = C.CDLL(jax_frontend.build_ffi(stub=True))
    L.ust_ffi_stub_frame_new.restype = C.c_void_p
    L.ust_ffi_stub_frame_new.argtypes = [C.c_void_p]
    L.ust_ffi_stub_frame_free.argtypes = [C.c_void_p]
    for fn in (L.ust_ffi_stub_add_arg, L.ust_ffi_stub_add_ret):
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    L.ust_ffi_stub_set_attr.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    L.ust_ffi_stub_error.restype = C.c_char_p
    L.ust_ffi_stub_error.argtypes = [C.c_void_p]
    L.ust_solve_helmholtz_ffi.argtypes = [C.c_void_p]
    L.ust_fwi_loss_grad_ffi.argtypes = [C.c_void_p]
    return L


def test_shim_compiles_and_exports_both_handlers():
    L = _lib()
    assert hasattr(L, "ust_solve_helmholtz_ffi") and hasattr(L, "ust_fwi_loss_grad_ffi")


def test_real_headers_are_required_for_the_jax_build():
    from waveforminversionust_b200 import jax_frontend
    if jax_frontend.xla_include_dir() is None:
        with pytest.raises(jax_frontend.FfiUnavailable):
            jax_frontend.build_ffi()


@pytest.mark.gpu
def test_handlers_forward_to_the_c_abi():
    import torch
    import waveforminversionust_b200 as w
    L = _lib()
    n, nelem = 52, 16
    geom, f, vel_true = small_case(n, nelem)
    dv = "cuda:0"
    dims = lambda *d: (C.c_int64 * len(d))(*d)

    def add(frame, fn, t, dt):
        fn(frame, C.c_void_p(t.data_ptr()), dt, t.dim(), dims(*t.shape))

    x = torch.as_tensor(geom.xi.astype(np.float32)).to(dv)
    vel = torch.as_tensor(vel_true.astype(np.float32)).to(dv)
    src = torch.as_tensor(geom.dense_src(np.complex64)).to(dv).reshape(n * n, nelem).contiguous()
    ft = torch.tensor([f], dtype=torch.float32, device=dv)
    for adj in (0, 1):
        out = torch.zeros_like(src)
        fr = L.ust_ffi_stub_frame_new(C.c_void_p(torch.cuda.current_stream().cuda_stream))
        for t, dt in ((x, F32), (x, F32), (vel, F32), (src, C64), (ft, F32), (torch.tensor([adj], dtype=torch.int32, device=dv), S32)):
            add(fr, L.ust_ffi_stub_add_arg, t, dt)
        add(fr, L.ust_ffi_stub_add_ret, out, C64)
        L.ust_ffi_stub_set_attr(fr, b"a0", geom.a0)
        L.ust_ffi_stub_set_attr(fr, b"L_PML", geom.L_PML)
        rc = L.ust_solve_helmholtz_ffi(fr)
        assert rc == 0, L.ust_ffi_stub_error(fr)
        L.ust_ffi_stub_frame_free(fr)
        torch.cuda.synchronize()
        want = w.solve_helmholtz(geom.xi, geom.yi, vel, src.reshape(n, n, nelem), float(np.float32(f)), geom.a0, geom.L_PML, bool(adj))
        assert rel(out.cpu().numpy().reshape(n, n, nelem), want.cpu().numpy()) < 1e-6
    # (loss, grad): loss arrives as two float32 words
    freqs = np.array([0.9 * f, f], dtype=np.float32)
    rec = np.ascontiguousarray(np.stack([observed_data(geom, float(fr_), vel_true) for fr_ in freqs]).astype(np.complex64))
    slow = torch.full((n, n), 1 / 1480.0, dtype=torch.float32, device=dv)
    rx_lin = (geom.y_idx * n + geom.x_idx).astype(np.int32)
    loss2 = torch.zeros(2, dtype=torch.float32, device=dv)
    grad = torch.zeros((n, n), dtype=torch.float32, device=dv)
    fr = L.ust_ffi_stub_frame_new(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    tens = [(slow, F32), (torch.as_tensor(rec).to(dv), C64), (torch.as_tensor(geom.src_lin).to(dv), S32), (torch.as_tensor(rx_lin).to(dv), S32),
            (torch.as_tensor(geom.mask_indices.astype(np.int32)).to(dv), S32), (x, F32), (x, F32), (torch.as_tensor(freqs).to(dv), F32)]
    for t, dt in tens:
        add(fr, L.ust_ffi_stub_add_arg, t, dt)
    add(fr, L.ust_ffi_stub_add_ret, loss2, F32)
    add(fr, L.ust_ffi_stub_add_ret, grad, F32)
    L.ust_ffi_stub_set_attr(fr, b"a0", geom.a0)
    L.ust_ffi_stub_set_attr(fr, b"L_PML", geom.L_PML)
    rc = L.ust_fwi_loss_grad_ffi(fr)
    assert rc == 0, L.ust_ffi_stub_error(fr)
    torch.cuda.synchronize()
    l_want, g_want = w.fwi_loss_function(slow, geom.xi, geom.yi, torch.as_tensor(rec).to(dv), geom.dense_src(), freqs.astype(np.float64),
                                         geom.a0, geom.L_PML, geom.tx_include, geom.ind_matlab, geom.mask_indices, geom.num_elements)
    l_got = float(loss2[0].double() + loss2[1].double())
    assert abs(l_got - float(l_want)) <= 1e-12 * abs(float(l_want)) and rel(grad.cpu().numpy(), g_want.cpu().numpy()) < 1e-6
    # a missing attribute is an error from the binding, not a crash
    L.ust_ffi_stub_frame_free(fr)
    fr = L.ust_ffi_stub_frame_new(None)
    assert L.ust_fwi_loss_grad_ffi(fr) != 0 and L.ust_ffi_stub_error(fr)
    L.ust_ffi_stub_frame_free(fr)
    w.clear_plans()
