#!/usr/bin/env python
"""Phase timestamps (ns since kernel entry, %globaltimer) of CTA (0,0,0) of the TMA-fed tcgen05 GEMM, printed by
ust_test_cgemm when UST_TC2_TRACE is set.  Slots: 1 setup done (barriers + TMEM alloc), 2 first TMA issued, 3 first
stage landed, 4 chunk 0 MMAs issued, 5 drain warps see D1 of chunk 0, 6 D1 handed back, 7 MMA warp sees it, 8 last
chunk's D1 hand-back seen, 9 last chunk issued, 10 D2 complete, 11 tile staged in smem, 12 written out, 13 planes
emitted, 14 TMEM freed."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waveforminversionust_b200 import _lib  # noqa: E402

os.environ["UST_TC2_TRACE"] = "1"
L = _lib.lib()
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N, K, with_cin) in [(128, 128, 64, True), (128, 128, 64, True), (128, 128, 512, False), (512, 256, 512, True), (2048, 2048, 64, True)]:
    mk = lambda *s: torch.complex(torch.randn(*s, generator=g, device="cuda"), torch.randn(*s, generator=g, device="cuda"))
    A, B, Cin = mk(M, K), mk(K, N), mk(M, N)
    out = torch.empty((M, N), dtype=torch.complex64, device="cuda")
    torch.cuda.synchronize()
    rc = L.ust_test_cgemm(3, 0, M, N, K, C.c_void_p(A.data_ptr()), K, C.c_void_p(B.data_ptr()), N,
                          C.c_void_p(Cin.data_ptr()) if with_cin else None, N, C.c_void_p(out.data_ptr()), N,
                          C.c_float(-1.0), 0, 0, 0, 0, None)
    _lib.check(rc, "ust_test_cgemm")
