#!/usr/bin/env python
"""CPU baseline of BASELINE.md section 3: the oracle (NumPy assembly + SciPy spsolve -> SuperLU, re-factorising on every call
exactly as the reference does) timed IN FULL on configs[0] (the shipped 301^2 dataset: one NCG iteration = forward + adjoint +
perturbation solve) and configs[1] (synthetic 256^2, 256 sources: forward + adjoint), complex64 and complex128, on this
machine's host cores (SuperLU is single-threaded).  Prints one JSON object; run it on the GPU box next to bench.py:

    python tools/cpu_baseline.py > gpurun_out/cpu_baseline.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import cfg1_inputs, observed_data  # noqa: E402
from oracle import fwi as ofwi  # noqa: E402
from oracle import helmholtz as oh  # noqa: E402
from waveforminversionust_b200 import geometry as G  # noqa: E402


def main():
    out = {"host_cores": os.cpu_count(), "threads_used": 1, "kind": "port",
           "what": "oracle port of the reference's CPU path (SciPy spsolve -> SuperLU gssv), timed in full"}
    g = np.load(os.path.join(ROOT, "tests", "golden", "cfg1_shipped.npz"))
    geom, rec = cfg1_inputs(g["rec"], g["x_circ"], g["y_circ"])
    for dtype in ("c64", "c128"):
        t0 = time.perf_counter()
        ofwi.nonlinear_conjugate_gradient_vectorized(geom.xi, geom.yi, geom.num_elements, rec, geom.dense_src(), geom.tx_include,
                                                     geom.ind_matlab, 1480.0, float(g["f"]), 1, geom.a0, geom.L_PML, geom.mask_indices,
                                                     dtype=dtype, reuse_factor=False)
        t = time.perf_counter() - t0
        out[f"cfg1_{dtype}"] = {"sec_per_ncg_iteration": t, "source_solves_per_sec": 3 * 256 / t,
                                "what": "301^2, 256 sources x 193 receivers, 350 kHz: forward + adjoint + perturbation solve, three factorisations"}
        print(f"cfg1 {dtype}: {t:.1f} s per NCG iteration", file=sys.stderr, flush=True)
    geom = G.ring_geometry(256, 256)
    f = G.frequency_for_grid(256)
    vel0 = G.blob_model(geom, dc=15.0, seed=99)
    src = geom.dense_src()
    for dtype in ("c64", "c128"):
        t0 = time.perf_counter()
        u = oh.solve_helmholtz(geom.xi, geom.yi, vel0, src, f, geom.a0, geom.L_PML, False, dtype=dtype)
        oh.solve_helmholtz(geom.xi, geom.yi, vel0, u, f, geom.a0, geom.L_PML, True, dtype=dtype)
        t = time.perf_counter() - t0
        out[f"cfg2_{dtype}"] = {"sec_forward_plus_adjoint": t, "source_solves_per_sec": 2 * 256 / t,
                                "what": "256^2, 256 sources, one frequency: forward + adjoint solve_helmholtz calls (two factorisations)"}
        print(f"cfg2 {dtype}: {t:.1f} s forward + adjoint", file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
