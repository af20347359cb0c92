#!/usr/bin/env python
"""Timing bounds of the factorisation chain: what a free pivot inversion / a free row panel would buy.

Runs the benchmark step (512^2, 256 sources) under the UST_EXP switches of factor.cuh -- the RESULTS of those runs are wrong
by construction (stale pivot inverses / row panels), only the launch structure and timing are kept -- for the classic and the
two-level Gauss-Jordan scheme at 16 and 2 frequencies.

    python tools/exp_bounds.py [variant ...]            (each variant runs in its own process: switches are read at plan creation)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "fused": {"UST_FUSE_RP": "1"},
    "deep": {"UST_DEEP": "1"},
    "riding": {},
    "riding_pdl0": {"UST_PDL_MIN_BATCH": "0"},
    "riding_pdl8": {"UST_PDL_MIN_BATCH": "8"},
    "riding_pdl16": {"UST_PDL_MIN_BATCH": "16"},
    "deep_g1": {"UST_DEEP": "1", "UST_GROUPS": "1"},
    "riding_g1": {"UST_DEEP": "0", "UST_GROUPS": "1"},
    "deep_g4": {"UST_DEEP": "1", "UST_GROUPS": "4"},
    "riding_g4": {"UST_DEEP": "0", "UST_GROUPS": "4"},
    "classic": {},
    "classic_nopivinv": {"UST_EXP": "1"},
    "classic_nopiv": {"UST_EXP": "2"},
    "classic_norowpanel": {"UST_EXP": "4"},
    "classic_nopiv_norowpanel": {"UST_EXP": "6"},
    "gj2": {"UST_GJ2": "1"},
    "gj2_nopivinv": {"UST_GJ2": "1", "UST_EXP": "1"},
    "gj2_nopiv": {"UST_GJ2": "1", "UST_EXP": "2"},
    "gj2_nopiv_norowpanel": {"UST_GJ2": "1", "UST_EXP": "6"},
}


def child(nfreq, groups):
    import bench
    import torch
    sys.argv = ["bench.py", "--nfreq", str(nfreq), "--no-cpu-baseline"] + (["--groups", str(groups)] if groups else [])
    a = bench.parse()
    geom, freqs, vel_true, vel0 = bench.workload(a)
    H = bench.Harness(a, torch, None, geom, freqs, vel_true, vel0, 0, 1, 0)
    for _ in range(2):
        H.step_dev()
    ms, _ = H.timed(H.step_dev, 3)
    plan = H.eng.plan
    plan.profile(True)
    H.step_dev()
    prof = plan.get_profile()
    plan.profile(False)
    out = {"ms_per_step": ms / 3, "classes": {k: [round(v[0], 2), v[1]] for k, v in prof.items() if v[1]}}
    print("RESULT " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    names = sys.argv[1:] or list(VARIANTS)
    for nfreq in (16, 2):
        for name in names:
            env = dict(os.environ, **VARIANTS[name])
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(nfreq), "0"], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if not line:
                print(nfreq, name, "FAILED", r.stderr[-800:])
                continue
            d = json.loads(line[-1][7:])
            c = d["classes"]
            fac = sum(c.get(k, [0, 0])[0] for k in ("schur", "gj_k0", "gj_panel", "gj_update", "gj_pivot") ) + (c.get("gj_rowpanel", [0, 0])[0] if "UST_GJ2" in VARIANTS[name] else 0)
            print(f"nfreq {nfreq:2d} {name:26s} step {d['ms_per_step']:7.1f} ms | one-chain profile: factor {fac:6.1f}  " +
                  "  ".join(f"{k} {v[0]:.1f}/{v[1]}" for k, v in c.items() if k in ("schur", "gj_k0", "gj_panel", "gj_update", "gj_pivot", "gj_rowpanel", "sweep_gemm", "tri_apply")), flush=True)
