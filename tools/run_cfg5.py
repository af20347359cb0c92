#!/usr/bin/env python
"""BASELINE configs[4]: synthetic 2048^2 grid, 512-element ring, multi-frequency L-BFGS end to end on the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_cfg5.py [--n 2048] [--nsrc 512] [--maxiter 2]

One frequency per GPU (a 2048^2 factorisation is 103 GB of operand planes: one at a time fits a 180 GB GPU next to the
34 GB of wavefields); the observed data come from the true model with this solver; every rank runs
distributed.run_lbfgs_sharded.  Rank 0 prints one JSON line (losses of every evaluation, seconds per evaluation, sound-speed
error against the true model before / after)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--nsrc", type=int, default=512)
    ap.add_argument("--maxiter", type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from waveforminversionust_b200 import geometry as G
    from waveforminversionust_b200.distributed import ShardedFWI, run_lbfgs_sharded
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    geom = G.ring_geometry(a.n, a.nsrc)
    f_hi = G.frequency_for_grid(a.n)
    freqs = np.linspace(f_hi * 300.0 / 596.0, f_hi, world) if world > 1 else np.array([f_hi * 300.0 / 596.0])
    vel_true = G.blob_model(geom)
    eng = ShardedFWI(geom, freqs, dtype="c64", device=local, rank=rank, world=world)
    plan, dv = eng.plan, torch.device(f"cuda:{local}")
    nt, ne, nl = geom.tx_include.size, geom.num_elements, len(eng.local)
    rec = torch.zeros((max(nl, 1), nt, ne), dtype=plan.tcplx, device=dv)
    t0 = time.perf_counter()
    plan.fwi_loss_grad(torch.as_tensor((1.0 / vel_true).astype(plan.real)).to(dv), rec, eng.local_freqs)
    amp = torch.as_tensor(G.source_amplitudes(nt)).to(dv, plan.tcplx)
    rx = torch.as_tensor((geom.y_idx * geom.Nx + geom.x_idx).astype(np.int64)).to(dv)
    for i in range(nl):
        U = plan.wavefield(i).reshape(geom.Ny * geom.Nx, nt)
        rec[i] = U[rx, :].T * amp[:, None]
        del U
    torch.cuda.synchronize()
    t_data = time.perf_counter() - t0
    hist = []
    t0 = time.perf_counter()
    vel = run_lbfgs_sharded(eng, rec, 1500.0, maxiter=a.maxiter, history=hist)
    torch.cuda.synchronize()
    t_opt = time.perf_counter() - t0
    inside = np.hypot(*np.meshgrid(geom.xi.astype(np.float64), geom.yi.astype(np.float64))) < 0.09
    e0 = float(np.sqrt(np.mean((1500.0 - vel_true[inside]) ** 2)))
    e1 = float(np.sqrt(np.mean((vel[inside] - vel_true[inside]) ** 2)))
    if rank == 0:
        print(json.dumps({"config": f"{a.n}x{a.n} grid, {a.nsrc}-element ring, {freqs.size} frequencies on {world} GPUs, L-BFGS maxiter={a.maxiter}",
                          "freq_khz": [round(float(f) / 1e3, 1) for f in freqs], "evaluations": len(hist),
                          "loss_ratio": [round(l / hist[0][0], 6) for l, _ in hist], "sec_per_evaluation": t_opt / max(len(hist), 1),
                          "sec_synthetic_data": t_data, "sound_speed_rms_error_m_s": [e0, e1], "device_bytes": plan.device_bytes,
                          "status": plan.status()}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
