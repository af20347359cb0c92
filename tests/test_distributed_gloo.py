"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frequency sharding and the single packed
all-reduce of (loss, grad).  The per-rank evaluator here is the oracle; on GPUs it is libustfwi.so."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waveforminversionust_b200 import distributed as D


def test_shard_frequencies_partitions_exactly():
    for nf in (1, 2, 5, 16, 17):
        for world in (1, 2, 3, 4, 8):
            parts = [D.shard_frequencies(nf, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(nf))
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1
    assert D.shard_frequencies(16, 3, 8) == [6, 7]


def test_pack_unpack_keeps_loss_precision():
    g = torch.randn(7, 9, dtype=torch.float32)
    loss = 5.334395123456789e-14
    l2, g2 = D.unpack_loss_grad(D.pack_loss_grad(loss, g), g.shape)
    assert torch.equal(g2, g) and abs(float(l2) - loss) / loss < 1e-13


def _worker(rank, world, port, tmp):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import observed_data, small_case
    from oracle import fwi as ofwi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, nelem = 36, 16
    geom, f0, vel_true = small_case(n, nelem, pml_cells=4.0)
    freqs = np.array([0.8, 0.9, 1.0]) * f0
    recs = [observed_data(geom, f, vel_true, seed=5 + i) for i, f in enumerate(freqs)]
    slow = np.full((n, n), 1 / 1480.0)
    tail = (geom.a0, geom.L_PML, geom.tx_include, geom.ind_matlab, geom.mask_indices, geom.num_elements)
    loss, grad = 0.0, np.zeros((n, n))
    for i in D.shard_frequencies(freqs.size, rank, world):
        l, g = ofwi.fwi_loss_and_grad(slow, geom.xi, geom.yi, recs[i], geom.dense_src(), freqs[i], *tail, dtype="c128")
        loss += l
        grad += g
    L, G = D.allreduce_loss_grad(loss, torch.as_tensor(grad))
    if rank == 0:
        np.savez(tmp, loss=float(L), grad=G.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_frequency_sharding_matches_single_process(tmp_path):
    from common import observed_data, rel, small_case
    from oracle import fwi as ofwi
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    n, nelem = 36, 16
    geom, f0, vel_true = small_case(n, nelem, pml_cells=4.0)
    freqs = np.array([0.8, 0.9, 1.0]) * f0
    slow = np.full((n, n), 1 / 1480.0)
    tail = (geom.a0, geom.L_PML, geom.tx_include, geom.ind_matlab, geom.mask_indices, geom.num_elements)
    loss, grad = 0.0, 0.0
    for i, f in enumerate(freqs):
        l, g = ofwi.fwi_loss_and_grad(slow, geom.xi, geom.yi, observed_data(geom, f, vel_true, seed=5 + i),
                                      geom.dense_src(), f, *tail, dtype="c128")
        loss += l
        grad = grad + g
    assert float(got["loss"]) == pytest.approx(loss, rel=1e-12)
    assert rel(got["grad"], grad) < 1e-12


def _worker_src(rank, world, port, tmp):
    """Source-block sharding: every rank evaluates the same frequency on its block of transmitters."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from common import observed_data, small_case
    from oracle import fwi as ofwi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, nelem = 36, 16
    geom, f, vel_true = small_case(n, nelem, pml_cells=4.0)
    rec = observed_data(geom, f, vel_true, seed=5)
    slow = np.full((n, n), 1 / 1480.0)
    tx = D.shard_frequencies(geom.tx_include.size, rank, world)  # this rank's transmitters
    l, g = ofwi.fwi_loss_and_grad(slow, geom.xi, geom.yi, rec[tx], geom.dense_src()[:, :, tx], f, geom.a0, geom.L_PML,
                                  geom.tx_include[tx], geom.ind_matlab, geom.mask_indices[tx], geom.num_elements, dtype="c128")
    L, G = D.allreduce_loss_grad(l, torch.as_tensor(g))
    if rank == 0:
        np.savez(tmp, loss=float(L), grad=G.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_source_block_sharding_matches_single_process(tmp_path):
    from common import observed_data, rel, small_case
    from oracle import fwi as ofwi
    out = str(tmp_path / "r0.npz")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_src, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    n, nelem = 36, 16
    geom, f, vel_true = small_case(n, nelem, pml_cells=4.0)
    rec = observed_data(geom, f, vel_true, seed=5)
    loss, grad = ofwi.fwi_loss_and_grad(np.full((n, n), 1 / 1480.0), geom.xi, geom.yi, rec, geom.dense_src(), f, geom.a0, geom.L_PML,
                                        geom.tx_include, geom.ind_matlab, geom.mask_indices, geom.num_elements, dtype="c128")
    assert float(got["loss"]) == pytest.approx(loss, rel=1e-12)
    assert rel(got["grad"], grad) < 1e-12
