"""time-domain synthesis (TimeDomainSimulation.m) and the frequency-continuation schedule.
CPU tests pin the oracle restatement against the closed form; GPU tests compare the CUDA path (ust_idtft, batched
per-frequency solves, staged NCG) with the oracle on identical seeded inputs."""
import numpy as np
import pytest

from common import observed_data, rel, small_case
from oracle import fwi as ofwi
from oracle import timedomain as otd


def test_hanning_is_matlab_hanning():
    # hanning(4) = [0.3455 0.9045 0.9045 0.3455] (MATLAB documentation example), symmetric, no zero end points
    assert np.allclose(otd.hanning(4), [0.3454915, 0.9045085, 0.9045085, 0.3454915], atol=1e-7)
    assert np.allclose(otd.hanning(81), otd.hanning(81)[::-1]) and otd.hanning(81).min() > 0


def test_oracle_idtft_closed_form_and_linearity():
    rng = np.random.default_rng(0)
    f = 1e5 + 5e3 * np.arange(6)
    resp = otd.hanning(6)
    time = np.linspace(0, 1.6e-4, 17)
    W = rng.standard_normal((3, 4, 6)) + 1j * rng.standard_normal((3, 4, 6))
    out = otd.idtft(W, f, resp, time, 5e3)
    ref = np.zeros((3, 4, 17), dtype=np.complex128)
    for t in range(17):
        for k in range(6):
            ref[:, :, t] += np.exp(2j * np.pi * f[k] * time[t]) * 5e3 * resp[k] * W[:, :, k]
    assert rel(out, ref) < 1e-14
    # a single spectral line comes back as that complex exponential
    one = np.zeros((1, 1, 6), dtype=np.complex128)
    one[0, 0, 2] = 2.0 - 1.0j
    assert np.allclose(otd.idtft(one, f, resp, time, 5e3)[0, 0], (2.0 - 1.0j) * 5e3 * resp[2] * np.exp(2j * np.pi * f[2] * time))
    W2 = rng.standard_normal((3, 4, 6)) + 1j * rng.standard_normal((3, 4, 6))
    assert rel(otd.idtft(W + 2 * W2, f, resp, time, 5e3), out + 2 * otd.idtft(W2, f, resp, time, 5e3)) < 1e-14


def test_continuation_stages_cover_the_band_low_to_high():
    import waveforminversionust_b200 as w
    f = np.array([4e5, 1e5, 3e5, 2e5, 6e5, 5e5, 7e5])
    st = w.continuation_stages(f, 3)
    assert [list(s) for s in st] == [list(s) for s in otd.continuation_stages(f, 3)]
    assert sorted(np.concatenate(st).tolist()) == list(range(7))
    tops = [f[s].max() for s in st]
    lows = [f[s].min() for s in st]
    assert all(tops[i] < lows[i + 1] for i in range(len(st) - 1))
    assert np.allclose(w.hanning(9), otd.hanning(9))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["c64", "c128"])
def test_idtft_matches_oracle(dtype):
    import torch
    import waveforminversionust_b200 as w
    rng = np.random.default_rng(1)
    ny, nx, nf, nt = 37, 41, 13, 50  # nt not a multiple of the kernel's time tile, ny*nx not a multiple of 256
    f = 1e5 + 5e3 * np.arange(nf)
    resp = otd.hanning(nf)
    time = np.linspace(0, 1.6e-4, nt)
    W = (rng.standard_normal((ny, nx, nf)) + 1j * rng.standard_normal((ny, nx, nf))).astype(np.complex64 if dtype == "c64" else np.complex128)
    ref = otd.idtft(W, f, resp, time, 5e3)
    got = w.idtft(W, f, resp, time, 5e3)
    assert got.shape == (ny, nx, nt) and rel(got, ref) < (2e-6 if dtype == "c64" else 1e-13)
    got_dev = w.idtft(torch.as_tensor(W).cuda(), f, resp, time)  # df defaults to f[1] - f[0]
    assert got_dev.is_cuda and rel(got_dev.cpu().numpy(), ref) < (2e-6 if dtype == "c64" else 1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["c64", "c128"])
def test_time_domain_simulation_matches_oracle(dtype):
    import waveforminversionust_b200 as w
    n = 56
    geom, f0, vel = small_case(n, 16)
    f = np.linspace(0.5 * f0, f0, 5)  # TimeDomainSimulation.m:29-32 (flow:df:fhigh), scaled to this grid
    resp = otd.hanning(f.size)
    time = np.linspace(0.0, 2 * 0.12 / 1500.0, 41)  # :49-51
    src = geom.dense_src(np.complex128)[:, :, 3]
    WFo, WTo = otd.time_domain_simulation(geom.xi, geom.yi, vel, src, f, resp, time, geom.a0, geom.L_PML, dtype="c128")
    WF, WT = w.time_domain_simulation(geom.xi, geom.yi, vel, src, f, resp, time, geom.a0, geom.L_PML, dtype=dtype, batch=2)
    ef, et = rel(WF, WFo), rel(WT, WTo)
    print(f"{dtype}: frequency stack {ef:.2e}, time-domain wavefield {et:.2e}")
    assert ef < (1e-5 if dtype == "c64" else 1e-9) and et < (1e-5 if dtype == "c64" else 1e-9)
    cd, cdo = w.channel_data(WT, geom.x_idx, geom.y_idx), otd.channel_data(WTo, geom.x_idx, geom.y_idx)
    assert cd.shape == (time.size, geom.num_elements) and rel(cd, cdo) < (2e-5 if dtype == "c64" else 1e-9)
    w.clear_plans()


@pytest.mark.gpu
def test_frequency_continuation_matches_staged_oracle_ncg():
    import waveforminversionust_b200 as w
    n, nelem, niter = 48, 16, 2
    geom, f0, vel_true = small_case(n, nelem)
    f = np.array([0.8 * f0, 0.6 * f0, f0])  # deliberately unsorted
    rec = np.stack([observed_data(geom, fk, vel_true) for fk in f])
    stages = w.continuation_stages(f, 3)
    assert [int(s[0]) for s in stages] == [1, 0, 2]
    hist = []
    VEL = w.frequency_continuation(geom.xi, geom.yi, geom.num_elements, rec, geom.dense_src(), geom.tx_include, geom.ind_matlab,
                                   1480.0, f, stages, niter, geom.a0, geom.L_PML, geom.mask_indices, dtype="c128", history=hist)
    velo = 1480.0
    for s in stages:
        k = int(s[0])
        velo, _, _, _, _ = ofwi.nonlinear_conjugate_gradient_vectorized(geom.xi, geom.yi, geom.num_elements, rec[k], geom.dense_src(),
                                                                        geom.tx_include, geom.ind_matlab, velo, f[k], niter, geom.a0,
                                                                        geom.L_PML, geom.mask_indices, dtype="c128")
    rms = float(np.sqrt(np.mean((VEL - velo) ** 2)))
    print(f"staged NCG: VEL RMS diff {rms:.3e} m/s; stage losses {[[round(h['loss'] / hs[0]['loss'], 3) for h in hs] for hs in hist]}")
    assert rms < 1e-3 and len(hist) == 3 and all(len(h) == niter for h in hist)
    assert all(hs[-1]["loss"] < hs[0]["loss"] for hs in hist)
    w.clear_plans()
