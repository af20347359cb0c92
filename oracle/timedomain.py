"""Oracle restatement of the reference's time-domain synthesis and frequency schedule
(TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py; parity unpinned: the reference ships these as MATLAB scripts
without stored outputs).

``Lecture19_Fwi/TimeDomainSimulation.m``: frequency axis ``flow:df:fhigh`` with a Hann response (``:29-33``), one forward
solve per frequency for one transmitting element (``:36-46``), then an inverse discrete-time Fourier transform on an
arbitrary time axis -- "Not an IFFT though!" (``:48-56``) -- and the channel data read at the ring elements (``:72-75``).
"""
from __future__ import annotations

import numpy as np

from .helmholtz import solve_helmholtz


def hanning(n):
    """MATLAB ``hanning(n)``: the n-point symmetric Hann window WITHOUT the zero end points,
    ``0.5*(1 - cos(2*pi*(1:n)'/(n+1)))`` (``TimeDomainSimulation.m:33``)."""
    k = np.arange(1, n + 1, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * k / (n + 1)))


def idtft(WVFIELD_F, f, resp_freq, time, df):
    """``IDTFT = exp(1i*2*pi*f.*time')*df; WVFIELD_T = pagemtimes(IDTFT, resp_freq .* permute(WVFIELD_F,[3,1,2]))``
    (``TimeDomainSimulation.m:54-56``): WVFIELD_F (Ny, Nx, Nf) -> WVFIELD_T (Ny, Nx, Nt), complex128 arithmetic."""
    f = np.asarray(f, dtype=np.float64)
    time = np.asarray(time, dtype=np.float64)
    W = np.exp(2j * np.pi * np.outer(time, f)) * df * np.asarray(resp_freq, dtype=np.float64)[None, :]  # (Nt, Nf)
    return np.einsum("tf,yxf->yxt", W, np.asarray(WVFIELD_F, dtype=np.complex128))


def time_domain_simulation(xi, yi, C, src, f, resp_freq, time, a0, L_PML, dtype="c128", stencil="python"):
    """``TimeDomainSimulation.m:36-56``: ``src`` is the (Ny, Nx) source of the transmitting element; returns
    (WVFIELD_F (Ny, Nx, Nf), WVFIELD_T (Ny, Nx, Nt))."""
    f = np.asarray(f, dtype=np.float64)
    WVFIELD_F = np.stack([solve_helmholtz(xi, yi, C, src[:, :, None], fk, a0, L_PML, False, dtype=dtype, stencil=stencil)[:, :, 0]
                          for fk in f], axis=2)
    df = float(f[1] - f[0]) if f.size > 1 else 1.0
    return WVFIELD_F, idtft(WVFIELD_F, f, resp_freq, time, df)


def channel_data(WVFIELD_T, x_idx, y_idx):
    """``channelData(t, e) = WVFIELD_T(y_idx(e), x_idx(e), t)`` (``TimeDomainSimulation.m:72-75``), 0-based indices here."""
    return np.asarray(WVFIELD_T)[np.asarray(y_idx), np.asarray(x_idx), :].T


def continuation_stages(f, nstages):
    """Low-to-high frequency schedule: the band is cut into ``nstages`` contiguous groups of (nearly) equal size in
    ascending frequency; stage s inverts group s starting from the result of stage s-1.  The reference only states the
    rule ("Use 0.1-0.6 MHz ... cycle skipping", ``SimulateData.m:29-30``, slides p.24) and has no multi-frequency code."""
    order = np.argsort(np.asarray(f, dtype=np.float64), kind="stable")
    return [np.sort(g) for g in np.array_split(order, int(nstages)) if g.size]
