"""
CPU oracle: HEALPix RING-ordered pixel centres (closed form; healpy is absent and is used
by the reference for plotting only, tools/visualization.py:13,274).  Test infrastructure
only.  Formulae: Gorski et al. 2005, restated in SURVEY.md A.4a.  The reference contract is
"1-D map of length 12*Nside^2 in RING order" (observables/dataset.py:225-232,
simulators/hammurabi.py:139-153).
"""
import numpy as np


def nside2npix(nside):
    return 12 * int(nside) ** 2


def pix2zphi_ring(nside, pix=None):
    """Returns (z = cos(colatitude), sin(colatitude), phi) for RING pixel indices."""
    nside = int(nside)
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    p = np.arange(npix, dtype=np.int64) if pix is None else np.asarray(pix, dtype=np.int64)
    z = np.empty(p.shape)
    sth = np.empty(p.shape)
    phi = np.empty(p.shape)

    north = p < ncap
    south = p >= npix - ncap
    equat = ~(north | south)

    # north polar cap
    q = p[north]
    i = (1 + np.floor(np.sqrt(1.0 + 2.0 * q)).astype(np.int64)) >> 1
    # integer sqrt correction (guards the float sqrt at large arguments)
    i = np.where(2 * i * (i - 1) > q, i - 1, i)
    i = np.where(2 * (i + 1) * i <= q, i + 1, i)
    j = q + 1 - 2 * i * (i - 1)
    tmp = (i * i) * (4.0 / npix)
    z[north] = 1.0 - tmp
    sth[north] = np.sqrt(tmp * (2.0 - tmp))
    phi[north] = (j - 0.5) * (np.pi / (2.0 * i))

    # equatorial belt
    q = p[equat] - ncap
    i = q // (4 * nside) + nside
    j = q % (4 * nside) + 1
    fodd = np.where(((i + nside) & 1) == 1, 1.0, 0.5)
    ze = (2 * nside - i) * (2.0 / (3.0 * nside))
    z[equat] = ze
    sth[equat] = np.sqrt((1.0 - ze) * (1.0 + ze))
    phi[equat] = (j - fodd) * (np.pi / (2.0 * nside))

    # south polar cap
    q = npix - p[south]
    i = (1 + np.floor(np.sqrt(2.0 * q - 1.0)).astype(np.int64)) >> 1
    i = np.where(2 * i * (i - 1) >= q, i - 1, i)
    i = np.where(2 * (i + 1) * i < q, i + 1, i)
    j = 4 * i + 1 - (q - 2 * i * (i - 1))
    tmp = (i * i) * (4.0 / npix)
    z[south] = -(1.0 - tmp)
    sth[south] = np.sqrt(tmp * (2.0 - tmp))
    phi[south] = (j - 0.5) * (np.pi / (2.0 * i))
    return z, sth, phi


def pix2ang_ring(nside, pix=None):
    z, sth, phi = pix2zphi_ring(nside, pix)
    return np.arctan2(sth, z), phi


def pix2vec_ring(nside, pix=None):
    """Unit vectors [npix,3]; Galactic l = phi, b = pi/2 - theta; x toward l=0,b=0."""
    z, sth, phi = pix2zphi_ring(nside, pix)
    return np.stack([sth * np.cos(phi), sth * np.sin(phi), z], axis=-1)
