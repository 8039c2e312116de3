#!/usr/bin/env python
"""
The one hammurabiX OUTPUT the reference repository records: tutorials/tutorial_hammurabi.ipynb (cell 14)
prints the synchrotron Stokes I map at 23 GHz, Nside 32, of hammurabiX's built-in WMAP LSA regular field
(b0 = 6 uG, psi0 = 27.9, psi1 = 1.3, chi0 = 24.6 deg) + analytic CRE model (alpha = 3, beta = theta = 0,
r0 = 5.6 kpc, z0 = 1.2 kpc, E0 = 20.5 GeV, j0 = 0.03) + YMW16 thermal electrons -- first and last three
pixels (RING): 0.08884023 0.08772291 0.08634578 ... 0.08866684 0.08728929 0.08849186  [mK: the dataset the
simulator was initialised with is in mK].  Total intensity does not depend on the thermal electrons, so these
six numbers pin the conventions of OUR line-of-sight integrator that only a hammurabiX run can confirm for
Stokes I: the CRE normalisation from (E0, j0), the K(p) prefactor, brightness temperature (K_RJ) units, the
LSA field definition, the integration limits (tests/data/test_ham.xml:122-131: oc_r_max 30 kpc, gc_r_max
20 kpc, |z| <= 10 kpc) -- all at once.

This script evaluates oracle/los.py's emissivity definition with oracle/fields.py's LSA and CRE formulas
sampled DIRECTLY at the ray samples (no grid, so no interpolation error enters) and prints the ratio to the
notebook's numbers.  CPU only; `tests/test_oracle_los.py::test_tutorial_hammurabi_pin` asserts the outcome.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import healpix as H   # noqa: E402
from oracle import los as L       # noqa: E402

NOTEBOOK_I_MK = {0: 0.08884023, 1: 0.08772291, 2: 0.08634578, -3: 0.08866684, -2: 0.08728929, -1: 0.08849186}
LSA = dict(b0=6.0, psi0=np.deg2rad(27.9), psi1=np.deg2rad(1.3), chi0=np.deg2rad(24.6))
CRE = dict(alpha=3.0, r0=5.6, z0=1.2, e0=20.5, j0=0.03)
OBSERVER = np.array([-8.3, 0.0, 0.006])


def lsa_at(x, y, z, b0, psi0, psi1, chi0, r0=8.0, z0=1.0, r_min=3.0, r_max=20.0):
    """oracle/fields.py:lsa_B at points."""
    R = np.sqrt(x * x + y * y)
    ok = (R > r_min) & (R < r_max)
    Rs = np.where(ok, R, 1.0)
    cp, sp = x / Rs, y / Rs
    psi = psi0 + psi1 * np.log(Rs / r0)
    chi = chi0 * np.tanh(z / z0)
    cc = np.cos(chi)
    br, bphi = b0 * np.sin(psi) * cc, b0 * np.cos(psi) * cc
    B = np.stack([br * cp - bphi * sp, br * sp + bphi * cp, b0 * np.sin(chi)], axis=-1)
    return np.where(ok[..., None], B, 0.0)


def cre_at(x, y, z, j0, r0, z0, r_sun):
    """oracle/fields.py:cre_analytic at points."""
    R = np.sqrt(x * x + y * y)
    ch = np.cosh(z / z0)
    return j0 * np.exp(-(R - r_sun) / r0) / (ch * ch)


def stokes_i_mk(pixels, nside=32, nu_ghz=23.0, dr=0.01, r_max=30.0, gc_r_max=20.0, gc_z_max=10.0,
                r_sun=8.3, lsa=LSA, cre=CRE):
    dirs = H.pix2vec_ring(nside)[np.asarray(pixels)]
    north, east = L.sky_basis(dirs)
    r = (np.arange(int(np.ceil(r_max / dr - 1e-9))) + 0.5) * dr
    pos = OBSERVER[None, None, :] + r[None, :, None] * dirs[:, None, :]
    x, y, z = pos[..., 0], pos[..., 1], pos[..., 2]
    inside = (np.sqrt(x * x + y * y) <= gc_r_max) & (np.abs(z) <= gc_z_max)
    B = lsa_at(x, y, z, **lsa)
    bn = (B * north[:, None, :]).sum(-1)
    be = (B * east[:, None, :]).sum(-1)
    bp2 = bn * bn + be * be
    p = cre['alpha']
    K = L.sync_coefficient(p, nu_ghz, cre['e0'])
    em = cre_at(x, y, z, cre['j0'], cre['r0'], cre['z0'], r_sun) * K * np.power(bp2, (p + 1.0) / 4.0) * (dr * L.KPC_CM)
    return 1.0e3 * (em * inside).sum(axis=1)


def main():
    npix = 12 * 32 ** 2
    pix = [k % npix for k in NOTEBOOK_I_MK]
    ours = stokes_i_mk(pix)
    f = L.rj_to_cmb(23.0)
    print('K_RJ -> K_CMB at 23 GHz: %.6f' % f)
    print('pixel   notebook [mK]   ours K_RJ [mK]  ratio     ours K_CMB [mK]  ratio')
    for (k, want), got in zip(NOTEBOOK_I_MK.items(), ours):
        print('%5d   %.8f      %.8f      %.5f   %.8f       %.5f' % (k % npix, want, got, got / want, got * f,
                                                                  got * f / want))


if __name__ == '__main__':
    main()
