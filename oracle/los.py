"""
CPU oracle, stage 2: line-of-sight integration of gridded fields to sky maps.

PARITY: STOKES I PINNED, THE REST UNPINNED BY THE REFERENCE.  IMAGINE delegates this arithmetic to the
external C++ program hammurabiX v2.4.2 (docker/Dockerfile:42, requirements.txt:9 `hampyx>=2.4.0`;
call sites simulators/hammurabi.py:83,207,212).  Its source is not under /root/reference
and no reference test exercises it, so this file *defines* the result.  It is pinned by
analytic closed forms (tests/test_oracle_los.py) and -- total intensity only -- by the one hammurabiX
output the reference records (tutorials/tutorial_hammurabi.ipynb cell 14: Stokes I of the built-in LSA +
analytic CRE models, reproduced to 3e-5 after the K_RJ -> K_CMB conversion `rj_to_cmb`;
tools/hammurabi_tutorial_pin.py, test_tutorial_hammurabi_pin).  Faraday depth sign and the Q/U convention
have no recorded output to compare with (docs/los_conventions.md).  What the reference
does fix, and this file follows:
  * observable keys / units: fd [rad m^-2], dm [pc cm^-3], sync I/Q/U/PI [K], PA [rad]
    (simulators/hammurabi.py:56, 214-227; observables/observable_dict.py:11-25, IAU Q/U)
  * field layout: C-order float64 [nx,ny,nz(,3)], box extents + vertex counts
    (simulators/hammurabi.py:229-273)
  * geometry defaults: observer (-8.3, 0, 0.006) kpc, oc_r_max = 30 kpc
    (tests/data/test_ham.xml:26-30, 122-131)
  * RM constant 0.812 rad m^-2 / (cm^-3 uG pc)  (tutorials/tutorial_simulator.ipynb:81)

Definition (frozen):
  sample j of pixel p:  r_j = (j + 1/2) dr,  x_j = o + r_j n(p),  j = 0 .. n_steps-1
  fields at x_j by trilinear interpolation on the vertex grid; zero outside the box
  B_par = -B.n (component toward the observer; RM > 0 for a field pointing at the observer)
  DM  = sum ne dl_pc                        dl_pc = 1e3 dr
  f_j = 0.812 ne B_par dl_pc ;  RM = sum f_j ;  phi_j = sum_{i<j} f_i + f_j/2
  sky basis at n: East = (-sin l, cos l, 0), North = (-sin b cos l, -sin b sin l, cos b)
  B_N = B.North, B_E = B.East, Bperp^2 = B_N^2 + B_E^2
  p = alpha - beta R - theta z   (CRE spectral index, test_ham.xml:356 comment)
  em_j = cre * K(p, nu) * (Bperp^2)^((p+1)/4) * dl_cm        (Rybicki & Lightman eq. 6.36,
         cre = flux normalisation at E0 in [GeV m^2 s sr]^-1, brightness temperature K_RJ)
  I = sum em_j ;  Q + iU = sum -em_j Pi(p) (B_N + i B_E)^2 / Bperp^2 * exp(2 i phi_j lambda^2)
  Pi(p) = (p+1)/(p+7/3)      (IAU: chi = psi_B + pi/2, psi_B from North through East)
"""
import numpy as np
from scipy.special import gamma as _gamma

# CGS constants (frozen; mirrored bit-for-bit in imagine_b200/constants.py)
C_LIGHT = 2.99792458e10        # cm / s
Q_E = 4.80320471e-10           # esu
M_E = 9.10938356e-28           # g
K_B = 1.38064852e-16           # erg / K
GEV_ERG = 1.6021766208e-3      # erg
KPC_CM = 3.0856775814913673e21  # cm
MEC2_ERG = M_E * C_LIGHT * C_LIGHT
RM_CONST = 0.812               # rad m^-2 per (cm^-3 uG pc)


def sync_coefficient(p, nu_ghz, e0_gev):
    """K(p, nu): brightness temperature [K] per (cre * (Bperp_uG^2)^((p+1)/4) * cm)."""
    p = np.asarray(p, dtype=np.float64)
    nu = nu_ghz * 1.0e9
    t_conv = C_LIGHT * C_LIGHT / (2.0 * K_B * nu * nu)
    ptot = (np.sqrt(3.0) * Q_E**3 / (2.0 * np.pi * MEC2_ERG * (p + 1.0))
            * _gamma(p / 4.0 + 19.0 / 12.0) * _gamma(p / 4.0 - 1.0 / 12.0)
            * (2.0 * np.pi * M_E * C_LIGHT * nu / (3.0 * Q_E)) ** (-(p - 1.0) / 2.0))
    c_gamma = ((4.0 * np.pi / C_LIGHT) * (e0_gev * GEV_ERG) ** p
               * MEC2_ERG ** (1.0 - p) / (GEV_ERG * 1.0e4))
    return t_conv * 0.5 * ptot * c_gamma * (1.0e-6) ** ((p + 1.0) / 2.0)


def pol_fraction(p):
    return (p + 1.0) / (p + 7.0 / 3.0)


H_PLANCK = 6.62607015e-27      # erg s
T_CMB = 2.725                  # K


def rj_to_cmb(nu_ghz, t_cmb=T_CMB):
    """dT_CMB / dT_RJ = (e^x - 1)^2 / (x^2 e^x), x = h nu / (k T_cmb): hammurabiX reports K-cmb
    (observables/observable_dict.py:15-19; pinned by tools/hammurabi_tutorial_pin.py)."""
    x = H_PLANCK * nu_ghz * 1.0e9 / (K_B * t_cmb)
    return (np.exp(x) - 1.0) ** 2 / (x * x * np.exp(x))


def lambda_sq_m2(nu_ghz):
    lam = (C_LIGHT * 1.0e-2) / (nu_ghz * 1.0e9)
    return lam * lam


class LosConfig:
    def __init__(self, observer=(-8.3, 0.0, 0.006), dr=0.1, r_max=30.0, n_steps=None,
                 nu_ghz=23.0, alpha=3.0, beta=0.0, theta=0.0, e0_gev=20.6):
        self.observer = np.asarray(observer, dtype=np.float64)
        self.dr = float(dr)
        self.n_steps = int(n_steps) if n_steps is not None else int(np.ceil(r_max / dr - 1e-9))
        self.nu_ghz, self.alpha, self.beta, self.theta, self.e0_gev = (
            float(nu_ghz), float(alpha), float(beta), float(theta), float(e0_gev))


def inv_steps(grid):
    """(n-1)/(hi-lo) per axis: shared verbatim by oracle and GPU."""
    return np.array([(int(grid.resolution[a]) - 1) / (grid.box[a, 1] - grid.box[a, 0])
                     for a in range(3)])


def sample_geometry(grid, cfg, dirs):
    """Positions, in-box mask, base cell and weights for all (pixel, step) samples."""
    n = dirs.reshape(-1, 3)
    r = (np.arange(cfg.n_steps, dtype=np.float64) + 0.5) * cfg.dr
    pos = cfg.observer[None, None, :] + r[None, :, None] * n[:, None, :]      # [m,J,3]
    inv = inv_steps(grid)
    u = (pos - grid.box[None, None, :, 0]) * inv[None, None, :]
    nmax = (grid.resolution - 1).astype(np.float64)
    inside = np.all((u >= 0.0) & (u <= nmax[None, None, :]), axis=-1)
    ui = np.where(inside[..., None], u, 0.0)
    i0 = np.minimum(ui.astype(np.int64), (grid.resolution - 2)[None, None, :])
    t = ui - i0
    return pos, inside, i0, t


def _trilinear(f, i0, t):
    """f [nx,ny,nz,(c)] at cells i0 [...,3] with weights t [...,3]; z first, then y, x."""
    ix, iy, iz = i0[..., 0], i0[..., 1], i0[..., 2]
    tx, ty, tz = t[..., 0], t[..., 1], t[..., 2]
    if f.ndim == 4:
        tx, ty, tz = tx[..., None], ty[..., None], tz[..., None]

    def zl(a, b):
        v0, v1 = f[ix + a, iy + b, iz], f[ix + a, iy + b, iz + 1]
        return v0 + tz * (v1 - v0)
    c00, c01, c10, c11 = zl(0, 0), zl(0, 1), zl(1, 0), zl(1, 1)
    c0 = c00 + ty * (c01 - c00)
    c1 = c10 + ty * (c11 - c10)
    return c0 + tx * (c1 - c0)


def sky_basis(dirs):
    n = dirs.reshape(-1, 3)
    sth = np.sqrt(n[:, 0] ** 2 + n[:, 1] ** 2)
    ok = sth > 0
    cph = np.where(ok, n[:, 0] / np.where(ok, sth, 1.0), 1.0)
    sph = np.where(ok, n[:, 1] / np.where(ok, sth, 1.0), 0.0)
    east = np.stack([-sph, cph, np.zeros_like(sph)], axis=-1)
    north = np.stack([-n[:, 2] * cph, -n[:, 2] * sph, sth], axis=-1)
    return north, east


def integrate(grid, cfg, dirs, B, ne=None, cre=None, pixel_chunk=2048):
    """
    B [Nb,nx,ny,nz,3], ne [Nn,nx,ny,nz] or None, cre [Nc,nx,ny,nz] or None with
    Nb/Nn/Nc in {1, N}.  Returns dict of [N, m] maps: 'dm', 'fd', 'I', 'Q', 'U'
    (entries requiring a missing field are omitted).
    """
    dirs = np.asarray(dirs, dtype=np.float64).reshape(-1, 3)
    m = dirs.shape[0]
    N = max(B.shape[0], 1 if ne is None else ne.shape[0], 1 if cre is None else cre.shape[0])
    out = {}
    if ne is not None:
        out['dm'] = np.zeros((N, m))
        out['fd'] = np.zeros((N, m))
    if cre is not None:
        for k in 'IQU':
            out[k] = np.zeros((N, m))
    dl_pc = cfg.dr * 1.0e3
    dl_cm = cfg.dr * KPC_CM
    lam2 = lambda_sq_m2(cfg.nu_ghz)
    const_p = (cfg.beta == 0.0 and cfg.theta == 0.0)
    north, east = sky_basis(dirs)

    for s in range(0, m, pixel_chunk):
        sl = slice(s, min(m, s + pixel_chunk))
        nd = dirs[sl]
        pos, inside, i0, t = sample_geometry(grid, cfg, nd)
        w = inside.astype(np.float64)
        if cre is not None:
            if const_p:
                p = np.float64(cfg.alpha)
            else:
                R = np.sqrt(pos[..., 0] ** 2 + pos[..., 1] ** 2)
                p = cfg.alpha - cfg.beta * R - cfg.theta * pos[..., 2]
            Kp = sync_coefficient(p, cfg.nu_ghz, cfg.e0_gev)
            Pi = pol_fraction(p)
            ex = (p + 1.0) / 4.0
        for r in range(N):
            Bs = _trilinear(B[r if B.shape[0] > 1 else 0], i0, t) * w[..., None]   # [mc,J,3]
            if ne is not None:
                nes = _trilinear(ne[r if ne.shape[0] > 1 else 0], i0, t) * w
                bpar = -(Bs[..., 0] * nd[:, None, 0] + Bs[..., 1] * nd[:, None, 1]
                         + Bs[..., 2] * nd[:, None, 2])
                f = RM_CONST * nes * bpar * dl_pc
                out['dm'][r, sl] = (nes * dl_pc).sum(axis=1)
                out['fd'][r, sl] = f.sum(axis=1)
                phi = np.cumsum(f, axis=1) - 0.5 * f
            else:
                phi = 0.0
            if cre is not None:
                cres = _trilinear(cre[r if cre.shape[0] > 1 else 0], i0, t) * w
                bn = (Bs * north[sl][:, None, :]).sum(axis=-1)
                be = (Bs * east[sl][:, None, :]).sum(axis=-1)
                bp2 = bn * bn + be * be
                pos_b = bp2 > 0.0
                safe = np.where(pos_b, bp2, 1.0)
                em = np.where(pos_b, cres * Kp * np.power(safe, ex) * dl_cm, 0.0)
                qc = -em * Pi * (bn * bn - be * be) / safe
                uc = -em * Pi * (2.0 * bn * be) / safe
                ang = 2.0 * phi * lam2
                ca, sa = np.cos(ang), np.sin(ang)
                out['I'][r, sl] = em.sum(axis=1)
                out['Q'][r, sl] = (qc * ca - uc * sa).sum(axis=1)
                out['U'][r, sl] = (qc * sa + uc * ca).sum(axis=1)
    return out


def in_box_samples(grid, cfg, dirs, pixel_chunk=4096):
    """Exact number of (pixel, step) samples that fall inside the box (reported by bench)."""
    dirs = np.asarray(dirs, dtype=np.float64).reshape(-1, 3)
    total = 0
    for s in range(0, dirs.shape[0], pixel_chunk):
        _, inside, _, _ = sample_geometry(grid, cfg, dirs[s:s + pixel_chunk])
        total += int(inside.sum())
    return total
