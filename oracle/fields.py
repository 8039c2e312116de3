"""
CPU oracle, stage 1: grid coordinates, analytic + random fields, same-type sum rule,
and the TestSimulator sampling stage.  NumPy restatement of the reference arithmetic
(test infrastructure only -- see oracle/__init__.py).  Canonical units: kpc, microgauss,
cm^-3, rad.

Reference lines restated (relative to /root/reference/imagine):
  grid coordinates        fields/grid.py:229-275   (np.mgrid with n*1j slices)
  cylindrical radius      fields/grid.py:124
  ConstantMagneticField   fields/basic_fields.py:29-40
  ConstantThermalElectrons fields/basic_fields.py:55-56
  ExponentialThermalElectrons fields/basic_fields.py:81-88
  RandomThermalElectrons  fields/basic_fields.py:111-127
  CosThermalElectronDensity fields/test_field.py:36-43
  NaiveGaussianMagneticField fields/test_field.py:84-96
  same-type sum           simulators/simulator.py:166-178
  TestSimulator.simulate  simulators/test_simulator.py:47-69

The GPU generates its Gaussian components with counter-based Philox4x32-10 (defined
here, `philox_normals`), so stochastic fields agree with the reference *in
distribution*; `*_reference` functions restate the reference's MT19937 stream and are
used for the CPU baseline timing and the distribution checks.
"""
import numpy as np

# field kinds -- must match include/imagine_b200.h (IMB_FIELD_*)
CONST_B, GAUSS_B, CONST_NE, EXP_NE, COS_NE, RANDOM_NE, CRE_ANALYTIC, CRE_CONST, LSA_B, JF12_B = range(10)

TYPE_OF_KIND = {CONST_B: 'magnetic_field', GAUSS_B: 'magnetic_field',
                CONST_NE: 'thermal_electron_density', EXP_NE: 'thermal_electron_density',
                COS_NE: 'thermal_electron_density', RANDOM_NE: 'thermal_electron_density',
                CRE_ANALYTIC: 'cosmic_ray_electron_density',
                CRE_CONST: 'cosmic_ray_electron_density', LSA_B: 'magnetic_field',
                JF12_B: 'magnetic_field'}
STOCHASTIC_KINDS = (GAUSS_B, RANDOM_NE)


class Grid:
    """UniformGrid, cartesian: vertex-centred, end-inclusive (grid.py:256-260)."""

    def __init__(self, box, resolution):
        self.box = np.asarray(box, dtype=np.float64).reshape(3, 2)
        self.resolution = np.asarray(resolution, dtype=np.int64).reshape(3)

    @property
    def shape(self):
        return tuple(int(n) for n in self.resolution)

    @property
    def ncell(self):
        return int(np.prod(self.resolution))

    def step(self, axis):
        # numpy's mgrid: step = (stop-start)/float(n-1) when n != 1, else 1
        n = int(self.resolution[axis])
        lo, hi = self.box[axis]
        return (hi - lo) / float(n - 1) if n != 1 else 1.0

    def axis_coords(self, axis):
        # mgrid evaluates index*step + start (two roundings, no fma)
        n = int(self.resolution[axis])
        return np.arange(n, dtype=np.float64) * self.step(axis) + self.box[axis, 0]

    def xyz(self):
        x, y, z = (self.axis_coords(a) for a in range(3))
        return x[:, None, None], y[None, :, None], z[None, None, :]


# ----------------------------------------------------------------------------- analytic
def constant_B(grid, Bx, By, Bz):
    out = np.empty(grid.shape + (3,))
    out[..., 0], out[..., 1], out[..., 2] = Bx, By, Bz
    return out


def constant_ne(grid, ne):
    return np.ones(grid.shape) * ne


def exponential_ne(grid, n0, Re, he):
    x, y, z = grid.xyz()
    R = np.sqrt(x**2 + y**2)
    return n0 * np.exp(-R / Re) * np.exp(-np.abs(z / he)) * np.ones(grid.shape)


def cos_ne(grid, n0, a, alpha, b, beta, c, gamma):
    x, y, z = grid.xyz()
    return ((n0) * (1 + np.cos(a * x + alpha))
            * (1 + np.cos(b * y + beta))
            * (1 + np.cos(c * z + gamma))) * np.ones(grid.shape)


def cre_analytic(grid, j0, r0, z0, r_sun):
    """
    NEW API surface (the reference's CosmicRayElectronDensityField is a stub,
    base_fields.py:104-113).  Gridded CRE flux normalisation at the reference energy E0,
    in [GeV m^2 s sr]^-1, WMAP-style disc: j0 * exp(-(R - r_sun)/r0) * sech^2(z/z0)
    (parameter names of fields/hamx/cre_analytic.py:19-27).
    """
    x, y, z = grid.xyz()
    R = np.sqrt(x**2 + y**2)
    ch = np.cosh(z / z0)
    return j0 * np.exp(-(R - r_sun) / r0) * (1.0 / (ch * ch)) * np.ones(grid.shape)


def lsa_B(grid, b0, psi0, psi1, chi0, r0=8.0, z0=1.0, r_min=3.0, r_max=20.0):
    """
    WMAP-3yr logarithmic-spiral-arm regular field (Page et al. 2007), the model behind
    hammurabiX's `breg_lsa`; the reference only forwards (b0, psi0, psi1, chi0) to the binary
    (fields/hamx/breg_lsa.py:20-31), so like the LOS integrator this is defined here, unpinned.
    Angles in rad.  B = b0 (sin psi cos chi R^ + cos psi cos chi phi^ + sin chi z^),
    psi = psi0 + psi1 ln(R/r0), chi = chi0 tanh(z/z0); zero outside r_min < R < r_max.
    """
    x, y, z = (v * np.ones(grid.shape) for v in grid.xyz())
    R = np.sqrt(x**2 + y**2)
    ok = (R > r_min) & (R < r_max)
    Rs = np.where(ok, R, 1.0)
    cp, sp = x / Rs, y / Rs
    psi = psi0 + psi1 * np.log(Rs / r0)
    chi = chi0 * np.tanh(z / z0)
    cc = np.cos(chi)
    br, bphi = b0 * np.sin(psi) * cc, b0 * np.cos(psi) * cc
    out = np.stack([br * cp - bphi * sp, br * sp + bphi * cp, b0 * np.sin(chi)], axis=-1)
    return np.where(ok[..., None], out, 0.0)


# spiral-arm boundaries on the negative x axis [kpc] and relative cross-sections of the eight arm
# regions (Jansson & Farrar 2012, ApJ 757, 14, section 5.1.1 / table 1)
JF12_RC = (5.1, 6.3, 7.1, 8.3, 9.8, 11.4, 12.7, 15.5)
JF12_F = (0.130, 0.165, 0.094, 0.122, 0.13, 0.118, 0.084, 0.156)
JF12_PITCH_DEG = 11.5


def jf12_B(grid, b_arm1, b_arm2, b_arm3, b_arm4, b_arm5, b_arm6, b_arm7, b_ring, h_disk, w_disk,
           Bn, Bs, rn, rs, wh, z0, B0_X, Xtheta_deg, rpc_X, r0_X):
    """
    Jansson & Farrar (2012) regular Galactic magnetic field, the model behind hammurabiX's `breg_jf12`;
    the reference only forwards the 20 parameters to the binary (fields/hamx/breg_jf12.py:17-39), so like
    LSA it is defined here from the paper (eqs. 5-9), unpinned by reference code.  Parameter names and
    order are the reference's.  Zero for r_spherical < 1 kpc and R > 20 kpc.
      disk (R 3-5 kpc): azimuthal ring b_ring; (R > 5): eight logarithmic spiral regions of pitch 11.5 deg,
        field b_i (5 kpc / R) along the spiral, b_8 fixed by flux conservation; x (1 - L(z, h_disk, w_disk))
      toroidal halo: exp(-|z|/z0) L(z, h_disk, w_disk) Bn (1 - L(R, rn, wh))  (north; Bs, rs south)
      X field: poloidal, elevation Xtheta outside r_p = rpc_X, b_X(r_p) = B0_X exp(-r_p / r0_X)
    with L(x, h, w) = 1 / (1 + exp(-2 (|x| - h) / w)).
    """
    x, y, z = (v * np.ones(grid.shape) for v in grid.xyz())
    R = np.sqrt(x**2 + y**2)
    rsph = np.sqrt(R**2 + z**2)
    Rs = np.where(R > 0, R, 1.0)
    cp, sp = x / Rs, y / Rs
    phi = np.arctan2(y, x)

    def L(v, h, w):
        return 1.0 / (1.0 + np.exp(-2.0 * (np.abs(v) - h) / w))

    pitch = np.deg2rad(JF12_PITCH_DEG)
    sin_i, cos_i = np.sin(pitch), np.cos(pitch)
    b = [b_arm1, b_arm2, b_arm3, b_arm4, b_arm5, b_arm6, b_arm7]
    b8 = -sum(f * bi for f, bi in zip(JF12_F[:7], b)) / JF12_F[7]
    b.append(b8)
    # disk
    lfd = L(z, h_disk, w_disk)
    # radius at which the spiral through (R, phi) crosses the negative x axis, folded into [rc_8/e^(2pi tan), rc_8]
    r_negx = R * np.exp(-(phi - np.pi) / np.tan(np.pi / 2.0 - pitch))
    for _ in range(2):
        r_negx = np.where(r_negx > JF12_RC[7], R * np.exp(-(phi + np.pi) / np.tan(np.pi / 2.0 - pitch)), r_negx)
        r_negx = np.where(r_negx > JF12_RC[7], R * np.exp(-(phi + 3.0 * np.pi) / np.tan(np.pi / 2.0 - pitch)), r_negx)
    arm = np.zeros(grid.shape)
    for i in range(7, -1, -1):
        arm = np.where(r_negx < JF12_RC[i], b[i], arm)
    bmag_spiral = arm * (5.0 / Rs)
    b_disk_r = np.where(R > 5.0, bmag_spiral * sin_i, 0.0)
    b_disk_phi = np.where(R > 5.0, bmag_spiral * cos_i, np.where(R > 3.0, b_ring, 0.0))
    b_disk_r, b_disk_phi = b_disk_r * (1.0 - lfd), b_disk_phi * (1.0 - lfd)
    # toroidal halo
    b_tor = np.exp(-np.abs(z) / z0) * lfd * np.where(z >= 0, Bn * (1.0 - L(R, rn, wh)), Bs * (1.0 - L(R, rs, wh)))
    # X field
    tan_x = np.tan(np.deg2rad(Xtheta_deg))
    rc_bound = rpc_X + np.abs(z) / tan_x
    outer = R >= rc_bound
    rp_out = R - np.abs(z) / tan_x
    rp_in = R * rpc_X / rc_bound
    rp = np.where(outer, rp_out, rp_in)
    bx_mag = B0_X * np.exp(-rp / r0_X) * np.where(outer, rp / Rs, (rp / Rs) ** 2)
    theta_x = np.where(outer, np.deg2rad(Xtheta_deg),
                       np.arctan2(np.abs(z), np.where(outer, 1.0, R - rp_in)))
    theta_x = np.where((~outer) & (R - rp_in <= 0), np.pi / 2.0, theta_x)
    b_x_r = bx_mag * np.cos(theta_x) * np.where(z >= 0, 1.0, -1.0)
    b_x_z = bx_mag * np.sin(theta_x)
    br = b_disk_r + b_x_r
    bphi = b_disk_phi + b_tor
    out = np.stack([br * cp - bphi * sp, br * sp + bphi * cp, b_x_z], axis=-1)
    ok = (rsph >= 1.0) & (R <= 20.0) & (R > 0)
    return np.where(ok[..., None], out, 0.0)


def cre_const(grid, j0):
    return np.ones(grid.shape) * j0


# ------------------------------------------------------------- reference RNG (MT19937)
def gaussian_B_reference(grid, a0, b0, seed):
    """scipy.stats.norm(loc,scale).rvs(size, random_state=seed) == this (bit-identical)."""
    rs = np.random.RandomState(int(seed))
    return rs.standard_normal(grid.shape + (3,)) * b0 + a0


def random_ne_reference(grid, mean, std, min_ne, seed):
    rs = np.random.RandomState(int(seed))
    out = rs.standard_normal(grid.shape) * std + mean
    if np.isfinite(min_ne):
        out[out < min_ne] = min_ne
    return out


# -------------------------------------------------------------- Philox4x32-10 (GPU RNG)
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011).  Inputs/outputs uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for r in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def _u52(hi, lo):
    """(hi, lo) uint32 -> double in (0,1): ((hi<<20 | lo>>12) + 0.5) * 2^-52, exact."""
    k = (hi.astype(np.uint64) << np.uint64(20)) | (lo.astype(np.uint64) >> np.uint64(12))
    return (k.astype(np.float64) + 0.5) * (2.0 ** -52)


def philox_normal_blocks(seed, stream, block_index):
    """
    Box-Muller pair for each Philox block.  counter = (block_lo, block_hi, stream, 0),
    key = (seed_lo, seed_hi).  Returns (z0, z1) = r*(cos, sin)(2*pi*u2), r = sqrt(-2 ln u1).
    """
    b = np.asarray(block_index, dtype=np.uint64)
    seed = int(seed)
    x0, x1, x2, x3 = philox4x32_10(b & _MASK, b >> np.uint64(32),
                                   np.full(b.shape, stream, dtype=np.uint64),
                                   np.zeros(b.shape, dtype=np.uint64),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u1, u2 = _u52(x0, x1), _u52(x2, x3)
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * u2  # in units of pi, exact
    return r * _cospi(ang), r * _sinpi(ang)


def _reduce_pi(t):
    # t in [0,2): exact reduction to quadrant q and f in [-0.25, 0.25]
    q = np.rint(t * 2.0)
    f = t - q * 0.5
    return q.astype(np.int64) & 3, f


def _cospi(t):
    q, f = _reduce_pi(t)
    c, s = np.cos(np.pi * f), np.sin(np.pi * f)
    return np.choose(q, [c, -s, -c, s])


def _sinpi(t):
    q, f = _reduce_pi(t)
    c, s = np.cos(np.pi * f), np.sin(np.pi * f)
    return np.choose(q, [s, c, -s, -c])


def philox_normals(seed, stream, count):
    """Flat stream: value n = element (n & 1) of block (n >> 1)."""
    nblk = (count + 1) // 2
    z0, z1 = philox_normal_blocks(seed, stream, np.arange(nblk, dtype=np.uint64))
    out = np.empty(2 * nblk)
    out[0::2], out[1::2] = z0, z1
    return out[:count]


def gaussian_B_philox(grid, a0, b0, seed, stream):
    z = philox_normals(seed, stream, 3 * grid.ncell).reshape(grid.shape + (3,))
    return z * b0 + a0


def random_ne_philox(grid, mean, std, min_ne, seed, stream):
    out = philox_normals(seed, stream, grid.ncell).reshape(grid.shape) * std + mean
    if np.isfinite(min_ne):
        out[out < min_ne] = min_ne
    return out


# ------------------------------------------------------------------ descriptor evaluator
def evaluate_fields(grid, descs, seeds, rng='philox'):
    """
    Mirror of the C-ABI ``imb_field_eval``.  ``descs`` = list of (kind, params, stream[, seed_row]);
    ``seeds`` = per-realisation seeds [n_real] or a table [n_seed_rows, n_real] (one row per
    FieldFactory seed list; ``seed_row`` picks the row).  Same-type fields are
    summed left to right (simulator.py:175-178).  Returns {field_type: array} where the
    array has a leading realisation axis of length len(seeds) if any contributing field
    is stochastic, else length 1.
    """
    seed_table = np.atleast_2d(np.asarray(seeds, dtype=np.uint64))
    nreal = seed_table.shape[1]
    descs = [tuple(d) + (0,) if len(d) == 3 else tuple(d) for d in descs]
    out = {}
    for ftype in dict.fromkeys(TYPE_OF_KIND[d[0]] for d in descs):
        mine = [d for d in descs if TYPE_OF_KIND[d[0]] == ftype]
        stochastic = any(d[0] in STOCHASTIC_KINDS for d in mine)
        reals = []
        for r in range(nreal if stochastic else 1):
            total = None
            for kind, p, stream, row in mine:
                seeds = seed_table[row]
                if kind == CONST_B:
                    f = constant_B(grid, *p[:3])
                elif kind == LSA_B:
                    f = lsa_B(grid, *p[:8])
                elif kind == JF12_B:
                    f = jf12_B(grid, *p[:20])
                elif kind == GAUSS_B:
                    f = (gaussian_B_philox(grid, p[0], p[1], seeds[r], stream) if rng == 'philox'
                         else gaussian_B_reference(grid, p[0], p[1], seeds[r]))
                elif kind == CONST_NE:
                    f = constant_ne(grid, p[0])
                elif kind == EXP_NE:
                    f = exponential_ne(grid, *p[:3])
                elif kind == COS_NE:
                    f = cos_ne(grid, *p[:7])
                elif kind == RANDOM_NE:
                    f = (random_ne_philox(grid, p[0], p[1], p[2], seeds[r], stream) if rng == 'philox'
                         else random_ne_reference(grid, p[0], p[1], p[2], seeds[r]))
                elif kind == CRE_ANALYTIC:
                    f = cre_analytic(grid, *p[:4])
                elif kind == CRE_CONST:
                    f = cre_const(grid, p[0])
                else:
                    raise ValueError(kind)
                total = f if total is None else total + f
            reals.append(total)
        out[ftype] = np.stack(reals)
    return out


# --------------------------------------------------------------------- TestSimulator
def test_simulator_sample(grid, B, ne, points, b_axis=1):
    """
    t = B[..., b_axis] * ne at the nearest vertex of each point (test_simulator.py:47-69;
    scipy RegularGridInterpolator(method='nearest'), default bounds_error=True).
    ``points`` is [P,3] (x,y,z in kpc).  Ties go to the lower vertex (norm_distance <= .5).
    """
    t = B[..., b_axis] * ne
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    idx = []
    for a in range(3):
        c = grid.axis_coords(a)
        p = points[:, a]
        if np.any(p < c[0]) or np.any(p > c[-1]):
            raise ValueError('One of the requested xi is out of bounds in dimension %d' % a)
        n = c.size
        if n == 1:
            idx.append(np.zeros(p.shape, dtype=np.int64))
            continue
        i = np.clip(np.searchsorted(c, p, side='right') - 1, 0, n - 2)
        yi = (p - c[i]) / (c[i + 1] - c[i])
        idx.append(np.where(yi <= .5, i, i + 1))
    return t[tuple(idx)]
