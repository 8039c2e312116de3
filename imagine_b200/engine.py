"""
Tensor-level Python face of the C ABI: torch tensors in (device buffers only), torch tensors
or floats out.  Every function here is a thin argument marshaller around one `imb_*` entry
point; the arithmetic lives in csrc/*.cu.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .constants import sync_coefficient


def _torch():
    import torch
    return torch


def _f64(t, name):
    torch = _torch()
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64):
        raise TypeError('%s must be a CUDA float64 tensor' % name)
    return t if t.is_contiguous() else t.contiguous()


class DeviceGrid:
    """UniformGrid geometry (fields/grid.py:189-275), cartesian, canonical units (kpc)."""

    def __init__(self, box, resolution):
        self.box = np.asarray(box, dtype=np.float64).reshape(3, 2)
        self.resolution = np.asarray(resolution, dtype=np.int64).reshape(3)
        self.c = L.make_grid(self.box, self.resolution)

    @property
    def shape(self):
        return tuple(int(n) for n in self.resolution)

    @property
    def ncell(self):
        return int(np.prod(self.resolution))


class LosSettings:
    """Observer geometry + radial sampling + synchrotron spectrum (imb_los_cfg)."""

    def __init__(self, observer=(-8.3, 0.0, 0.006), dr=0.1, r_max=30.0, n_steps=None,
                 nu_ghz=23.0, alpha=3.0, beta=0.0, theta=0.0, e0_gev=20.6):
        self.observer = tuple(float(v) for v in observer)
        self.dr = float(dr)
        self.n_steps = int(n_steps) if n_steps is not None else int(np.ceil(r_max / dr - 1e-9))
        self.nu_ghz, self.alpha, self.beta, self.theta, self.e0_gev = (
            float(nu_ghz), float(alpha), float(beta), float(theta), float(e0_gev))

    def with_frequency(self, nu_ghz):
        return LosSettings(self.observer, self.dr, None, self.n_steps, nu_ghz, self.alpha,
                           self.beta, self.theta, self.e0_gev)

    def c_struct(self, out_cols_per_dest=0, out_dest_stride=0, peer_delta=None):
        cfg = L.LosCfg()
        cfg.out_cols_per_dest, cfg.out_dest_stride = int(out_cols_per_dest), int(out_dest_stride)
        if peer_delta is not None:
            cfg.n_peers = len(peer_delta)
            for d, v in enumerate(peer_delta):
                cfg.peer_delta[d] = int(v)
        for a in range(3):
            cfg.observer[a] = self.observer[a]
        cfg.dr, cfg.n_steps, cfg.nu_ghz = self.dr, self.n_steps, self.nu_ghz
        cfg.alpha, cfg.beta, cfg.theta, cfg.e0_gev = self.alpha, self.beta, self.theta, self.e0_gev
        cfg.sync_k = sync_coefficient(self.alpha, self.nu_ghz, self.e0_gev)
        return cfg


# --------------------------------------------------------------------------------- stage 1
_FIELD_ROWS = {}


def field_rows(descs, n_real):
    """Rows per field type (imb_field_rows): a function of the kinds and n_real only -- memoised, a sampler
    asks the same question ten thousand times."""
    key = (tuple(int(d[0]) for d in descs), int(n_real))
    rows = _FIELD_ROWS.get(key)
    if rows is None:
        arr, n = L.make_field_descs(descs)
        lib = L.load()
        rows = _FIELD_ROWS[key] = tuple(lib.imb_field_rows(arr, n, t, n_real) for t in range(3))
    return rows


FIELD_NAMES = ('magnetic_field', 'thermal_electron_density', 'cosmic_ray_electron_density')
_FIELD_TYPE = {}


def field_type(kind):
    """0 magnetic field, 1 thermal electrons, 2 cosmic-ray electrons (what imb_field_rows says)."""
    t = _FIELD_TYPE.get(kind)
    if t is None:
        rows = field_rows([(kind, [0.0], 0, 0)], 1)
        t = _FIELD_TYPE[kind] = next(i for i in range(3) if rows[i] > 0)
    return t


INTERLEAVED_STRIDES = (24, 48, 96)


def pick_b_stride(n_real):
    """Widest realisation interleave a block of `n_real` realisations fills without padding:
    96 / 48 / 24 = [group][cell][3][32 / 16 / 8] (lane-per-realisation LOS kernel, los_rl.cu); any other
    count takes the 8-wide layout with the last group padded (idle lanes)."""
    for rl in (32, 16, 8):
        if n_real % rl == 0:
            return 3 * rl
    return 24


def b_shape(rows, ncell, b_stride):
    """Shape of the magnetic-field buffer for a layout code."""
    if b_stride in (24, 48, 96):
        rl = b_stride // 3
        return ((rows + rl - 1) // rl, ncell, 3, rl)    # [group][cell][x|y|z][realisation % RL]
    return (rows, ncell, b_stride)


def b_stride_of(B):
    """Layout code of a magnetic-field buffer (inverse of `b_shape`)."""
    if B.dim() == 4:
        return 3 * int(B.shape[-1])
    return int(B.shape[-1])


def field_shapes(grid, descs, n_real, b_stride=4):
    """Buffer shapes `field_eval` will use/expect: {name: shape}, plus the effective b_stride."""
    rows = field_rows(descs, n_real)
    if b_stride == 'auto':
        b_stride = pick_b_stride(n_real)
    if b_stride in INTERLEAVED_STRIDES and rows[0] != n_real:
        b_stride = 4
    shapes = {}
    for t, name in enumerate(FIELD_NAMES):
        if rows[t] == 0:
            continue
        shapes[name] = b_shape(rows[0], grid.ncell, b_stride) if t == 0 else (rows[t], grid.ncell)
    return shapes, b_stride


def field_eval(grid, descs, seeds, n_real=None, b_stride=4, out=None):
    """
    Evaluate and sum all fields (imb_field_eval).  Returns dict with any of
    'magnetic_field' [rows, ncell, b_stride] (or an interleaved `b_shape`), 'thermal_electron_density' [rows, ncell],
    'cosmic_ray_electron_density' [rows, ncell]; rows is n_real when a stochastic field of the
    type is present and 1 otherwise.  `descs` entries are (kind, params, philox_stream[, seed_row]);
    `seeds` is [n_real] or [n_seed_rows, n_real].
    """
    torch = _torch()
    dev = L.init()
    seeds = np.ascontiguousarray(np.atleast_2d(np.asarray(seeds, dtype=np.uint64)))
    if n_real is None:
        n_real = int(seeds.shape[1])
    if seeds.ndim != 2 or seeds.shape[1] != n_real:
        raise ValueError('need one seed per realisation (rows = seed lists)')
    arr, n = L.make_field_descs(descs)
    rows = field_rows(descs, n_real)
    names = ('magnetic_field', 'thermal_electron_density', 'cosmic_ray_electron_density')
    bufs = {}
    if b_stride == 'auto':
        b_stride = pick_b_stride(n_real)
    if b_stride in INTERLEAVED_STRIDES and rows[0] != n_real:
        b_stride = 4          # interleaving realisations only makes sense for a stochastic B
    for t, name in enumerate(names):
        if rows[t] == 0:
            continue
        shape = b_shape(rows[0], grid.ncell, b_stride) if t == 0 else (rows[t], grid.ncell)
        if out is not None and name in out:
            buf = out[name]
            if tuple(buf.shape) != shape:
                raise ValueError('preallocated %s has shape %s, need %s' % (name, tuple(buf.shape), shape))
        else:
            buf = torch.empty(shape, dtype=torch.float64, device='cuda:%d' % dev)
        bufs[name] = buf
    B, ne, cre = (bufs.get(k) for k in names)
    L.check(L.load().imb_field_eval(
        C.byref(grid.c), arr, n, seeds.ctypes.data_as(C.c_void_p), int(seeds.shape[0]), n_real,
        L.ptr(B), rows[0], b_stride, L.ptr(ne), rows[1], L.ptr(cre), rows[2], L.stream_ptr()))
    return bufs


def field_eval_lanes(grid, descs, seeds, n_real, b_stride, out_B, lane_offset):
    """
    imb_field_eval_lanes: the magnetic field of one parameter point into lanes [lane_offset, lane_offset +
    b_stride / 3) of the wide interleaved buffer `out_B` [1, ncell, 3, RL_out] shared by several points.
    """
    L.init()
    seeds = np.ascontiguousarray(np.atleast_2d(np.asarray(seeds, dtype=np.uint64)))
    if seeds.shape[1] != n_real:
        raise ValueError('need one seed per realisation (rows = seed lists)')
    out_B = _f64(out_B, 'out_B')
    if out_B.dim() != 4 or out_B.shape[0] != 1 or out_B.shape[1] != grid.ncell or out_B.shape[2] != 3:
        raise ValueError('out_B must be [1, ncell, 3, RL]')
    arr, n = L.make_field_descs(descs)
    L.check(L.load().imb_field_eval_lanes(
        C.byref(grid.c), arr, n, seeds.ctypes.data_as(C.c_void_p), int(seeds.shape[0]), int(n_real),
        L.ptr(out_B), int(b_stride), 3 * int(out_B.shape[3]), int(lane_offset), L.stream_ptr()))
    return out_B


def test_simulate(grid, B, ne, n_real, points, b_axis=1, scale=1.0):
    """TestSimulator.simulate for all realisations: [n_real, P] (imb_test_simulate)."""
    torch = _torch()
    L.init()
    B, ne = _f64(B, 'B'), _f64(ne, 'ne')
    pts = _f64(points, 'points').reshape(-1, 3)
    out = torch.empty((n_real, pts.shape[0]), dtype=torch.float64, device=B.device)
    L.check(L.load().imb_test_simulate(
        C.byref(grid.c), L.ptr(B), B.shape[0], B.shape[-1], L.ptr(ne), ne.shape[0], n_real,
        int(b_axis), L.ptr(pts), pts.shape[0], float(scale), L.ptr(out), L.stream_ptr()))
    return out


# --------------------------------------------------------------------------------- stage 2
def healpix_ring_dirs(nside):
    torch = _torch()
    dev = L.init()
    out = torch.empty((12 * int(nside) ** 2, 3), dtype=torch.float64, device='cuda:%d' % dev)
    L.check(L.load().imb_healpix_ring_dirs(int(nside), L.ptr(out), L.stream_ptr()))
    return out


def los_integrate(grid, settings, dirs, B, ne=None, cre=None, n_real=1,
                  want=('dm', 'fd', 'I', 'Q', 'U'), out=None, out_layout=None):
    """
    Integrate every realisation along `dirs` [m,3] (imb_los_integrate).  Returns a dict of
    [n_real, m] tensors for the requested quantities.  With `out_layout = (cols_per_dest,
    dest_stride)` the tensors in `out` are base pointers of a pixel-sharded send buffer
    (pixel p of realisation r at (p // cols) * dest_stride + r * cols + p % cols); a third entry
    `peer_delta` (list of P offsets in doubles) switches on the fused NVLink exchange, where the
    block of destination d lands directly in rank d's symmetric receive buffer.
    """
    torch = _torch()
    L.init()
    dirs = _f64(dirs, 'dirs').reshape(-1, 3)
    m = dirs.shape[0]
    B = _f64(B, 'B')
    ne = None if ne is None else _f64(ne, 'ne')
    cre = None if cre is None else _f64(cre, 'cre')
    res = {}
    for k in ('dm', 'fd', 'I', 'Q', 'U'):
        if k in want:
            if out is not None and k in out:
                res[k] = out[k]
            else:
                res[k] = torch.empty((n_real, m), dtype=torch.float64, device=B.device)
    cfg = settings.c_struct(*(out_layout or (0, 0)))
    interleaved = B.dim() == 4                      # realisation-interleaved layouts (b_shape)
    L.check(L.load().imb_los_integrate(
        C.byref(grid.c), C.byref(cfg), L.ptr(dirs), m, L.ptr(B),
        n_real if interleaved else B.shape[0], b_stride_of(B),
        L.ptr(ne), 0 if ne is None else ne.shape[0], L.ptr(cre), 0 if cre is None else cre.shape[0],
        n_real, L.ptr(res.get('dm')), L.ptr(res.get('fd')), L.ptr(res.get('I')),
        L.ptr(res.get('Q')), L.ptr(res.get('U')), L.stream_ptr()))
    return res


def pol_derive(Q, U, want_pi=True, want_pa=True):
    torch = _torch()
    L.init()
    Q, U = _f64(Q, 'Q'), _f64(U, 'U')
    PI = torch.empty_like(Q) if want_pi else None
    PA = torch.empty_like(Q) if want_pa else None
    L.check(L.load().imb_pol_derive(L.ptr(Q), L.ptr(U), Q.numel(), L.ptr(PI), L.ptr(PA),
                                    L.stream_ptr()))
    return PI, PA


# --------------------------------------------------------------------------------- stage 3
def ensemble_mean(sims):
    torch = _torch()
    L.init()
    sims = _f64(sims, 'sims')
    N, m = sims.shape
    out = torch.empty((1, m), dtype=torch.float64, device=sims.device)
    L.check(L.load().imb_ensemble_mean(L.ptr(sims), N, m, L.ptr(out), L.stream_ptr()))
    return out


def ensemble_var(sims):
    torch = _torch()
    L.init()
    sims = _f64(sims, 'sims')
    N, m = sims.shape
    out = torch.empty((m,), dtype=torch.float64, device=sims.device)
    L.check(L.load().imb_ensemble_var(L.ptr(sims), N, m, L.ptr(out), L.stream_ptr()))
    return out


_EST = {'empirical': L.EST_EMPIRICAL, 'oas': L.EST_OAS, 'diagonal': L.EST_DIAGONAL}
_ROUTE = {'auto': L.ROUTE_AUTO, 'gram': L.ROUTE_GRAM, 'dense': L.ROUTE_DENSE}


def mcov(sims, estimator='oas'):
    """(mean [1,m], cov [m,m], rho) = empirical_mcov / oas_mcov / diagonal_mcov on device."""
    torch = _torch()
    L.init()
    sims = _f64(sims, 'sims')
    N, m = sims.shape
    mean = torch.empty((1, m), dtype=torch.float64, device=sims.device)
    cov = torch.empty((m, m), dtype=torch.float64, device=sims.device)
    rho = C.c_double(0.0)
    L.check(L.load().imb_mcov(L.ptr(sims), N, m, _EST[estimator], L.ptr(mean), L.ptr(cov),
                              C.byref(rho), L.stream_ptr()))
    return mean, cov, rho.value


def _cov_kind(cov):
    if cov is None:
        return L.COV_NONE
    if cov.dim() == 1:
        return L.COV_VARIANCE
    if cov.dim() == 2 and cov.shape[0] == cov.shape[1]:
        return L.COV_DENSE
    raise ValueError('covariance must be [m] (variance) or [m,m]')


def loglike_ensemble(sims, meas, cov, estimator='oas', route='auto',
                     use_trace_approximation=False, return_info=False):
    L.init()
    sims, meas, cov = _f64(sims, 'sims'), _f64(meas, 'meas').reshape(-1), _f64(cov, 'cov')
    N, m = sims.shape
    if meas.numel() != m or cov.shape[0] != m:
        raise ValueError('shape mismatch between simulations, measurements and covariance')
    out, info = C.c_double(0.0), (C.c_int32 * 2)()
    L.check(L.load().imb_loglike_ensemble(
        L.ptr(sims), N, m, L.ptr(meas), L.ptr(cov), _cov_kind(cov), _EST[estimator], _ROUTE[route],
        int(bool(use_trace_approximation)), C.byref(out), info, L.stream_ptr()))
    return (out.value, (info[0], info[1])) if return_info else out.value


def loglike_ensemble_batch(sims, meas, var, estimator='oas', return_info=False):
    """
    Sum-ready per-key values of EnsembleLikelihood for K observables of identical shape with
    diagonal measurement covariances (imb_loglike_ensemble_batch): lists of tensors in, list of
    floats out, one device synchronisation.
    """
    L.init()
    K = len(sims)
    sims = [_f64(t, 'sims') for t in sims]
    meas = [_f64(t, 'meas').reshape(-1) for t in meas]
    var = [_f64(t, 'var').reshape(-1) for t in var]
    N, m = sims[0].shape
    for s_, y_, v_ in zip(sims, meas, var):
        if tuple(s_.shape) != (N, m) or y_.numel() != m or v_.numel() != m:
            raise ValueError('batched keys must share one [N, m] shape')
    arr = lambda ts: (C.c_void_p * K)(*[t.data_ptr() for t in ts])
    out, info = (C.c_double * K)(), (C.c_int32 * (2 * K))()
    L.check(L.load().imb_loglike_ensemble_batch(K, arr(sims), N, m, arr(meas), arr(var), _EST[estimator],
                                                out, info, L.stream_ptr()))
    vals = [out[k] for k in range(K)]
    return (vals, [(info[2 * k], info[2 * k + 1]) for k in range(K)]) if return_info else vals


def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


KMAX_BATCH = 32          # keys (observables x parameter points) per batched launch (gram_batch.cu:KMAX)


def _check_key_views(tensors, name):
    """The phased / asynchronous entry points take raw row pointers (views into exchange buffers are
    legitimate): every tensor must be CUDA float64 with unit stride along the pixels."""
    torch = _torch()
    for t in tensors:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64):
            raise TypeError('%s must be CUDA float64 tensors' % name)
        if t.dim() >= 1 and t.numel() > 0 and t.stride(-1) != 1:
            raise ValueError('%s must be contiguous along the pixel axis' % name)


def _check_out(out, flags, K):
    torch = _torch()
    if K < 1 or K > KMAX_BATCH:
        raise ValueError('between 1 and %d observables per batched call' % KMAX_BATCH)
    if not (out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() >= K):
        raise TypeError('out must be a contiguous CUDA float64 tensor with at least K elements')
    if flags is not None and not (flags.is_cuda and flags.dtype == torch.int32 and flags.is_contiguous()
                                  and flags.numel() >= K):
        raise TypeError('flags must be a contiguous CUDA int32 tensor with at least K elements')


def gram_phase1(sims, N, m_local, layout, meas, G):
    """imb_gram_phase1: partial Gram of the local pixels of every key into G [K, N, N]."""
    L.init()
    rpb, bstride, rstride = layout
    _check_key_views(sims, 'sims'); _check_key_views(meas, 'meas'); _check_key_views([G], 'G')
    if not 1 <= len(sims) <= KMAX_BATCH:
        raise ValueError('between 1 and %d observables per batched call' % KMAX_BATCH)
    L.check(L.load().imb_gram_phase1(len(sims), _ptr_array(sims), N, m_local, rpb, bstride, rstride,
                                     _ptr_array(meas), L.ptr(G), L.stream_ptr()))


def gram_phase2(sims, N, m_local, layout, meas, var, m_total, G, P2, estimator='oas'):
    """imb_gram_phase2: shrinkage from the (all-reduced) G, then the partial P2 [K, N*N+N+4]."""
    L.init()
    rpb, bstride, rstride = layout
    _check_key_views(sims, 'sims'); _check_key_views(meas, 'meas'); _check_key_views(var, 'var')
    _check_key_views([G, P2], 'G / P2')
    if not 1 <= len(sims) <= KMAX_BATCH:
        raise ValueError('between 1 and %d observables per batched call' % KMAX_BATCH)
    L.check(L.load().imb_gram_phase2(len(sims), _ptr_array(sims), N, m_local, rpb, bstride, rstride,
                                     _ptr_array(meas), _ptr_array(var), int(m_total), _EST[estimator],
                                     L.ptr(G), L.ptr(P2), L.stream_ptr()))


def gram_phase3(K, N, m_local, G, P2):
    """imb_gram_phase3: per-key log-likelihoods from the (all-reduced) P2; list of floats."""
    L.init()
    out, info = (C.c_double * K)(), (C.c_int32 * (2 * K))()
    L.check(L.load().imb_gram_phase3(K, N, m_local, L.ptr(G), L.ptr(P2), out, info, L.stream_ptr()))
    for k in range(K):
        if info[2 * k + 1] != 0:
            raise np.linalg.LinAlgError('Gram route not applicable for key %d (info %d): non-positive '
                                        'shrunk diagonal; gather the maps and use the dense route'
                                        % (k, info[2 * k + 1]))
    return [out[k] for k in range(K)]


# ---- collectives behind the C ABI (csrc/comm.cu): for hosts without torch.distributed -------------------
def comm_unique_id():
    """128-byte NCCL id (bytes); rank 0 creates it, the host ships it to the other ranks."""
    L.init()
    buf = (C.c_char * 128)()
    L.check(L.load().imb_comm_unique_id(buf))
    return bytes(buf)


def comm_init(nranks, rank, unique_id):
    L.init()
    buf = (C.c_char * 128).from_buffer_copy(unique_id)
    L.check(L.load().imb_comm_init(int(nranks), int(rank), buf))


def comm_finalize():
    L.check(L.load().imb_comm_finalize())


def comm_size():
    """Ranks of the library's own communicator (0: not initialised)."""
    return int(L.load().imb_comm_size())


def alltoall(send, recv):
    """imb_alltoall: block r of `send` [P, ...] to rank r, block q of `recv` from rank q (same stream)."""
    _check_key_views([send, recv], 'send / recv')
    P = send.shape[0]
    if send.shape != recv.shape or send.data_ptr() == recv.data_ptr() or send.numel() % P:
        raise ValueError('send and recv must be distinct [P, ...] tensors of one shape')
    L.check(L.load().imb_alltoall(L.ptr(send), L.ptr(recv), send.numel() // P, L.stream_ptr()))
    return recv


def allreduce_sum(t):
    """In-place SUM all-reduce of a CUDA float64 tensor over the library's own communicator."""
    _check_key_views([t], 'tensor')
    L.check(L.load().imb_allreduce_sum(L.ptr(t), t.numel(), L.stream_ptr()))
    return t


def loglike_ensemble_sharded(sims, N, m_local, layout, meas, var, m_total, G, P2, estimator='oas'):
    """imb_loglike_ensemble_sharded: the three Gram phases with their two all-reduces in ONE library call."""
    L.init()
    rpb, bstride, rstride = layout
    K = len(sims)
    _check_key_views(sims, 'sims'); _check_key_views(meas, 'meas'); _check_key_views(var, 'var')
    _check_key_views([G, P2], 'G / P2')
    out, info = (C.c_double * K)(), (C.c_int32 * (2 * K))()
    L.check(L.load().imb_loglike_ensemble_sharded(K, _ptr_array(sims), N, m_local, rpb, bstride, rstride,
                                                  _ptr_array(meas), _ptr_array(var), int(m_total),
                                                  _EST[estimator], L.ptr(G), L.ptr(P2), out, info,
                                                  L.stream_ptr()))
    for k in range(K):
        if info[2 * k + 1] != 0:
            raise np.linalg.LinAlgError('Gram route not applicable for key %d (info %d)' % (k, info[2 * k + 1]))
    return [out[k] for k in range(K)]


def loglike_ensemble_batch_async(sims, meas, var, out, flags=None, estimator='oas'):
    """As `loglike_ensemble_batch`, but values go to the device tensor `out` [K] and nothing syncs."""
    L.init()
    K = len(sims)
    sims = [_f64(t, 'sims') for t in sims]
    meas = [_f64(t, 'meas').reshape(-1) for t in meas]
    var = [_f64(t, 'var').reshape(-1) for t in var]
    _check_out(out, flags, K)
    N, m = sims[0].shape
    for s_, y_, v_ in zip(sims, meas, var):
        if tuple(s_.shape) != (N, m) or y_.numel() != m or v_.numel() != m:
            raise ValueError('every observable of a batch must have the same [N, m] shape')
    L.check(L.load().imb_loglike_ensemble_batch_async(K, _ptr_array(sims), N, m, _ptr_array(meas),
                                                      _ptr_array(var), _EST[estimator], L.ptr(out),
                                                      L.ptr(flags), L.stream_ptr()))


def gram_phase3_async(K, N, m_local, G, P2, out, flags=None):
    """imb_gram_phase3_async: per-key values into the device tensor `out` [K]; no synchronisation."""
    L.init()
    _check_key_views([G, P2], 'G / P2')
    _check_out(out, flags, K)
    L.check(L.load().imb_gram_phase3_async(K, N, m_local, L.ptr(G), L.ptr(P2), L.ptr(out), L.ptr(flags),
                                           L.stream_ptr()))


def loglike_ensemble_diagonal(sims, meas, var):
    L.init()
    sims, meas, var = _f64(sims, 'sims'), _f64(meas, 'meas').reshape(-1), _f64(var, 'var')
    N, m = sims.shape
    if meas.numel() != m or var.numel() != m:
        raise ValueError('shape mismatch between simulations, measurements and variance')
    out = C.c_double(0.0)
    L.check(L.load().imb_loglike_ensemble_diagonal(L.ptr(sims), N, m, L.ptr(meas), L.ptr(var),
                                                   C.byref(out), L.stream_ptr()))
    return out.value


def loglike_simple(sims, meas, cov=None, return_info=False):
    L.init()
    sims, meas = _f64(sims, 'sims'), _f64(meas, 'meas').reshape(-1)
    cov = None if cov is None else _f64(cov, 'cov')
    N, m = sims.shape
    if meas.numel() != m or (cov is not None and cov.shape[0] != m):
        raise ValueError('shape mismatch between simulations, measurements and covariance')
    out, info = C.c_double(0.0), (C.c_int32 * 2)()
    L.check(L.load().imb_loglike_simple(L.ptr(sims), N, m, L.ptr(meas), L.ptr(cov), _cov_kind(cov),
                                        C.byref(out), info, L.stream_ptr()))
    return (out.value, (info[0], info[1])) if return_info else out.value


def gather_columns(x, idx):
    """mask_obs / mask_var on device: x[:, idx]."""
    torch = _torch()
    L.init()
    x = _f64(x, 'x')
    squeeze = x.dim() == 1
    x2 = x.reshape(1, -1) if squeeze else x
    out = torch.empty((x2.shape[0], idx.numel()), dtype=torch.float64, device=x.device)
    L.check(L.load().imb_gather_columns(L.ptr(x2), x2.shape[0], x2.shape[1], L.ptr(idx), idx.numel(),
                                        L.ptr(out), L.stream_ptr()))
    return out.reshape(-1) if squeeze else out


def gather_cov(cov, idx):
    """mask_cov on device: cov[idx][:, idx]."""
    torch = _torch()
    L.init()
    cov = _f64(cov, 'cov')
    out = torch.empty((idx.numel(), idx.numel()), dtype=torch.float64, device=cov.device)
    L.check(L.load().imb_gather_cov(L.ptr(cov), cov.shape[0], L.ptr(idx), idx.numel(), L.ptr(out),
                                    L.stream_ptr()))
    return out
