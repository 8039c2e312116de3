"""
ctypes binding of libimagine_b200.so (the C ABI declared in include/imagine_b200.h).

There is NO CPU fallback: if the shared library is missing or the device is not a B200-class
GPU, importing callers get a loud error.  PyTorch is used only as the device allocator and
stream provider; every signature below is plain pointers and sizes.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# IMB_LIB_PATH: A/B runs of tuning builds (tools/build_variant.sh); the product is the in-tree default
LIB_PATH = os.environ.get('IMB_LIB_PATH') or os.path.join(HERE, 'libimagine_b200.so')

IMB_OK, IMB_EINVAL, IMB_ECUDA, IMB_ENOMEM, IMB_ENOTINIT, IMB_EBOUNDS, IMB_ESINGULAR, IMB_EUNSUPPORTED, IMB_ESTATE = range(9)

# field kinds (include/imagine_b200.h)
(FIELD_CONST_B, FIELD_GAUSS_B, FIELD_CONST_NE, FIELD_EXP_NE, FIELD_COS_NE, FIELD_RANDOM_NE,
 FIELD_CRE_ANALYTIC, FIELD_CRE_CONST, FIELD_LSA_B, FIELD_JF12_B) = range(10)
COV_NONE, COV_VARIANCE, COV_DENSE = range(3)
EST_EMPIRICAL, EST_OAS, EST_DIAGONAL = range(3)
ROUTE_AUTO, ROUTE_GRAM, ROUTE_DENSE = range(3)
INFO_ROUTE_GRAM, INFO_ROUTE_DENSE, INFO_CHOLESKY, INFO_LU, INFO_NEGATIVE_DET = 1, 2, 4, 8, 16
MAX_FIELDS = 16


class ImbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class Grid(C.Structure):
    _fields_ = [('lo', C.c_double * 3), ('hi', C.c_double * 3), ('n', C.c_int32 * 3),
                ('reserved', C.c_int32)]


class FieldDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('stream', C.c_int32), ('seed_row', C.c_int32),
                ('reserved', C.c_int32), ('p', C.c_double * 24)]


class LosCfg(C.Structure):
    _fields_ = [('observer', C.c_double * 3), ('dr', C.c_double), ('n_steps', C.c_int32),
                ('out_cols_per_dest', C.c_int32), ('nu_ghz', C.c_double), ('alpha', C.c_double),
                ('beta', C.c_double), ('theta', C.c_double), ('e0_gev', C.c_double),
                ('sync_k', C.c_double), ('out_dest_stride', C.c_int64), ('n_peers', C.c_int32),
                ('reserved', C.c_int32), ('peer_delta', C.c_int64 * 16)]


_P = C.c_void_p
_SIGNATURES = {
    'imb_init': (C.c_int, [C.c_int]),
    'imb_finalize': (C.c_int, []),
    'imb_version': (C.c_int, []),
    'imb_strerror': (C.c_char_p, [C.c_int]),
    'imb_last_cuda_error': (C.c_char_p, []),
    'imb_launch_count': (C.c_int64, []),
    'imb_field_eval': (C.c_int, [C.POINTER(Grid), C.POINTER(FieldDesc), C.c_int, _P, C.c_int, C.c_int,
                                 _P, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int, _P]),
    'imb_field_eval_lanes': (C.c_int, [C.POINTER(Grid), C.POINTER(FieldDesc), C.c_int, _P, C.c_int, C.c_int,
                                       _P, C.c_int, C.c_int, C.c_int, _P]),
    'imb_field_rows': (C.c_int, [C.POINTER(FieldDesc), C.c_int, C.c_int, C.c_int]),
    'imb_test_simulate': (C.c_int, [C.POINTER(Grid), _P, C.c_int, C.c_int, _P, C.c_int, C.c_int,
                                    C.c_int, _P, C.c_int, C.c_double, _P, _P]),
    'imb_healpix_ring_dirs': (C.c_int, [C.c_int, _P, _P]),
    'imb_los_integrate': (C.c_int, [C.POINTER(Grid), C.POINTER(LosCfg), _P, C.c_int, _P, C.c_int,
                                    C.c_int, _P, C.c_int, _P, C.c_int, C.c_int,
                                    _P, _P, _P, _P, _P, _P]),
    'imb_pol_derive': (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    'imb_ensemble_mean': (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    'imb_ensemble_var': (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    'imb_mcov': (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_double), _P]),
    'imb_loglike_ensemble': (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int32), _P]),
    'imb_loglike_ensemble_batch': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_int,
                                             C.POINTER(C.c_double), C.POINTER(C.c_int32), _P]),
    'imb_gram_phase1': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P, _P]),
    'imb_gram_phase2': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P,
                                  C.c_int64, C.c_int, _P, _P, _P]),
    'imb_gram_phase3': (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, C.POINTER(C.c_double),
                                  C.POINTER(C.c_int32), _P]),
    'imb_loglike_ensemble_batch_async': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    'imb_gram_phase3_async': (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    'imb_loglike_ensemble_diagonal': (C.c_int, [_P, C.c_int, C.c_int, _P, _P,
                                                C.POINTER(C.c_double), _P]),
    'imb_loglike_simple': (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int,
                                     C.POINTER(C.c_double), C.POINTER(C.c_int32), _P]),
    'imb_gather_columns': (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    'imb_gather_cov': (C.c_int, [_P, C.c_int, _P, C.c_int, _P, _P]),
    'imb_comm_unique_id': (C.c_int, [_P]),
    'imb_comm_init': (C.c_int, [C.c_int, C.c_int, _P]),
    'imb_comm_finalize': (C.c_int, []),
    'imb_comm_size': (C.c_int, []),
    'imb_comm_rank': (C.c_int, []),
    'imb_allreduce_sum': (C.c_int, [_P, C.c_int64, _P]),
    'imb_alltoall': (C.c_int, [_P, _P, C.c_int64, _P]),
    'imb_loglike_ensemble_sharded': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                               _P, _P, C.c_int64, C.c_int, _P, _P, C.POINTER(C.c_double),
                                               C.POINTER(C.c_int32), _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_device = None


def load():
    """dlopen the library and declare every signature; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'libimagine_b200.so is not built (%s). Run `python -c "import __graft_entry__ as g; '
            'g.build()"` or `python -m imagine_b200._build`. There is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code):
    if code == IMB_OK:
        return
    lib = load()
    msg = lib.imb_strerror(code).decode()
    if code in (IMB_ECUDA, IMB_ENOMEM, IMB_EUNSUPPORTED):
        msg += ': ' + lib.imb_last_cuda_error().decode()
    if code == IMB_EBOUNDS:
        raise ValueError('One of the requested xi is out of bounds (' + msg + ')')
    if code == IMB_ESINGULAR:
        raise np.linalg.LinAlgError('Singular matrix')
    if code == IMB_EINVAL:
        raise ValueError('libimagine_b200: ' + msg)
    raise ImbError(code, 'libimagine_b200: ' + msg)


def init(device=None):
    """Initialise the library on a CUDA device (default: torch's current device)."""
    global _device
    if device is None and _device is not None:
        return _device                      # hot path: already initialised
    import torch
    if not torch.cuda.is_available():
        raise ImbError(IMB_ECUDA, 'imagine_b200 needs a CUDA device (B200, sm_100a); '
                                  'no CPU fallback exists')
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    if _device == device:
        return device
    check(load().imb_init(device))
    _device = device
    return device


def current_device():
    return _device


def launch_count():
    return int(load().imb_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(_device if _device is not None else 0))


def make_grid(box, resolution):
    g = Grid()
    box = np.asarray(box, dtype=np.float64).reshape(3, 2)
    res = np.asarray(resolution, dtype=np.int64).reshape(3)
    for a in range(3):
        g.lo[a], g.hi[a], g.n[a] = float(box[a, 0]), float(box[a, 1]), int(res[a])
    return g


def make_field_descs(descs):
    """descs: iterable of (kind, params, stream[, seed_row]) -> ctypes array."""
    descs = list(descs)
    if len(descs) > MAX_FIELDS * 3:
        raise ValueError('too many fields')
    arr = (FieldDesc * max(1, len(descs)))()
    for i, d in enumerate(descs):
        kind, params, stream = d[0], d[1], d[2]
        arr[i].kind, arr[i].stream = int(kind), int(stream)
        arr[i].seed_row = int(d[3]) if len(d) > 3 else 0
        for k, v in enumerate(params):
            arr[i].p[k] = float(v)
    return arr, len(descs)
