"""not-gpu: the C-ABI library builds/loads and exports every symbol include/imagine_b200.h declares."""
import os
import re

import pytest

from imagine_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _build.build()           # no-op when up to date; nvcc cross-compiles without a GPU
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, 'include', 'imagine_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(imb_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_version_and_strerror(lib):
    assert lib.imb_version() == 100
    assert lib.imb_strerror(0) == b'ok'
    assert b'bounds' in lib.imb_strerror(_lib.IMB_EBOUNDS)


def test_field_rows_host_only(lib):
    descs = [(_lib.FIELD_CONST_B, [0, 1, 0], 0), (_lib.FIELD_GAUSS_B, [0, 2], 1),
             (_lib.FIELD_EXP_NE, [.01, 3, .1], 2)]
    arr, n = _lib.make_field_descs(descs)
    assert [lib.imb_field_rows(arr, n, t, 16) for t in range(3)] == [16, 1, 0]


def test_compute_entry_points_refuse_without_init(lib):
    # no GPU here: every compute call must fail loudly, never fall back
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    assert lib.imb_healpix_ring_dirs(4, C.c_void_p(16), None) == _lib.IMB_ENOTINIT
    with pytest.raises(_lib.ImbError):
        _lib.init()


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(_lib.Grid) == 64
    assert C.sizeof(_lib.FieldDesc) == 208
    assert C.sizeof(_lib.LosCfg) == 232


def test_field_layout_selection_host_only(lib):
    """engine.pick_b_stride / b_shape / b_stride_of / field_shapes: the layout follows the shard size."""
    import torch
    from imagine_b200 import engine

    class G:                      # field_shapes only reads ncell
        ncell = 10
    assert [engine.pick_b_stride(n) for n in (256, 64, 32, 48, 16, 8, 24, 12, 4, 1)] == \
        [96, 96, 96, 48, 48, 24, 24, 24, 24, 24]
    for bs, rows in ((96, 64), (96, 33), (48, 16), (24, 8), (24, 5), (4, 3), (3, 3)):
        shape = engine.b_shape(rows, G.ncell, bs)
        assert engine.b_stride_of(torch.empty(shape, dtype=torch.float64)) == bs
        rl = {96: 32, 48: 16, 24: 8}.get(bs)
        if rl:                    # whole groups: lanes beyond `rows` exist and are zero-filled by the kernel
            assert shape[0] * rl >= rows and (shape[0] - 1) * rl < rows
    stochastic = [(_lib.FIELD_CONST_B, [0, 1, 0], 0), (_lib.FIELD_GAUSS_B, [0, 2], 1), (_lib.FIELD_EXP_NE, [.01, 3, .1], 2)]
    shapes, bs = engine.field_shapes(G, stochastic, 32, 'auto')
    assert bs == 96 and shapes['magnetic_field'] == (1, 10, 3, 32) and shapes['thermal_electron_density'] == (1, 10)
    # interleaving needs one B row per realisation: a deterministic B falls back to the padded rows
    shapes, bs = engine.field_shapes(G, stochastic[:1] + stochastic[2:], 32, 'auto')
    assert bs == 4 and shapes['magnetic_field'] == (1, 10, 4)
