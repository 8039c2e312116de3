"""
GPU Simulator drop-ins.

`B200Simulator` takes the place of `imagine.simulators.Hammurabi`
(simulators/hammurabi.py:29-346: observables 'fd', 'dm', 'sync'; units K / rad m^-2 /
pc cm^-3; RING-ordered HEALPix maps of length 12 Nside^2) without the XML / temp-file /
subprocess round trip: fields are evaluated on the device grid and integrated along the
lines of sight by the CUDA kernels.  `B200TestSimulator` mirrors
`imagine.simulators.TestSimulator` (simulators/test_simulator.py:19-69).

Both override `Simulator.__call__` -- still the public entry `Pipeline._get_observables`
uses (pipelines/pipeline.py:850) -- to evaluate the WHOLE ensemble in one batched pass instead
of the reference's N x K Python loop (simulator.py:411-421), and return a `Simulations`
dictionary of device-resident `DeviceObservable`s.  `simulate(key, coords_dict,
realization_id, output_units)` stays valid for the base-class loop and for foreign callers:
it integrates the host fields found in `self.fields` for one realisation.

Field plugins that expose `device_descriptor()` (api/fields.py) are fused into one launch;
any other `Field` subclass still works -- its `get_data` arrays are uploaded and summed on
the device.
"""
import logging

import numpy as np

from . import units as u
from ._bases import Simulator, Simulations, Masks
from .device_observable import DeviceObservable, ShardedDeviceObservable
from .fields import descriptor_of, device_grid
from .. import engine
from ..constants import rj_to_cmb
from ..hotpath import PeerExchange, exchange_pixel_shards, ensure_library_comm, gather_rows, shard_bounds

FIELD_UNITS = {'magnetic_field': u.microgauss, 'thermal_electron_density': u.cm ** -3}
_NAMES = ('magnetic_field', 'thermal_electron_density', 'cosmic_ray_electron_density')
log = logging.getLogger(__name__)


def _torch():
    import torch
    return torch


def _host_values(q, unit):
    """Plain float64 array of a field returned by Field.get_data, in canonical units."""
    if isinstance(q, u.Quantity):
        return np.asarray(q.to_value(unit) if unit is not None else q.value, dtype=np.float64)
    return np.asarray(q, dtype=np.float64)


def _accepted_fields(sim, field_list):
    """Fields `sim` consumes after the grid checks of Simulator.prepare_fields (simulator.py:128-147)."""
    wanted = (*sim.required_field_types, *sim.optional_field_types)
    sim.grid, sim.grids = None, None
    used = []
    for f in field_list:
        if f.type not in wanted:
            continue
        if f.grid is not None and sim.allowed_grid_types is not None:
            assert f.grid.grid_type in sim.allowed_grid_types, 'Grid type not allowed'
        if sim.grid is None:
            sim.grid = f.grid
        assert sim.grid is f.grid, 'Multiple grids when a common grid is required'
        used.append(f)
    assert set(sim.required_field_types) <= {f.type for f in used}, 'Missing required field'
    return used


class _DeviceFieldEvaluator:
    """Evaluates a list of Field plugins for a block of realisations into device buffers."""

    def __init__(self, b_stride):
        self.b_stride = b_stride
        self.effective_b_stride = 4
        self._bufs = {}

    def _evaluate_by_reference(self, sim, fields, grid, lo, hi):
        """
        Field lists with DEPENDENCIES: the reference's own `Simulator.prepare_fields`
        (simulator.py:98-180: dependency sort, string and class dependencies, same-type sum) builds the
        host fields of every realisation; they are uploaded row by row.  Slow path, exact semantics.
        """
        torch = _torch()
        n_local = hi - lo
        dg = device_grid(grid)
        rows = {}
        for i in range(lo, hi):
            Simulator.prepare_fields(sim, fields, i)
            for ftype, data in sim.fields.items():
                if ftype not in _NAMES:
                    continue
                rows.setdefault(ftype, []).append(_host_values(data, FIELD_UNITS.get(ftype)))
        out = {}
        for ftype, host_rows in rows.items():
            host = np.stack(host_rows)
            if ftype == 'magnetic_field':
                dev = torch.zeros((n_local, dg.ncell, 4), dtype=torch.float64, device='cuda')
                dev[:, :, :3] = torch.from_numpy(host.reshape(n_local, dg.ncell, 3)).cuda()
            else:
                dev = torch.from_numpy(host.reshape(n_local, dg.ncell)).cuda()
            out[ftype] = dev
        self.effective_b_stride = 4
        return dg, out

    def evaluate(self, fields, grid, lo, hi, sim=None, all_fields=None):
        torch = _torch()
        n_local = hi - lo
        dg = device_grid(grid)
        if sim is not None and any(getattr(f, 'dependencies_list', None) for f in fields):
            return self._evaluate_by_reference(sim, all_fields if all_fields is not None else fields, grid, lo, hi)
        descs, seed_rows, foreign = [], [], []
        for f in fields:
            d = descriptor_of(f)
            if d is None:
                foreign.append(f)
                continue
            row = 0
            if f.stochastic_field:
                seed_rows.append(np.asarray(f.ensemble_seeds, dtype=np.uint64)[lo:hi])
                row = len(seed_rows) - 1
            descs.append((d[0], d[1], d[2], row))
        seeds = np.stack(seed_rows) if seed_rows else np.zeros((1, n_local), dtype=np.uint64)
        b_stride = engine.pick_b_stride(n_local) if self.b_stride == 'auto' else self.b_stride
        if b_stride in engine.INTERLEAVED_STRIDES and any(f.type == 'magnetic_field' for f in foreign):
            b_stride = 4              # host arrays are summed in the row-major padded layout
        out = {}
        if descs:
            # device buffers persist across evaluations (a Pipeline calls this ~10^4 times)
            shapes, b_stride = engine.field_shapes(dg, descs, n_local, b_stride)
            for name, shape in shapes.items():
                buf = self._bufs.get(name)
                if buf is None or tuple(buf.shape) != shape:
                    buf = torch.empty(shape, dtype=torch.float64, device='cuda')
                    self._bufs[name] = buf
                out[name] = buf
            out = engine.field_eval(dg, descs, seeds, n_real=n_local, b_stride=b_stride, out=out)
        self.effective_b_stride = b_stride if b_stride not in engine.INTERLEAVED_STRIDES else 4
        # foreign plugins: host get_data -> upload -> device sum (simulator.py:166-178 semantics)
        for f in foreign:
            unit = FIELD_UNITS.get(f.type, f.units)
            rows = n_local if f.stochastic_field else 1
            host = np.stack([_host_values(f.get_data(lo + i), unit) for i in range(rows)])
            if f.type == 'magnetic_field':
                dev = torch.zeros((rows, dg.ncell, self.effective_b_stride), dtype=torch.float64, device='cuda')
                dev[:, :, :3] = torch.from_numpy(host.reshape(rows, dg.ncell, 3)).cuda()
            else:
                dev = torch.from_numpy(host.reshape(rows, dg.ncell)).cuda()
            cur = out.get(f.type)
            out[f.type] = dev if cur is None else cur + dev       # broadcasts 1 <-> n_local rows
        return dg, out


class B200Simulator(Simulator):
    """
    Faraday depth, dispersion measure and synchrotron Stokes maps on the GPU.

    Parameters
    ----------
    measurements : Measurements
        keys ('fd', None, Nside, None), ('dm', None, Nside, None),
        ('sync', freq_GHz, Nside, 'I'|'Q'|'U'|'PI'|'PA')   (observable_dict.py:11-25)
    observer : 3 floats [kpc]           (hammurabi XML grid/observer, test_ham.xml:26-30)
    radial_step, radial_max : [kpc]     (oc_r_res, oc_r_max, test_ham.xml:122-131)
    cre_alpha, cre_beta, cre_theta, cre_E0 : CRE spectrum p = alpha - beta R - theta z at E0 [GeV]
                                        (fields/hamx/cre_analytic.py:19-27)
    rank, world, group : ensemble sharding over torch.distributed ranks (pipeline.py:737-744)
    masks : Masks or None -- applied INSIDE the simulator like `Hammurabi(measurements, masks=...)`
               (hammurabi.py:281-325: one identical mask covering every observable): masked pixels are not
               ray-marched at all and come back as zeros in the full-size maps
    temperature : 'cmb' (default) | 'rj' -- Stokes I, Q, U, PI as thermodynamic CMB temperature, which is what
               hammurabiX reports and the reference's key convention names ("K-cmb", observable_dict.py:15-19),
               or as Rayleigh-Jeans brightness temperature, which is what the kernels integrate: a factor
               (e^x - 1)^2 / (x^2 e^x), x = h nu / k T_cmb (1.01375 at 23 GHz) applied with the unit scale of the
               key.  With 'cmb' the recorded hammurabiX output of tutorials/tutorial_hammurabi.ipynb is
               reproduced to 3e-5 (tools/hammurabi_tutorial_pin.py, tests/test_oracle_los.py)
    exchange : 'auto' | 'peer' | 'alltoall' | 'allgather' -- how the maps reach the likelihood on several ranks
               (hotpath.py); 'alltoall' returns pixel-sharded observables
    """
    SIMULATED_QUANTITIES = ['fd', 'dm', 'sync']
    REQUIRED_FIELD_TYPES = ['magnetic_field']
    OPTIONAL_FIELD_TYPES = ['thermal_electron_density', 'cosmic_ray_electron_density']
    ALLOWED_GRID_TYPES = ['cartesian']

    def __init__(self, measurements, observer=(-8.3, 0.0, 0.006), radial_step=0.1, radial_max=30.0,
                 cre_alpha=3.0, cre_beta=0.0, cre_theta=0.0, cre_E0=20.6, b_stride='auto',
                 rank=0, world=1, group=None, exchange='auto', masks=None, temperature='cmb'):
        super().__init__(measurements)
        if temperature not in ('rj', 'cmb'):
            raise ValueError("temperature must be 'rj' (Rayleigh-Jeans brightness temperature) or 'cmb'")
        self.temperature = temperature
        self._kept = {}
        self.masks = masks
        self.exchange = exchange
        self.exchange_note = None          # set when 'auto' falls back from the fused NVLink exchange
        self._send = None
        self._peer, self._peer_shape = None, None
        self.los_kwargs = dict(observer=tuple(float(u.value_in(v, u.kpc)) for v in observer),
                               dr=float(u.value_in(radial_step, u.kpc)),
                               r_max=float(u.value_in(radial_max, u.kpc)),
                               alpha=float(cre_alpha), beta=float(cre_beta), theta=float(cre_theta),
                               e0_gev=float(cre_E0))
        self.b_stride = b_stride if b_stride == 'auto' else int(b_stride)
        self._evaluator = None
        self.rank, self.world, self.group = int(rank), int(world), group
        self._dirs = {}
        self._row_cache = (None, {})
        for key in self.observables:
            name, freq, nside, tag = key
            if nside == 'tab':
                raise NotImplementedError('Tabular datasets not yet supported!')   # hammurabi.py:340-341
            if name == 'sync':
                if tag not in ('I', 'Q', 'U', 'PI', 'PA'):
                    raise ValueError('unrecognised synchrotron tag %r' % (tag,))
                if freq is None:
                    raise ValueError('synchrotron observables need a frequency')

    # -- masks applied while simulating (hammurabi.py:281-325) -------------------------------------
    @property
    def masks(self):
        return self._masks

    @masks.setter
    def masks(self, masks):
        if masks is not None:
            assert isinstance(masks, Masks)
            mask_keys = [k for k in masks.keys() if k[0] in self.simulated_quantities]
            for obs in self.observables:
                assert obs in mask_keys, 'All observables must be covered by the masks'
            mask_data = masks[mask_keys[0]].data
            for k in mask_keys:
                assert np.array_equal(mask_data, masks[k].data), 'All the masks must be identical'
        self._masks = masks
        self._kept = {}

    def _kept_pixels(self, nside):
        """(indices of the unmasked pixels [int64, device], their directions) for one Nside, or None."""
        if self._masks is None:
            return None
        if nside not in self._kept:
            torch = _torch()
            key = next(k for k in self._masks.keys() if k[0] in self.simulated_quantities)
            keep = np.flatnonzero(np.asarray(self._masks[key].data[0]).astype(bool))
            idx = torch.from_numpy(keep.astype(np.int64)).cuda()
            self._kept[nside] = (idx, self._dirs_for(nside).index_select(0, idx).contiguous())
        return self._kept[nside]

    # -- unit handling (hammurabi.py:214-227) ---------------------------------------------------
    @staticmethod
    def _units(key):
        if key[0] == 'sync':
            if key[3] in ('I', 'Q', 'U', 'PI'):
                return u.K
            if key[3] == 'PA':
                return u.rad
            raise ValueError
        if key[0] == 'fd':
            return u.rad / u.m / u.m
        if key[0] == 'dm':
            return u.pc / u.cm ** 3
        raise ValueError

    def _scale(self, key):
        out_unit = self.output_units.get(key)
        s = 1.0 if out_unit is None else u.unit_scale(self._units(key), out_unit)
        if getattr(self, 'temperature', 'cmb') == 'cmb' and key[0] == 'sync' and key[3] in ('I', 'Q', 'U', 'PI'):
            s *= rj_to_cmb(float(key[1]))            # K_RJ -> K_CMB, what hammurabiX reports
        return s

    def _dirs_for(self, nside):
        if nside not in self._dirs:
            self._dirs[nside] = engine.healpix_ring_dirs(int(nside))
        return self._dirs[nside]

    # -- the integrator for a block of realisations --------------------------------------------------
    def _integrate(self, dg, fields, n_real, sharded=None):
        """
        {key: tensor [n_real, m]} for every registered observable.  With `sharded = (views, layout)`
        the LOS kernel writes each key straight into its slot of the destination-major send buffer
        (views[key] = base pointer, layout = (cols_per_dest, dest_stride)); returns {}.
        """
        torch = _torch()
        B = fields['magnetic_field']
        ne = fields.get('thermal_electron_density')
        cre = fields.get('cosmic_ray_electron_density')
        # group by (nside, frequency): one launch each; fd/dm ride with the first group of their nside
        groups = {}
        for key in self.observables:
            name, freq, nside, tag = key
            groups.setdefault(int(nside), {}).setdefault(None if name != 'sync' else float(freq), []).append(key)
        result = {}
        for nside, by_freq in groups.items():
            dirs = self._dirs_for(nside)
            kept = self._kept_pixels(nside)
            m_full = dirs.shape[0]
            if kept is not None:
                dirs = kept[1]
            freqs = [f for f in by_freq if f is not None]
            plain = by_freq.get(None, [])
            launches = [(f, by_freq[f] + (plain if i == 0 else [])) for i, f in enumerate(freqs)]
            if not freqs:
                launches = [(None, plain)]
            for freq, keys in launches:
                want = set()
                for k in keys:
                    if k[0] in ('fd', 'dm'):
                        want.add(k[0])
                    elif k[3] in ('I', 'Q', 'U'):
                        want.add(k[3])
                    else:
                        want.update(('Q', 'U'))
                if want & {'fd', 'dm'} and ne is None:
                    raise AssertionError("'fd'/'dm' need a thermal_electron_density field")
                if want & {'I', 'Q', 'U'} and cre is None:
                    raise AssertionError("'sync' needs a cosmic_ray_electron_density field")
                kw = dict(self.los_kwargs)
                if freq is not None:
                    kw['nu_ghz'] = freq
                settings = engine.LosSettings(**kw)
                if sharded is not None:
                    views, layout = sharded
                    out = {(k[0] if k[0] in ('fd', 'dm') else k[3]): views[k] for k in keys}
                    engine.los_integrate(dg, settings, dirs, B, ne,
                                         cre if want & {'I', 'Q', 'U'} else None, n_real=n_real,
                                         want=tuple(want), out=out, out_layout=layout)
                    continue
                maps = engine.los_integrate(dg, settings, dirs, B, ne,
                                            cre if want & {'I', 'Q', 'U'} else None,
                                            n_real=n_real, want=tuple(want))
                derived = None
                for k in keys:
                    if k[0] in ('fd', 'dm'):
                        t = maps[k[0]]
                    elif k[3] in ('I', 'Q', 'U'):
                        t = maps[k[3]]
                    else:
                        if derived is None:
                            derived = engine.pol_derive(maps['Q'], maps['U'])
                        t = derived[0] if k[3] == 'PI' else derived[1]
                    s = self._scale(k)
                    t = t if s == 1.0 else t * s
                    if kept is not None:                   # masked pixels were not marched: zeros
                        full = torch.zeros((t.shape[0], m_full), dtype=t.dtype, device=t.device)
                        full.index_copy_(1, kept[0], t)
                        t = full
                    result[k] = t
        return result

    # -- batched public entry -----------------------------------------------------------------------
    def __call__(self, field_list):
        self.register_ensemble_size(field_list)
        fields = _accepted_fields(self, field_list)
        N = self._ensemble_size
        lo, hi = shard_bounds(N, self.rank, self.world)
        if self._evaluator is None:
            self._evaluator = _DeviceFieldEvaluator(self.b_stride)
        dg, dev_fields = self._evaluator.evaluate(fields, self.grid, lo, hi, sim=self, all_fields=field_list)
        if self._can_shard_pixels(dev_fields):
            return self._call_pixel_sharded(dg, dev_fields, N, hi - lo)
        maps = self._integrate(dg, dev_fields, hi - lo)
        sims = Simulations()
        for key in self.observables:
            block = maps[key]
            if self.world > 1:
                block = gather_rows(block, N, self.group)
            unit = self.output_units.get(key) or self._units(key)
            sims.append(name=key, data=DeviceObservable(block, unit, self.output_coords[key],
                                                        self.output_type[key]),
                        coords=self.output_coords[key], otype=self.output_type[key])
        return sims

    # -- pixel-sharded exchange (multi-rank fast path) -------------------------------------------------
    def _can_shard_pixels(self, dev_fields):
        if self.world <= 1 or self.exchange not in ('auto', 'alltoall', 'peer') or self._masks is not None:
            return False
        nsides = {int(k[2]) for k in self.observables}
        direct = all(k[0] in ('fd', 'dm') or k[3] in ('I', 'Q', 'U') for k in self.observables)
        per_real = any(t.dim() == 4 or t.shape[0] > 1 for t in dev_fields.values())
        ok = len(nsides) == 1 and direct and per_real and (12 * nsides.pop() ** 2) % self.world == 0
        if not ok and self.exchange in ('alltoall', 'peer'):
            raise ValueError('exchange="alltoall" needs one Nside divisible over the ranks, only '
                             'fd/dm/I/Q/U observables and at least one stochastic field')
        return ok

    def _call_pixel_sharded(self, dg, dev_fields, N, n_local):
        """
        The LOS kernel writes destination-major [P, K, n_local, m/P]; one all-to-all; the result is a
        `Simulations` of `ShardedDeviceObservable`s that `B200EnsembleLikelihood` consumes in place.
        """
        torch = _torch()
        K = len(self.observables)
        m = 12 * int(self.observables[0][2]) ** 2
        mp = m // self.world
        shape = (self.world, K, n_local, mp)
        if self._peer is None and self.exchange == 'peer' and self._peer_shape != shape:
            self._peer_shape = shape
            try:                                   # fused NVLink exchange (hotpath.PeerExchange)
                self._peer = PeerExchange(shape, self.rank, self.group)
            except Exception as exc:
                if self.exchange == 'peer':
                    raise
                self.exchange_note = 'all-to-all: symmetric memory unavailable (%s: %s)' % (type(exc).__name__, exc)
                if self.rank == 0:
                    log.warning('B200Simulator exchange fallback -- %s', self.exchange_note)
        still_valid = None
        if self._peer is not None and tuple(self._peer.recv.shape) == shape:
            recv = self._peer.recv
            self._peer_generation = generation = getattr(self, '_peer_generation', 0) + 1
            still_valid = lambda: self._peer_generation == generation
            views = {key: recv[self.rank, k] for k, key in enumerate(self.observables)}
            self._integrate(dg, dev_fields, n_local, sharded=(views, (mp, 0, self._peer.peer_delta)))
            self._peer.barrier()
        else:
            if self._send is None or tuple(self._send.shape) != shape:
                self._send = torch.empty(shape, dtype=torch.float64, device='cuda')
            send = self._send
            views = {key: send[0, k] for k, key in enumerate(self.observables)}
            self._integrate(dg, dev_fields, n_local, sharded=(views, (mp, K * n_local * mp)))
            ensure_library_comm(self.rank, self.world, self.group)     # once; then the library's all-to-all
            recv = exchange_pixel_shards(send, self.group)
        sims = Simulations()
        for k, key in enumerate(self.observables):
            s = self._scale(key)
            if s != 1.0:
                recv[:, k].mul_(s)
            unit = self.output_units.get(key) or self._units(key)
            sims.append(name=key,
                        data=ShardedDeviceObservable(recv, k, N, m, self.rank, self.world, self.group, unit,
                                                     self.output_coords[key], self.output_type[key],
                                                     still_valid=still_valid),
                        coords=self.output_coords[key], otype=self.output_type[key])
        return sims

    # -- per-realisation contract (simulator.py:370-393) -----------------------------------------------
    def simulate(self, key, coords_dict, realization_id, output_units):
        torch = _torch()
        cached_id, rows = self._row_cache
        if cached_id != realization_id or key not in rows:
            assert self.fields is not None and 'magnetic_field' in self.fields, 'prepare_fields first'
            dg = device_grid(self.grid)
            dev = {}
            for name in _NAMES:
                if name not in self.fields:
                    continue
                host = _host_values(self.fields[name], FIELD_UNITS.get(name))
                if name == 'magnetic_field':
                    t = torch.zeros((1, dg.ncell, 4), dtype=torch.float64, device='cuda')
                    t[:, :, :3] = torch.from_numpy(host.reshape(1, dg.ncell, 3)).cuda()
                else:
                    t = torch.from_numpy(host.reshape(1, dg.ncell)).cuda()
                dev[name] = t
            rows = {k: v[0].cpu().numpy() for k, v in self._integrate(dg, dev, 1).items()}
            self._row_cache = (realization_id, rows)
        unit = self.output_units.get(key) or self._units(key)
        return rows[key] * unit

    def __getstate__(self):
        state = dict(self.__dict__)
        state['_dirs'] = {}
        state['_kept'] = {}
        state['_evaluator'] = None
        state['_send'] = None
        state['_peer'], state['_peer_shape'] = None, None
        state['_row_cache'] = (None, {})
        state['group'] = None
        return state


class B200TestSimulator(Simulator):
    """t(x,y,z) = B_axis * n_e at the nearest grid vertex of each tabulated point (no integration)."""
    SIMULATED_QUANTITIES = ['test']
    REQUIRED_FIELD_TYPES = ['magnetic_field', 'thermal_electron_density']
    ALLOWED_GRID_TYPES = ['cartesian']
    __test__ = False

    def __init__(self, measurements, LoS_axis='y'):
        super().__init__(measurements)
        if LoS_axis == 'y':
            self.B_axis = 1
        elif LoS_axis == 'z':
            self.B_axis = 2
        else:
            raise ValueError

    def _points(self, coords_dict):
        torch = _torch()
        pts = np.array([np.atleast_1d(u.value_in(coords_dict[c], u.kpc)) for c in ('x', 'y', 'z')],
                       dtype=np.float64).T
        return torch.from_numpy(np.ascontiguousarray(pts)).cuda()

    def _scale(self, key):
        out_unit = self.output_units.get(key)
        native = u.microgauss * u.cm ** -3
        return 1.0 if out_unit is None else u.unit_scale(native, out_unit)

    def __call__(self, field_list):
        self.register_ensemble_size(field_list)
        fields = _accepted_fields(self, field_list)
        N = self._ensemble_size
        dg, dev = _DeviceFieldEvaluator(3).evaluate(fields, self.grid, 0, N, sim=self, all_fields=field_list)
        sims = Simulations()
        for key in self.observables:
            block = engine.test_simulate(dg, dev['magnetic_field'], dev['thermal_electron_density'], N,
                                         self._points(self.output_coords[key]), b_axis=self.B_axis,
                                         scale=self._scale(key))
            unit = self.output_units.get(key) or (u.microgauss * u.cm ** -3)
            sims.append(name=key, data=DeviceObservable(block, unit, self.output_coords[key],
                                                        self.output_type[key]),
                        coords=self.output_coords[key], otype=self.output_type[key])
        return sims

    def simulate(self, key, coords_dict, realization_id, output_units):
        torch = _torch()
        dg = device_grid(self.grid)
        B = _host_values(self.fields['magnetic_field'], u.microgauss).reshape(1, dg.ncell, 3)
        ne = _host_values(self.fields['thermal_electron_density'], u.cm ** -3).reshape(1, dg.ncell)
        native = u.microgauss * u.cm ** -3
        scale = 1.0 if output_units is None else u.unit_scale(native, output_units)
        row = engine.test_simulate(dg, torch.from_numpy(B).cuda(), torch.from_numpy(ne).cuda(), 1,
                                   self._points(coords_dict), b_axis=self.B_axis, scale=scale)
        return row[0].cpu().numpy() * (output_units if output_units is not None else native)


def dump_hammurabi_fields(fields, directory, prefix=''):
    """
    hammurabiX-compatible field I/O (simulators/hammurabi.py:229-273): every field type as raw
    C-order float64 (`ndarray.tofile`), one file per type named like the XML `fieldio` entries
    (breg / tereg / cre), plus the box / vertex-count keys the wrapper copies into the XML.  Lets a
    real hammurabiX install elsewhere integrate the very same grids for an external cross-check of
    the LOS definition.  `fields` = {field_type: Quantity or ndarray [nx,ny,nz(,3)]} for ONE
    realisation (e.g. `simulator.fields` after `prepare_fields`); returns {hamx_key: (path, meta)}.
    """
    import os
    names = {'magnetic_field': 'breg', 'thermal_electron_density': 'tereg',
             'cosmic_ray_electron_density': 'cre'}
    out = {}
    for ftype, data in fields.items():
        if ftype not in names:
            continue
        arr = np.ascontiguousarray(_host_values(data, FIELD_UNITS.get(ftype)), dtype=np.float64)
        path = os.path.join(directory, '%s%s.bin' % (prefix, names[ftype]))
        arr.tofile(path)
        out[names[ftype]] = (path, {'nx': arr.shape[0], 'ny': arr.shape[1], 'nz': arr.shape[2]})
    return out
