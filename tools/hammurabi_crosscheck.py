#!/usr/bin/env python
"""
External cross-check of the LOS definition (docs/los_conventions.md) against a real hammurabiX binary.

  python tools/hammurabi_crosscheck.py [--nside 16] [--grid 32] [--freq 23] [--dr 0.1] [--outdir DIR]
                                       [--hamx PATH]

1. evaluates ONE realisation of the benchmark fields (exponential n_e disc, constant + Gaussian B,
   analytic CRE) on the GPU and integrates it with B200Simulator;
2. writes the B and n_e grids in hammurabiX's external-field format (raw C-order float64,
   imagine/simulators/hammurabi.py:229-273 -> `dump_hammurabi_fields`) and a parameter file with the same
   box, vertex counts, observer, radial limits and step, ONE shell at the observable's Nside, the analytic
   CRE model with the same parameters (tests/data/test_ham.xml is the template for every tag used);
3. runs `hamx <params.xml>` when a binary is found, reads its maps (raw float64) and prints, per
   observable, max |rel diff|, the correlation coefficient and the fitted scale -- a sign flip (rows 9 / 11
   of the conventions table) shows up as correlation -1, a normalisation mismatch (rows 14 / 16) as a
   scale != 1.
Without a binary it stops after step 2; the directory can be carried to a machine that has hammurabiX.
"""
import argparse
import os
import shutil
import subprocess
import sys
import xml.etree.ElementTree as ET

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BOX = [[-20.0, 20.0], [-20.0, 20.0], [-4.0, 4.0]]
OBSERVER = (-8.3, 0.0, 0.006)


def _val(parent, tag, value):
    ET.SubElement(parent, tag, value=str(value))


def write_params(path, nside, grid_n, freq, dr, r_max, cre):
    root = ET.Element('root')
    obs = ET.SubElement(root, 'observable')
    ET.SubElement(obs, 'dm', cue='1', filename='dm.bin', nside=str(nside))
    ET.SubElement(obs, 'faraday', cue='1', filename='fd.bin', nside=str(nside))
    ET.SubElement(obs, 'sync', cue='1', freq=str(freq), filename='sync.bin', nside=str(nside))
    ET.SubElement(root, 'mask', cue='0', filename='mask.bin')
    fio = ET.SubElement(root, 'fieldio')
    ET.SubElement(fio, 'breg', read='1', write='0', filename='breg.bin')
    ET.SubElement(fio, 'brnd', read='0', write='0', filename='brnd.bin')
    ET.SubElement(fio, 'tereg', read='1', write='0', filename='tereg.bin')
    ET.SubElement(fio, 'ternd', read='0', write='0', filename='ternd.bin')
    ET.SubElement(fio, 'cre', read='0', write='0', filename='cre.bin')
    grid = ET.SubElement(root, 'grid')
    o = ET.SubElement(grid, 'observer')
    for tag, v in zip('xyz', OBSERVER):
        _val(o, tag, v)
    for name in ('box_breg', 'box_tereg'):
        b = ET.SubElement(grid, name)
        for tag, n in zip(('nx', 'ny', 'nz'), grid_n):
            _val(b, tag, n)
        for ax, (lo, hi) in zip('xyz', BOX):
            _val(b, ax + '_min', lo)
            _val(b, ax + '_max', hi)
    shell = ET.SubElement(grid, 'shell')
    layer = ET.SubElement(shell, 'layer', type='auto')
    auto = ET.SubElement(layer, 'auto')
    _val(auto, 'shell_num', 1)
    _val(auto, 'nside_sim', nside)
    _val(shell, 'oc_r_min', 0.0)
    _val(shell, 'oc_r_max', r_max)
    _val(shell, 'gc_r_min', 0.0)
    _val(shell, 'gc_r_max', 40.0)
    _val(shell, 'gc_z_min', -10.0)
    _val(shell, 'gc_z_max', 10.0)
    _val(shell, 'oc_r_res', dr)
    mf = ET.SubElement(root, 'magneticfield')
    ET.SubElement(mf, 'regular', cue='0', type='unif')        # B comes from breg.bin
    ET.SubElement(mf, 'random', cue='0', type='local', seed='0')
    te = ET.SubElement(root, 'thermalelectron')
    ET.SubElement(te, 'regular', cue='0', type='unif')
    ET.SubElement(te, 'random', cue='0', type='global', seed='0')
    c = ET.SubElement(root, 'cre', cue='1', type='analytic')
    an = ET.SubElement(c, 'analytic')
    for tag in ('alpha', 'beta', 'theta', 'r0', 'z0', 'E0', 'j0'):
        _val(an, tag, cre[tag])
    ET.ElementTree(root).write(path, xml_declaration=True, encoding='utf-8')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nside', type=int, default=16)
    ap.add_argument('--grid', type=int, default=32)
    ap.add_argument('--freq', type=float, default=23.0)
    ap.add_argument('--dr', type=float, default=0.1)
    ap.add_argument('--r-max', type=float, default=30.0)
    ap.add_argument('--outdir', default='hammurabi_crosscheck')
    ap.add_argument('--hamx', default=shutil.which('hamx'))
    args = ap.parse_args()
    import imagine_b200.api as api
    from imagine_b200.api.simulators import dump_hammurabi_fields
    u = api.units
    os.makedirs(args.outdir, exist_ok=True)
    n = [args.grid] * 3
    grid = api.UniformGrid(box=[[lo * u.kpc, hi * u.kpc] for lo, hi in BOX], resolution=n)
    cre = dict(alpha=3.0, beta=0.0, theta=0.0, r0=5.0, z0=1.0, E0=20.6, j0=0.0217)
    fields = [api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.01 * u.cm ** -3,
                                                                'scale_radius': 3 * u.kpc,
                                                                'scale_height': 0.1 * u.kpc}, ensemble_size=1),
              api.ConstantMagneticField(grid, parameters={'Bx': 0 * u.microgauss, 'By': 1 * u.microgauss,
                                                          'Bz': 0 * u.microgauss}, ensemble_size=1),
              api.NaiveGaussianMagneticField(grid, parameters={'a0': 0 * u.microgauss, 'b0': 2 * u.microgauss},
                                             ensemble_size=1, ensemble_seeds=[233]),
              api.AnalyticCosmicRayElectrons(grid, parameters={'j0': cre['j0'], 'r0': cre['r0'] * u.kpc,
                                                               'z0': cre['z0'] * u.kpc, 'r_sun': 8.3 * u.kpc},
                                             ensemble_size=1)]
    m = 12 * args.nside ** 2
    keys = {'dm': ('dm', None, args.nside, None), 'fd': ('fd', None, args.nside, None),
            'I': ('sync', args.freq, args.nside, 'I'), 'Q': ('sync', args.freq, args.nside, 'Q'),
            'U': ('sync', args.freq, args.nside, 'U')}
    M = api.Measurements()
    for k in keys.values():
        M.append(name=k, data=np.zeros((1, m)), otype='HEALPix')
    sim = api.B200Simulator(M, observer=OBSERVER, radial_step=args.dr, radial_max=args.r_max,
                            cre_alpha=cre['alpha'], cre_E0=cre['E0'],
                            temperature='cmb')          # hammurabiX reports K_CMB (docs/los_conventions.md)
    maps = {q: sim(fields)[k].data[0] for q, k in keys.items()}
    np.savez(os.path.join(args.outdir, 'b200_maps.npz'), **maps)
    sim.prepare_fields(fields, 0)                       # host copies of the summed fields, realisation 0
    dumped = dump_hammurabi_fields(sim.fields, args.outdir)
    params = os.path.join(args.outdir, 'params.xml')
    write_params(params, args.nside, n, args.freq, args.dr, args.r_max, cre)
    print('wrote', params, 'and', sorted(p for p, _ in dumped.values()))
    if not args.hamx:
        print('no hamx binary on PATH: inputs written; run `hamx params.xml` in %s where hammurabiX exists '
              'and re-run this script with --hamx to compare' % args.outdir)
        return
    subprocess.run([args.hamx, 'params.xml'], cwd=args.outdir, check=True)
    got = {'dm': np.fromfile(os.path.join(args.outdir, 'dm.bin')),
           'fd': np.fromfile(os.path.join(args.outdir, 'fd.bin'))}
    sync = np.fromfile(os.path.join(args.outdir, 'sync.bin'))
    if sync.size >= 3 * m:                              # hammurabiX writes I, Q, U one after the other
        got.update(I=sync[:m], Q=sync[m:2 * m], U=sync[2 * m:3 * m])
    for q, ours in maps.items():
        if q not in got or got[q].size != m:
            print('%-3s no hammurabiX map of the expected size' % q)
            continue
        theirs = got[q]
        scale = float(np.dot(ours, theirs) / max(np.dot(ours, ours), 1e-300))
        corr = float(np.corrcoef(ours, theirs)[0, 1])
        rel = float(np.abs(theirs - ours).max() / max(np.abs(ours).max(), 1e-300))
        print('%-3s corr %+.6f  scale(theirs/ours) %.6f  max|diff|/max %.3e' % (q, corr, scale, rel))


if __name__ == '__main__':
    main()
