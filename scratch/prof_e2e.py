import cProfile, pstats, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagine_b200.api as api
from imagine_b200 import _lib
_lib.init(0)
u = api.units
G, nside, N = 32, 16, 8          # tiny GPU work: what remains is host overhead
m = 12 * nside ** 2
grid = api.UniformGrid(box=[[-20, 20] * u.kpc, [-20, 20] * u.kpc, [-4, 4] * u.kpc], resolution=[G] * 3)
keys = [('fd', None, nside, None), ('dm', None, nside, None), ('sync', 23.0, nside, 'I'), ('sync', 23.0, nside, 'Q'), ('sync', 23.0, nside, 'U')]
units = [u.rad / u.m ** 2, u.pc / u.cm ** 3, u.K, u.K, u.K]
M = api.Measurements()
for k, un in zip(keys, units):
    M.append(name=k, data=np.random.rand(1, m) * un, cov_data=np.ones(m) * un ** 2, otype='HEALPix')
sim = api.B200Simulator(M); lik = api.B200EnsembleLikelihood(M)
seeds = list(range(1, N + 1))
def step(reset=True):
    fields = [api.ExponentialThermalElectrons(grid, parameters={'central_density': 0.01 * u.cm ** -3, 'scale_radius': 3.0 * u.kpc, 'scale_height': 0.1 * u.kpc}, ensemble_size=N),
              api.ConstantMagneticField(grid, parameters={'Bx': 0 * u.microgauss, 'By': 1 * u.microgauss, 'Bz': 0 * u.microgauss}, ensemble_size=N),
              api.NaiveGaussianMagneticField(grid, parameters={'a0': 0 * u.microgauss, 'b0': 2 * u.microgauss}, ensemble_size=N, ensemble_seeds=seeds),
              api.AnalyticCosmicRayElectrons(grid, parameters={'j0': 0.0217, 'r0': 5.0 * u.kpc, 'z0': 1.0 * u.kpc, 'r_sun': 8.3 * u.kpc}, ensemble_size=N)]
    if reset: lik._reset_device_state()
    return float(lik(sim(fields)))
for _ in range(5): step()
torch.cuda.synchronize()
for reset in (True, False):
    t0 = time.perf_counter()
    for _ in range(200): step(reset)
    print('reset=%s: %.3f ms/step' % (reset, (time.perf_counter() - t0) / 200 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
