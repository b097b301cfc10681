import sys, json, numpy as np
sys.path.insert(0, ".")
from metrotrpl_b200 import _capi
import bench
ctx = _capi.Context(0)
t = np.linspace(0, 200, 81)
out = {}
for nx in (128, 256, 512):
    x = (np.arange(nx) + 0.5) * (311.0 / nx)
    sim = {"lengths": [311.0] * 2, "nx": [nx] * 2, "meas_types": ["TRPL", "TRPL"], "num_meas": 2}
    ini = [2e16 * np.exp(-x / 100.0), 2e17 * np.exp(-x / 100.0)]
    prob = _capi.pack_problem(sim, ini, [t] * 2, [np.full(len(t), 20.0)] * 2, [np.full(len(t), 0.05)] * 2)
    n = 4096
    params = _capi.pack_params(bench.draw_states(n, seed=6), bench.IDX, bench.UNITS)
    aux = _capi.default_aux(n, 2, [1.0] * 2)
    ctx.set_problem(prob)
    opts = _capi.make_opts(RTOL=1e-7)
    ms = []
    for _ in range(4):
        ll, st, ns, _c = ctx.loglik_batch(params, aux, opts); ms.append(ctx.last_kernel_ms())
    k = float(np.mean(ms[1:]))
    out[nx] = {"kernel_ms": k, "sims_per_s": 2 * n / k * 1e3, "node_steps_per_s": float(ns.sum()) * nx / k * 1e3, "mean_steps": float(ns[..., 0].mean())}
print(json.dumps(out))
