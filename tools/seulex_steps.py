"""Developer helper (GPU box): accepted / rejected step statistics of the extrapolation integrator
against RODAS4 on the first n random parameter sets of the benchmark list."""
import sys

import numpy as np

sys.path.insert(0, ".")
from metrotrpl_b200 import _capi          # noqa: E402
import bench                              # noqa: E402
from tests import parity_cases as pc      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = _capi.Context(0)
g, prob, _, _ = pc.staub_problem()
params = _capi.pack_params(bench.draw_states(4096, seed=20261018), bench.IDX, bench.UNITS)[:n]
aux = _capi.default_aux(n, 6, [1.0] * 6)
ctx.set_problem(prob)
out = {}
for name, flag in (("rodas4", _capi.OPT_NO_EXPLICIT), ("seulex", _capi.OPT_EXTRAPOLATION | _capi.OPT_CTA_PER_TRAJ)):
    ll, st, ns, cur = ctx.loglik_batch(params, aux, _capi.make_opts(RTOL=1e-7, flags=flag), want_curves=True)
    out[name] = (ns.reshape(-1, 2), cur)
    a, r = ns[..., 0].ravel(), ns[..., 1].ravel()
    print(f"{name}: accepted mean {a.mean():.0f} max {a.max()}  rejected mean {r.mean():.1f} max {r.max()}  total max {(a + r).max()}")
    worst = np.argsort(-(a + r))[:8]
    print("   worst trajectories (index: accepted+rejected):", [(int(i), int(a[i]), int(r[i])) for i in worst])
ra, sa = out["rodas4"][0], out["seulex"][0]
print("ratio of total steps (rodas4 / seulex): mean", (ra.sum(1) / sa.sum(1)).mean(), "min", (ra.sum(1) / sa.sum(1)).min())
with np.errstate(all="ignore"):
    T, C = out["rodas4"][1], out["seulex"][1]
    ok = T > 1e-12 * T.max(axis=1, keepdims=True)
    print("max curve difference seulex vs rodas4 (top 12 decades of each set's curves):", np.abs(np.where(ok, C / T - 1, 0)).max())
