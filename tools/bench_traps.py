"""Developer helper (GPU box): throughput of the BASELINE configs[3] variant - traps model, IRF
convolution, nx = 256 - on a batch of parameter sets spread around the golden fixture's states."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metrotrpl_b200 import _capi  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "traps_irf.npz"))
names = [str(n) for n in g["names"]]
idx = {n: i for i, n in enumerate(names)}
t = g["t"]
nx = int(g["nx"])
tables = {520: (g["moments"], g["t_irf"])}
sim = {"lengths": list(g["lengths"]), "nx": [nx] * 2, "meas_types": ["TRPL"] * 2, "num_meas": 2}
prob = _capi.pack_problem(sim, g["inis"], [t] * 2, list(g["vals"]), list(g["uncs"]), model="traps",
                          ini_mode="fluence", irf_convolution=[520, 520], irf_tables=tables)
n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(1)
base = g["states"][rng.integers(0, len(g["states"]), n_sets)]
jit = np.ones_like(base)
act = [idx[n] for n in names if n not in ("n0", "eps", "Tm", "m")]
jit[:, act] = 10 ** rng.uniform(-0.1, 0.1, size=(n_sets, len(act)))
params = _capi.pack_params(base * jit, idx, g["units"], model="traps")
aux = _capi.default_aux(n_sets, 2, [1.0] * 2)
ctx = _capi.Context(0)
ctx.set_problem(prob)
opts = _capi.make_opts(RTOL=1e-7)
ctx.loglik_batch(params, aux, opts)
ms = []
for _ in range(3):
    ll, st, ns, _ = ctx.loglik_batch(params, aux, opts)
    ms.append(ctx.last_kernel_ms())
k = float(np.mean(ms))
print(json.dumps({"workload": "configs[3]: traps model + IRF convolution, nx=256", "sets": n_sets, "curves_per_set": 2,
                  "times_per_curve": int(len(t)), "kernel_ms": k, "sims_per_s": 2 * n_sets / (k * 1e-3),
                  "mean_steps": float(ns[..., 0].mean()), "frac_failed": float(np.mean((st & 7) != 0)),
                  "finite_logll": float(np.mean(np.isfinite(ll[..., 0])))}))
