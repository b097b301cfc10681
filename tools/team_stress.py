"""Developer helper (GPU box): race hunt for the two- and four-warp team kernels.  The configs[3] batch (traps + IRF,
nx = 256) and a padded 'std' grid (nx = 200) are launched repeatedly under random permutations of the
parameter sets and random explicit queue orders; every launch must reproduce the first one bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metrotrpl_b200 import _capi  # noqa: E402
import bench  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = _capi.Context(0)
rng = np.random.default_rng(11)
g = np.load(os.path.join(ROOT, "tests", "golden", "traps_irf.npz"))
names = [str(n) for n in g["names"]]
idx = {n: i for i, n in enumerate(names)}
nx = int(g["nx"])
sim = {"lengths": list(g["lengths"]), "nx": [nx] * 2, "meas_types": ["TRPL"] * 2, "num_meas": 2}
prob = _capi.pack_problem(sim, g["inis"], [g["t"]] * 2, list(g["vals"]), list(g["uncs"]), model="traps",
                          ini_mode="fluence", irf_convolution=[520, 520], irf_tables={520: (g["moments"], g["t_irf"])})
n_sets = 2048
base = g["states"][rng.integers(0, len(g["states"]), n_sets)]
jit = np.ones_like(base)
act = [idx[n] for n in names if n not in ("n0", "eps", "Tm", "m")]
jit[:, act] = 10 ** rng.uniform(-0.1, 0.1, size=(n_sets, len(act)))
cases = [("traps+irf nx=256", prob, _capi.pack_params(base * jit, idx, g["units"], model="traps"),
          _capi.default_aux(n_sets, 2, [1.0] * 2))]
t = np.linspace(0, 200, 81)
x = (np.arange(200) + 0.5) * (311.0 / 200)
sim2 = {"lengths": [311.0] * 3, "nx": [200] * 3, "meas_types": ["TRPL", "TRTS", "TRPL"], "num_meas": 3}
ini2 = [2e16 * np.exp(-x / 100.0), 2e17 * np.exp(-x / 100.0), 5e15 * np.ones(200)]
prob2 = _capi.pack_problem(sim2, ini2, [t] * 3, [np.full(len(t), 20.0)] * 3, [np.full(len(t), 0.05)] * 3)
cases.append(("std nx=200 (padded)", prob2, _capi.pack_params(bench.draw_states(n_sets, seed=5), bench.IDX, bench.UNITS),
              _capi.default_aux(n_sets, 3, [1.0] * 3)))
x4 = (np.arange(400) + 0.5) * (311.0 / 400)
sim4 = {"lengths": [311.0] * 2, "nx": [400] * 2, "meas_types": ["TRPL", "TRTS"], "num_meas": 2}
ini4 = [2e16 * np.exp(-x4 / 100.0), 2e17 * np.exp(-x4 / 100.0)]
prob4 = _capi.pack_problem(sim4, ini4, [t] * 2, [np.full(len(t), 20.0)] * 2, [np.full(len(t), 0.05)] * 2)
cases.append(("std nx=400 (padded, four-warp team)", prob4,
              _capi.pack_params(bench.draw_states(n_sets, seed=6), bench.IDX, bench.UNITS),
              _capi.default_aux(n_sets, 2, [1.0] * 2)))
opts = _capi.make_opts(RTOL=1e-7)
for label, pb, params, aux in cases:
    ctx.set_problem(pb)
    ll0, st0, ns0, cur0 = ctx.loglik_batch(params, aux, opts, want_curves=True)
    bad = 0
    for r in range(reps):
        perm = rng.permutation(n_sets)
        ctx.set_queue_order(rng.permutation(n_sets * pb.n_meas) if r % 2 else None)
        ll, st, ns, cur = ctx.loglik_batch(params[perm], aux[perm], opts, want_curves=True)
        same = (np.array_equal(ll, ll0[perm], equal_nan=True) and np.array_equal(ns, ns0[perm]) and
                np.array_equal(st, st0[perm]) and np.array_equal(cur, cur0[perm], equal_nan=True))
        bad += 0 if same else 1
    ctx.set_queue_order(None)
    print(f"{label}: {reps} permuted launches of {n_sets * pb.n_meas} trajectories, {bad} differ; "
          f"failed flags {int(np.sum((st0 & 7) != 0))}, mean steps {ns0[..., 0].mean():.1f}")
    assert bad == 0
