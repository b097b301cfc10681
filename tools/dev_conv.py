"""Developer helper (GPU box): the self-convergence test's statistics with the worst cases listed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tests import parity_cases as pc
from metrotrpl_b200 import _capi
ctx = _capi.Context(0)
g, prob, params, aux = pc.staub_problem()
states = bench.draw_states(4096, seed=99)
names = [str(n) for n in g["names"]]
idx = {n: i for i, n in enumerate(names)}
P = _capi.pack_params(states, idx, g["units"])
A = np.repeat(aux[:1], 4096, axis=0)
ctx.set_problem(prob)
_, st7, ns7, c7 = ctx.loglik_batch(P, A, _capi.make_opts(RTOL=1e-7), want_curves=True)
_, st9, ns9, c9 = ctx.loglik_batch(P, A, _capi.make_opts(RTOL=1e-9), want_curves=True)
_, st8, ns8, c8 = ctx.loglik_batch(P, A, _capi.make_opts(RTOL=1e-8), want_curves=True)
c7 = c7.reshape(4096, 6, -1); c9 = c9.reshape(4096, 6, -1); c8 = c8.reshape(4096, 6, -1)
def stats(ca, cb):
    with np.errstate(all="ignore"):
        win6 = cb > 1e-6 * cb[:, :, :1]; win3 = cb > 1e-3 * cb[:, :, :1]
        err = np.abs(ca / cb - 1); efold = np.log(np.maximum(cb[:, :, :1] / cb, 1.0))
    return np.where(win6, err / (1.0 + efold), 0.0), np.where(win3, err, 0.0)
scaled, top3 = stats(c7, c9)
s89, t89 = stats(c8, c9)
print("7v9: max err per e-fold", scaled.max(), "top3", top3.max(), "| 8v9:", s89.max(), t89.max())
per = scaled.max(axis=2)
order = np.argsort(-per.ravel())[:8]
for o in order:
    s, m = divmod(int(o), 6)
    k = int(np.argmax(scaled[s, m]))
    print(f"set {s} meas {m}: scaled {per[s, m]:.2e} at t-index {k} (S/S0 {c9[s, m, k] / c9[s, m, 0]:.2e}) top3 {top3[s, m].max():.2e} "
          f"steps7 {ns7[s, m].tolist()} steps9 {ns9[s, m].tolist()} st {st7[s, m]} 8v9 {s89[s, m].max():.2e}")
print("99.9 percentile of per-curve scaled error", np.percentile(per, 99.9), "median", np.median(per))
ws = sorted(set(int(o) // 6 for o in order))
np.savez(os.path.join(ROOT, "gpurun_out", "worst.npz"), sets=np.array(ws), states=states[ws], c7=c7[ws], c9=c9[ws])
