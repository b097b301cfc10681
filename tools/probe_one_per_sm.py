import json, sys, os
import numpy as np
sys.path.insert(0, ".")
from metrotrpl_b200 import _capi
import bench
from tests import parity_cases as pc
ctx = _capi.Context(0)
g, prob, _, _ = pc.staub_problem()
params = _capi.pack_params(bench.draw_states(4096, seed=20261018), bench.IDX, bench.UNITS)
aux = _capi.default_aux(4096, 6, [1.0] * 6)
ctx.set_problem(prob)
flag = _capi.OPT_EXTRAPOLATION | _capi.OPT_CTA_PER_TRAJ
opts = _capi.make_opts(RTOL=1e-7, flags=flag | _capi.OPT_NO_EXPLICIT)
for n in (25, 32, 40, 48, 64):
    row = {"n_traj": n * 6}
    for order in (False, True):
        ms = []
        ctx.set_queue_order(None)
        for rep in range(6):
            ctx.upload(params[:n], aux[:n]); ctx.run_resident(opts)
            ll, st, ns, _ = ctx.download(); ms.append(ctx.last_kernel_ms())
            if order: ctx.set_queue_order(np.argsort(-ns.sum(axis=-1).ravel(), kind="stable"))
        row["ordered_ms" if order else "default_ms"] = float(np.median(ms[2:]))
    row["steps_max"] = int(ns.sum(axis=-1).max())
    print(json.dumps(row), flush=True)
