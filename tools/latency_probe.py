"""Latency of small batches: RODAS4 one warp / one CTA per trajectory, and the order-6 extrapolation
integrator one warp / one CTA per trajectory.
usage (GPU box): python tools/latency_probe.py [n_sets ...]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from metrotrpl_b200 import _capi          # noqa: E402
import bench                              # noqa: E402
from tests import parity_cases as pc      # noqa: E402

ctx = _capi.Context(0)
g, prob, _, _ = pc.staub_problem()        # the six staub curves (nx = 128) and a measurement
params = _capi.pack_params(bench.draw_states(4096, seed=20261018), bench.IDX, bench.UNITS)
aux = _capi.default_aux(4096, 6, [1.0] * 6)
ctx.set_problem(prob)
out = []
for n in [int(a) for a in sys.argv[1:]] or [1, 8, 32, 64, 128, 256, 1024, 4096]:
    row = {"n_sets": n, "n_traj": n * prob.n_meas}
    for name, flag in (("warp", 0), ("cta", _capi.OPT_CTA_PER_TRAJ), ("seulex_warp", _capi.OPT_EXTRAPOLATION),
                       ("seulex_cta", _capi.OPT_EXTRAPOLATION | _capi.OPT_CTA_PER_TRAJ)):
        opts = _capi.make_opts(RTOL=1e-7, flags=flag | _capi.OPT_NO_EXPLICIT)
        ms = []
        for rep in range(6):
            ctx.upload(params[:n], aux[:n])
            ctx.run_resident(opts)
            ll, st, ns, _ = ctx.download()
            ms.append(ctx.last_kernel_ms())
        row[name + "_ms"] = float(np.median(ms[1:]))
        row[name + "_steps_max"] = int(ns.sum(axis=-1).max())
        row[name + "_steps_mean"] = float(ns.sum(axis=-1).mean())
    row["speedup_cta"] = row["warp_ms"] / row["cta_ms"]
    row["speedup_seulex_cta"] = row["warp_ms"] / row["seulex_cta_ms"]
    out.append(row)
    print(json.dumps(row), flush=True)
