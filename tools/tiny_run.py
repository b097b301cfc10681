"""Developer helper: a tiny batch of the headline instantiation (nx=128, std) and of the traps
model (nx=256) - the thing to run under compute-sanitizer."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from metrotrpl_b200 import _capi
from tests import parity_cases as pc
ctx = _capi.Context(0)
g, prob, params, aux = pc.staub_problem()
ctx.set_problem(prob)
ll, st, ns, cur = ctx.loglik_batch(params[:8], aux[:8], _capi.make_opts(RTOL=1e-7), want_curves=True)
print("std nx=128:", ll[:, :, 0].sum(axis=1)[:3], ns[..., 0].mean())
print(pc.check_traps_irf(lambda p, P, A, o, w: (ctx.set_problem(p), ctx.loglik_batch(P, A, o, want_curves=w))[1]))
ctx.close()
