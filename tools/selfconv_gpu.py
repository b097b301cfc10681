import sys; sys.path.insert(0,".")
import numpy as np, bench
from tests import parity_cases as pc
from metrotrpl_b200 import _capi
g, prob, params, aux = pc.staub_problem()
states = bench.draw_states(4096, seed=99)
names=[str(n) for n in g["names"]]; idx={n:i for i,n in enumerate(names)}
P=_capi.pack_params(states, idx, g["units"]); A=np.repeat(aux[:1],4096,axis=0)
ctx=_capi.Context(0); ctx.set_problem(prob)
_,st7,ns7,c7 = ctx.loglik_batch(P,A,_capi.make_opts(RTOL=1e-7),want_curves=True)
_,st9,ns9,c9 = ctx.loglik_batch(P,A,_capi.make_opts(RTOL=1e-9),want_curves=True)
c7=c7.reshape(4096,6,-1); c9=c9.reshape(4096,6,-1)
win=c9>1e-6*c9[:,:,:1]
with np.errstate(all="ignore"):
    err=np.where(win,np.abs(c7/c9-1),0)
bad=np.argwhere(err>0.1)
print("n bad", len(bad))
for a,b,k in bad[:8]:
    print(a,b,k,"c7",c7[a,b,max(0,k-2):k+2],"c9",c9[a,b,max(0,k-2):k+2],"st",st7[a,b],st9[a,b],"steps",ns7[a,b],ns9[a,b], "t", g["t"][max(0,k-2):k+2])
    print(dict(zip(names,states[a])))
efold=np.log(np.maximum(c9[:,:,:1]/np.maximum(c9,1e-300),1))
scaled=np.where(win, err/(1+efold), 0)
scaled[err>0.1]=0
print("max scaled err", scaled.max(), "max err top3", np.where(c9>1e-3*c9[:,:,:1], err*(err<0.1), 0).max())
