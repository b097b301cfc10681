import numpy as np
import proto_nq
from proto_nq import *
g = np.load("/root/repo/tests/golden/staub6.npz")
names=[str(n) for n in g["names"]]; idx={n:i for i,n in enumerate(names)}
t=g["t"]
def make_scale(qf, mode="rms"):
    def scale_vec(p, y, rtol, atol):
        N,P,_,_=unpack(p,y)
        sc=np.empty_like(y)
        sc[0::2]=atol+rtol*np.abs(N)
        sc[1::2]=(atol+rtol*np.maximum(np.abs(N),np.abs(P)))*qf
        return sc
    return scale_vec
for qf in [1.0, 10.0, 100.0, 1e4]:
    proto_nq.scale_vec=make_scale(qf)
    tot=0; worst=0
    for s in [0,1,3,5,8,12,16]:
        for m in [0,1,4,5]:
            p=make_par(g["states"][s]*g["units"],idx,g["lengths"][m],128)
            y=np.zeros(256); y[0::2]=g["ini"][m]*1e-21+p.n0
            st={}
            out=integrate(p,y,t,rtol=1e-7,atol=1e-20,stats=st)
            tot+=st["nsteps"]; worst=max(worst,np.max(np.abs(out/g["pl_tight"][s,m]-1)))
    print(f"Q scale x{qf:g}: steps {tot} worst err {worst:.2e}", flush=True)
