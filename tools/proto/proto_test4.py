import sys, time
import numpy as np
import proto_nq
from proto_nq import *
g = np.load("/root/repo/tests/golden/staub6.npz")
names = [str(n) for n in g["names"]]; idx = {n:i for i,n in enumerate(names)}
t = g["t"]
s = int(sys.argv[1]); m = int(sys.argv[2]); rtol=float(sys.argv[3])
st = g["states"][s]; print(dict(zip(names, st)))
p = make_par(st*g["units"], idx, g["lengths"][m], 128)
y = np.zeros(256); y[0::2] = g["ini"][m]*1e-21 + p.n0
# instrument: log steps
log=[]
orig = proto_nq.PL_of
def PLlog(p_, y_):
    v = orig(p_, y_); log.append((v, y_[0::2].max(), y_[0::2].min())); return v
proto_nq.PL_of = PLlog
stt={}
out = integrate(p, y, t, rtol=rtol, atol=1e-20, stats=stt)
print(stt)
pl = np.array([l[0] for l in log]); 
print("PL(0)", pl[0])
# decades vs step index
for k in range(0, len(pl), max(1,len(pl)//40)):
    print(k, f"{pl[k]/pl[0]:.3e}", f"Nmax {log[k][1]:.3e} Nmin {log[k][2]:.3e}")
ref = g["pl_tight"][s,m]
print(np.c_[t[::10], out[::10]/out[0], ref[::10]/ref[0]])
