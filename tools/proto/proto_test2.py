import sys, time
import numpy as np
from proto_nq import *
names = "n0 p0 mu_n mu_p ks Cn Cp Sf Sb tauN tauP eps Tm m".split()
units = np.array([1e-21,1e-21,1e5,1e5,1e12,1e33,1e33,0.01,0.01,1,1,1,1,1])
guess = np.array([1e8,3e15,20,20,4.8e-11,4.4e-29,4.4e-29,10,10,511,871,10,300,1.0])
idx = {n:i for i,n in enumerate(names)}
ini = np.loadtxt("/root/reference/Inputs/staub_MAPI_threepower_twothick_input.csv", delimiter=",")
d=np.loadtxt("/root/reference/Inputs/real_staub_aug_corr_renoised.csv",delimiter=",")
t = d[:141,0]; t = t[t<=2000]
lengths=[311,2000,311,2000,311,2000]
s = guess*units
for rtol in [1e-5,1e-6,1e-7,1e-8]:
  for m in range(6):
    p = make_par(s, idx, lengths[m], 128)
    dN = ini[m]*1e-21
    y = np.zeros(256); y[0::2] = dN + p.n0
    st={}
    t0=time.perf_counter()
    out = integrate(p, y, t, rtol=rtol, atol=1e-18, stats=st)
    el=time.perf_counter()-t0
    ref_d = np.load(f"/tmp/work/pl_{m}_1e-07.npy"); ref_t = np.load(f"/tmp/work/pl_{m}_1e-10.npy")
    print(f"rtol {rtol:g} m{m} steps {st['nsteps']} rej {st['nrej']}  vs tight {np.max(np.abs(out/ref_t-1)):.2e}  vs default {np.max(np.abs(out/ref_d-1)):.2e}  ref d-vs-t {np.max(np.abs(ref_d/ref_t-1)):.2e}  {el:.1f}s")
