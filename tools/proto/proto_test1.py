import sys, time
sys.path.insert(0, "/root/reference")
import numpy as np
from forward_solver import dydt_numba
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
# 1. RHS check vs reference on a perturbed state
rng = np.random.default_rng(0)
for m in [0,1]:
    p = make_par(s, idx, lengths[m], 128)
    dN = ini[m]*1e-21
    N = (dN + p.n0)*(1+0.01*rng.standard_normal(128)); P = (dN + p.p0)
    # make P consistent with zero net charge
    P = P + (np.sum(N-p.n0) - np.sum(P-p.p0))/128
    rho = (P-p.p0)-(N-p.n0)
    Q = np.concatenate(([0],np.cumsum(rho)))
    E = p.Lam*p.dx*Q
    yref = np.concatenate([N,P,E])
    dref = dydt_numba(0.0, yref, 128, p.dx, p.n0,p.p0,p.mun,p.mup,p.ks,p.Cn,p.Cp,p.Sf,p.Sb,p.tauN,p.tauP,p.Lam,p.Tm)
    y = np.empty(256); y[0::2]=N; y[1::2]=Q[1:]
    f, J = rhs(p, y, True)
    print("fN err", np.max(np.abs(f[0::2]-dref[:128])/np.abs(dref[:128]).max()))
    print("fQ err", np.max(np.abs(f[1::2]*p.Lam*p.dx - dref[257:])/np.abs(dref[257:]).max()))
    # FD jacobian check
    Jfd = np.zeros_like(J)
    for k in range(256):
        hh = 1e-6*max(abs(y[k]), 1e-9)
        yp = y.copy(); yp[k]+=hh; ym=y.copy(); ym[k]-=hh
        Jfd[:,k] = (rhs(p,yp)-rhs(p,ym))/(2*hh)
    Jfd[:,255]=0
    print("J err", np.max(np.abs(J-Jfd))/np.max(np.abs(J)), np.max(np.abs(J-Jfd)/(np.abs(J)+1e-3*np.max(np.abs(J)))))
    # bandwidth check
    r,c = np.nonzero(Jfd); print("bw", np.max(np.abs(r-c)))
