import numpy as np, sys
from proto_nq import *
from proto_pcr import factor, solve
names = "n0 p0 mu_n mu_p ks Cn Cp Sf Sb tauN tauP eps Tm m".split()
units = np.array([1e-21,1e-21,1e5,1e5,1e12,1e33,1e33,0.01,0.01,1,1,1,1,1])
idx = {n:i for i,n in enumerate(names)}
lo = np.array([1e8,1e14,1,1,1e-11,1e-29,1e-29,1e-4,1e-4,1,1,10,300,1.0])
hi = np.array([1e8,1e16,100,100,1e-9,1e-27,1e-27,1e4,1e4,1500,3000,10,300,1.0])
rng = np.random.default_rng(1)
ini = np.loadtxt("/root/reference/Inputs/staub_MAPI_threepower_twothick_input.csv", delimiter=",")
lengths=[311,2000,311,2000,311,2000]
worst = 0
for trial in range(300):
    st = 10**rng.uniform(np.log10(lo), np.log10(hi)); s = st*units
    m = rng.integers(6)
    NPL = 4; L = 128
    p = make_par(s, idx, lengths[m], L)
    dN = ini[m]*1e-21 * 10**rng.uniform(-3, 0)
    # random-ish state: perturb N, and a random charge field
    N = (dN + p.n0)*np.exp(0.3*rng.standard_normal(L))
    y = np.zeros(2*L); y[0::2]=N
    Q = 1e-3*N.mean()*np.cumsum(rng.standard_normal(L)); Q -= np.linspace(0,1,L)*Q[-1]; Q[-1]=0
    y[1::2] = Q * rng.choice([0,1,1e-2])
    Nn,P,_,_ = unpack(p,y)
    if P.min() <= 0: continue
    f, J = rhs(p, y, True)
    h = 10**rng.uniform(-6, 3)
    M = np.eye(2*L)/(0.25*h) - J
    r = rng.standard_normal(2*L)*np.abs(f).max()
    xref = np.linalg.solve(M, r)
    # blocks
    A = np.zeros((L,2,2)); B=np.zeros((L,2,2)); C=np.zeros((L,2,2))
    for i in range(L):
        B[i] = M[2*i:2*i+2, 2*i:2*i+2]
        if i>0: A[i] = M[2*i:2*i+2, 2*i-2:2*i]
        if i<L-1: C[i] = M[2*i:2*i+2, 2*i+2:2*i+4]
    F = factor(A.reshape(32,NPL,2,2), B.reshape(32,NPL,2,2), C.reshape(32,NPL,2,2), NPL)
    x = solve(F, r.reshape(32,NPL,2), NPL).reshape(-1)
    # error in the scaled norm
    sc = np.abs(xref).reshape(L,2).max(axis=0)
    e = (np.abs(x-xref).reshape(L,2)/sc).max()
    res = np.abs(M@x - r).max()/np.abs(r).max()
    cond = np.linalg.cond(M)
    worst = max(worst, e)
    if e > 1e-9: print(trial, f"h={h:.2e} err={e:.2e} res={res:.2e} cond={cond:.2e}")
print("worst", worst)
