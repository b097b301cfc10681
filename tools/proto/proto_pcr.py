"""Numpy prototype of the warp-parallel block-tridiagonal solve (partition + PCR), lanes = axis 0."""
import numpy as np

def inv2(M):
    a,b,c,d = M[...,0,0],M[...,0,1],M[...,1,0],M[...,1,1]
    det = a*d-b*c
    R = np.empty_like(M)
    R[...,0,0]=d/det; R[...,0,1]=-b/det; R[...,1,0]=-c/det; R[...,1,1]=a/det
    return R

def shf(X, s):
    """value from lane l-s (s>0: shfl_up; s<0: shfl_down); out of range -> own value"""
    Y = X.copy()
    if s > 0: Y[s:] = X[:-s]
    elif s < 0: Y[:s] = X[-s:]
    return Y

def factor(A, B, C, NPL):
    """A,B,C: (32, NPL, 2, 2) sub/diag/super blocks. Returns factor dict."""
    F = {}
    ni = NPL - 1
    Dinv = np.zeros((32, max(ni,1), 2, 2)); Lm = np.zeros((32, max(ni,1), 2, 2))
    if ni > 0:
        D = B[:,0].copy(); Dinv[:,0] = inv2(D)
        for j in range(1, ni):
            Lm[:,j] = A[:,j] @ Dinv[:,j-1]
            D = B[:,j] - Lm[:,j] @ C[:,j-1]
            Dinv[:,j] = inv2(D)
        # spikes: V = T^-1 [A_0;0..], W = T^-1 [0..;C_{ni-1}]
        def tsolve(R):  # R (32, ni, 2, 2) matrix rhs
            Wk = R.copy()
            for j in range(1, ni): Wk[:,j] = Wk[:,j] - Lm[:,j] @ Wk[:,j-1]
            X = np.zeros_like(Wk)
            X[:,ni-1] = Dinv[:,ni-1] @ Wk[:,ni-1]
            for j in range(ni-2, -1, -1): X[:,j] = Dinv[:,j] @ (Wk[:,j] - C[:,j] @ X[:,j+1])
            return X
        RV = np.zeros((32, ni, 2, 2)); RV[:,0] = A[:,0]
        RW = np.zeros((32, ni, 2, 2)); RW[:,ni-1] = C[:,ni-1]
        V = tsolve(RV); W = tsolve(RW)
        Az = A[:,NPL-1]; Bz = B[:,NPL-1]; Cz = C[:,NPL-1]
        V0n = shf(V[:,0], -1); W0n = shf(W[:,0], -1)
        Ra = -Az @ V[:,ni-1]
        Rb = Bz - Az @ W[:,ni-1] - Cz @ V0n
        Rc = -Cz @ W0n
        F.update(V=V, W=W, Az=Az, Cz=Cz)
    else:
        Ra = A[:,0].copy(); Rb = B[:,0].copy(); Rc = C[:,0].copy()
    F.update(Dinv=Dinv, Lm=Lm, C=C, ni=ni)
    lane = np.arange(32)
    al = []; ga = []
    s = 1
    while s < 32:
        Rbi = inv2(Rb)
        alpha = -Ra @ shf(Rbi, s); gamma = -Rc @ shf(Rbi, -s)
        alpha[lane < s] = 0; gamma[lane + s > 31] = 0
        Ra_n = alpha @ shf(Ra, s); Rc_n = gamma @ shf(Rc, -s)
        Rb = Rb + alpha @ shf(Rc, s) + gamma @ shf(Ra, -s)
        Ra, Rc = Ra_n, Rc_n
        al.append(alpha); ga.append(gamma); s *= 2
    F.update(al=al, ga=ga, Rbinv=inv2(Rb))
    return F

def mv(M, v): return np.einsum('...ij,...j->...i', M, v)

def solve(F, r, NPL):
    """r: (32, NPL, 2)"""
    ni = F["ni"]; x = np.zeros_like(r)
    if ni > 0:
        Dinv, Lm, C = F["Dinv"], F["Lm"], F["C"]
        w = r[:, :ni].copy()
        for j in range(1, ni): w[:,j] = w[:,j] - mv(Lm[:,j], w[:,j-1])
        g = np.zeros_like(w)
        g[:,ni-1] = mv(Dinv[:,ni-1], w[:,ni-1])
        for j in range(ni-2, -1, -1): g[:,j] = mv(Dinv[:,j], w[:,j] - mv(C[:,j], g[:,j+1]))
        rr = r[:,NPL-1] - mv(F["Az"], g[:,ni-1]) - mv(F["Cz"], shf(g[:,0], -1))
    else:
        rr = r[:,0].copy()
    s = 1; k = 0
    while s < 32:
        rr = rr + mv(F["al"][k], shf(rr, s)) + mv(F["ga"][k], shf(rr, -s))
        s *= 2; k += 1
    z = mv(F["Rbinv"], rr)
    x[:,NPL-1] = z
    if ni > 0:
        zl = shf(z, 1); zl[0] = 0
        for j in range(ni):
            x[:,j] = g[:,j] - mv(F["V"][:,j], zl) - mv(F["W"][:,j], z)
    return x
