"""Prototype: integrate in transformed variables y_i = (ln N_i, Q_{i+1}/N_i) with RODAS4.
Block-diagonal change of variables keeps the Jacobian block tridiagonal."""
import sys, time
import numpy as np
from scipy.linalg import solve_banded
import proto_nq
from proto_nq import *
from check_rodas_coeffs import rodas4

def to_y(u):
    N=u[0::2]; Q=u[1::2]
    y=np.empty_like(u); y[0::2]=np.log(N); y[1::2]=Q/N
    return y
def to_u(y):
    v=y[0::2]; w=y[1::2]
    N=np.exp(v); u=np.empty_like(y); u[0::2]=N; u[1::2]=w*N
    return u
def rhs_y(p, y, want_jac=False):
    u=to_u(y); N=u[0::2]; Q=u[1::2]; w=y[1::2]
    if not want_jac:
        f=rhs(p,u)
    else:
        f,J=rhs(p,u,True)
    fN=f[0::2]; fQ=f[1::2]
    g=np.empty_like(y)
    g[0::2]=fN/N                    # dv/dt
    g[1::2]=(fQ - w*fN)/N           # dw/dt = (Q' N - Q N')/N^2 = (fQ - w fN)/N
    if not want_jac: return g
    n=len(y)
    # du/dy block diag: dN/dv = N ; dQ/dv = w N ; dQ/dw = N
    # G(y) = D(u) f(u), D = [[1/N, 0],[-w/N, 1/N]] per block
    # dG/dy = D J (du/dy) + (dD/dy) f
    Dm=np.zeros((n,n)); Um=np.zeros((n,n))
    for i in range(n//2):
        a=2*i; b=2*i+1
        Dm[a,a]=1/N[i]; Dm[b,a]=-w[i]/N[i]; Dm[b,b]=1/N[i]
        Um[a,a]=N[i]; Um[b,a]=w[i]*N[i]; Um[b,b]=N[i]
    Jy=Dm@J@Um
    for i in range(n//2):
        a=2*i; b=2*i+1
        # d/dv of (fN/N) explicit through 1/N: -fN/N ; d/dv of (fQ - w fN)/N: -(fQ - w fN)/N ; d/dw: -fN/N
        Jy[a,a]+= -fN[i]/N[i]
        Jy[b,a]+= -(fQ[i]-w[i]*fN[i])/N[i]
        Jy[b,b]+= -fN[i]/N[i]
    return g,Jy

def integrate_y(p, u0, tout, rtol=1e-7, stats=None, hist_out=None):
    A, C, g, m, mhat = rodas4()
    t=0.0; y=to_y(u0); tend=tout[-1]; n=len(y)
    nsteps=nrej=0
    u=to_u(y)
    f0=rhs(p,u); pl=PL_of(p,u); dpl=dPL_of(p,u,f0)
    hist=[(t,pl,dpl)]; out=np.zeros(len(tout)); out[0]=pl; io=1
    gy=rhs_y(p,y)
    Nn,P,_,_=unpack(p,u)
    def scale(y):
        u=to_u(y); Nn,P,_,_=unpack(p,u)
        sc=np.empty_like(y); sc[0::2]=rtol                    # absolute on ln N == relative on N
        sc[1::2]=rtol*np.maximum(1.0, np.abs(P)/np.abs(Nn))   # error in Q relative to max(N,P): w = Q/N
        return sc
    sc=scale(y)
    d1=np.sqrt(np.mean((gy/sc)**2)); h=min(0.01/max(d1,1e-300)*np.sqrt(np.mean((1/sc)**2))**0 ,1e-3)
    h=min(1e-3, 0.01*np.sqrt(np.mean((np.maximum(np.abs(y),1)/sc)**2))/d1)
    errold=1e-4; hacc=h; first=True
    while io<len(tout):
        h=min(h,tend-t)
        gy,J=rhs_y(p,y,True)
        M=np.eye(n)/(g*h)-J
        ab=to_banded(M,4,4)
        U=np.zeros((6,n))
        ok=True
        for i in range(6):
            if i==0: fi=gy
            else:
                yi=y+A[i,:i]@U[:i]
                fi=rhs_y(p,yi)
            r=fi+(C[i,:i]/h)@U[:i]
            U[i]=solve_banded((4,4),ab,r)
        ynew=y+m@U
        sc=scale(y)
        err=np.sqrt(np.mean((U[5]/sc)**2))
        if not np.isfinite(err): err=1e10
        fac=max(0.2,min(6.0,err**0.25/0.9)); hnew=h/fac
        if err<=1.0:
            nsteps+=1
            if not first:
                facgus=(hacc/h)*(err**2/errold)**0.25/0.9
                facgus=max(1/6.0,min(5.0,facgus)); fac=max(fac,facgus); hnew=h/fac
            first=False; hacc=h; errold=max(1e-2,err)
            t+=h; y=ynew
            u=to_u(y); fu=rhs(p,u); pl=PL_of(p,u); dpl=dPL_of(p,u,fu)
            hist.append((t,pl,dpl))
            if hist_out is not None: hist_out.append((t,h,pl))
            while io<len(tout) and tout[io]<=t*(1+1e-14):
                out[io]=hermite_eval(hist,tout[io]); io+=1
            h=hnew
        else:
            nrej+=1; h=hnew; first=True
    if stats is not None: stats["nsteps"]=nsteps; stats["nrej"]=nrej
    return out

if __name__=="__main__":
    g = np.load("/root/repo/tests/golden/staub6.npz")
    names=[str(n) for n in g["names"]]; idx={n:i for i,n in enumerate(names)}
    t=g["t"]
    tot_a=tot_b=0
    for s in [0,1,3,4,5,7,8,12,16]:
        for m in [0,1,4,5]:
            p=make_par(g["states"][s]*g["units"],idx,g["lengths"][m],128)
            u0=np.zeros(256); u0[0::2]=g["ini"][m]*1e-21+p.n0
            sa={}; sb={}
            oa=integrate(p,u0,t,rtol=1e-7,atol=1e-20,stats=sa)
            ob=integrate_y(p,u0,t,rtol=1e-7,stats=sb)
            T=g["pl_tight"][s,m]; win=T>1e-3*T[0]
            ea=np.abs(oa/T-1)[win].max(); eb=np.abs(ob/T-1)[win].max()
            tot_a+=sa["nsteps"]; tot_b+=sb["nsteps"]
            print(f"s{s} m{m}: linear steps {sa['nsteps']:4d} err {ea:.1e} | log-var steps {sb['nsteps']:4d} rej {sb['nrej']} err {eb:.1e}", flush=True)
    print("total", tot_a, tot_b)
