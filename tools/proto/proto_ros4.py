import numpy as np, sys, time
from scipy.linalg import solve_banded
import proto_nq
from proto_nq import *
from check_rodas_coeffs import to_classic, order_residuals

def ros4_lstab():
    g=0.57282
    A=np.zeros((4,4)); C=np.zeros((4,4))
    A[1,0]=2.0
    A[2,0]=0.1867943637803922e+01; A[2,1]=0.2344449711399156e+00
    A[3,:]=A[2,:]
    C[1,0]=-0.7137615036412310e+01
    C[2,0]=0.2580708087951457e+01; C[2,1]=0.6515950076447975e+00
    C[3,0]=-0.2137148994382534e+01; C[3,1]=-0.3214669691237626e+00; C[3,2]=-0.6949742501781779e+00
    m=np.array([0.2255570073418735e+01,0.2870493262186792e+00,0.4353179431840180e+00,0.1093502252409163e+01])
    e=np.array([-0.2815431932141155e+00,-0.7276199124938920e-01,-0.1082196201495311e+00,-0.1093502252409163e+01])
    return A,C,g,m,e

A,C,g,m,e=ros4_lstab()
al,G,b=to_classic(A,C,g,m)
print({k:f"{v:.2e}" for k,v in order_residuals(al,G,b).items()})
al,G,bh=to_classic(A,C,g,m-e)
print("embedded (m-e):",{k:f"{v:.2e}" for k,v in order_residuals(al,G,bh).items()})

def integrate_ros4(p, y0, tout, rtol=1e-7, atol=1e-20, stats=None):
    A,C,g,m,e=ros4_lstab()
    t=0.0; y=y0.copy(); tend=tout[-1]; n=len(y); nsteps=nrej=0
    f0=rhs(p,y); pl=PL_of(p,y); dpl=dPL_of(p,y,f0)
    hist=[(t,pl,dpl)]; out=np.zeros(len(tout)); out[0]=pl; io=1
    sc=scale_vec(p,y,rtol,atol)
    d0=np.sqrt(np.mean((y/sc)**2)); d1=np.sqrt(np.mean((f0/sc)**2))
    h=min(0.01*d0/d1,1e-3)
    first=True; errold=1e-4; hacc=h
    while io<len(tout):
        h=min(h,tend-t)
        f0,J=rhs(p,y,True)
        ab=to_banded(np.eye(n)/(g*h)-J)
        U=np.zeros((4,n))
        fi=f0
        for i in range(4):
            if i==1 or i==2:
                fi=rhs(p,y+A[i,:i]@U[:i])
            r=fi+(C[i,:i]/h)@U[:i]
            U[i]=solve_banded((3,3),ab,r)
        ynew=y+m@U
        errv=e@U
        sc=scale_vec(p,np.where(np.abs(ynew)>np.abs(y),ynew,y),rtol,atol)
        err=np.sqrt(np.mean((errv/sc)**2))
        if not np.isfinite(err): err=1e10
        fac=max(0.2,min(6.0,err**0.25/0.9)); hnew=h/fac
        if err<=1.0:
            nsteps+=1
            t+=h; y=ynew
            fn=rhs(p,y); pl=PL_of(p,y); dpl=dPL_of(p,y,fn)
            hist.append((t,pl,dpl))
            while io<len(tout) and tout[io]<=t*(1+1e-14):
                out[io]=hermite_eval(hist,tout[io]); io+=1
            h=hnew
        else:
            nrej+=1; h=hnew
    if stats is not None: stats["nsteps"]=nsteps; stats["nrej"]=nrej
    return out

if __name__=="__main__":
    gld=np.load("/root/repo/tests/golden/staub6.npz")
    names=[str(n) for n in gld["names"]]; idx={n:i for i,n in enumerate(names)}
    t=gld["t"]
    for rtol in (1e-7,1e-6,1e-5):
      ta=tb=0; wa=wb=0
      for s in [0,3,5,8,16]:
        for mm in [0,5]:
            p=make_par(gld["states"][s]*gld["units"],idx,gld["lengths"][mm],128)
            u0=np.zeros(256); u0[0::2]=gld["ini"][mm]*1e-21+p.n0
            sa={}; sb={}
            oa=integrate(p,u0,t,rtol=rtol,atol=1e-20,stats=sa)
            ob=integrate_ros4(p,u0,t,rtol=rtol,stats=sb)
            T=gld["pl_tight"][s,mm]
            ea=np.abs(oa/T-1).max(); eb=np.abs(ob/T-1).max()
            ta+=sa["nsteps"]; tb+=sb["nsteps"]; wa=max(wa,ea); wb=max(wb,eb)
      print(f"rtol {rtol:g}: RODAS4 steps {ta} worst err {wa:.1e} | ROS4 steps {tb} worst err {wb:.1e}", flush=True)
