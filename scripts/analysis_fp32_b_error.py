"""Design input: where the remaining error of b = sum J^T r on the fp32 path of the n x n
calibration case comes from.  numpy emulation of the residual in different arithmetic / parameter-set precisions against
fp64, J held exact; scaled as in the tests, max |db_i| / (sqrt(H_ii) sqrt(sum r^T r)).  CPU only.

Result (100 000 observations, 0.5 px noise):
    fp32 residual (today)                          1.2e-05
    fp64 arithmetic, float-rounded set             1.1e-05    -> not the arithmetic
    fp64 arithmetic, fp64 set of float(x)          1.1e-05    -> not the set alone: float(x) rounds cx, fx too
    fp32 arithmetic, set = float(set(x64))         1.2e-05
    fp32 + lo-correction of fx fy cx cy            2.0e-06    -> the ~600-valued intrinsics (ulp 6e-5 px) are the cause;
                                                                 r -= fx_lo xd + cx_lo costs 4 FMA per observation
"""
import numpy as np, sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from common import camera_consts
from oracle import oracle_py as orc
f32=np.float32
Cm=camera_consts()[12:]; C=Cm.reshape(4,4)
X_GT = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027, 586.0, 722.0, 638.0, 323.0, -0.12, 0.05, 0.001, -0.0007, 0.01])
rng=np.random.default_rng(3); n=100000
pts=np.column_stack([rng.uniform(2,5,n),rng.uniform(-1,1,n),rng.uniform(-.5,1,n)]).astype(f32).astype(np.float64)
def setof(x):
    T=orc.so3_convert6dof(x[:6]); TC=(T@C)[:3]; return np.concatenate([TC.reshape(-1), x[6:]])
def resid(s,P,pix,dt):
    s=s.astype(dt); P=P.astype(dt)
    p=np.column_stack([P,np.ones(len(P),dt)])@s[:12].reshape(3,4).T
    xn=p[:,0]/p[:,2]; yn=p[:,1]/p[:,2]; r2=xn*xn+yn*yn
    rad=1+r2*(s[16]+r2*(s[17]+r2*s[20])); xy2=2*xn*yn
    xd=xn*rad+s[18]*xy2+s[19]*(2*xn*xn+r2); yd=yn*rad+s[18]*(2*yn*yn+r2)+s[19]*xy2
    return np.column_stack([pix[:,0].astype(dt)-(s[12]*xd+s[14]), pix[:,1].astype(dt)-(s[13]*yd+s[15])])
pix=(-resid(setof(X_GT),pts,np.zeros((n,2)),np.float64)+rng.normal(0,.5,(n,2))).astype(f32).astype(np.float64)
x=X_GT*(1+0.002*np.cos(np.arange(15)))
J=np.zeros((n,2,15))
for j in range(15):
    h=1e-6*abs(x[j]); xp=x.copy(); xm=x.copy(); xp[j]+=h; xm[j]-=h
    J[:,:,j]=(resid(setof(xp),pts,pix,np.float64)-resid(setof(xm),pts,pix,np.float64))/(2*h)
r64=resid(setof(x),pts,pix,np.float64)
xf=x.astype(f32).astype(np.float64)
r32=resid(setof(xf),pts,pix,f32).astype(np.float64)          # fp32 path today: float x, float set, float arithmetic
r64_fx=resid(setof(xf),pts,pix,np.float64)                    # fp64 residual from the float-rounded x (what setup stages)
r64_fset=resid(setof(xf).astype(f32).astype(np.float64),pts,pix,np.float64)  # fp64 arithmetic, float-rounded SET
H=np.einsum('nop,noq->pq',J,J); d=np.sqrt(np.diag(H)); s=(r64**2).sum()
bref=np.einsum('nop,no->p',J,r64)
for name,r in (("fp32 residual (today)",r32),("fp64 arithmetic, float-rounded set",r64_fset),("fp64 arithmetic, fp64 set of float(x)",r64_fx)):
    b=np.einsum('nop,no->p',J,r); print("%-42s max |db| / (d sqrt(s)) = %.2e   |ds|/s = %.2e"%(name,np.max(np.abs(b-bref)/(d*np.sqrt(s))),abs((r**2).sum()-s)/s))
r32_x64=resid(setof(x).astype(f32).astype(np.float64),pts,pix,f32).astype(np.float64)
r64_set32=resid(setof(x).astype(f32).astype(np.float64),pts,pix,np.float64)
for name,r in (("fp32 arithmetic, set = float(set(x64))",r32_x64),("fp64 arithmetic, set = float(set(x64))",r64_set32)):
    b=np.einsum('nop,no->p',J,r); print("%-42s max |db| / (d sqrt(s)) = %.2e   |ds|/s = %.2e"%(name,np.max(np.abs(b-bref)/(d*np.sqrt(s))),abs((r**2).sum()-s)/s))
def resid_hilo(sx,P,pix):
    s=sx.astype(f32); lo=(sx-s.astype(np.float64)).astype(f32); P=P.astype(f32)
    p=np.column_stack([P,np.ones(len(P),f32)])@s[:12].reshape(3,4).T
    xn=p[:,0]/p[:,2]; yn=p[:,1]/p[:,2]; r2=xn*xn+yn*yn
    rad=1+r2*(s[16]+r2*(s[17]+r2*s[20])); xy2=2*xn*yn
    xd=xn*rad+s[18]*xy2+s[19]*(2*xn*xn+r2); yd=yn*rad+s[18]*(2*yn*yn+r2)+s[19]*xy2
    r0=pix[:,0].astype(f32)-(s[12]*xd+s[14]); r1=pix[:,1].astype(f32)-(s[13]*yd+s[15])
    r0=r0-(lo[12]*xd+lo[14]); r1=r1-(lo[13]*yd+lo[15])
    return np.column_stack([r0,r1]).astype(np.float64)
r=resid_hilo(setof(x),pts,pix)
b=np.einsum('nop,no->p',J,r); print("%-42s max |db| / (d sqrt(s)) = %.2e   |ds|/s = %.2e"%("fp32 + lo-correction of fx fy cx cy",np.max(np.abs(b-bref)/(d*np.sqrt(s))),abs((r**2).sum()-s)/s))
