"""Design input for the n x n calibration case: can phase 2 of wide_pass_kernel (C += J^T B over 32
observations, K = 64) run on tensor cores?  Emulates tf32 operands (10-bit mantissa, fp32 accumulation per K = 64 group,
fp64 fold every 8 groups as the kernel does) on the Jacobian of the pinhole + distortion model, whose columns span three
orders of magnitude, and reports max |dH| / sqrt(H_ii H_jj) against fp64.  CPU only (numpy + the oracle's so3 helper).

Result (this script, 65 536 observations):
    fp32 FFMA (current)                4.0e-08
    1x tf32                            1.9e-05     -> misses the 1e-5 tolerance
    3x tf32 (hi*hi + hi*lo + lo*hi)    6.0e-08     -> as good as fp32 FFMA
"""
import numpy as np, sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from common import camera_consts
from oracle import oracle_py as orc
Cm=camera_consts()[12:]; C=Cm.reshape(4,4)
X_GT = np.array([-0.01, 0.02, -0.06, 0.018, -0.0013, 0.027, 586.0, 722.0, 638.0, 323.0, -0.12, 0.05, 0.001, -0.0007, 0.01])
rng=np.random.default_rng(3); n=65536
pts=np.column_stack([rng.uniform(2,5,n),rng.uniform(-1,1,n),rng.uniform(-.5,1,n)])
def setof(x):
    T=orc.so3_convert6dof(x[:6]); TC=(T@C)[:3]; return np.concatenate([TC.reshape(-1), x[6:]])
def resid(s,P,pix):
    p=np.column_stack([P,np.ones(len(P))])@s[:12].reshape(3,4).T
    xn=p[:,0]/p[:,2]; yn=p[:,1]/p[:,2]; r2=xn*xn+yn*yn
    rad=1+r2*(s[16]+r2*(s[17]+r2*s[20])); xy2=2*xn*yn
    xd=xn*rad+s[18]*xy2+s[19]*(2*xn*xn+r2); yd=yn*rad+s[18]*(2*yn*yn+r2)+s[19]*xy2
    return np.column_stack([pix[:,0]-(s[12]*xd+s[14]), pix[:,1]-(s[13]*yd+s[15])])
pix=-resid(setof(X_GT),pts,np.zeros((n,2)))+rng.normal(0,.5,(n,2))
x=X_GT*(1+0.002*np.cos(np.arange(15)))
J=np.zeros((n,2,15))
for j in range(15):
    h=1e-6*abs(x[j]); xp=x.copy(); xm=x.copy(); xp[j]+=h; xm[j]-=h
    J[:,:,j]=(resid(setof(xp),pts,pix)-resid(setof(xm),pts,pix))/(2*h)
A=J.reshape(n*2,15).astype(np.float32)          # rows = (obs, output), what phase 2 multiplies: C += A^T A
def tf32(a):
    b=a.astype(np.float32).view(np.uint32).astype(np.uint64)
    b=(b+0x1000)&0xFFFFE000  # round to nearest (ties away) at 13 dropped bits
    return b.astype(np.uint32).view(np.float32)
Href=A.astype(np.float64).T@A.astype(np.float64)
d=np.sqrt(np.diag(Href)); S=np.outer(d,d)
def accumulate(prod_fn, K=64, flush=8):
    acc64=np.zeros((15,15)); acc32=np.zeros((15,15),np.float32); g=0
    for i in range(0,A.shape[0],K):
        acc32=(acc32+prod_fn(A[i:i+K])).astype(np.float32); g+=1
        if g%flush==0: acc64+=acc32; acc32[:]=0
    return acc64+acc32
fp32=lambda a: (a.T@a).astype(np.float32)
def t1(a): h=tf32(a); return (h.astype(np.float64).T@h.astype(np.float64)).astype(np.float32)
def t3(a):
    h=tf32(a); l=tf32(a-h); h64=h.astype(np.float64); l64=l.astype(np.float64)
    return (h64.T@h64+h64.T@l64+l64.T@h64).astype(np.float32)
for name,fn in (("fp32 FFMA (current)",fp32),("1x tf32",t1),("3x tf32 (hi*hi + hi*lo + lo*hi)",t3)):
    Hh=accumulate(fn); print("%-34s max |dH| / sqrt(Hii Hjj) = %.2e"%(name,np.max(np.abs(Hh-Href)/S)))
