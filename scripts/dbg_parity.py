import sys, numpy as np
sys.path.insert(0, ".")
from tests.helpers import oracle_walnutspy, close
from tests.test_gpu_walnutspy_parity import run_cuda
rng = np.random.default_rng(3)
n = 6
q0 = np.empty((n, 11))
q0[:, 0] = 3.0 * rng.standard_normal(n)
q0[:, 1:] = np.exp(0.5 * q0[:, :1]) * rng.standard_normal((n, 10))
out,state=run_cuda("funnel",q0,"R2P",0.3,0.3,12,100,1234)
dr,dg=oracle_walnutspy("funnel",q0,"R2P",0.3,0.3,12,100,1234,list(range(n)))
g=out["diag"]; d=out["draws"]
np.set_printoptions(linewidth=250,precision=5,suppress=False)
for c in range(n):
    err=np.max(np.abs(d[:,c]-dr[:,c])/np.maximum(1,np.abs(dr[:,c])),axis=1)
    bad=np.nonzero(err>1e-10)[0]
    print("chain",c,"first bad it",bad[:1], "max err before", err[:bad[0]].max() if len(bad) and bad[0]>0 else err.max())
    if len(bad):
        it=bad[0]
        for k in range(max(0,it-1),it+1):
            print(" it",k); print("  cuda",g[k,c]); print("  orcl",dg[k,c]); print("  err",err[k])
