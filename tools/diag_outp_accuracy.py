import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from conftest import load_golden
from htk_b200.estep import ForwardBackward
from htk_b200 import synth
from htk_b200.flat import flatten
from oracle import oracle_lib as O
hs = synth.make_tied_triphone_set(n_states=600, M=8, n_phys=400, n_logical=400, n_centre=20, seed=21, spread=0.2)
fm = flatten(hs)
feats, labs = synth.sample_corpus(fm, n_utts=2, T=600, Q=60, seed=3)
feat=feats[0]; states=np.arange(fm.J,dtype=np.int32)
# exact float64
x=feat.astype(np.float64)
mean=fm.mean.astype(np.float64); iv=fm.ivar.astype(np.float64); gc=fm.gConst.astype(np.float64)
ex=np.zeros((len(x),fm.J))
for s in range(fm.J):
    o,e=fm.stateMixOff[s],fm.stateMixOff[s+1]
    g=fm.mixGauss[o:e]
    d=x[:,None,:]-mean[g][None]
    lp=-0.5*(gc[g][None]+np.sum(d*d*iv[g][None],axis=2))+fm.mixLogWt[o:e].astype(np.float64)[None]
    m=lp.max(1); ex[:,s]=m+np.log(np.exp(lp-m[:,None]).sum(1))
orc=O.state_loglik(fm,feat,states).astype(np.float64)
print('oracle(float ref) vs exact: max %.2e mean %.2e'%(np.abs(orc-ex).max(), np.abs(orc-ex).mean()))
for k in (1,2):
    fb=ForwardBackward(fm,gmm_kernel=k); got=fb.OutP(feat,states).astype(np.float64); fb.close()
    print('kernel',k,'vs exact: max %.2e mean %.2e rms %.2e | vs oracle: max %.2e mean %.2e'%(np.abs(got-ex).max(),np.abs(got-ex).mean(),np.sqrt(((got-ex)**2).mean()),np.abs(got-orc).max(),np.abs(got-orc).mean()))
fb=ForwardBackward(fm,gmm_kernel=2); got=fb.OutP(feat,states).astype(np.float64); fb.close()
e=got-ex
print('TC signed error: mean %.3e std %.3e min %.3e max %.3e'%(e.mean(), e.std(), e.min(), e.max()))
for lo,hi in ((-40,-0),(-60,-40),(-80,-60),(-120,-80),(-1e9,-120)):
    m=(ex>=lo)&(ex<hi)
    if m.sum(): print('  b in [%g,%g): n=%d mean err %.3e std %.3e'%(lo,hi,m.sum(),e[m].mean(),e[m].std()))
# per-frame common shift does not matter for posteriors: remove per-frame mean
e2=e-e.mean(1,keepdims=True)
print('TC error after removing the per-frame mean: std %.3e max %.3e'%(e2.std(), np.abs(e2).max()))
