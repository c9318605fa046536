"""CPU emulation of K1's 3xTF32 folded DFT: IEEE round-to-nearest vs round-toward-zero fp32 accumulation,
and three orders of the K dimension. With truncation the emulation reproduces the error distribution
measured on the B200 (median ~6e-7, p99.99 ~2e-5, max ~1e-4 in the log10 domain), i.e. the tensor core's
fp32 accumulate truncates; reordering K does not help. (numpy only; ~1 minute.)"""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle.logmel import mel_filterbank, hann_periodic, power_spectrogram_f64
def rn_tf32(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xffffe000).astype(np.uint32).view(np.float32)
def trunc32(x64):   # round toward zero to fp32
    f = x64.astype(np.float32)
    bad = np.abs(f.astype(np.float64)) > np.abs(x64)
    f = np.where(bad, np.nextafter(f, np.float32(0)), f)
    return f
rng=np.random.default_rng(5)
n=16000*20; t=np.arange(n)/16000.0
a=0.1*rng.standard_normal(n)
for h in range(1,11): a+=(0.3/h)*np.sin(2*np.pi*220*h*t)*(0.5+0.5*np.sin(2*np.pi*3*t))
a[int(0.9*n):]=0; a=a.astype(np.float32)
pad=np.pad(a,(200,200),mode='reflect'); F=n//160
idx=(np.arange(F)*160)[:,None]+np.arange(400)[None,:]
fr=pad[idx]; nn=np.arange(1,201)
e64=fr[:,nn].astype(np.float64)+fr[:,400-nn]; o64=fr[:,nn].astype(np.float64)-fr[:,400-nn]
e=e64.astype(np.float32); o=o64.astype(np.float32)
w=hann_periodic()[nn]; k=np.arange(201); ph=(np.outer(nn,k)%400)
C=(w[:,None]*np.cos(2*np.pi*ph/400)); C[199]*=0.5
S=-(w[:,None]*np.sin(2*np.pi*ph/400)); S[199]=0
def split64(x64):
    hi=rn_tf32(x64.astype(np.float32)); lo=rn_tf32((x64-hi.astype(np.float64)).astype(np.float32)); return hi,lo
Chi,Clo=split64(C); Shi,Slo=split64(S)
ehi=rn_tf32(e); elo=rn_tf32(e-ehi); ohi=rn_tf32(o); olo=rn_tf32(o-ohi)
W=mel_filterbank().astype(np.float64)
want=np.log10(np.maximum(W@power_spectrogram_f64(a),1e-10))
def run(order, trunc):
    def gemm(ahi,alo,bhi,blo):
        acc=np.zeros((F,201),np.float32)
        for j in range(25):
            sl=order[8*j:8*j+8]
            for A,B in ((ahi,blo),(alo,bhi),(ahi,bhi)):
                s=acc.astype(np.float64)+(A[:,sl].astype(np.float64)@B[sl].astype(np.float64))
                acc=trunc32(s) if trunc else s.astype(np.float32)
        return acc
    re=gemm(ehi,elo,Chi,Clo); im=gemm(ohi,olo,Shi,Slo)
    P=(re.astype(np.float64)**2+im.astype(np.float64)**2)
    got=np.log10(np.maximum(W@P.T,1e-10)); e_=np.abs(got-want)
    return e_.max(), np.quantile(e_,0.9999), np.median(e_)
nat=np.arange(200)
strided=np.array([j+25*m for j in range(25) for m in range(8)])
sym=np.array([x for j in range(100) for x in (j,199-j)])   # pair window edge with window centre
for name,order in (("natural",nat),("strided 25",strided),("edge/centre pairs",sym)):
    for trunc in (False,True):
        print(name, "trunc" if trunc else "RN   ", ["%.2e"%v for v in run(order,trunc)])
