"""Numerics of the fast float-step variants of the quantised Linear vs the reference chain: Q4_K, [256,3072], 64 tokens.
Emulated in numpy/torch on the CPU: reference = fp16 per-op rounding chain -> cast to act -> fp32-accumulated matmul -> round to act."""
import numpy as np, torch
rng=np.random.default_rng(0)
N,K,M=256,3072,64
nb=N*K//256
d=(rng.normal(0,0.01,nb)).astype(np.float16); dmin=(rng.normal(0,0.01,nb)).astype(np.float16)
sc=rng.integers(0,64,(nb,8)); mn=rng.integers(0,64,(nb,8)); q=rng.integers(0,16,(nb,8,32))
def f16(x): return x.astype(np.float16)
# reference chain fp16
D16=f16(d[:,None].astype(np.float32)*sc.astype(np.float32))   # fp16(d*sc) (d fp16 * sc fp16 exact product rounded)
M16=f16(dmin[:,None].astype(np.float32)*mn.astype(np.float32))
t=f16(D16[:,:,None].astype(np.float32)*q.astype(np.float32))
Wref16=f16(t.astype(np.float32)-M16[:,:,None].astype(np.float32))
ideal=(d[:,None].astype(np.float64)*sc)[:,:,None]*q-(dmin[:,None].astype(np.float64)*mn)[:,:,None]
fma=D16[:,:,None].astype(np.float64)*q-M16[:,:,None].astype(np.float64)   # exact fused value
Wfast16=f16(fma)
def tb(x): return torch.from_numpy(np.ascontiguousarray(x))
def bf(x): return tb(x.astype(np.float32)).to(torch.bfloat16)
def rel(a,b): return float(torch.linalg.norm(a.double()-b.double())/torch.linalg.norm(b.double()))
for act in ("f16","bf16"):
    dt=torch.float16 if act=="f16" else torch.bfloat16
    X=torch.randn(M,K,dtype=torch.float32).to(dt)
    def lin(W):  # W torch any dtype [N,K]; fp32 accumulate (double used as proxy), output rounded to act
        return (X.double()@W.double().t()).to(dt)
    Wi=tb(ideal.reshape(N,K))
    yi=X.double()@Wi.t()
    Wr=tb(Wref16.reshape(N,K)).to(dt)
    yr=lin(Wr)
    cands={
      "fp16 single-FMA W (then cast to act)":tb(Wfast16.reshape(N,K)).to(dt),
      "fp16 single-FMA W kept fp16 (mixed MMA)":tb(Wfast16.reshape(N,K)),
      "unrounded W (scale after MMA)":tb(fma.reshape(N,K)),
    }
    if act=="bf16":
        Db=bf(D16).double().numpy(); Mb=bf(M16).double().numpy()
        cands["bf16-native HFMA2.BF16"]=bf(Db[:,:,None]*q-Mb[:,:,None]).reshape(N,K)
        cands["fp32 FMA -> bf16 once"]=bf(fma).reshape(N,K)
    print(act,"reference vs ideal:",rel(yr,yi))
    for k,W in cands.items():
        y=lin(W); print(f"  {k:45s} vs ref {rel(y,yr):.2e}  vs ideal {rel(y,yi):.2e}   W vs Wref {rel(W.double(),Wr.double()):.2e}")
