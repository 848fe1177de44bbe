import sys, time
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
from oracle_lib import Ref, ref_align, Oracle, oracle_breaking_points
from emu_lib import EmuAligner
r=Ref(); e=EmuAligner(); o=Oracle()
ACGT=np.frombuffer(b"ACGT",dtype=np.uint8)
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def mutate(t, err):
    out=[]
    for c in t:
        x=rng.random()
        if x<err/3: out.append(ACGT[rng.integers(4)])
        elif x<2*err/3: continue
        elif x<err: out.append(c); out.append(ACGT[rng.integers(4)])
        else: out.append(c)
    return bytes(out) if out else b"A"
t0=time.time(); n_cases=0; bad=0
while time.time()-t0 < float(sys.argv[2]) if len(sys.argv)>2 else 300:
    shape=rng.integers(0,6)
    if shape==0: m=int(rng.integers(1,400)); 
    elif shape==1: m=int(rng.integers(400,4000))
    elif shape==2: m=int(rng.integers(4000,12000))
    elif shape==3: m=int(rng.integers(12000,30000))
    elif shape==4: m=int(rng.integers(1,3000))
    else: m=int(rng.integers(2000,9000))
    t=rng.choice(ACGT,size=m).tobytes()
    if shape==4:  # unrelated, skewed shapes
        q=rng.choice(ACGT,size=int(rng.integers(1,12000))).tobytes()
    else:
        err=float(rng.choice([0.0,0.01,0.05,0.12,0.2,0.35,0.6]))
        q=mutate(np.frombuffer(t,dtype=np.uint8),err)
        if rng.random()<0.15: q=q[:max(1,len(q)//int(rng.integers(2,6)))]   # truncated query
        if rng.random()<0.1: q=q+rng.choice(ACGT,size=int(rng.integers(1,3000))).tobytes()
    ops,score,cig=ref_align(r,q,t)
    for guess in (-1, int(score*rng.uniform(0,1.0)), score, int(score*rng.uniform(1.0,3.0))+int(rng.integers(0,50))):
        W=int(rng.choice([1,13,500,1000])); qf=int(rng.integers(0,1000)); tb=int(rng.integers(0,100000))
        a,sa,depth,leaves=e.align(q,t,qf,tb,W,guess=guess)
        want=oracle_breaking_points(o,ops,qf,tb,tb+len(t),W)
        ok = sa==score and a.shape==ops.shape and (a==ops).all() and e.cigar==cig and e.breaking_points.shape==want.shape and (e.breaking_points==want).all()
        n_cases+=1
        if not ok:
            bad+=1; print("MISMATCH", len(q), len(t), guess, score, sa, flush=True)
            open(f"/tmp/fuzz_fail_{n_cases}.txt","wb").write(q+b"\n"+t+b"\n")
print("cases", n_cases, "bad", bad, "seconds", round(time.time()-t0))
