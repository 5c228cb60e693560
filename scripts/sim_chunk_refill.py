"""Scalar mirror of the -DVR_TRACE_CHUNK refill logic of k_trace (kernels.cu): warps claim the queue in chunks of 32, keep\ntwo chunks of entries, hand entries to idle lanes by rank. Checks that every queue entry is traced exactly once and that\nevery warp terminates, for queue lengths around the chunk boundaries and refill thresholds 12 / 20 / 32.\n  python scripts/sim_chunk_refill.py"""
import random
NO=0xFFFFFFFF
def run(n, n_warps, seed, threshold=12):
    rnd=random.Random(seed)
    cursor=[0]
    def atomic_add(k):
        v=cursor[0]; cursor[0]+=k; return v
    processed=[0]*n
    class W: pass
    warps=[]
    def load(base):
        return [ (base+l if base+l<n else NO) for l in range(32)]
    for w in range(n_warps):
        x=W(); base=atomic_add(64)
        x.cur=load(base); x.nxt=load(base+32)
        x.claim = atomic_add(32) if base+64<n else n
        x.used=0; x.have=[False]*32; x.left=[0]*32; x.exhausted=False; x.done=False
        warps.append(x)
    def refill(x):
        if x.exhausted: return
        need=[not h for h in x.have]
        while any(need):
            avail=32-x.used
            invalid=False
            rank=0
            for l in range(32):
                if need[l]:
                    r=rank; rank+=1
                    take = (not x.have[l]) and r<avail
                    e = x.cur[(x.used+r)&31]
                    if take and e!=NO:
                        processed[e]+=1; x.have[l]=True; x.left[l]=rnd.randint(1,40)
                    if take and e==NO: invalid=True
            if invalid:
                x.exhausted=True; break
            wanted=sum(need)
            x.used+=min(wanted,avail)
            if x.used<32: break
            x.cur=x.nxt; x.used=0
            base=x.claim
            x.nxt=load(base)
            if base<n: x.claim=atomic_add(32)
            need=[not h for h in x.have]
    steps=0
    while not all(x.done for x in warps):
        x=rnd.choice(warps)
        if x.done: continue
        steps+=1
        assert steps<10_000_000
        refill(x)
        if not any(x.have): x.done=True; continue
        # inner loop until live==0 or (not exhausted and live<threshold)
        while True:
            for l in range(32):
                if x.have[l]:
                    x.left[l]-=1
                    if x.left[l]<=0: x.have[l]=False
            live=sum(x.have)
            if live==0 or (not x.exhausted and live<threshold): break
    assert all(p==1 for p in processed), (n, n_warps, [i for i,p in enumerate(processed) if p!=1][:10])
for n in [0,1,31,32,33,63,64,65,95,96,97,1000,4096,5000,100000]:
    for nw in [1,2,7,64]:
        for seed in range(3):
            run(n,nw,seed, threshold=[12,20,32][seed])
print("ok")
