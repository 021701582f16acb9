"""List-scheduling model of the 64 x 64 tiles of a GEMM on 148 SMs x 4 resident CTAs (processor sharing inside an SM): dispatch orders
for products with triangular operands (k range of a tile depends on its row / column)."""
import heapq, itertools
def simulate(tiles, P=148, slots=4, per_cta_max=1.0, overhead=0.03):
    # tiles: list of work (in k-tile units of 64), in dispatch order; SM rate 1 unit of work per unit time shared equally among resident CTAs,
    # each CTA limited to per_cta_max of the SM rate. overhead: fixed per-CTA cost (units) for prologue/epilogue.
    # event-driven processor sharing per SM.
    sms=[[] for _ in range(P)]  # list of remaining work
    t=0.0; nxt=0; n=len(tiles)
    # initial fill round-robin
    def fill():
        nonlocal nxt
        progressed=True
        while nxt<n and progressed:
            progressed=False
            # choose SM with fewest residents (ties lowest idx)
            best=min(range(P), key=lambda s:(len(sms[s]), s))
            if len(sms[best])<slots:
                sms[best].append(tiles[nxt]+overhead); nxt+=1; progressed=True
    fill()
    while any(sms):
        # time to next completion
        dt=min((min(r)/min(per_cta_max,1.0/len(r)) for r in sms if r))
        for r in sms:
            if r:
                rate=min(per_cta_max,1.0/len(r))
                for i in range(len(r)): r[i]-=rate*dt
        t+=dt
        for s in range(P): sms[s]=[x for x in sms[s] if x>1e-9]
        fill()
    return t
T=32
def order_raster(work):
    # current raster: GROUP_M=16 over tm, inner tn... lin = blockIdx.x + gridDim.x*blockIdx.y ; grp = lin // (16*T); r = lin - grp*16*T; tm = grp*16 + r%16; tn = r//16
    out=[]
    for lin in range(T*T):
        grp=lin//(16*T); r=lin-grp*16*T; tm=grp*16+r%16; tn=r//16
        w=work(tm,tn)
        if w>0: out.append(w)
    return out
cases={
 "LOWER|KLO (Sinv=W'W)": lambda tm,tn: (T-tm) if tm>=tn else 0,
 "LOWER|KHI_N (t1 Linv')": lambda tm,tn: (tn+1) if tm>=tn else 0,
 "KHI_M (Linv dX)": lambda tm,tn: (tm+1),
 "full": lambda tm,tn: T,
 "LOWER (syrk)": lambda tm,tn: T if tm>=tn else 0,
}
for name,w in cases.items():
    tiles=order_raster(w)
    tot=sum(tiles)
    a=simulate(tiles); b=simulate(sorted(tiles,reverse=True))
    print(f"{name:28s} tiles {len(tiles):4d} work {tot:6d} ideal {tot/148:6.1f}  raster {a:6.1f}  longest-first {b:6.1f}")
print("--- snake / balanced orders")
def snake(tiles,P=148):
    s=sorted(tiles,reverse=True); out=[]
    for r in range(0,len(s),P):
        chunk=s[r:r+P]
        out+= chunk if (r//P)%2==0 else chunk[::-1]
    return out
for name,w in cases.items():
    tiles=order_raster(w); tot=sum(tiles)
    c=simulate(snake(tiles))
    print(f"{name:28s} ideal {tot/148:6.1f} snake {c:6.1f}")
