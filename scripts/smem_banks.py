"""Bank-conflict search for the shared-memory layout of the stage kernel (developer tool):
part 1 searches (pad_j, pad_k) per order; part 2 checks the XOR swizzles used for n = 8, 16."""
import itertools
def conflicts(n, pj, pk, split=2):
    SJ = n + pj; SK = SJ*n + pk
    NP = n*n
    nthr = ((NP*split + 31)//32)*32
    tot = {}
    for name in ('r','s','t'):
        worst = 0; sumdeg = 0; cnt = 0
        for w in range(nthr//32):
            for half in range(2):
                lanes = [w*32 + half*16 + l for l in range(16)]
                for m in range(n):
                    addrs = set()
                    for t in lanes:
                        p = t % NP; h = t // NP
                        if h >= split: continue
                        a, b = p % n, p // n
                        if name == 'r': ad = m + SJ*a + SK*b      # (j=a,k=b)
                        elif name == 's': ad = a + SJ*m + SK*b    # (i=a,k=b)
                        else: ad = a + SJ*b + SK*m                # (i=a,j=b)
                        addrs.add(ad)
                    banks = {}
                    for ad in addrs: banks[ad % 16] = banks.get(ad % 16, 0) + 1
                    deg = max(banks.values()) if banks else 1
                    worst = max(worst, deg); sumdeg += deg; cnt += 1
        tot[name] = (worst, sumdeg / cnt)
    return tot
for n in range(2, 17):
    best = None
    for pj, pk in itertools.product(range(0, 4), range(0, 17)):
        c = conflicts(n, pj, pk)
        score = sum(v[1] for v in c.values())
        size = (n + pj) * n * n + pk * n
        key = (round(score, 3), size)
        if best is None or key < best[0]:
            best = (key, pj, pk, c)
    print(n, best[1], best[2], best[0], {k: (v[0], round(v[1], 2)) for k, v in best[3].items()})
def at8(i,j,k): return (i ^ ((j>>1) + 4*(k&1))) + 8*(j ^ (k&1)) + 64*k
def check(n, at, split=2):
    NP=n*n; nthr=((NP*split+31)//32)*32
    res={}
    for name in ('r','s','t','lin'):
        worst=0; tot=0; cnt=0
        for w in range(nthr//32):
            for half in range(2):
                lanes=[w*32+half*16+l for l in range(16)]
                for m in range(n if name!='lin' else 1):
                    addrs=set()
                    for t in lanes:
                        p=t%NP; h=t//NP
                        if name=='lin':
                            node=t
                            if node>=n**3: continue
                            addrs.add(at(node%n,(node//n)%n,node//(n*n))); continue
                        if h>=split: continue
                        a,b=p%n,p//n
                        if name=='r': ad=at(m,a,b)
                        elif name=='s': ad=at(a,m,b)
                        else: ad=at(a,b,m)
                        addrs.add(ad)
                    banks={}
                    for ad in addrs: banks[ad%16]=banks.get(ad%16,0)+1
                    deg=max(banks.values()) if banks else 1
                    worst=max(worst,deg); tot+=deg; cnt+=1
        res[name]=(worst,round(tot/cnt,2))
    return res
print(8, check(8, at8))
# bijectivity
s=set(at8(i,j,k) for i in range(8) for j in range(8) for k in range(8)); print(len(s), min(s), max(s))
def at16(i,j,k): return (i ^ j) + 16*j + 256*k
print(16, check(16, at16, 1))
def at12(i,j,k): return i + 13*j + 156*k
print(12, check(12, at12))
