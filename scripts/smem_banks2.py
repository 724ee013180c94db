"""Bank-conflict search for the slab kernel's shared-memory layout (developer tool).
Lane mappings as in stage_slab.cu: r/s items r = lane index within the (g, pencil) block,
g = r // (N*KB), p = r % (N*KB), pa = p % N, pb = p // N; plane-mapped linear accesses
(thread = node (pi,pj), NPL planes at a time).  64-bit accesses: conflicts are counted per
half-warp (16 lanes) as the maximum number of distinct addresses per 8-byte bank."""
import itertools, sys

def ks_for(n): return 1 if n <= 8 else (2 if n <= 12 else 4)
def r32(x): return (x + 31) // 32 * 32

def degree(addrs):
    banks = {}
    for a in set(addrs):
        banks.setdefault(a % 16, set()).add(a)
    return max((len(v) for v in banks.values()), default=1)

def score(n, SJ, SK, KB):
    SC = SK * KB
    blk = r32(2 * n * KB)
    res = {}
    for name in ("r", "s"):
        tot = cnt = 0
        for w0 in range(0, blk, 16):
            lanes = range(w0, w0 + 16)
            for m in range(n):
                ad = []
                for r in lanes:
                    g, p = divmod(r, n * KB)
                    if g > 1: continue
                    pa, pb = p % n, p // n
                    base = (3 * g) * SC
                    ad.append(base + (m + SJ * pa + SK * pb if name == "r" else pa + SJ * m + SK * pb))
                if ad:
                    tot += degree(ad); cnt += 1
        res[name] = tot / cnt
    # linear (plane mapping): thread t -> node pnd = t % N2 (pi = pnd % n, pj = pnd // n), plane slot t // N2
    N2 = n * n
    nt = blk * 2 if blk * 2 <= 320 else blk  # approx
    tot = cnt = 0
    for w0 in range(0, max(nt, 32), 16):
        ad = []
        for t in range(w0, w0 + 16):
            ps, pnd = divmod(t, N2)
            if ps >= max(1, nt // N2): continue
            pi, pj = pnd % n, pnd // n
            kl = ps % KB
            ad.append(pi + SJ * pj + SK * kl + (ps // KB) * SC)
        if ad:
            tot += degree(ad); cnt += 1
    res["lin"] = tot / cnt
    if ks_for(n) == 1:  # t pattern from smem: thread (pa,pb)=(i,j), lanes over p = i + n*j (g blocks of N2)
        tot = cnt = 0
        for w0 in range(0, r32(2 * N2), 16):
            for m in range(n):
                ad = []
                for r in range(w0, w0 + 16):
                    g, p = divmod(r, N2)
                    if g > 1: continue
                    ad.append(3 * g * SC + p % n + SJ * (p // n) + SK * m)
                if ad:
                    tot += degree(ad); cnt += 1
        res["t"] = tot / cnt
    return res

for n in ([int(a) for a in sys.argv[1:]] or range(3, 17)):
    if n in (8, 16): continue
    KS = ks_for(n); KB = -(-n // KS)
    best = []
    for pj, pk in itertools.product(range(0, 4), range(0, 16)):
        SJ = n + pj; SK = SJ * n + pk
        sc = score(n, SJ, SK, KB)
        cost = 2 * sc["r"] + 2 * sc["s"] + sc["lin"] + sc.get("t", 0) * 2  # ~ accesses per node
        best.append((round(cost, 3), SK, pj, pk, {k: round(v, 2) for k, v in sc.items()}))
    best.sort()
    print(n, "KB", KB, "best", best[0], "| size-min", min(best, key=lambda b: (round(b[0] * 1.0, 1), b[1])))
