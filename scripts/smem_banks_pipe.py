"""Bank-conflict model of the pipelined kernel's s-pencil phase (developer tool): wavefronts per
64-bit shared-memory access of a warp for the lane -> (pencil, group) mappings and the H/E skew."""
import sys

def pad_j(n): return 3 if n in (6, 14) else (1 if n == 12 else 0)
def pad_k(n): return 3 if n in (3, 4, 7) else (7 if n == 10 else 0)
def r32(x): return (x + 31) // 32 * 32
def even(x): return (x + 1) // 2 * 2

def lay(n):
    swz = n in (8, 16)
    sj = n if swz else n + pad_j(n)
    sk = n * n if swz else sj * n + pad_k(n)
    def at(i, j, k):
        if n == 8: return (i ^ ((j >> 1) + 4 * (k & 1))) + 8 * (j ^ (k & 1)) + 64 * k
        if n == 16: return (i ^ j) + 16 * j + 256 * k
        return i + sj * j + sk * k
    return at, sk

def wavefronts(addrs):
    """addrs: list of 32 (or fewer) double addresses (None = inactive) -> wavefronts of the warp access"""
    tot = 0
    for h in range(0, 32, 16):
        banks = {}
        for a in addrs[h:h + 16]:
            if a is None: continue
            banks.setdefault(a % 16, set()).add(a)
        tot += max((len(v) for v in banks.values()), default=0)
    return tot

def analyse(n, ks, ig, he):
    at, sk = lay(n)
    kb = (n + ks - 1) // ks
    sc = sk * kb
    xl = even(n * n * kb + 2)
    blk = r32(2 * n * kb)
    tot_u = tot_c = ideal = 0
    for w0 in range(0, blk, 32):
        lanes = []
        for r in range(w0, w0 + 32):
            if r >= 2 * n * kb: lanes.append(None); continue
            if ig:
                pa, q = r % n, r // n
                g, pb = q & 1, q >> 1
            else:
                g, p = r // (n * kb), r % (n * kb)
                pa, pb = p % n, p // n
            lanes.append((g, pa, pb))
        for m in range(n):
            ua = [None if l is None else (3 * sc + he if l[0] else 0) + at(l[1], m, l[2]) for l in lanes]
            ca = [None if l is None else l[1] + n * m + n * n * l[2] for l in lanes]
            tot_u += wavefronts(ua); tot_c += wavefronts(ca)
            ideal += (sum(1 for l in lanes if l is not None) + 15) // 16
    return tot_u / ideal, tot_c / ideal

for n in range(6, 17):
    ks = 1 if n <= 8 else 2 if n <= 10 else 3 if n <= 11 else 4 if n <= 12 else 5 if n <= 13 else 8
    base = analyse(n, ks, False, 0)
    best = min(((analyse(n, ks, True, he), he) for he in range(16)), key=lambda t: t[0][0] + t[0][1])
    print(f"n={n} ks={ks}: plain U x{base[0]:.2f} cof x{base[1]:.2f} | IG best he={best[1]} U x{best[0][0]:.2f} cof x{best[0][1]:.2f}")
