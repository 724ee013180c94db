"""Bank-conflict model of the pipelined kernel's shared-memory access patterns (developer tool).
For an order n it searches the padding of the U/R layout (Lay<N>: i + SJ*j + SK*k), the H/E skew
HE and the lane mappings for the smallest weighted number of wavefronts per 64-bit access.
usage: python scripts/smem_banks_pipe.py n [ks]"""
import itertools
import sys


def r32(x): return (x + 31) // 32 * 32


def wavefronts(addrs):
    tot = 0
    for h in range(0, len(addrs), 16):
        banks = {}
        for a in addrs[h:h + 16]:
            if a is None:
                continue
            banks.setdefault(a % 16, set()).add(a)
        tot += max((len(v) for v in banks.values()), default=0)
    return tot


def ideal(addrs):
    return sum((sum(1 for a in addrs[h:h + 16] if a is not None) + 15) // 16 for h in range(0, len(addrs), 16))


def model(n, ks, pj, pk, he, ig_s=True, nt=None, split=2):
    kb = (n + ks - 1) // ks
    if n == 8 and pj == 0 and pk == 0:
        at = lambda i, j, k: (i ^ ((j >> 1) + 4 * (k & 1))) + 8 * (j ^ (k & 1)) + 64 * k
        sk = 64
    elif n == 16 and pj == 0 and pk == 0:
        at = lambda i, j, k: (i ^ j) + 16 * j + 256 * k
        sk = 256
    else:
        sj = n + pj
        sk = sj * n + pk
        at = lambda i, j, k: i + sj * j + sk * k
    sc = sk * kb
    comp = lambda g: (3 * sc + he) if g else 0
    blk = r32(2 * n * kb)
    res = {}

    def acc(name, weight, lanes_list):
        w = i_ = 0
        for lanes in lanes_list:
            w += wavefronts(lanes); i_ += ideal(lanes)
        res[name] = (weight, w / max(i_, 1))

    def rs_lanes(ig):
        out = []
        for w0 in range(0, blk, 32):
            lanes = []
            for r in range(w0, w0 + 32):
                if r >= 2 * n * kb:
                    lanes.append(None); continue
                if ig:
                    pa, q = r % n, r // n
                    g, pb = q & 1, q >> 1
                else:
                    g, p = r // (n * kb), r % (n * kb)
                    pa, pb = p % n, p // n
                lanes.append((g, pa, pb))
            out.append(lanes)
        return out

    # P1: r-pencils (plain mapping): U read at(m, j, k), R write at(o, j, k)
    l1 = rs_lanes(False)
    acc("P1 U/R", 18, [[None if l is None else comp(l[0]) + at(m, l[1], l[2]) for l in ls] for ls in l1 for m in range(n)])
    # P2: s-pencils: U read / R rw at(i, m, k); cof linear
    l2 = rs_lanes(ig_s)
    acc("P2 U/R", 24, [[None if l is None else comp(l[0]) + at(l[1], m, l[2]) for l in ls] for ls in l2 for m in range(n)])
    acc("P2 cof", 14, [[None if l is None else l[1] + n * m + n * n * l[2] for l in ls] for ls in l2 for m in range(n)])
    # P4: t-pencils (KS == 1): lanes (g, i, j): U read / R rmw at(i, j, m)
    tblk = r32(2 * n * n)
    lt = []
    for w0 in range(0, tblk, 32):
        lanes = []
        for r in range(w0, w0 + 32):
            g, p = r // (n * n), r % (n * n)
            lanes.append(None if g > 1 else (g, p % n, p // n))
        lt.append(lanes)
    acc("P4 U/R", 24, [[None if l is None else comp(l[0]) + at(l[1], l[2], m) for l in ls] for ls in lt for m in range(kb)])
    # P0 / P5: plane mapped: lanes (pi, pj) of plane kl of component c
    lp = [[(p % n, p // n) if p < n * n else None for p in range(w0, w0 + 32)] for w0 in range(0, r32(n * n), 32)]
    acc("P0/P5", 18, [[None if l is None else comp(c >= 3) + c % 3 * sc + at(l[0], l[1], kl) for l in ls]
                      for ls in lp for c in (0, 3) for kl in range(kb)])
    # flux: x faces (i = 0 / n-1, lanes over j then k), y faces (j = 0 / n-1, lanes over i then k)
    fx = []
    for side in (0, n - 1):
        pts = [(side, r % n, r // n) for r in range(n * kb)]
        fx += [pts[w0:w0 + 32] for w0 in range(0, len(pts), 32)]
        pts = [(r % n, side, r // n) for r in range(n * kb)]
        fx += [pts[w0:w0 + 32] for w0 in range(0, len(pts), 32)]
    acc("flux xy", 18 * 4.0 / n, [[at(*l) for l in ls] for ls in fx])
    tot = sum(w * f for w, f in res.values()) / sum(w for w, f in res.values())
    return tot, res, sc


if __name__ == "__main__":
    n = int(sys.argv[1])
    ks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    best = []
    for pj, pk, he in itertools.product(range(0, 4), range(0, 16), range(0, 16)):
        t, res, sc = model(n, ks, pj, pk, he)
        best.append((t, sc, pj, pk, he, res))
    best.sort(key=lambda x: (round(x[0], 3), x[1]))
    for t, sc, pj, pk, he, res in best[:6]:
        print(f"n={n} pad_j={pj} pad_k={pk} HE={he}: x{t:.3f}  SC={sc}  " +
              " ".join(f"{k}:{v[1]:.2f}" for k, v in res.items()))
    cur = {6: (3, 0), 14: (3, 0), 12: (1, 0), 3: (0, 3), 4: (0, 3), 7: (0, 3), 10: (0, 7)}.get(n, (0, 0))
    hecur = {6: 14, 7: 3, 8: 8, 9: 10, 10: 5, 11: 15, 13: 12, 14: 10, 15: 9}.get(n, 0)
    t, res, sc = model(n, ks, cur[0], cur[1], hecur)
    print(f"current pad_j={cur[0]} pad_k={cur[1]} HE={hecur}: x{t:.3f} SC={sc} " +
          " ".join(f"{k}:{v[1]:.2f}" for k, v in res.items()))
