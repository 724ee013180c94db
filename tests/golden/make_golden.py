"""Generates the mesh fixtures under tests/golden/ from the reference's own test inputs.

Run ONLY in the build container (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py

The fixtures hold INPUT data only (element corner coordinates, BC strings, the .map
partition codes and global vertex ids); the reference ships no golden OUTPUT arrays
(SURVEY.md section 4).  Formats followed:
  .re2  src/nek5_connect2.F:1717-1749 (header), :1820-1886 + :2196-2238 (mesh records),
        :1887-1969 (curve records), :1971-2069 + :2125-2164 (BC records)
  .rea  "**MESH DATA**" ASCII block (src/nek5_connect2.F:652-744 rdmesh2; BCs :842-1019)
  .map  src/nek5_connect11.F:892-1044 (7-int header; per element: partition code + 2^d
        global vertex ids)
"""
import os
import re
import struct

import numpy as np

REF = "/root/reference/tests"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_map(path, ncrn):
    with open(path) as f:
        hdr = [int(x) for x in f.readline().split()]
        nel = hdr[0]
        rows = [[int(x) for x in f.readline().split()] for _ in range(nel)]
    a = np.array(rows, dtype=np.int64)
    assert a.shape == (nel, 1 + ncrn)
    return a[:, 0].copy(), a[:, 1:].copy()


def read_re2(path):
    b = open(path, "rb").read()
    hdr = b[:80].decode()
    assert hdr.startswith("#v002"), hdr
    nel, ndim, nelv = int(hdr[5:14]), int(hdr[14:17]), int(hdr[17:27])
    (tag,) = struct.unpack("<f", b[80:84])
    assert abs(tag - 6.54321) < 1e-5, "byte-swapped re2 not handled"
    ncrn = 2 ** ndim
    nw = 1 + ndim * ncrn
    off = 84
    rec = np.frombuffer(b, dtype="<f8", count=nel * nw, offset=off).reshape(nel, nw)
    off += nel * nw * 8
    xc = rec[:, 1:1 + ncrn].copy()
    yc = rec[:, 1 + ncrn:1 + 2 * ncrn].copy()
    zc = rec[:, 1 + 2 * ncrn:1 + 3 * ncrn].copy() if ndim == 3 else np.zeros_like(xc)
    (ncurve,) = struct.unpack("<d", b[off:off + 8])
    off += 8 + int(ncurve) * 8 * 8
    assert int(ncurve) == 0, "curved sides not handled"
    (nbc,) = struct.unpack("<d", b[off:off + 8])
    off += 8
    cbc = [["E  "] * (2 * ndim) for _ in range(nel)]
    for _ in range(int(nbc)):
        vals = struct.unpack("<7d", b[off:off + 56])
        cb = b[off + 56:off + 59].decode()
        off += 64
        eg, f = int(vals[0]), int(vals[1])
        cbc[eg - 1][f - 1] = cb
    return ndim, xc, yc, zc, cbc


def read_rea_mesh(path):
    lines = open(path).read().splitlines()
    i = next(k for k, l in enumerate(lines) if "**MESH DATA**" in l)
    nel, ndim = int(lines[i + 1].split()[0]), int(lines[i + 1].split()[1])
    ncrn = 2 ** ndim
    xc = np.zeros((nel, ncrn)); yc = np.zeros((nel, ncrn)); zc = np.zeros((nel, ncrn))
    k = i + 2
    for e in range(nel):
        assert "ELEMENT" in lines[k]
        k += 1
        if ndim == 3:
            for half in range(2):
                xc[e, 4 * half:4 * half + 4] = [float(v) for v in lines[k].split()]
                yc[e, 4 * half:4 * half + 4] = [float(v) for v in lines[k + 1].split()]
                zc[e, 4 * half:4 * half + 4] = [float(v) for v in lines[k + 2].split()]
                k += 3
        else:
            xc[e] = [float(v) for v in lines[k].split()]
            yc[e] = [float(v) for v in lines[k + 1].split()]
            k += 2
    j = next(q for q, l in enumerate(lines) if "FLUID   BOUNDARY CONDITIONS" in l)
    cbc = [["E  "] * (2 * ndim) for _ in range(nel)]
    for q in range(nel * 2 * ndim):
        l = lines[j + 1 + q]
        cb = l[1:4]
        m = re.match(r"\s*(\d+)\s+(\d+)", l[4:])
        cbc[int(m.group(1)) - 1][int(m.group(2)) - 1] = cb
    return ndim, xc, yc, zc, cbc


def read_rea_curves(path, nel):
    """"***** CURVED SIDE DATA *****" block of a .rea (src/nek5_connect2.F rdcurve, nelgt < 1000:
    format (I3,I3,5G14.6,1X,A1)): per curved side IEDGE, IEL, CURVE(1:5), CCURVE.  Returns
    ccurve[nel][12] (1-char flags, ' ' = straight) and curve[nel,12,5]."""
    lines = open(path).read().splitlines()
    i = next(k for k, l in enumerate(lines) if "CURVED SIDE DATA" in l)
    ncurve = int(lines[i + 1].split()[0])
    ccurve = [[" "] * 12 for _ in range(nel)]
    curve = np.zeros((nel, 12, 5))
    for q in range(ncurve):
        t = lines[i + 2 + q].split()
        edge, e = int(t[0]), int(t[1])
        curve[e - 1, edge - 1] = [float(v) for v in t[2:7]]
        ccurve[e - 1][edge - 1] = t[7]
    return ccurve, curve


def save(name, ndim, xc, yc, zc, cbc, part, vertex, **extra):
    np.savez_compressed(os.path.join(OUT, name), ndim=ndim, xc=xc, yc=yc, zc=zc,
                        cbc=np.array(cbc), part=part, vertex=vertex, **extra)
    print(name, "nel", xc.shape[0], "bcs", sorted(set(np.array(cbc).ravel())))


if __name__ == "__main__":
    nd, xc, yc, zc, cbc = read_re2(f"{REF}/3dboxper/3dboxper.re2")
    part, vert = read_map(f"{REF}/3dboxper/3dboxper.map", 2 ** nd)
    save("3dboxper_mesh.npz", nd, xc, yc, zc, cbc, part, vert)
    nd, xc, yc, zc, cbc = read_rea_mesh(f"{REF}/3dboxpec/3dboxpec.rea")
    part, vert = read_map(f"{REF}/3dboxpec/3dboxpec.map", 2 ** nd)
    save("3dboxpec_mesh.npz", nd, xc, yc, zc, cbc, part, vert)
    # tests/cylwave: 50 hexahedra (5 layers of 10) in a cylinder of radius 4, circular-arc sides
    nd, xc, yc, zc, cbc = read_rea_mesh(f"{REF}/cylwave/cylwave.rea")
    part, vert = read_map(f"{REF}/cylwave/cylwave.map", 2 ** nd)
    ccurve, curve = read_rea_curves(f"{REF}/cylwave/cylwave.rea", xc.shape[0])
    save("cylwave_mesh.npz", nd, xc, yc, zc, cbc, part, vert, ccurve=np.array(ccurve),
         curve=curve)
