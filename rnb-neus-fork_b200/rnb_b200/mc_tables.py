"""Marching-cubes case tables, generated (not transcribed) at import time.

For each of the 256 sign configurations of a cell the isosurface polygons are found by tracing the cell's six faces:
a face whose corner signs change across two of its edges contributes one segment between the two crossings; a face
with four crossings (signs alternate around it) is ambiguous and is resolved by ONE rule that depends only on the face's
own corner signs -- every inside corner is cut off separately -- so two cells sharing a face always make the same
choice and the extracted surface is watertight.  Segments link into closed loops (every crossed cube edge lies on exactly
two faces), each loop is fanned into triangles and oriented so that normals point from the inside (field > threshold)
to the outside.  Vertices sit on cube edges, like PyMCubes / Lorensen-Cline; the triangulation of a polygon may differ
from a hand-made table, the surface does not.

Conventions: corner i has offsets (i & 1, (i >> 1) & 1, (i >> 2) & 1) along (x, y, z); edge e joins EDGE_CORNERS[e];
case index bit i is set when corner i is inside.  TRI_TABLE[case] = up to 5 triangles as edge triples, padded with -1.
"""
from __future__ import annotations

import numpy as np

CORNER_OFFSETS = np.array([[(i >> a) & 1 for a in range(3)] for i in range(8)], dtype=np.int64)
EDGE_CORNERS = []
for _axis in range(3):
    for _c in range(8):
        if not (_c >> _axis) & 1:
            EDGE_CORNERS.append((_c, _c | (1 << _axis)))
EDGE_CORNERS = np.array(EDGE_CORNERS, dtype=np.int64)            # [12, 2]; edges 4a .. 4a+3 run along axis a
_EDGE_ID = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGE_CORNERS)}


def _edge(a, b):
    return _EDGE_ID[(min(a, b), max(a, b))]


def _faces():
    """the six faces as 4 corners in cyclic order"""
    out = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in range(2):
            base = side << axis
            out.append([base, base | (1 << u), base | (1 << u) | (1 << v), base | (1 << v)])
    return out


FACES = _faces()


def _case_loops(case):
    inside = [(case >> i) & 1 for i in range(8)]
    adj = {}

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)

    for f in FACES:
        s = [inside[c] for c in f]
        crossings = [k for k in range(4) if s[k] != s[(k + 1) % 4]]          # crossing k lies on edge (f[k], f[k+1])
        if len(crossings) == 2:
            k0, k1 = crossings
            link(_edge(f[k0], f[(k0 + 1) % 4]), _edge(f[k1], f[(k1 + 1) % 4]))
        elif len(crossings) == 4:
            for k in range(4):
                if s[k]:                                                       # cut off every inside corner
                    link(_edge(f[(k - 1) % 4], f[k]), _edge(f[k], f[(k + 1) % 4]))
    loops, seen = [], set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [n for n in adj[cur]]
            assert len(nxt) == 2, (case, cur, nxt)
            # walk along the neighbour we did not come from (the two entries may be equal only for degenerate 2-cycles)
            step = nxt[0] if nxt[0] != prev else nxt[1]
            if prev is None:
                step = nxt[0]
            if step == start:
                break
            loop.append(step)
            seen.add(step)
            prev, cur = cur, step
        loops.append(loop)
    return loops, inside


def _build():
    mid = (CORNER_OFFSETS[EDGE_CORNERS[:, 0]] + CORNER_OFFSETS[EDGE_CORNERS[:, 1]]) * 0.5
    tri = -np.ones((256, 16), dtype=np.int8)
    count = np.zeros(256, dtype=np.int32)
    for case in range(256):
        loops, inside = _case_loops(case)
        out = []
        for loop in loops:
            assert len(loop) >= 3, (case, loop)
            pts = mid[loop]
            ctr = pts.mean(0)
            n = np.zeros(3)
            for k in range(len(loop)):
                n += np.cross(pts[k] - ctr, pts[(k + 1) % len(loop)] - ctr)
            # desired direction: from the inside corner to the outside corner of every crossed edge
            want = np.zeros(3)
            for e in loop:
                a, b = EDGE_CORNERS[e]
                d = (CORNER_OFFSETS[b] - CORNER_OFFSETS[a]).astype(float)
                want += d if inside[a] else -d
            if np.dot(n, want) < 0:
                loop = loop[::-1]
            for k in range(1, len(loop) - 1):
                out += [loop[0], loop[k], loop[k + 1]]
        assert len(out) <= 15, (case, len(out))
        tri[case, :len(out)] = out
        count[case] = len(out) // 3
    return tri, count


TRI_TABLE, TRI_COUNT = _build()
