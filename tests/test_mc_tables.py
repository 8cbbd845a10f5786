"""CPU checks of the generated marching-cubes tables (rnb_b200/mc_tables.py)."""
import numpy as np

from rnb_b200 import mc_tables as M


def test_table_shape_and_symmetry():
    assert M.TRI_TABLE.shape == (256, 16) and M.TRI_COUNT.max() == 5
    assert M.TRI_COUNT[0] == 0 and M.TRI_COUNT[255] == 0
    for case in range(256):
        used = M.TRI_TABLE[case][M.TRI_TABLE[case] >= 0]
        assert len(used) == 3 * M.TRI_COUNT[case]
        # every vertex lies on an edge whose corners have different signs
        for e in used:
            a, b = M.EDGE_CORNERS[e]
            assert ((case >> a) & 1) != ((case >> b) & 1)
        # the crossed edges are exactly the vertices used
        crossed = {e for e, (a, b) in enumerate(M.EDGE_CORNERS) if ((case >> a) & 1) != ((case >> b) & 1)}
        assert set(int(e) for e in used) == crossed


def test_cpu_extraction_is_watertight():
    from collections import Counter
    n = 14
    g = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    for u, chi in ((0.6 - np.sqrt(X ** 2 + Y ** 2 + Z ** 2), 2), (0.28 - np.sqrt((np.sqrt(X ** 2 + Y ** 2) - 0.6) ** 2 + Z ** 2), 0)):
        verts, tris = {}, []
        for x in range(n - 1):
            for y in range(n - 1):
                for z in range(n - 1):
                    case = sum(1 << i for i, (ox, oy, oz) in enumerate(M.CORNER_OFFSETS) if u[x + ox, y + oy, z + oz] > 0)
                    t = M.TRI_TABLE[case]
                    for k in range(0, 3 * M.TRI_COUNT[case], 3):
                        ids = []
                        for e in t[k:k + 3]:
                            a, b = M.EDGE_CORNERS[e]
                            key = (tuple(np.array([x, y, z]) + M.CORNER_OFFSETS[a]), tuple(np.array([x, y, z]) + M.CORNER_OFFSETS[b]))
                            ids.append(verts.setdefault(key, len(verts)))
                        tris.append(ids)
        de = Counter()
        for a, b, c in tris:
            for e in ((a, b), (b, c), (c, a)):
                de[e] += 1
        und = Counter(tuple(sorted(e)) for e in de)
        assert max(de.values()) == 1 and set(und.values()) == {2}
        assert len(verts) - len(und) + len(tris) == chi
