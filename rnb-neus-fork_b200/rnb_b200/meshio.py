"""Mesh export for validate_mesh (reference exp_runner.py:567-578: vertices scaled by scale_mats_np, then
`trimesh.Trimesh(vertices, triangles).export('….ply')`; trimesh is not a dependency here).  Binary little-endian PLY
with float32 vertices and int32 triangle indices -- the layout trimesh writes."""
from __future__ import annotations

import numpy as np


def write_ply(path, vertices, triangles):
    v = np.ascontiguousarray(np.asarray(vertices), dtype="<f4").reshape(-1, 3)
    t = np.ascontiguousarray(np.asarray(triangles), dtype="<i4").reshape(-1, 3)
    if t.size and (t.min() < 0 or t.max() >= len(v)):
        raise ValueError("write_ply: triangle index out of range")
    header = ("ply\nformat binary_little_endian 1.0\n"
              f"element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {len(t)}\nproperty list uchar int vertex_indices\nend_header\n")
    faces = np.empty(len(t), dtype=[("n", "u1"), ("i", "<i4", (3,))])
    faces["n"] = 3
    faces["i"] = t
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(v.tobytes())
        f.write(faces.tobytes())


def read_ply(path):
    """Reads back what write_ply wrote (tests, round trips)."""
    with open(path, "rb") as f:
        nv = nf = None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                nv = int(line.split()[-1])
            elif line.startswith("element face"):
                nf = int(line.split()[-1])
            elif line == "end_header":
                break
        v = np.frombuffer(f.read(nv * 12), dtype="<f4").reshape(nv, 3)
        faces = np.frombuffer(f.read(nf * 13), dtype=[("n", "u1"), ("i", "<i4", (3,))])
    return v.copy(), faces["i"].copy()
