"""pytest configuration: markers, paths and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rnb-neus-fork_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cosine(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


# ---- parity margins: the tolerance-bearing asserts of the GPU tests go through check(), which records how far inside
# the bound the measured value is; the table is written at session end (gpurun_out/parity_margins.md on the GPU box)
# so every loosened tolerance has its measured error next to it (profiles/r02_parity_margins.md is a committed copy).
_MARGINS = []


def check(name, value, tol):
    value = float(value)
    _MARGINS.append((name, value, float(tol)))
    assert value < tol, (name, value, tol)


def pytest_sessionfinish(session, exitstatus):
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not _MARGINS or not os.path.isdir(out_dir):
        return
    lines = ["| check | measured | bound | measured / bound |", "|---|---|---|---|"]
    for name, v, t in _MARGINS:
        lines.append(f"| {name} | {v:.3e} | {t:.2e} | {v / t:.2f} |")
    with open(os.path.join(out_dir, "parity_margins.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
