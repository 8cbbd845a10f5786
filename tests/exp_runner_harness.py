"""TEST INFRASTRUCTURE: run the reference's exp_runner.py UNCHANGED on top of either `models` package.

    python tests/exp_runner_harness.py --models {reference|dropin} --ref-root DIR --work DIR --mode train_rnb [--is_continue]

What the harness supplies (nothing in exp_runner.py / models/dataset.py is edited or copied):
  * import shims for the four third-party modules this image lacks: `pyhocon` (a HOCON-subset parser that covers
    confs/*.conf), `trimesh` (Trimesh(...).export -> binary PLY), `mcubes` (marching_cubes -> this library's device
    extractor; PyMCubes 0.1.6 is what the reference pins, README.md:36) and `icecream`;
  * a recording `torch.utils.tensorboard.SummaryWriter` so the six scalars train_rnb logs per iteration
    (exp_runner.py:266-274) can be read back as JSON;
  * a tiny synthetic case directory in the layout models/dataset.py:130-170 loads (cameras.npz, mask/ normal/ albedo/
    PNG maps of an analytic sphere) and a conf derived from confs/wmask_rnb.conf with a short schedule;
  * the import path: `--models reference` puts the reference checkout first, `--models dropin` puts
    rnb-neus-fork_b200/ first and lets `models.dataset` fall through to the reference's own file.
Then `runpy.run_path(exp_runner.py, run_name="__main__")` with exp_runner's own command line.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rnb-neus-fork_b200")


# ------------------------------------------------------------------------------------------------ pyhocon subset
class ConfigTree(dict):
    def _walk(self, key, create=False):
        node = self
        parts = key.split(".")
        for p in parts[:-1]:
            if p not in node:
                if not create:
                    raise KeyError(key)
                dict.__setitem__(node, p, ConfigTree())
            node = dict.__getitem__(node, p)
        return node, parts[-1]

    def __getitem__(self, key):
        node, last = self._walk(key)
        return dict.__getitem__(node, last)

    def __setitem__(self, key, value):
        node, last = self._walk(key, create=True)
        dict.__setitem__(node, last, value)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def get_int(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else int(v)

    def get_float(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else float(v)

    def get_bool(self, key, default=None):
        v = self.get(key, default)
        if isinstance(v, str):
            return v.strip().lower() in ("true", "yes", "on")
        return None if v is None else bool(v)

    def get_string(self, key, default=None):
        v = self.get(key, default)
        return None if v is None else str(v)

    def get_list(self, key, default=None):
        return self.get(key, default)

    def get_config(self, key, default=None):
        return self.get(key, default)


def _scalar(tok):
    t = tok.strip().strip(",").strip()
    if len(t) >= 2 and t[0] == t[-1] and t[0] in "\"'":
        return t[1:-1]
    low = t.lower()
    if low in ("true", "false"):
        return low == "true"
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t)
    except ValueError:
        return t


def parse_hocon(text):
    """key = value | key { ... } | key = [a, b, ...] (possibly over several lines); `#` and `//` comments."""
    lines = []
    for raw in text.splitlines():
        line = re.sub(r"(#|//).*$", "", raw).strip()
        if line:
            lines.append(line)
    root = ConfigTree()
    stack = [root]
    i = 0
    while i < len(lines):
        line = lines[i]
        i += 1
        if line in ("}", "},"):
            stack.pop()
            continue
        m = re.match(r"^([A-Za-z0-9_.\-]+)\s*[=:]?\s*\{$", line)
        if m:
            child = ConfigTree()
            dict.__setitem__(stack[-1], m.group(1), child)
            stack.append(child)
            continue
        m = re.match(r"^([A-Za-z0-9_.\-]+)\s*[=:]\s*(.*)$", line)
        if not m:
            raise ValueError(f"hocon shim: cannot parse {line!r}")
        key, val = m.group(1), m.group(2).strip()
        if val.startswith("["):
            buf = val
            while "]" not in buf:
                buf += " " + lines[i]
                i += 1
            inner = buf[buf.index("[") + 1: buf.rindex("]")]
            dict.__setitem__(stack[-1], key, [_scalar(t) for t in inner.split(",") if t.strip()])
        else:
            dict.__setitem__(stack[-1], key, _scalar(val))
    return root


def install_shims(log_path):
    ph = types.ModuleType("pyhocon")

    class ConfigFactory:
        @staticmethod
        def parse_string(text):
            return parse_hocon(text)

        @staticmethod
        def parse_file(path):
            return parse_hocon(open(path).read())

    ph.ConfigFactory, ph.ConfigTree = ConfigFactory, ConfigTree
    sys.modules["pyhocon"] = ph

    ic = types.ModuleType("icecream")
    ic.ic = lambda *a, **k: None
    sys.modules["icecream"] = ic

    tm = types.ModuleType("trimesh")

    class Trimesh:
        def __init__(self, vertices=None, faces=None, vertex_colors=None, **kw):
            self.vertices, self.faces, self.vertex_colors = vertices, faces, vertex_colors

        def export(self, path):
            sys.path.insert(0, PKG)
            from rnb_b200 import meshio
            meshio.write_ply(path, self.vertices, self.faces)

    tm.Trimesh = Trimesh
    sys.modules["trimesh"] = tm

    mc = types.ModuleType("mcubes")

    def marching_cubes(u, threshold):
        import torch
        if PKG not in sys.path:
            sys.path.insert(0, PKG)
        from rnb_b200 import grid
        return grid.marching_cubes_device(torch.as_tensor(u, device="cuda"), threshold)

    mc.marching_cubes = marching_cubes
    sys.modules["mcubes"] = mc

    import torch.utils.tensorboard as tb
    records = []

    class SummaryWriter:
        def __init__(self, log_dir=None, **kw):
            self.log_dir = log_dir

        def add_scalar(self, tag, value, step=None, **kw):
            records.append((tag, float(value), int(step) if step is not None else -1))
            if tag == "Statistics/weight_max":          # the last of the six scalars of an iteration
                json.dump(records, open(log_path, "w"))

        def close(self):
            pass

    tb.SummaryWriter = SummaryWriter


# ------------------------------------------------------------------------------------------------ synthetic case
def write_case(case_dir, n_views=4, H=48, W=64, radius=0.6, cam_dist=3.0):
    """bearPNG-shaped case (reference README.md:63-79): cameras.npz + mask/ normal/ albedo/ 8-bit PNG maps of a sphere."""
    import cv2 as cv
    import numpy as np
    for d in ("mask", "normal", "albedo"):
        os.makedirs(os.path.join(case_dir, d), exist_ok=True)
    focal = 0.5 * W / math.tan(math.asin(radius / cam_dist) / 0.7)
    K = np.eye(4)
    K[0, 0] = K[1, 1] = focal
    K[0, 2], K[1, 2] = 0.5 * (W - 1), 0.5 * (H - 1)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    dirs_cam = np.stack([(xs - K[0, 2]) / focal, (ys - K[1, 2]) / focal, np.ones_like(xs)], -1)
    dirs_cam /= np.linalg.norm(dirs_cam, axis=-1, keepdims=True)
    cams = {}
    for v in range(n_views):
        phi = 2.0 * math.pi * v / n_views
        c = np.array([cam_dist * math.cos(phi) * math.cos(0.35), cam_dist * math.sin(phi) * math.cos(0.35), cam_dist * math.sin(0.35)])
        z = -c / np.linalg.norm(c)
        x = np.cross(z, np.array([0.0, 0.0, 1.0]))
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        pose = np.eye(4)
        pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = x, y, z, c          # camera -> world
        cams[f"world_mat_{v}"] = K @ np.linalg.inv(pose)
        cams[f"scale_mat_{v}"] = np.eye(4)
        R = pose[:3, :3]
        d = dirs_cam @ R.T
        b = d @ c
        disc = b * b - (c @ c - radius * radius)
        hit = disc > 0
        t = -b - np.sqrt(np.clip(disc, 0, None))
        p = c + d * t[..., None]
        n_world = np.where(hit[..., None], p / radius, 0.0)
        n_cam = n_world @ R                                                    # R^T n
        # models/dataset.py:57-64 decodes normal = 2 img - 1 with y and z flipped
        enc = (n_cam * np.array([1.0, -1.0, -1.0]) + 1.0) * 0.5
        enc = np.where(hit[..., None], enc, 0.5)
        alb = np.where(hit[..., None], 0.65 + 0.3 * np.sin(3.0 * p + np.array([0.0, 1.0, 2.0])), 0.0)
        to8 = lambda a: np.clip(np.round(a * 255.0), 0, 255).astype(np.uint8)
        cv.imwrite(os.path.join(case_dir, "mask", f"{v:03d}.png"), to8(hit.astype(np.float64)))
        cv.imwrite(os.path.join(case_dir, "normal", f"{v:03d}.png"), cv.cvtColor(to8(enc), cv.COLOR_RGB2BGR))
        cv.imwrite(os.path.join(case_dir, "albedo", f"{v:03d}.png"), cv.cvtColor(to8(alb), cv.COLOR_RGB2BGR))
    np.savez(os.path.join(case_dir, "cameras.npz"), **cams)


def write_conf(ref_root, path, exp_dir, case_dir, end_iter, warm_up_iter, batch_size, conf_name="wmask_rnb.conf"):
    text = open(os.path.join(ref_root, "confs", conf_name)).read()

    def setv(key, val):
        nonlocal text
        text, n = re.subn(rf"(^\s*{key}\s*=\s*)[^\n]*", lambda m: m.group(1) + str(val), text, count=1, flags=re.M)
        assert n == 1, key
    setv("base_exp_dir", exp_dir)
    setv("data_dir", case_dir + "/")
    setv("end_iter", end_iter)
    setv("warm_up_iter", warm_up_iter)
    setv("batch_size", batch_size)
    setv("save_freq", end_iter)
    setv("val_freq", 10 ** 9)
    setv("val_mesh_freq", 10 ** 9)
    setv("report_freq", max(1, end_iter // 2))
    open(path, "w").write(text)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", choices=["reference", "dropin"], required=True)
    ap.add_argument("--ref-root", required=True)
    ap.add_argument("--work", required=True)
    ap.add_argument("--mode", default="train_rnb")
    ap.add_argument("--conf", default="wmask_rnb.conf")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warm-up-iter", type=int, default=10)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--is_continue", action="store_true")
    ap.add_argument("--no_albedo", action="store_true")
    a = ap.parse_args()
    ref_root = os.path.abspath(a.ref_root)
    work = os.path.abspath(a.work)
    case_dir = os.path.join(work, "data", "sphere")
    exp_dir = os.path.join(work, "exp_" + a.models)
    os.makedirs(exp_dir, exist_ok=True)
    if not os.path.isfile(os.path.join(case_dir, "cameras.npz")):
        write_case(case_dir)
    conf_path = os.path.join(work, f"{a.models}_{a.conf}")
    write_conf(ref_root, conf_path, exp_dir, case_dir, a.iters, a.warm_up_iter, a.batch, a.conf)
    install_shims(os.path.join(exp_dir, f"scalars_{a.mode}.json"))
    # import path: who provides `models.fields` / `models.renderer`
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (ROOT, PKG, os.path.join(ROOT, "tests"))]
    if a.models == "reference":
        sys.path.insert(0, ref_root)
    else:
        sys.path.insert(0, ref_root)
        sys.path.insert(0, PKG)
        import models                                      # the drop-in package ...
        assert os.path.abspath(os.path.dirname(models.__file__)) == os.path.join(PKG, "models")
        models.__path__.append(os.path.join(ref_root, "models"))     # ... `models.dataset` falls through to the reference's file
    import torch
    torch.manual_seed(0)                                   # exp_runner does not seed the network initialisation
    os.chdir(ref_root)                                     # `general.recording = [./, ./models]` is relative to the checkout
    argv = [os.path.join(ref_root, "exp_runner.py"), "--mode", a.mode, "--conf", conf_path, "--case", "sphere"]
    if a.is_continue:
        argv.append("--is_continue")
    if a.no_albedo:
        argv.append("--no_albedo")
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")
    import models.fields as mf
    print("HARNESS_DONE models.fields from", os.path.abspath(mf.__file__))


if __name__ == "__main__":
    main()
