"""The reference's own exp_runner.py, UNCHANGED, on the drop-in `models` package (north_star: "so exp_runner.py train_rnb and
validate_mesh run unchanged").  tests/exp_runner_harness.py supplies only what the image lacks (pyhocon /
trimesh / mcubes / icecream shims, a recording SummaryWriter, a synthetic case directory).  Needs the reference checkout:
/root/reference in the build container, or the copy oracle/stage_reference.py stages under baseline/_ref for the GPU box;
skipped when neither exists.

  1. `--mode train_rnb` for 20 iterations (10 warm-up + 10 regular, 512 rays, confs/wmask_rnb.conf) with the REFERENCE's
     models package and with the drop-in: the per-iteration losses must agree to 1e-2 (same seeds: exp_runner reseeds
     every iteration, exp_runner.py:170), both runs end with the checkpoint and the 512^3 mesh exp_runner writes;
  2. `--mode validate_mesh --is_continue` on the drop-in: resumes from the checkpoint the REFERENCE run wrote
     (checkpoints are interchangeable) and extracts the mesh again.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "tests", "exp_runner_harness.py")


def run(work, models, mode, *extra):
    t0 = time.time()
    p = subprocess.run([sys.executable, HARNESS, "--models", models, "--ref-root", ref_loader.REFERENCE_ROOT, "--work", work,
                        "--mode", mode, *extra], capture_output=True, text=True, timeout=1500)
    tail = "\n".join((p.stdout + "\n" + p.stderr).strip().splitlines()[-25:])
    assert p.returncode == 0, f"exp_runner ({models}, {mode}) failed:\n{tail}"
    assert "HARNESS_DONE" in p.stdout
    return time.time() - t0, p.stdout


def scalars(work, models, mode="train_rnb"):
    rec = json.load(open(os.path.join(work, "exp_" + models, f"scalars_{mode}.json")))
    out = {}
    for tag, val, step in rec:
        out.setdefault(tag, {})[step] = val
    return out


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present (run oracle/stage_reference.py)")
def test_exp_runner_unchanged_on_reference_and_dropin(tmp_path):
    work = str(tmp_path)
    t_ref, out_ref = run(work, "reference", "train_rnb")
    assert os.path.join(ref_loader.REFERENCE_ROOT, "models", "fields.py") in out_ref
    t_new, out_new = run(work, "dropin", "train_rnb")
    assert os.path.join(ROOT, "rnb-neus-fork_b200", "models", "fields.py") in out_new
    a, b = scalars(work, "reference"), scalars(work, "dropin")
    lines = [f"exp_runner.py --mode train_rnb, 20 iterations x 512 rays + validate_mesh(512^3): reference models {t_ref:.1f} s, "
             f"drop-in {t_new:.1f} s (wall, incl. interpreter start and data loading)",
             "iter  loss(reference)  loss(drop-in)   rel.diff   eikonal(ref)  eikonal(drop-in)"]
    worst = 0.0
    for step in sorted(a["Loss/loss"]):
        la, lb = a["Loss/loss"][step], b["Loss/loss"][step]
        rel = abs(lb - la) / abs(la)
        worst = max(worst, rel)
        lines.append(f"{step:4d}  {la:14.6f}  {lb:14.6f}  {rel:9.2e}  {a['Loss/eikonal_loss'][step]:12.6f}  {b['Loss/eikonal_loss'][step]:12.6f}")
    assert len(a["Loss/loss"]) == 20 and len(b["Loss/loss"]) == 20
    for tag in ("Loss/color_loss", "Loss/eikonal_loss", "Statistics/s_val", "Statistics/cdf", "Statistics/weight_max"):
        assert set(a[tag]) == set(b[tag]), tag            # the six TensorBoard reads of exp_runner.py:266-274
    assert worst < 1e-2, "\n".join(lines)
    # both runs wrote what exp_runner writes: the checkpoint at end_iter and the mesh of the final validate_mesh
    from rnb_b200 import meshio
    meshes = {}
    for m in ("reference", "dropin"):
        exp = os.path.join(work, "exp_" + m)
        assert os.path.isfile(os.path.join(exp, "checkpoints", "ckpt_000020.pth")), m
        v, t = meshio.read_ply(os.path.join(exp, "meshes", "00000020.ply"))
        meshes[m] = (v, t)
        r = np.linalg.norm(v, axis=1)
        lines.append(f"{m}: mesh {len(v)} vertices / {len(t)} triangles, mean radius {r.mean():.4f} (geometric init: 0.5)")
        assert len(t) > 1000 and abs(r.mean() - 0.5) < 0.05
    nv = [len(meshes[m][0]) for m in ("reference", "dropin")]
    assert abs(nv[0] - nv[1]) / nv[0] < 0.02
    # ---- checkpoints are interchangeable: the drop-in resumes from the REFERENCE run's checkpoint and meshes it
    import shutil
    shutil.rmtree(os.path.join(work, "exp_dropin", "checkpoints"))
    shutil.copytree(os.path.join(work, "exp_reference", "checkpoints"), os.path.join(work, "exp_dropin", "checkpoints"))
    os.remove(os.path.join(work, "exp_dropin", "meshes", "00000020.ply"))
    t_mesh, _ = run(work, "dropin", "validate_mesh", "--is_continue")
    v2, t2 = meshio.read_ply(os.path.join(work, "exp_dropin", "meshes", "00000020.ply"))
    lines.append(f"drop-in --mode validate_mesh --is_continue on the reference's checkpoint: {len(v2)} vertices in {t_mesh:.1f} s wall")
    assert abs(len(v2) - nv[0]) / nv[0] < 0.02
    lines.append(f"worst relative loss difference over 20 iterations: {worst:.2e}")
    log = "\n".join(lines)
    print(log)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        open(os.path.join(out_dir, "r02_exp_runner_smoke.log"), "w").write(log + "\n")
