"""TEST / MEASUREMENT INFRASTRUCTURE: make the unmodified reference travel to the GPU box.

    python oracle/stage_reference.py          # /root/reference -> baseline/_ref/RNb-NeuS-fork  (git-ignored)

`/root/reference` exists only in the build container; `baseline/_ref/` is git-ignored but NOT gpurun-ignored, so a staged
copy rides along with the snapshot (the mechanism the bench contract gives the reference arm; the reference is a
directory of scripts with no setup.py, so "pip install --target baseline/_ref" has nothing to install and the tree is
copied as it is).  On the GPU box it lets
  * tests/test_exp_runner_smoke.py run the reference's own exp_runner.py, unchanged, on both `models` packages,
  * bench.py time the REAL reference on the host cores (cpu_baseline.kind = "reference") and through PyTorch-CUDA
    (cuda_reference.kind = "reference") instead of oracle/torch_port.py.
Nothing under baseline/_ref is imported by the product path; oracle/ref_loader.py is the only door to it.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("RNB_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref", "RNb-NeuS-fork")


def main():
    if not os.path.isfile(os.path.join(SRC, "exp_runner.py")):
        print(f"stage_reference: no reference checkout at {SRC}; nothing staged")
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    keep = lambda d, names: [n for n in names if n in (".git", "assets", "logs", "__pycache__")]
    shutil.copytree(SRC, DST, ignore=keep)
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f"staged {n} files of {SRC} under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
