#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — recipe that stages the UNMODIFIED reference under ``oracle/_ref/``.

    python oracle/make_ref.py            # no-op when /root/reference is absent (GPU box: uses the shipped copy)

The reference (115dkk/Impulcifer-pip313) is pure Python, so "building" it is a file copy: the packages the
BRIR pipeline imports (``core``, ``autoeq``, ``infra``, ``i18n``) and the bundled data files its estimator /
target stages open (``data/*.wav`` sweeps, ``data/*.csv`` targets).  Nothing is edited.  ``oracle/_ref/`` is
git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so the copy
travels to the GPU box where

  * ``tests/test_pipeline_dropin_gpu.py`` runs the reference's own ``BRIRPipeline`` twice — once as is (CPU),
    once with ``impulcifer_pip313_b200.install()`` active — and compares the WAV files it writes, and
  * ``bench.py --impl reference`` times the reference's own hot path under its own process pool
    (``core.parallel_utils.parallel_map``), ``cpu_baseline.kind = "reference"``.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s reference / cpu_baseline legs may touch ``oracle/``.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("IMPULCIFER_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

PACKAGES = ("core", "autoeq", "infra", "i18n")
DATA_GLOBS = (".wav", ".csv")
TOP_FILES = ("LICENSE", "pyproject.toml")


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "core", "pipeline.py"))


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "core")):
        print(f"make_ref: {SRC} not present; {'using the staged copy' if staged() else 'nothing staged'}")
        return 0
    tmp = DST + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache")
    for pkg in PACKAGES:
        shutil.copytree(os.path.join(SRC, pkg), os.path.join(tmp, pkg), ignore=ignore)
    os.makedirs(os.path.join(tmp, "data"))
    for name in sorted(os.listdir(os.path.join(SRC, "data"))):
        p = os.path.join(SRC, "data", name)
        if os.path.isfile(p) and name.endswith(DATA_GLOBS):
            shutil.copy2(p, os.path.join(tmp, "data", name))
    for name in TOP_FILES:
        p = os.path.join(SRC, name)
        if os.path.isfile(p):
            shutil.copy2(p, os.path.join(tmp, name))
    shutil.rmtree(DST, ignore_errors=True)
    os.rename(tmp, DST)
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    print(f"make_ref: staged {n} files of the reference under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
