#!/usr/bin/env python
"""Install the UNMODIFIED reference (xyin-anl/MapsTorch) into the git-ignored ``baseline/_ref/``.

    python tools/install_reference.py [/root/reference]

``bench.py --impl reference`` imports ``mapstorch`` from there to time the reference's own CPU code on the
GPU box (``/root/reference`` itself does not travel with the repo snapshot, ``baseline/_ref`` does: it is
git-ignored but not gpurun-ignored).  The offline pip install needs a writable source tree (setuptools
writes ``*.egg-info`` / ``build``), so the read-only checkout is copied to a temporary directory first;
``--no-deps`` because h5py / marimo / matplotlib are absent from the image and not needed by the fit path
(``mapstorch.opt``, ``.map``, ``.bkg``, ``.default``, ``.constant`` import with torch, numpy, scipy, tqdm).
No reference source enters the git history.
"""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
TARGET = ROOT / "baseline" / "_ref"


def install(source="/root/reference", force=False):
    source = Path(source)
    if (TARGET / "mapstorch" / "opt.py").exists() and not force:
        return TARGET
    if not (source / "mapstorch" / "opt.py").exists():
        raise FileNotFoundError(f"{source} does not hold the reference checkout")
    if TARGET.exists():
        shutil.rmtree(TARGET)
    TARGET.parent.mkdir(parents=True, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = Path(tmp) / "src"
        shutil.copytree(source, work)
        for path in work.rglob("*"):
            path.chmod(path.stat().st_mode | 0o200)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
               "/opt/wheelhouse", "--no-deps", "--target", str(TARGET), str(work)]
        done = subprocess.run(cmd, capture_output=True, text=True)
        if done.returncode != 0:
            # pip unavailable / build backend missing: the package is pure Python, a plain copy is the same install
            shutil.copytree(work / "mapstorch", TARGET / "mapstorch", dirs_exist_ok=True)
            (TARGET / "INSTALL_NOTE.txt").write_text("pip install failed, package copied verbatim:\n" + done.stderr[-2000:])
    return TARGET


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    print(install(args[0] if args else "/root/reference", force="--force" in sys.argv))
