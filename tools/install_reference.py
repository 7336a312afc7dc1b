#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with the snapshot).

    python tools/install_reference.py

Used only by measurement / golden-generation tools (tests/golden/make_golden_kmeans.py, bench.py's `ref_gpu`
section): the reference's own Triton k-means and its FlashInfer launcher must run on a B200, and
/root/reference does not exist there.  Nothing in the product package imports baseline/_ref.

1. `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target baseline/_ref
   /root/reference` (the contract's recipe).  In this image it fails: the build backend `hatchling` is not in
   the wheelhouse.
2. Fallback = what that wheel would contain for the hot path: the pure-Python package tree `svg/` minus the
   git submodules (`svg/kernels/3rdparty`, 339 MB) and the deprecated `*_orig` model copies, copied verbatim.
"""
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"


def main():
    if not (REF / "svg").is_dir():
        print("reference checkout not mounted; nothing to do")
        return 0
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                        "/opt/wheelhouse", "--no-deps", "--target", str(DST), str(REF)], capture_output=True, text=True)
    how = "pip"
    if r.returncode != 0 or not (DST / "svg").is_dir():
        how = "copy (pip failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:120] + ")"
        shutil.copytree(REF / "svg", DST / "svg",
                        ignore=shutil.ignore_patterns("3rdparty", "*_orig", "__pycache__", "*.so", "csrc", "include"))
    (DST / "INSTALL.txt").write_text(f"installed from {REF} by tools/install_reference.py via {how}\n")
    n = sum(1 for _ in (DST / "svg").rglob("*.py"))
    print(f"baseline/_ref: {n} python files via {how}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
