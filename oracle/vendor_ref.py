"""Recipe that makes the UNMODIFIED reference travel to the GPU box (TEST / BENCH INFRASTRUCTURE ONLY).

    python oracle/vendor_ref.py        # run in the build container, where /root/reference exists

Copies the reference's `models/` and `util/` Python packages byte for byte into `baseline/_ref/` -- a directory that is
git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so `bench.py --impl
reference` (the reference's own FCGANModel.optimize_parameters on the box's host cores) and `bench.py --impl eager_cuda`
(the same unmodified modules moved to the GPU: PyTorch eager + cuDNN, the competitive bar of SURVEY 8d) can import it on a
box where /root/reference does not exist.  `__graft_entry__.build()` runs this when /root/reference is present.
A MANIFEST with the sha256 of every copied file is written next to the copy so that "unmodified" can be checked.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SGK_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PACKAGES = ("models", "util")


def vendor(verbose=False):
    if not os.path.isfile(os.path.join(SRC, "models", "networks.py")):
        return False
    lines = []
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                src = os.path.join(root, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                lines.append("%s  %s" % (hashlib.sha256(open(src, "rb").read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(sorted(lines, key=lambda l: l.split("  ")[1])) + "\n")
    if verbose:
        print("vendored %d files of the reference into %s" % (len(lines), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor(verbose=True) else 1)
