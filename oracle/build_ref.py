"""Recipe for oracle/_ref: the UNMODIFIED reference package, copied file by file from the read-only checkout so that it
travels to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored).  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py          # in the build container, where /root/reference exists

The reference is pure Python (numpy / scipy / scikit-learn), so "building" it is copying `ssp_bayes_opt/`; nothing is
edited.  Its `__init__` imports `nengo` unconditionally (ssp_bayes_opt/__init__.py:15-16); `oracle/ref_loader.py` stubs that
module at import time, as SURVEY.md section 8c describes.  `MANIFEST.txt` records the sha256 of every file copied.
Used by: bench.py --impl reference / cpu_baseline (kind "reference"), and tests that compare the oracle restatement with
the live code when the copy is present.
"""
import hashlib
import os
import shutil
import sys

REF = os.environ.get("SSPBO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def build(verbose=True) -> bool:
    src = os.path.join(REF, "ssp_bayes_opt")
    if not os.path.isdir(src):
        if verbose:
            print(f"oracle/_ref: no reference checkout at {REF}; keeping whatever is there")
        return os.path.isdir(os.path.join(DST, "ssp_bayes_opt"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(src, os.path.join(DST, "ssp_bayes_opt"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    lines = []
    for root, _, files in os.walk(os.path.join(DST, "ssp_bayes_opt")):
        for f in sorted(files):
            p = os.path.join(root, f)
            lines.append(f"{hashlib.sha256(open(p, 'rb').read()).hexdigest()}  {os.path.relpath(p, DST)}")
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as fh:
        fh.write(f"# copied unmodified from {src}\n" + "\n".join(sorted(lines)) + "\n")
    if verbose:
        print(f"oracle/_ref: {len(lines)} files copied from {src}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
