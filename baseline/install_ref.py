"""Put the UNMODIFIED reference where it can travel to the GPU box: baseline/_ref/ (git-ignored, not
gpurun-ignored), as BASELINE.md §3 plans.  The reference has no setup.py / pyproject, so there is nothing
for pip to install: its .py files are copied byte for byte (no edits, no re-formatting).  Run in the build
container (the only place /root/reference exists); __graft_entry__.build() calls install() when it can.

    python baseline/install_ref.py
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("PCV_REFERENCE", "/root/reference")


def install(force=False):
    """-> path of baseline/_ref, or None when the reference is not present here and no copy exists."""
    if os.path.isdir(DST) and os.path.exists(os.path.join(DST, "models", "pivotcvae.py")) and not force:
        return DST
    if not os.path.isdir(SRC):
        return None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".git")]
        for f in files:
            if f.endswith(".py") or f in ("LICENSE", "README.md"):
                rel = os.path.relpath(os.path.join(root, f), SRC)
                os.makedirs(os.path.dirname(os.path.join(DST, rel)) or DST, exist_ok=True)
                shutil.copyfile(os.path.join(root, f), os.path.join(DST, rel))
    return DST


if __name__ == "__main__":
    print(install(force=True))
