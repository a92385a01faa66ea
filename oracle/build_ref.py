"""Recipe for ``oracle/_ref``: an UNMODIFIED copy of the reference's own hot-path modules (TEST / BASELINE
INFRASTRUCTURE ONLY).

    python oracle/build_ref.py          # needs /root/reference; writes oracle/_ref/ (git-ignored)

The reference is pure Python (no build step): the "build" is a byte-for-byte copy of the files its training
step imports -- ``mpgan/``, ``gapt/``, ``train.py``, ``setup_training.py``, ``plotting.py`` (imported by
train.py) and the argument file of the published ``mp_g`` model -- into ``oracle/_ref/``, which is listed in
.gitignore (reference sources never enter this repo's history) but NOT in .gpurunignore, so the copy travels to
the GPU box like the built ``.so``.  ``oracle/ref_loader.py`` imports it (with the ``jetnet`` / ``matplotlib``
packages it never calls on this path stubbed out); ``bench.py --impl reference`` and the ``gpu_eager`` baseline
time it, and ``MANIFEST.json`` records the sha256 of every copied file next to the source's so "unmodified" can
be checked.
"""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("MPGAN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["train.py", "setup_training.py", "plotting.py", "gen.py", "trained_models/mp_g/args.txt"]
DIRS = ["mpgan", "gapt"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=True) -> bool:
    """Copies the files; returns False (and leaves an existing copy alone) when the reference is not present."""
    if not os.path.isdir(REF):
        if verbose:
            print(f"build_ref: {REF} not present; keeping the existing oracle/_ref "
                  f"({'found' if os.path.isdir(DST) else 'absent'})")
        return False
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    todo = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(REF, d))):
            if name.endswith(".py"):
                todo.append(os.path.join(d, name))
    for rel in todo:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {"sha256": _sha(dst), "source_sha256": _sha(src)}
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"build_ref: copied {len(todo)} files into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() or os.path.isdir(DST) else 1)
