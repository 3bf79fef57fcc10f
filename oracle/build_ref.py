"""Recipe for oracle/_ref: the UNMODIFIED reference package, runnable where /root/reference does not exist.

TEST INFRASTRUCTURE -- not product code.  hopwise is pure Python, so "building" the reference is a copy: the
`hopwise/` package is taken file by file from where it lies under /root/reference into oracle/_ref/hopwise
(git-ignored, so no reference source ever enters this repository's history, but not gpurun-ignored, so it travels to
the GPU box like the built .so files), together with three stub modules for logging-only dependencies that this
image lacks (colorama, colorlog, texttable: imported at hopwise/utils/logger.py:25-26 and utils/utils.py:29, used
only for coloured log lines and a FLOPs table).  oracle/_ref/MANIFEST.json records the sha256 of every copied file,
and `verify()` re-checks the copy against the source tree when that is present, so "unmodified" is checkable.

Who may use oracle/_ref: tests/, __graft_entry__.smoke() and bench.py's reference arm (`--impl reference` runs the
real hopwise KGTrainer on the host cores).  Nothing under hopwise_b200/ imports it.

    python oracle/build_ref.py            # build (no-op when up to date)
    python oracle/build_ref.py --verify   # compare oracle/_ref with /root/reference
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("HOPWISE_REFERENCE_ROOT", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")
MANIFEST = os.path.join(REF_DIR, "MANIFEST.json")

# the reference imports these at module level for coloured logging / a parameter table only
STUBS = {
    "colorama.py": '"""stub (oracle/build_ref.py): hopwise only calls colorama.init()."""\n\n\n'
                   "def init(*args, **kwargs):\n    return None\n",
    "colorlog.py": '"""stub (oracle/build_ref.py): hopwise only builds a ColoredFormatter for its stream handler."""\n'
                   "import logging\n\n\n"
                   "class ColoredFormatter(logging.Formatter):\n"
                   "    def __init__(self, fmt=None, datefmt=None, log_colors=None, **kwargs):\n"
                   "        super().__init__((fmt or '%(message)s').replace('%(log_color)s', ''), datefmt)\n",
    "texttable.py": '"""stub (oracle/build_ref.py): hopwise only uses Texttable in get_flops()."""\n\n\n'
                    "class Texttable:\n"
                    "    def __init__(self, *args, **kwargs):\n        self.rows = []\n\n"
                    "    def add_rows(self, rows, header=True):\n        self.rows += list(rows)\n\n"
                    "    def set_cols_align(self, *a):\n        pass\n\n"
                    "    def set_cols_valign(self, *a):\n        pass\n\n"
                    "    def draw(self):\n        return '\\n'.join(' | '.join(map(str, r)) for r in self.rows)\n",
}
SKIP_DIRS = {"__pycache__"}


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def source_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hopwise"))


def built() -> bool:
    return os.path.exists(MANIFEST) and os.path.isdir(os.path.join(REF_DIR, "hopwise"))


def _walk(root):
    for dirpath, dirs, files in os.walk(root):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        for f in sorted(files):
            if f.endswith((".pyc", ".pyo")):
                continue
            full = os.path.join(dirpath, f)
            yield os.path.relpath(full, root), full


def build(force: bool = False) -> str | None:
    """Copy the reference package into oracle/_ref.  Returns the directory, or None when neither the source tree nor
    a previous build exists (the GPU box only ever uses the copy that travelled with the snapshot)."""
    if not source_available():
        return REF_DIR if built() else None
    src_root = os.path.join(REFERENCE_ROOT, "hopwise")
    if built() and not force:
        try:
            old = json.load(open(MANIFEST))
            if all(os.path.getmtime(full) <= old.get("built_at", 0) for _, full in _walk(src_root)):
                return REF_DIR
        except Exception:
            pass
    dst_root = os.path.join(REF_DIR, "hopwise")
    if os.path.isdir(dst_root):
        shutil.rmtree(dst_root)
    files = {}
    for rel, full in _walk(src_root):
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(full, dst)
        files["hopwise/" + rel] = _sha(dst)
    for name, text in STUBS.items():
        with open(os.path.join(REF_DIR, name), "w") as f:
            f.write(text)
    import time

    json.dump({"source": src_root, "built_at": time.time(), "n_files": len(files), "sha256": files,
               "stubs": sorted(STUBS)}, open(MANIFEST, "w"), indent=0)
    return REF_DIR


def verify() -> int:
    """Number of files of oracle/_ref/hopwise that differ from the source tree (0 = an unmodified copy)."""
    man = json.load(open(MANIFEST))
    bad = 0
    for rel, digest in man["sha256"].items():
        here = os.path.join(REF_DIR, rel)
        if not os.path.exists(here) or _sha(here) != digest:
            bad += 1
            continue
        src = os.path.join(REFERENCE_ROOT, rel)
        if source_available() and (not os.path.exists(src) or _sha(src) != digest):
            bad += 1
    return bad


if __name__ == "__main__":
    if "--verify" in sys.argv:
        n = verify()
        print(f"{n} file(s) differ")
        sys.exit(1 if n else 0)
    print(build(force="--force" in sys.argv))
