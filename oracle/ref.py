"""Import the unmodified reference (hopwise) from oracle/_ref -- TEST INFRASTRUCTURE, see oracle/build_ref.py.

oracle/_ref is built in the build container from /root/reference and travels to the GPU box with the snapshot;
`import_ref()` puts it (package + the three logging stubs) first on sys.path.  Used by tests/, smoke() and
bench.py --impl reference only.
"""

from __future__ import annotations

import os
import sys

from . import build_ref

REF_DIR = build_ref.REF_DIR


def ref_available() -> bool:
    return build_ref.built() or build_ref.source_available()


def import_ref():
    """Returns the `hopwise` module of oracle/_ref (building the copy first where the source tree exists)."""
    if not build_ref.built():
        if build_ref.build() is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is not present")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import hopwise

    if not os.path.abspath(hopwise.__file__).startswith(os.path.abspath(REF_DIR)):
        raise RuntimeError(f"another hopwise is already imported from {hopwise.__file__}")
    return hopwise
