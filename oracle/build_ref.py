#!/usr/bin/env python
"""Recipe for ``oracle/_ref/`` (TEST / BENCH INFRASTRUCTURE, never product code).

The reference is pure Python: its cross-correlation path is the single file
``xcorr/util/cc.py`` (numpy only).  "Building" it means placing that file,
UNMODIFIED, where the GPU box can load it -- ``/root/reference`` does not exist
there, ``oracle/_ref/`` travels with the snapshot (git-ignored, not
gpurun-ignored).  Run by ``__graft_entry__.build()`` whenever ``/root/reference``
is present; the copy is byte-identical (sha256 recorded next to it) and is never
committed.

    python oracle/build_ref.py          # -> oracle/_ref/xcorr_util_cc.py + MANIFEST.json
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("XCORR_REFERENCE", "/root/reference")
SOURCES = {"xcorr_util_cc.py": "xcorr/util/cc.py"}          # the hot path of SURVEY 8a (C1-C4)
OUT = os.path.join(HERE, "_ref")


def build():
    """Returns the manifest, or None when the reference tree is absent (GPU box:
    the prebuilt copy is used as is)."""
    if not os.path.isdir(REF_ROOT):
        return None
    os.makedirs(OUT, exist_ok=True)
    manifest = {}
    for name, rel in SOURCES.items():
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[name] = {"source": rel, "sha256": hashlib.sha256(f.read()).hexdigest()}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return manifest


def load():
    """The reference's own ``util/cc.py`` as a module, or None if ``oracle/_ref``
    was never built.  Only ``cc(..., unbiased=False)`` and ``lag`` are usable on numpy >= 1.24
    (``weight`` needs the removed ``np.int``, util/cc.py:137) -- which is all
    ``core/process.py:155-162`` calls."""
    path = os.path.join(OUT, "xcorr_util_cc.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("xcorr_reference_util_cc", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    m = build()
    print(json.dumps(m, indent=1) if m else f"{REF_ROOT} not present; kept the existing oracle/_ref")
