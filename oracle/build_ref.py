#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the reference's OWN CPU path, compiled where it lies.  TEST / BENCH
INFRASTRUCTURE ONLY.

The reference has no native CPU kernel for this path (OPS/src/cpu/ms_deform_attn_cpu.cpp:22-45
raises); what it executes on a CPU is ``ms_deform_attn_core_pytorch``
(OPS/functions/ms_deform_attn_func.py:52-72, OPS = connectomics/model/maskformer_block/
pixel_decoder/ops).  That file is Python, so "compiling the reference from its own sources" means
byte-compiling it: this script runs ``py_compile`` on the file under ``/root/reference`` and writes
only the resulting ``.pyc`` (a binary, like a ``.so``) plus a small provenance record into the
git-ignored ``oracle/_ref/``.  No reference source is copied into the repository.

``/root/reference`` exists in the authoring container only; the built ``oracle/_ref/`` travels to
the GPU box with the snapshot, where ``bench.py --impl reference`` and ``cpu_baseline`` execute
it (``kind: "reference"``) and ``tests/test_oracle.py`` checks the in-repo restatement against it.
``__graft_entry__.build()`` calls ``build()`` when the reference is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PCTRANS_REFERENCE", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "connectomics", "model", "maskformer_block", "pixel_decoder", "ops",
                        "functions", "ms_deform_attn_func.py")
OUT_DIR = os.path.join(HERE, "_ref")
# not named *.pyc: snapshot tools treat that suffix as interpreter junk and drop it
PYC = os.path.join(OUT_DIR, "ms_deform_attn_func.bytecode")
META = os.path.join(OUT_DIR, "provenance.json")


def build(force: bool = False):
    """Returns the path of the byte-compiled reference module, or None when the reference tree is
    not available here (then an existing oracle/_ref/ is left as it is)."""
    if not os.path.exists(REF_FILE):
        return PYC if os.path.exists(PYC) else None
    with open(REF_FILE, "rb") as f:
        digest = hashlib.sha256(f.read()).hexdigest()
    tag = f"{sys.version_info.major}.{sys.version_info.minor}"
    try:
        with open(META) as f:
            meta = json.load(f)
        if not force and os.path.exists(PYC) and meta.get("sha256") == digest and meta.get("python") == tag:
            return PYC
    except (OSError, ValueError):
        pass
    os.makedirs(OUT_DIR, exist_ok=True)
    py_compile.compile(REF_FILE, cfile=PYC, doraise=True)
    with open(META, "w") as f:
        json.dump({"source": REF_FILE, "sha256": digest, "python": tag,
                   "what": "py_compile of the reference's ms_deform_attn_func.py (bytecode only)"}, f, indent=1)
    return PYC


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
