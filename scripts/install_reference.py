#!/usr/bin/env python
"""Copies the reference's own hot-path modules, UNMODIFIED, into git-ignored baseline/_ref so that
`bench.py --impl reference` can time the real thing on a box where /root/reference does not exist
(baseline/_ref is git-ignored but travels with gpurun and with the round-end snapshot).

    python scripts/install_reference.py [/root/reference]

Copied: util/fourier.py, util/resampling.py, util/timing.py (byte for byte).  Added next to them: an
empty `soundfile` stand-in (util/resampling.py imports it at module level; the functions the bench
calls -- np_rfft_pick, speed_to_pos, sinc_wrapper_mt -- never touch it) and MANIFEST.json with the
SHA-256 of every copied file.  The reference is pure Python (pip has nothing to build or install;
its requirements.txt pins nothing), so this copy IS the installation.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("util/fourier.py", "util/resampling.py", "util/timing.py")


def install(ref="/root/reference", dst=None):
    dst = dst or os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(ref):
        return None
    os.makedirs(os.path.join(dst, "util"), exist_ok=True)
    manifest = {"source": ref, "files": {}}
    for rel in FILES:
        src = os.path.join(ref, rel)
        shutil.copyfile(src, os.path.join(dst, rel))
        manifest["files"][rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(dst, "soundfile.py"), "w") as f:
        f.write('"""Stand-in for the soundfile package (not installed in this image): the reference imports it at module\n'
                'level, the benchmarked functions never use it."""\n\n\nclass SoundFile:\n    def __init__(self, *a, **k):\n'
                '        raise RuntimeError("soundfile is not available in the benchmark harness")\n')
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    return dst


if __name__ == "__main__":
    out = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("installed to", out if out else "(reference tree not found: nothing done)")
