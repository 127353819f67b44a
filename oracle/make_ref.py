#!/usr/bin/env python
"""Vendor the UNMODIFIED Python reference's hot-path packages into oracle/_ref/src (git-ignored,
NOT gpurun-ignored: it travels to the GPU box like a built .so, where /root/reference does not
exist).  TEST / BASELINE INFRASTRUCTURE ONLY: bench.py's `--impl reference` arm and `cpu_baseline`
leg time it; nothing under sdrterm_b200/ may import it.  The reference has no compiled sources, so
"building" it is a byte-for-byte copy of src/dsp and src/misc plus a manifest of their hashes.

    python oracle/make_ref.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def make(ref_root: str = '/root/reference') -> str | None:
    src = os.path.join(ref_root, 'src')
    if not os.path.isdir(os.path.join(src, 'dsp')):
        return None
    dst = os.path.join(HERE, '_ref', 'src')
    manifest = {}
    for pkg in ('dsp', 'misc'):
        d = os.path.join(dst, pkg)
        if os.path.isdir(d):
            shutil.rmtree(d)
        shutil.copytree(os.path.join(src, pkg), d, ignore=shutil.ignore_patterns('__pycache__'))
        for root, _, files in os.walk(d):
            for f in sorted(files):
                p = os.path.join(root, f)
                manifest[os.path.relpath(p, dst)] = hashlib.sha256(open(p, 'rb').read()).hexdigest()
    with open(os.path.join(HERE, '_ref', 'MANIFEST.json'), 'w') as fh:
        json.dump({'source': ref_root, 'files': manifest}, fh, indent=1)
    return dst


if __name__ == '__main__':
    out = make(*sys.argv[1:2])
    print(out or 'reference tree not found: nothing vendored')
