#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference, made runnable where /root/reference does not exist.

    python oracle/make_ref.py            (run in the build container; __graft_entry__.build() calls it)

TEST INFRASTRUCTURE.  The reference is pure Python, so "building" it is copying its package directory
`/root/reference/prospect` to `oracle/_ref/prospect` (byte for byte; a manifest with the sha256 of every file is
written next to it).  `oracle/_ref/` is git-ignored -- no reference source enters the history -- but it is not
gpurun-ignored, so it travels to the GPU box like the built .so files.  Users: `bench.py --impl reference` (the CPU
arm: the reference's own Scheduler in `mode: threaded` on all host cores), the `cpu_baseline` leg, and the tests
that drive the reference's own tasks over the CUDA kernel class (tests/test_gpu_dropin.py).  Nothing under
prospect_public_b200/ imports it.

The two packages the reference imports but this image lacks (getdist, matplotlib; no network) are served by the
import shims under tests/golden/_shims -- they are ours, and committed.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/prospect'
DST = os.path.join(HERE, '_ref')


def build(force=False):
    """Copy the reference package into oracle/_ref/ (no-op when the source tree is absent, e.g. on the GPU box).
    Returns the path of the copy or None."""
    target = os.path.join(DST, 'prospect')
    if not os.path.isdir(SRC):
        return target if os.path.isdir(target) else None
    if os.path.isdir(target) and not force:
        return target
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(SRC, target, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    manifest = {}
    for root, _, files in os.walk(target):
        for fn in sorted(files):
            p = os.path.join(root, fn)
            with open(p, 'rb') as f:
                manifest[os.path.relpath(p, DST)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': SRC, 'files': manifest}, f, indent=1)
    return target


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
