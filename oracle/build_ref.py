"""Recipe for `oracle/_ref/`: the UNMODIFIED reference package, so that it can travel.

    python -m oracle.build_ref            # dev container only (/root/reference must exist)

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python: "building" it
means copying the source files of `/root/reference/ppgs` (every `*.py`, byte for byte, no edits)
to `oracle/_ref/ppgs/`.  `oracle/_ref/` is git-ignored (reference sources never enter this
repository's history) but NOT gpurun-ignored, so the copy ships to the GPU box with the
snapshot, where `/root/reference` does not exist.  Consumers: `bench.py --impl reference`
(`cpu_baseline.kind = "reference"`) and `bench.py`'s `torch_gpu_baseline` leg, through
`oracle/ref_arm.py` + `oracle/refshim.py` (stand-ins for the reference's absent third-party
imports).  The product package never imports it.
"""
import hashlib
import os
import shutil
import sys

SOURCE = os.environ.get('PPGS_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, '_ref')


def build(verbose=True):
    src = os.path.join(SOURCE, 'ppgs')
    if not os.path.isdir(src):
        if verbose:
            print(f'oracle/build_ref: {src} not found; keeping the existing oracle/_ref (if any)')
        return False
    dst = os.path.join(TARGET, 'ppgs')
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    digest, count = hashlib.sha256(), 0
    for root, dirs, files in os.walk(src):
        dirs.sort()
        for name in sorted(files):
            if not name.endswith('.py'):
                continue   # assets (30 MB of partition lists, notebooks, audio) are not on the path
            path = os.path.join(root, name)
            rel = os.path.relpath(path, src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(path, out)
            digest.update(rel.encode())
            digest.update(open(path, 'rb').read())
            count += 1
    with open(os.path.join(TARGET, 'MANIFEST.txt'), 'w') as f:
        f.write(f'source {src}\nfiles {count}\nsha256 {digest.hexdigest()}\n')
    if verbose:
        print(f'oracle/build_ref: copied {count} files to {dst} (sha256 {digest.hexdigest()[:16]})')
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
