"""Stage the UNMODIFIED reference's Python sources for the hot path into ``oracle/_ref/`` (test infrastructure).

The reference is pure Python, so "building" it is copying its two packages (``design_opt``, ``khrylib``: .py files and
the .yml configs) from ``/root/reference`` -- no file is edited.  ``oracle/_ref/`` is git-ignored (it never enters the
history) but travels with gpurun snapshots, so the GPU box can time the real reference (``bench.py --impl reference``)
although ``/root/reference`` does not exist there.  Run by ``__graft_entry__.build()`` when the reference tree is present.

    python oracle/stage_ref.py [--src /root/reference]
"""
import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
KEEP = ('.py', '.yml', '.yaml')


def stage(src: str = '/root/reference', dst: str = DST) -> int:
    if not os.path.isdir(os.path.join(src, 'design_opt')):
        return 0
    n = 0
    for pkg in ('design_opt', 'khrylib'):
        for root, dirs, files in os.walk(os.path.join(src, pkg)):
            dirs[:] = [d for d in dirs if d not in ('__pycache__', 'assets')]
            rel = os.path.relpath(root, src)
            for f in files:
                if f.endswith(KEEP):
                    os.makedirs(os.path.join(dst, rel), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(dst, rel, f))
                    n += 1
    with open(os.path.join(dst, 'STAGED_FROM'), 'w') as fh:
        fh.write(f'{src}\n{n} files copied verbatim by oracle/stage_ref.py\n')
    return n


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference')
    print(stage(ap.parse_args().src), 'files staged into', DST)
