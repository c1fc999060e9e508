#!/usr/bin/env python
"""Recipe for oracle/_ref — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (musikisomorphie/implicit-normalizing-flows) is pure Python.  This recipe COMPILES its `lib/`
package, from the sources where they lie under /root/reference, to CPython bytecode:

    python oracle/build_ref.py            # /root/reference/lib/**/*.py -> oracle/_ref/lib/**/*.pyc

Only the compiled .pyc files are written (sourceless layout: `pkg/module.pyc`, which the import system loads
without the .py) — no reference source text enters the repository or its history; `oracle/_ref/` is git-ignored
but not gpurun-ignored, so the compiled reference travels to the GPU box (same image, same interpreter) with
the snapshot like the built .so files do.  `oracle/ref_runner.py` imports it behind the two import shims the
reference needs on a current stack (`torch._six`, `termcolor` — SURVEY.md section 8c); it is the checker in
tests and the CPU arm of bench.py (`cpu_baseline.kind = "reference"`), never part of the product."""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('IMPFLOW_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(HERE, '_ref')
STAMP = os.path.join(DST, 'BUILT_FROM')


def _stamp():
    parts = [sys.version.split()[0]]
    for root, _, files in sorted(os.walk(os.path.join(SRC, 'lib'))):
        for f in sorted(files):
            if f.endswith('.py'):
                st = os.stat(os.path.join(root, f))
                parts.append('%s:%d:%d' % (os.path.relpath(os.path.join(root, f), SRC), st.st_size, int(st.st_mtime)))
    return '\n'.join(parts)


def build(verbose=True):
    """Compile <reference>/lib into oracle/_ref/lib when the reference checkout is present.  Returns the
    destination, or None when there is neither a checkout nor a prebuilt copy (GPU box: the prebuilt one is used)."""
    src = os.path.join(SRC, 'lib')
    prebuilt = os.path.isfile(os.path.join(DST, 'lib', 'implicit_flow.pyc'))
    if not os.path.isdir(src):
        return DST if prebuilt else None
    stamp = _stamp()
    if prebuilt and os.path.isfile(STAMP) and open(STAMP).read() == stamp:
        return DST
    shutil.rmtree(DST, ignore_errors=True)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d != '__pycache__']
        out_dir = os.path.join(DST, os.path.relpath(root, SRC))
        os.makedirs(out_dir, exist_ok=True)
        for f in files:
            if f.endswith('.py'):
                py_compile.compile(os.path.join(root, f), cfile=os.path.join(out_dir, f + 'c'), dfile=os.path.join(
                    'reference', os.path.relpath(os.path.join(root, f), SRC)), doraise=True, quiet=1)
                n += 1
    with open(STAMP, 'w') as f:
        f.write(stamp)
    if verbose:
        print('oracle/_ref: compiled %d modules of %s to bytecode' % (n, src))
    return DST


if __name__ == '__main__':
    out = build()
    print(out if out else 'no reference checkout at %s and no prebuilt oracle/_ref' % SRC)
    sys.exit(0 if out else 1)
