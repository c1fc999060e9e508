#!/usr/bin/env python
"""Recipe for oracle/_ref — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (musikisomorphie/implicit-normalizing-flows) is pure Python.  This recipe COMPILES its `lib/`
package, from the sources where they lie under /root/reference, to CPython bytecode:

    python oracle/build_ref.py            # /root/reference/lib/**/*.py -> oracle/_ref/reference_lib.bin

Only compiled bytecode is written: one archive (zip container, sourceless layout `lib/pkg/module.pyc`, loaded by
the standard zipimport machinery without any .py) — no reference source text enters the repository or its
history; `oracle/_ref/` is git-ignored but not gpurun-ignored, so the compiled reference travels to the GPU box
(same image, same interpreter) with the snapshot like the built .so files do.  `oracle/ref_runner.py` imports it behind the two import shims the
reference needs on a current stack (`torch._six`, `termcolor` — SURVEY.md section 8c); it is the checker in
tests and the CPU arm of bench.py (`cpu_baseline.kind = "reference"`), never part of the product."""
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get('IMPFLOW_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(HERE, '_ref')
STAMP = os.path.join(DST, 'BUILT_FROM')
ARCHIVE = os.path.join(DST, 'reference_lib.bin')


def _stamp():
    parts = [sys.version.split()[0]]
    for root, _, files in sorted(os.walk(os.path.join(SRC, 'lib'))):
        for f in sorted(files):
            if f.endswith('.py'):
                st = os.stat(os.path.join(root, f))
                parts.append('%s:%d:%d' % (os.path.relpath(os.path.join(root, f), SRC), st.st_size, int(st.st_mtime)))
    return '\n'.join(parts)


def build(verbose=True):
    """Compile <reference>/lib into oracle/_ref/reference_lib.bin when the reference checkout is present.  Returns
    the archive path, or None when there is neither a checkout nor a prebuilt archive (GPU box: the prebuilt one is
    used as it is)."""
    src = os.path.join(SRC, 'lib')
    prebuilt = os.path.isfile(ARCHIVE)
    if not os.path.isdir(src):
        return ARCHIVE if prebuilt else None
    stamp = _stamp()
    if prebuilt and os.path.isfile(STAMP) and open(STAMP).read() == stamp:
        return ARCHIVE
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    n = 0
    with tempfile.TemporaryDirectory() as tmp, zipfile.ZipFile(ARCHIVE, 'w', zipfile.ZIP_DEFLATED) as zf:
        empty = os.path.join(tmp, 'empty.py')
        open(empty, 'w').close()
        init_pyc = os.path.join(tmp, 'empty.pyc')
        py_compile.compile(empty, cfile=init_pyc, dfile='reference/lib/__init__.py', doraise=True)
        zf.write(init_pyc, 'lib/__init__.pyc')          # `lib` is a namespace directory in the reference
        for root, dirs, files in os.walk(src):
            dirs[:] = [d for d in dirs if d != '__pycache__']
            rel = os.path.relpath(root, SRC)
            for f in sorted(files):
                if f.endswith('.py'):
                    out = os.path.join(tmp, 'm%d.pyc' % n)
                    py_compile.compile(os.path.join(root, f), cfile=out, dfile=os.path.join('reference', rel, f),
                                       doraise=True, quiet=1)
                    zf.write(out, os.path.join(rel, f + 'c'))
                    n += 1
    with open(STAMP, 'w') as f:
        f.write(stamp)
    if verbose:
        print('oracle/_ref: compiled %d modules of %s to bytecode (%s)' % (n, src, os.path.basename(ARCHIVE)))
    return ARCHIVE


if __name__ == '__main__':
    out = build()
    print(out if out else 'no reference checkout at %s and no prebuilt oracle/_ref' % SRC)
    sys.exit(0 if out else 1)
