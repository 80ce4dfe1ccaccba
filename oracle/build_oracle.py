"""ORACLE -- TEST INFRASTRUCTURE ONLY.  gcc build of the plain-C oracle primitives (oracle/fastdepth_oracle.c ->
oracle/libfastdepth_oracle.so).  Called by ``__graft_entry__.build()`` and by the test session fixture; the product
package never refers to it (tests/test_layout.py).  ``python oracle/build_oracle.py`` builds by hand."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'fastdepth_oracle.c')
LIB = os.path.join(HERE, 'libfastdepth_oracle.so')


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = [os.environ.get('CC', 'gcc'), '-O2', '-fPIC', '-shared', '-o', LIB, SRC, '-lm']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('oracle build failed: %s\n%s' % (' '.join(cmd), r.stdout))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
