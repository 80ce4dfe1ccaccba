"""The oracle is test infrastructure: the product path must never import it."""
import os
import re

from conftest import ROOT

PRODUCT = ['models.py', 'imagenet', 'fastdepth_b200']


def product_files():
    for p in PRODUCT:
        full = os.path.join(ROOT, p)
        if os.path.isfile(full):
            yield full
        else:
            for d, _, fs in os.walk(full):
                for f in fs:
                    if f.endswith(('.py', '.cu', '.cuh', '.h')):
                        yield os.path.join(d, f)


def test_product_never_touches_the_oracle():
    pat = re.compile(r'^\s*(from|import)\s+oracle\b|oracle/|fastdepth_oracle', re.M)
    for f in product_files():
        assert not pat.search(open(f).read()), f


def test_no_compat_layers():
    pat = re.compile(r'^\s*(import|from)\s+(triton|tilelang|tvm)\b|torch\.compile\(', re.M)
    for f in product_files():
        assert not pat.search(open(f).read()), f


def test_required_layout():
    for p in ('bench.py', '__graft_entry__.py', 'include/fastdepth_b200.h', 'oracle/fastdepth_oracle.py',
              'tests/golden/make_golden.py', 'DESIGN.md', 'INTEGRATION.md'):
        assert os.path.exists(os.path.join(ROOT, p)), p
