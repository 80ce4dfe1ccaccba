"""Host-only check of the fused block kernel's shared-memory / pipeline planner (no GPU needed): every block of
the stock and the NetAdapt-pruned network, at 224x224 b64, 480x640 b16 and tiny shapes, must get a plan that fits
the 227 KB of a B200 SM, keeps the TMEM accumulator pair within 512 columns and never falls back to narrow MMAs."""
import ctypes

import pytest

from fastdepth_b200 import _lib, synthetic

KEYS = ('ok', 'splits', 'n_cta', 'items', 'kblocks', 's_in', 's_a', 's_b', 'bn', 'nb', 'b_resident', 'epi_groups',
        'n_stg', 'smem_bytes', 'tmem_cols', 'in_stage_stride', 'nacc', 'epi_colsplit')
STRIDES = (2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1)


def plan(ks, stride, h, w, n, cin, cout, head=0):
    lib = _lib.load()
    out = (ctypes.c_int * 18)()
    _lib.check(lib.fd_debug_block_plan(ks, stride, h, w, n, cin, cout, head, out, 18))
    return dict(zip(KEYS, out))


def blocks(widths, h, w):
    enc, dec = widths
    hh, ww = h // 2, w // 2
    for i in range(1, 14):
        hh, ww = hh // STRIDES[i], ww // STRIDES[i]
        yield ('conv%d' % i, 3, STRIDES[i], hh, ww, enc[i - 1], enc[i], 0)
    c = enc[13]
    for j, co in enumerate(dec, start=1):
        yield ('decode_conv%d' % j, 5, 1, hh, ww, c, co, 1 if j == 5 else 0)
        c, hh, ww = co, hh * 2, ww * 2


@pytest.mark.parametrize('widths', [synthetic.STOCK_WIDTHS, synthetic.PRUNED_WIDTHS], ids=['stock', 'pruned'])
@pytest.mark.parametrize('shape', [(64, 224, 224), (16, 480, 640), (2, 64, 96), (1, 32, 32), (512, 224, 224)])
def test_every_block_gets_a_valid_plan(built_lib, widths, shape):
    n, h, w = shape
    for name, ks, stride, hh, ww, cin, cout, head in blocks(widths, h, w):
        p = plan(ks, stride, hh, ww, n, cin, cout, head)
        assert p['ok'] == 1, (name, p)
        assert p['smem_bytes'] <= 227 * 1024, (name, p)
        assert p['nacc'] in (1, 2) and p['nacc'] * p['n_cta'] <= p['tmem_cols'] <= 512 and p['n_cta'] % 16 == 0, (name, p)
        assert p['nacc'] == 2 or (p['epi_colsplit'] == 1 and p['epi_groups'] == 2), (name, p)   # one accumulator: both groups drain it
        assert not p['epi_colsplit'] or (p['epi_groups'] == 2 and p['n_cta'] > 64 and not head), (name, p)
        assert p['n_cta'] * p['splits'] >= cout and (p['splits'] == 1 or p['n_cta'] % 64 == 0), (name, p)
        assert p['s_in'] >= 1 and 2 <= p['s_a'] <= 4 and p['bn'] * p['nb'] >= p['n_cta'], (name, p)
        assert p['bn'] >= min(64, p['n_cta']), (name, p)            # no narrow MMAs
        assert p['epi_groups'] in (1, 2) and (head or p['n_stg'] in (p['epi_groups'], 2 * p['epi_groups'])), (name, p)
        if p['b_resident']:
            assert p['splits'] == 1 and p['s_b'] == p['kblocks'] * p['nb'] <= 16


def test_stock_b64_plans_snapshot(built_lib):
    """The metric configuration: high-res blocks keep their weights resident and a deep A ring."""
    p = plan(3, 1, 112, 112, 64, 32, 64)          # conv1
    assert p['b_resident'] == 1 and p['s_a'] == 4 and p['epi_groups'] == 2 and p['items'] == 6272
    p = plan(3, 1, 14, 14, 64, 512, 512)          # conv7..11
    # one 512-column accumulator per tile: the depthwise half and the input tile are not repeated per split
    assert p['splits'] == 1 and p['n_cta'] == 512 and p['nacc'] == 1 and p['bn'] == 256 and p['epi_colsplit'] == 1
    p = plan(5, 1, 112, 112, 64, 64, 32, head=1)  # decode_conv5 + folded head
    assert p['n_stg'] == 0 and p['splits'] == 1
