"""NYU val pre-processing (reference dataloaders/nyu.py:48-59) as one gather."""
import numpy as np
import pytest
import torch

from fastdepth_b200 import preprocess
from oracle import fastdepth_oracle as orc


def _sample(seed, h=480, w=640):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (rng.random((h, w), dtype=np.float32) * 9.0 + 0.5)


def test_index_maps_shape_and_monotone():
    rows, cols = preprocess.nyu_val_index_maps()
    assert rows.shape == (224,) and cols.shape == (224,)
    assert (np.diff(rows) > 0).all() and (np.diff(cols) > 0).all()
    # centre crop of the 250x333 intermediate: 11 rows / 14-15 columns dropped per side -> source margins
    assert 20 <= rows[0] <= 24 and 455 <= rows[-1] <= 459
    assert 26 <= cols[0] <= 32 and 606 <= cols[-1] <= 612


@pytest.mark.parametrize('out_hw', [(224, 224), (192, 256)])
def test_gather_tables_equal_three_step_chain(out_hw):
    """One gather through the composed tables == Resize -> CenterCrop -> Resize done step by step (bit-exact)."""
    rgb, depth = _sample(1)
    x, t = orc.nyu_val_transform(rgb, depth, out_hw)
    rows, cols = preprocess.nyu_val_index_maps(480, 640, out_hw)
    xg = torch.from_numpy((rgb[rows][:, cols].astype('float') / 255).transpose(2, 0, 1).copy()).float()
    tg = torch.from_numpy(depth[rows][:, cols].copy()).unsqueeze(0)
    assert x.shape == (3, *out_hw) and t.shape == (1, *out_hw)
    assert torch.equal(x, xg) and torch.equal(t, tg)
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float32, torch.float16, torch.bfloat16])
def test_gpu_gather_bit_exact(dtype):
    n = 3
    samples = [_sample(10 + i) for i in range(n)]
    rgb = torch.from_numpy(np.stack([s[0] for s in samples])).cuda()
    depth = torch.from_numpy(np.stack([s[1] for s in samples])).cuda()
    x, t = preprocess.nyu_val_transform(rgb, depth, (224, 224), dtype)
    assert x.shape == (n, 3, 224, 224) and x.dtype == dtype and t.shape == (n, 1, 224, 224) and t.dtype == torch.float32
    for i, (r, d) in enumerate(samples):
        xo, to = orc.nyu_val_transform(r, d)
        assert torch.equal(x[i].cpu(), xo.to(dtype)), f'image {i}'
        assert torch.equal(t[i].cpu(), to)


@pytest.mark.gpu
def test_gpu_gather_rgb_only_and_errors():
    r, _ = _sample(5)
    rgb = torch.from_numpy(r[None]).cuda()
    x, t = preprocess.nyu_val_transform(rgb, None)
    assert t is None and torch.equal(x[0].cpu(), orc.nyu_val_transform(r, np.zeros((480, 640), np.float32))[0])
    from fastdepth_b200 import _lib
    lib = _lib.load()
    with pytest.raises(RuntimeError):
        _lib.check(lib.fd_nyu_val_gather(rgb.data_ptr(), None, None, None, 1, 480, 640, 224, 224, 0, x.data_ptr(), None, 0, None))


@pytest.mark.gpu
def test_gpu_preprocess_feeds_forward():
    """Raw frame -> gather -> fused forward -> metrics, all on device, equals the oracle chain."""
    import models
    from fastdepth_b200 import synthetic
    r, d = _sample(7)
    sd = synthetic.synthetic_state_dict(synthetic.STOCK_WIDTHS, seed=0)
    m = models.MobileNetSkipAdd((224, 224), pretrained=False)
    m.load_state_dict(sd); m.eval().cuda()
    x, t = preprocess.nyu_val_transform(torch.from_numpy(r[None]).cuda(), torch.from_numpy(d[None]).cuda())
    with torch.no_grad():
        y = m(x)
    xo, _ = orc.nyu_val_transform(r, d)
    yo = orc.skipadd_forward(sd, xo[None])
    from conftest import rel_err
    assert rel_err(y.cpu(), yo) <= 1e-3
