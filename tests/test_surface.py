"""Drop-in surface of models.py / imagenet/mobilenet.py."""
import io
import os
import pickle

import numpy as np
import pytest
import torch
import torch.nn as nn

import imagenet.mobilenet
import models
from conftest import GOLDEN
from fastdepth_b200 import plan, synthetic


def expected_keys():
    keys = []
    bn = ['weight', 'bias', 'running_mean', 'running_var', 'num_batches_tracked']
    keys += ['conv0.0.weight'] + ['conv0.1.' + b for b in bn]
    for i in range(1, 14):
        keys += ['conv%d.0.weight' % i] + ['conv%d.1.%s' % (i, b) for b in bn]
        keys += ['conv%d.3.weight' % i] + ['conv%d.4.%s' % (i, b) for b in bn]
    for j in range(1, 6):
        keys += ['decode_conv%d.0.0.weight' % j] + ['decode_conv%d.0.1.%s' % (j, b) for b in bn]
        keys += ['decode_conv%d.1.0.weight' % j] + ['decode_conv%d.1.1.%s' % (j, b) for b in bn]
    keys += ['decode_conv6.0.weight'] + ['decode_conv6.1.' + b for b in bn]
    return keys


def test_state_dict_schema_and_child_order():
    m = models.MobileNetSkipAdd((224, 224), pretrained=False)
    assert list(m.state_dict().keys()) == expected_keys()          # 228 entries, reference order
    assert len(expected_keys()) == 228
    assert [n for n, _ in m.named_children()] == ['conv%d' % i for i in range(14)] + \
        ['decode_conv%d' % j for j in range(1, 7)]
    assert m.output_size == (224, 224)
    assert sum(p.numel() for p in m.parameters()) == 3960257 or sum(p.numel() for p in m.parameters()) > 3.9e6


def test_synthetic_state_dict_loads_strictly():
    for widths in (synthetic.STOCK_WIDTHS, synthetic.PRUNED_WIDTHS):
        m = models.MobileNetSkipAdd((224, 224), pretrained=False, widths=widths)
        m.load_state_dict(synthetic.synthetic_state_dict(widths), strict=True)


def test_no_cpu_fallback_and_eval_only():
    m = models.MobileNetSkipAdd((64, 64), pretrained=False).eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.rand(1, 3, 64, 64))
    m.train()
    with pytest.raises(RuntimeError, match='inference-only'):
        m(torch.rand(1, 3, 64, 64))


def test_pickle_roundtrip_like_main_py():
    """main.py:49-57 loads whole-module pickles; unpickling skips __init__ and must not carry
    the engine."""
    m = models.MobileNetSkipAdd((224, 224), pretrained=False).eval()
    m.__dict__['_fd_engine'] = object()
    buf = io.BytesIO()
    torch.save({'model': m, 'epoch': 0}, buf)
    buf.seek(0)
    ck = torch.load(buf, weights_only=False)
    m2 = ck['model']
    assert isinstance(m2, models.MobileNetSkipAdd) and '_fd_engine' not in m2.__dict__
    assert list(m2.state_dict().keys()) == expected_keys()


def test_config1_plumbing_mobilenet_nnconv5_cpu():
    """BASELINE config 1: MobileNet-NNConv5 (dense decoder) batch 1 fp32 on CPU, plain PyTorch."""
    m = models.MobileNet('nnconv5', (224, 224), pretrained=False).eval()
    assert sum(p.numel() for p in m.parameters()) > 20e6            # dense 5x5 decoder
    with torch.no_grad():
        y = m(torch.rand(1, 3, 64, 64))
    assert y.shape == (1, 1, 64, 64)
    mdw = models.MobileNet('nnconv5dw', (224, 224), pretrained=False).eval()
    with torch.no_grad():
        assert mdw(torch.rand(1, 3, 32, 32)).shape == (1, 1, 32, 32)


def test_decoder_factory_strings():
    assert isinstance(models.choose_decoder('nnconv5'), models.NNConv)
    with pytest.raises(NotImplementedError):
        models.choose_decoder('upproj')
    with pytest.raises(AssertionError):
        models.choose_decoder('bogus')


def test_weights_init_quirk():
    """weights_init on an nn.Sequential is a no-op (reference models.py:699-704)."""
    seq = models.pointwise(8, 8)
    before = seq[0].weight.clone()
    models.weights_init(seq)
    assert torch.equal(seq[0].weight, before)
    models.weights_init(seq[0])
    assert not torch.equal(seq[0].weight, before)


def test_encoder_surface():
    e = imagenet.mobilenet.MobileNet()
    assert len(e.model) == 15 and isinstance(e.model[14], nn.AvgPool2d)
    assert e.model[6][0].stride == (2, 2) and e.model[7][0].stride == (1, 1)
    with torch.no_grad():
        assert e.eval()(torch.rand(1, 3, 224, 224)).shape == (1, 1000)


def test_describe_and_bn_folding():
    m = models.MobileNetSkipAdd((224, 224), pretrained=False, widths=synthetic.PRUNED_WIDTHS)
    m.load_state_dict(synthetic.synthetic_state_dict(synthetic.PRUNED_WIDTHS))
    m.eval()
    descs, weights, names = plan.describe(m)
    assert len(descs) == 20 and names[14] == 'decode_conv1'
    assert [d['stride'] for d in descs[:14]] == [2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1]
    assert [d['skip_src'] for d in descs[14:19]] == [-1, 5, 3, 1, -1]
    assert [d['ksize'] for d in descs[14:19]] == [5] * 5 and all(d['upsample'] for d in descs[14:19])
    assert descs[0]['act'] == 1 and descs[14]['act'] == 0 and descs[19]['c_in'] == 16
    # folded affine == BatchNorm eval
    bn = m.conv3[1]
    x = torch.randn(2, bn.num_features, 3, 3)
    s, b = plan.fold_bn(bn)
    want = bn(x)
    got = x * torch.from_numpy(s).view(1, -1, 1, 1) + torch.from_numpy(b).view(1, -1, 1, 1)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    assert weights[2][0].shape == (56, 9) and weights[2][3].shape == (88, 56)


def test_skipconcat_surface():
    """reference models.py:734-814: same children / key schema as SkipAdd, wider decoder inputs; plan marks concat skips."""
    m = models.MobileNetSkipConcat((224, 224), pretrained=False)
    assert list(m.state_dict().keys()) == expected_keys()
    assert [getattr(m, 'decode_conv%d' % j)[0][0].in_channels for j in range(1, 6)] == [1024, 512, 512, 256, 128]
    m.load_state_dict(synthetic.synthetic_state_dict(skip='concat'), strict=True)
    descs, _, _ = plan.describe(m.eval())
    assert [(d['skip_src'], d.get('skip_mode', 0)) for d in descs[14:19]] == [(-1, 0), (5, 1), (3, 1), (1, 1), (-1, 0)]
    assert [d['c_in'] for d in descs[14:20]] == [1024, 512, 512, 256, 128, 32]


def test_forward_lanes_surface_without_gpu():
    """fastdepth_b200.engine.ForwardLanes is plain host bookkeeping until a batch arrives: it can be built anywhere, refuses
    nonsense, and -- like the module -- has no CPU fallback."""
    from fastdepth_b200.engine import ForwardLanes
    m = models.MobileNetSkipAdd((64, 96), pretrained=False).eval()
    lanes = ForwardLanes(m, lanes=3)
    assert len(lanes) == 3 and all(e.options.get('pdl') == 0 for e in lanes.engines)      # PDL stays off on concurrent lanes
    assert ForwardLanes(m, lanes=1).engines[0].options == {}
    with pytest.raises(ValueError):
        ForwardLanes(m, lanes=0)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            lanes.engines[0].plan_for(torch.rand(1, 3, 64, 96))
