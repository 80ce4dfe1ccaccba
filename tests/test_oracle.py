"""Pin the oracle (oracle/fastdepth_oracle.py) against vectors the LIVE reference produced
(tests/golden/make_golden.py ran reference models.py:654-732 and metrics.py:31-95)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from fastdepth_b200 import synthetic
from oracle import fastdepth_oracle as orc

FIXTURES = ['skipadd_stock_2x64x96', 'skipadd_pruned_2x64x96', 'skipadd_stock_1x224x224']


def _load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


@pytest.mark.parametrize('name', FIXTURES)
def test_oracle_forward_matches_reference(name):
    fx = _load(name)
    widths = (tuple(int(v) for v in fx['widths_enc']), tuple(int(v) for v in fx['widths_dec']))
    n, h, w = (int(v) for v in fx['shape'])
    sd = synthetic.synthetic_state_dict(widths, seed=int(fx['wseed']))
    x = synthetic.synthetic_input(n, h, w, seed=int(fx['xseed']))
    stages = {}
    y = orc.skipadd_forward(sd, x, stages=stages)
    want = torch.from_numpy(fx['output'])
    assert y.shape == want.shape
    assert (want == 0).float().mean() < 0.5          # not a dead output
    assert rel_err(y, want) < 1e-4              # fp32 summation-order drift only (fp64 tie-break below)
    # stage-wise samples localise any drift
    for key in [k[6:-4] for k in fx.files if k.startswith('stage/') and k.endswith('/idx')]:
        got = stages[key].reshape(-1)[torch.from_numpy(fx['stage/%s/idx' % key])]
        ref = torch.from_numpy(fx['stage/%s/val' % key])
        assert tuple(stages[key].shape) == tuple(fx['stage/%s/shape' % key]), key
        assert torch.allclose(got, ref, rtol=3e-4, atol=3e-4), key    # fp32 summation-order drift


def test_oracle_nnconv_dw_matches_reference():
    """MobileNet + NNConv(5, dw) without skips against the live reference's output."""
    fx = _load('nnconv5dw_stock_2x64x96')
    n, h, w = (int(v) for v in fx['shape'])
    sd = synthetic.to_mobilenet_keys(synthetic.synthetic_state_dict(seed=int(fx['wseed'])))
    y = orc.nnconv_dw_forward(sd, synthetic.synthetic_input(n, h, w, seed=int(fx['xseed'])))
    assert rel_err(y, torch.from_numpy(fx['output'])) < 1e-4


def test_oracle_skipconcat_matches_reference():
    """MobileNetSkipConcat against the live reference's output."""
    fx = _load('skipconcat_stock_2x64x96')
    n, h, w = (int(v) for v in fx['shape'])
    sd = synthetic.synthetic_state_dict(seed=int(fx['wseed']), skip='concat')
    y = orc.skipconcat_forward(sd, synthetic.synthetic_input(n, h, w, seed=int(fx['xseed'])))
    assert rel_err(y, torch.from_numpy(fx['output'])) < 1e-4


def test_oracle_fp64_agrees_with_fp32():
    sd = synthetic.synthetic_state_dict(seed=3)
    x = synthetic.synthetic_input(1, 32, 64, seed=5)
    a = orc.skipadd_forward(sd, x)
    b = orc.skipadd_forward(sd, x, dtype=torch.float64)
    assert rel_err(a, b) < 1e-4


def test_relu6_clamp_is_exercised():
    """The synthetic recipe must actually hit the upper clamp, otherwise ReLU6 is untested."""
    sd = synthetic.synthetic_state_dict(seed=1)
    x = synthetic.synthetic_input(1, 64, 64, seed=0)
    stages = {}
    orc.skipadd_forward(sd, x, stages=stages)
    sat = (stages['conv5'] >= 6.0).float().mean().item()
    assert 2e-4 < sat < 0.5


def test_metrics_known_answer():
    fx = _load('metrics_known_answer')
    names = [str(n) for n in fx['names']]
    got = orc.evaluate_one(fx['pred_sub4'], fx['depth_sub4'])
    for k, v in zip(names, fx['sub4_values']):
        assert got[k] == pytest.approx(float(v), rel=2e-5, abs=1e-9), k
    full = dict(zip(names, fx['full_values']))
    assert full['rmse'] == pytest.approx(618.001, abs=2e-3)
    assert full['delta1'] == pytest.approx(0.7773, abs=1e-4)


def test_metrics_are_averaged_per_image():
    fx = _load('metrics_known_answer')
    names = [str(n) for n in fx['names']]
    avg, n = orc.average_per_image(fx['multi_out'], fx['multi_tgt'])
    assert n == 3
    for k, v in zip(names, fx['multi_avg']):
        assert avg[k] == pytest.approx(float(v), rel=2e-5), k
    pooled = orc.evaluate_one(fx['multi_out'], fx['multi_tgt'])
    assert abs(pooled['rmse'] - avg['rmse']) > 1e-3              # pooling pixels is a different number


# ---------------------------------------------------------------------------------------------------------------------
# the plain-C restatement of the arithmetic (oracle/fastdepth_oracle.c + oracle/c_oracle.py: loops, no PyTorch operator)
# against the same vectors of the live reference
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', FIXTURES)
def test_c_oracle_forward_matches_reference(name, built_lib):
    from oracle import c_oracle
    fx = _load(name)
    widths = (tuple(int(v) for v in fx['widths_enc']), tuple(int(v) for v in fx['widths_dec']))
    n, h, w = (int(v) for v in fx['shape'])
    sd = synthetic.synthetic_state_dict(widths, seed=int(fx['wseed']))
    x = synthetic.synthetic_input(n, h, w, seed=int(fx['xseed']))
    stages = {}
    y = torch.from_numpy(c_oracle.forward(sd, x, skip='add', stages=stages))
    want = torch.from_numpy(fx['output'])
    assert y.shape == want.shape and rel_err(y, want) < 1e-4
    for key in [k[6:-4] for k in fx.files if k.startswith('stage/') and k.endswith('/idx')]:
        if key not in stages:
            continue                                       # the C composition records the named children only
        got = torch.from_numpy(stages[key]).reshape(-1)[torch.from_numpy(fx['stage/%s/idx' % key])]
        assert tuple(stages[key].shape) == tuple(fx['stage/%s/shape' % key]), key
        assert torch.allclose(got, torch.from_numpy(fx['stage/%s/val' % key]), rtol=3e-4, atol=3e-4), key


def test_c_oracle_skipconcat_and_no_skip_match_reference(built_lib):
    from oracle import c_oracle
    fx = _load('skipconcat_stock_2x64x96')
    n, h, w = (int(v) for v in fx['shape'])
    sd = synthetic.synthetic_state_dict(seed=int(fx['wseed']), skip='concat')
    y = torch.from_numpy(c_oracle.forward(sd, synthetic.synthetic_input(n, h, w, seed=int(fx['xseed'])), skip='concat'))
    assert rel_err(y, torch.from_numpy(fx['output'])) < 1e-4
    fx = _load('nnconv5dw_stock_2x64x96')
    n, h, w = (int(v) for v in fx['shape'])
    sd = orc.to_skipadd_keys(synthetic.to_mobilenet_keys(synthetic.synthetic_state_dict(seed=int(fx['wseed']))))
    y = torch.from_numpy(c_oracle.forward(sd, synthetic.synthetic_input(n, h, w, seed=int(fx['xseed'])), skip=None))
    assert rel_err(y, torch.from_numpy(fx['output'])) < 1e-4


def test_c_oracle_agrees_with_torch_oracle_in_fp64(built_lib):
    """Two independent restatements, one answer: double-accumulating C loops vs the PyTorch-primitive oracle run in fp64."""
    from oracle import c_oracle
    sd = synthetic.synthetic_state_dict(synthetic.PRUNED_WIDTHS, seed=5)
    x = synthetic.synthetic_input(1, 64, 64, seed=9)
    a = torch.from_numpy(c_oracle.forward(sd, x)).double()
    b = orc.skipadd_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, x.double(), dtype=torch.float64)
    assert rel_err(a, b) < 2e-5            # the C path rounds every layer's output to fp32 like the reference does
