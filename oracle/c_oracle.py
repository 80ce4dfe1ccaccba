"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Composition of the plain-C primitives of oracle/fastdepth_oracle.c into
``MobileNetSkipAdd.forward`` / ``MobileNetSkipConcat.forward`` (reference models.py:706-732, 789-814) on a state_dict.

numpy + ctypes only: no PyTorch operator takes part, so this oracle shares no operator library with the reference (which
delegates its arithmetic to PyTorch) nor with oracle/fastdepth_oracle.py (which restates the path with PyTorch primitives).
Loops in C, so the 2x64x96 and 1x224x224 golden cases take a fraction of a second / a few seconds.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libfastdepth_oracle.so')
ENCODER_STRIDES = (2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1)   # reference imagenet/mobilenet.py:41-54
BN_EPS = 1e-5
_F = ctypes.POINTER(ctypes.c_float)
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('oracle/libfastdepth_oracle.so is missing: run `python oracle/build_oracle.py`')
        lib = ctypes.CDLL(LIB_PATH)
        i = ctypes.c_int
        lib.fo_conv_dense.argtypes = [_F, _F, _F, i, i, i, i, i, i, i, i]
        lib.fo_conv_depthwise.argtypes = [_F, _F, _F, i, i, i, i, i, i, i]
        lib.fo_conv_pointwise.argtypes = [_F, _F, _F, i, i, i, i]
        lib.fo_bn_act.argtypes = [_F, _F, _F, _F, _F, ctypes.c_float, i, i, i, i]
        lib.fo_upsample2x.argtypes = [_F, _F, i, i, i]
        lib.fo_add.argtypes = [_F, _F, ctypes.c_size_t]
        for f in ('fo_conv_dense', 'fo_conv_depthwise', 'fo_conv_pointwise', 'fo_bn_act', 'fo_upsample2x', 'fo_add'):
            getattr(lib, f).restype = None
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(_F)


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, 'detach') else t, dtype=np.float32)


def _bn_act(x, sd, prefix, act):
    n, c, h, w = x.shape
    load().fo_bn_act(_p(x), _p(_np(sd[prefix + '.running_mean'])), _p(_np(sd[prefix + '.running_var'])),
                     _p(_np(sd[prefix + '.weight'])), _p(_np(sd[prefix + '.bias'])), BN_EPS, act, n, c, h * w)
    return x


def _dense(x, w, stride, pad):
    n, ci, h, wd = x.shape
    co, _, k, _ = w.shape
    out = np.empty((n, co, (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1), np.float32)
    load().fo_conv_dense(_p(x), _p(w), _p(out), n, ci, h, wd, co, k, stride, pad)
    return out


def _depthwise(x, w, stride):
    n, c, h, wd = x.shape
    k = w.shape[-1]
    pad = (k - 1) // 2
    out = np.empty((n, c, (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1), np.float32)
    load().fo_conv_depthwise(_p(x), _p(w), _p(out), n, c, h, wd, k, stride, pad)
    return out


def _pointwise(x, w):
    n, ci, h, wd = x.shape
    co = w.shape[0]
    out = np.empty((n, co, h, wd), np.float32)
    load().fo_conv_pointwise(_p(x), _p(np.ascontiguousarray(w.reshape(co, ci))), _p(out), n, ci, h * wd, co)
    return out


def _upsample(x):
    n, c, h, w = x.shape
    out = np.empty((n, c, 2 * h, 2 * w), np.float32)
    load().fo_upsample2x(_p(x), _p(out), n * c, h, w)
    return out


def forward(sd, x, skip='add', stages=None):
    """skip='add': MobileNetSkipAdd.forward; 'concat': MobileNetSkipConcat.forward; None: no skips (NNConv5 depthwise decoder
    behind MobileNet, state_dict already renamed with fastdepth_oracle.to_skipadd_keys).  Returns [N,1,H,W] float32."""
    x = _np(x)
    x = _bn_act(_dense(x, _np(sd['conv0.0.weight']), 2, 1), sd, 'conv0.1', 2)          # conv_bn + ReLU6, mobilenet.py:22-27
    keep = {}
    if stages is not None:
        stages['conv0'] = x.copy()
    for i in range(1, 14):                                                                 # conv_dw, mobilenet.py:29-38
        x = _bn_act(_depthwise(x, _np(sd['conv%d.0.weight' % i]), ENCODER_STRIDES[i]), sd, 'conv%d.1' % i, 2)
        x = _bn_act(_pointwise(x, _np(sd['conv%d.3.weight' % i])), sd, 'conv%d.4' % i, 2)
        if stages is not None:
            stages['conv%d' % i] = x.copy()
        if i in (1, 3, 5):                                                                 # models.py:714-719
            keep[i] = x
    after = {4: 1, 3: 3, 2: 5}                                                             # models.py:724-729 / 806-811
    for j in range(1, 6):
        x = _bn_act(_depthwise(x, _np(sd['decode_conv%d.0.0.weight' % j]), 1), sd, 'decode_conv%d.0.1' % j, 1)
        x = _bn_act(_pointwise(x, _np(sd['decode_conv%d.1.0.weight' % j])), sd, 'decode_conv%d.1.1' % j, 1)
        x = _upsample(x)
        if skip == 'concat' and stages is not None:
            stages['decode_conv%d' % j] = x.copy()                                         # the block's own slice
        if skip is not None and j in after:
            if skip == 'add':
                load().fo_add(_p(x), _p(keep[after[j]]), x.size)
            else:
                x = np.ascontiguousarray(np.concatenate((x, keep[after[j]]), 1))
        if skip != 'concat' and stages is not None:
            stages['decode_conv%d' % j] = x.copy()
    x = _bn_act(_pointwise(x, _np(sd['decode_conv6.0.weight'])), sd, 'decode_conv6.1', 1)  # pointwise(32, 1), models.py:698,731
    if stages is not None:
        stages['decode_conv6'] = x.copy()
    return x
