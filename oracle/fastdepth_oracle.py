"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement of the reference's algorithm for the hot path
``MobileNetSkipAdd.forward`` (reference models.py:706-732) and of its metric
(reference metrics.py:31-55).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file; nothing
under ``fastdepth_b200/`` or ``models.py`` does (tests/test_layout.py enforces it).

Where the arithmetic lives: the reference delegates every operation to PyTorch
(``nn.Conv2d``, ``nn.BatchNorm2d``, ``nn.ReLU/ReLU6``, ``F.interpolate``, ``+``), a
third-party dependency whose source is not under /root/reference (README.md:19 names
v0.4.1 in prose; no lock file).  The restatement therefore spells each stage out with the
same primitive semantics -- cross-correlation with symmetric zero padding, eval-mode BN
``(x-mean)/sqrt(var+eps)*gamma+beta`` with eps=1e-5, ReLU6=clamp(0,6), nearest x2 =
``in[y//2, x//2]`` -- as explicit functional calls on a *state_dict* (no nn.Module), and is
PINNED against the live reference module imported read-only from /root/reference:
``tests/golden/make_golden.py`` runs the reference's own ``models.MobileNetSkipAdd`` and
``metrics.Result`` in the build container and commits the vectors under ``tests/golden``;
``tests/test_oracle.py`` checks this file against them.  (The reference itself has no
tests or golden vectors for this path -- SURVEY.md section 4 / 8c.)

Run with ``dtype=torch.float64`` for a tie-breaking higher-precision answer.  An operator-library-free twin of this
file (plain C loops + numpy/ctypes composition) lives in ``oracle/fastdepth_oracle.c`` / ``oracle/c_oracle.py`` and is
pinned against the same golden vectors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

ENCODER_STRIDES = (2, 1, 2, 1, 2, 1, 2, 1, 1, 1, 1, 1, 2, 1)   # reference imagenet/mobilenet.py:41-54
BN_EPS = 1e-5                                                    # nn.BatchNorm2d default


def _bn(x, sd, p, dtype):
    """Eval-mode BatchNorm2d (reference imagenet/mobilenet.py:25,32,36; models.py:57,66,73)."""
    g = sd[p + '.weight'].to(dtype); b = sd[p + '.bias'].to(dtype)
    m = sd[p + '.running_mean'].to(dtype); v = sd[p + '.running_var'].to(dtype)
    inv = g / torch.sqrt(v + BN_EPS)
    return x * inv.view(1, -1, 1, 1) + (b - m * inv).view(1, -1, 1, 1)


def stem(x, sd, dtype):
    """conv_bn(3, C0, stride 2) + ReLU6 (reference imagenet/mobilenet.py:22-27, 41)."""
    w = sd['conv0.0.weight'].to(dtype)
    return _bn(F.conv2d(x, w, None, stride=2, padding=1), sd, 'conv0.1', dtype).clamp(0.0, 6.0)


def encoder_dw(x, sd, i, dtype):
    """first half of conv_dw: dw3x3(stride, pad 1, groups=C) + BN + ReLU6
    (reference imagenet/mobilenet.py:31-33)."""
    w = sd['conv%d.0.weight' % i].to(dtype)
    y = F.conv2d(x, w, None, stride=ENCODER_STRIDES[i], padding=1, groups=w.shape[0])
    return _bn(y, sd, 'conv%d.1' % i, dtype).clamp(0.0, 6.0)


def encoder_pw(x, sd, i, dtype):
    """second half of conv_dw: pw1x1 + BN + ReLU6 (reference imagenet/mobilenet.py:35-37)."""
    w = sd['conv%d.3.weight' % i].to(dtype)
    return _bn(F.conv2d(x, w), sd, 'conv%d.4' % i, dtype).clamp(0.0, 6.0)


def decoder_dw(x, sd, j, dtype):
    """depthwise(C, 5): dw5x5 s1 p2 + BN + ReLU (reference models.py:61-68, 683-697)."""
    w = sd['decode_conv%d.0.0.weight' % j].to(dtype)
    k = w.shape[-1]
    y = F.conv2d(x, w, None, stride=1, padding=(k - 1) // 2, groups=w.shape[0])
    return _bn(y, sd, 'decode_conv%d.0.1' % j, dtype).clamp_min(0.0)


def decoder_pw(x, sd, j, dtype):
    """pointwise(C, C'): 1x1 + BN + ReLU (reference models.py:70-75, 683-697)."""
    w = sd['decode_conv%d.1.0.weight' % j].to(dtype)
    return _bn(F.conv2d(x, w), sd, 'decode_conv%d.1.1' % j, dtype).clamp_min(0.0)


def upsample2x(x):
    """F.interpolate(scale_factor=2, mode='nearest'): out[y,x] = in[y//2, x//2]
    (reference models.py:723)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def head(x, sd, dtype):
    """decode_conv6 = pointwise(C,1) (reference models.py:698, 731)."""
    w = sd['decode_conv6.0.weight'].to(dtype)
    return _bn(F.conv2d(x, w), sd, 'decode_conv6.1', dtype).clamp_min(0.0)


@torch.no_grad()
def skipadd_forward(sd, x, dtype=torch.float32, stages=None):
    """MobileNetSkipAdd.forward (reference models.py:706-732) on a state_dict.

    ``stages``: optional dict that receives every named child's output
    (``conv0..conv13``, ``decode_conv1..5`` = AFTER upsample(+skip), ``decode_conv6``), plus
    ``conv{i}.dw`` / ``decode_conv{j}.dw`` / ``decode_conv{j}.pw`` intermediates, all NCHW."""
    x = x.to(dtype)
    keep = {}
    x = stem(x, sd, dtype)
    if stages is not None:
        stages['conv0'] = x
    for i in range(1, 14):
        d = encoder_dw(x, sd, i, dtype)
        x = encoder_pw(d, sd, i, dtype)
        if stages is not None:
            stages['conv%d.dw' % i] = d
            stages['conv%d' % i] = x
        if i in (1, 3, 5):                       # reference models.py:714-719
            keep[i] = x
    add_after = {4: 1, 3: 3, 2: 5}               # reference models.py:724-729
    for j in range(1, 6):
        d = decoder_dw(x, sd, j, dtype)
        p = decoder_pw(d, sd, j, dtype)
        x = upsample2x(p)
        if j in add_after:
            x = x + keep[add_after[j]]
        if stages is not None:
            stages['decode_conv%d.dw' % j] = d
            stages['decode_conv%d.pw' % j] = p
            stages['decode_conv%d' % j] = x
    x = head(x, sd, dtype)
    if stages is not None:
        stages['decode_conv6'] = x
    return x


@torch.no_grad()
def skipconcat_forward(sd, x, dtype=torch.float32, stages=None):
    """MobileNetSkipConcat.forward (reference models.py:789-814): identical to skipadd_forward except that the saved
    encoder outputs are CONCATENATED after the upsampled decoder output (torch.cat((x, x1), 1), l.806-811)."""
    x = stem(x.to(dtype), sd, dtype)
    keep = {}
    if stages is not None:
        stages['conv0'] = x
    for i in range(1, 14):
        x = encoder_pw(encoder_dw(x, sd, i, dtype), sd, i, dtype)
        if stages is not None:
            stages['conv%d' % i] = x
        if i in (1, 3, 5):
            keep[i] = x
    cat_after = {4: 1, 3: 3, 2: 5}
    for j in range(1, 6):
        x = upsample2x(decoder_pw(decoder_dw(x, sd, j, dtype), sd, j, dtype))
        if stages is not None:
            stages['decode_conv%d' % j] = x            # the block's own slice (before the concatenation)
        if j in cat_after:
            x = torch.cat((x, keep[cat_after[j]]), 1)
    x = head(x, sd, dtype)
    if stages is not None:
        stages['decode_conv6'] = x
    return x


def to_skipadd_keys(sd):
    """state_dict of ``models.MobileNet(decoder='nnconv5dw')`` (keys ``mobilenet.<i>.*``, ``decoder.conv<j>.*``,
    reference models.py:441, 229-244) renamed to the MobileNetSkipAdd schema used by the functions above."""
    out = {}
    for k, v in sd.items():
        if k.startswith('mobilenet.'):
            i, rest = k[len('mobilenet.'):].split('.', 1)
            out['conv%s.%s' % (i, rest)] = v
        elif k.startswith('decoder.conv'):
            j, rest = k[len('decoder.conv'):].split('.', 1)
            out['decode_conv%s.%s' % (j, rest)] = v
    return out


@torch.no_grad()
def nnconv_dw_forward(sd, x, dtype=torch.float32, stages=None):
    """MobileNet.forward with the depthwise NNConv decoder (reference models.py:457-460 -> 253-270): the same 14
    encoder blocks, five (depthwise 5x5 + pointwise) blocks each followed by nearest x2, pointwise(32,1) -- no skips.
    ``sd`` uses the MobileNet key schema."""
    sk = to_skipadd_keys(sd)
    x = stem(x.to(dtype), sk, dtype)
    if stages is not None:
        stages['mobilenet.0'] = x
    for i in range(1, 14):
        x = encoder_pw(encoder_dw(x, sk, i, dtype), sk, i, dtype)
        if stages is not None:
            stages['mobilenet.%d' % i] = x
    for j in range(1, 6):
        x = upsample2x(decoder_pw(decoder_dw(x, sk, j, dtype), sk, j, dtype))
        if stages is not None:
            stages['decoder.conv%d' % j] = x
    return head(x, sk, dtype)


# --------------------------------------------------------------------------------------
# NYU val pre-processing (reference dataloaders/nyu.py:48-59)  -- PARITY UNPINNED for this function:
# the reference does the arithmetic with scipy.misc.imresize (dataloaders/transforms.py:337-339), removed in SciPy 1.3
# and absent here, so the reference's own transform cannot be executed.  scipy 1.2's imresize is restated from its
# published source: toimage(arr[, mode='F']) -> PIL.Image.resize(size, resample=NEAREST) -> fromimage, with a float
# `size` meaning "fraction of the current size" (size = (array(im.size) * size).astype(int)).
# --------------------------------------------------------------------------------------
def _imresize_nearest(arr, size):
    from PIL import Image
    im = Image.fromarray(arr, mode='F') if arr.ndim == 2 else Image.fromarray(arr)
    if isinstance(size, float):
        size = (int(im.size[0] * size), int(im.size[1] * size))          # (W, H)
    else:
        size = (size[1], size[0])
    return np.asarray(im.resize(size, resample=Image.NEAREST))


def nyu_val_transform(rgb_u8, depth, out_hw=(224, 224), iheight=480.0):
    """NYUDataset.val_transform (dataloaders/nyu.py:48-59) + ToTensor (dataloaders/dataloader.py:104-109) for ONE
    sample: rgb_u8 [H,W,3] uint8, depth [H,W] float32 -> (input [3,oh,ow] float32 in [0,1], target [1,oh,ow] float32)."""
    def chain(a):
        a = _imresize_nearest(a, 250.0 / iheight)                       # transforms.Resize(250.0 / iheight)
        h, w = a.shape[0], a.shape[1]
        i, j = int(round((h - 228) / 2.0)), int(round((w - 304) / 2.0))  # transforms.CenterCrop((228, 304))
        a = a[i:i + 228, j:j + 304]
        return _imresize_nearest(np.ascontiguousarray(a), tuple(out_hw))  # transforms.Resize(self.output_size)
    rgb = np.asarray(chain(rgb_u8), dtype='float') / 255                 # np.asfarray(rgb_np, dtype='float') / 255
    x = torch.from_numpy(rgb.transpose((2, 0, 1)).copy()).float()
    t = torch.from_numpy(chain(np.asarray(depth, dtype=np.float32)).copy()).float().unsqueeze(0)
    return x, t


# --------------------------------------------------------------------------------------
# metrics (reference metrics.py:31-55, 71-95)
# --------------------------------------------------------------------------------------
METRIC_NAMES = ('irmse', 'imae', 'mse', 'rmse', 'mae', 'absrel', 'lg10', 'delta1', 'delta2', 'delta3')


def evaluate_one(output, target):
    """Result.evaluate (reference metrics.py:31-55) on ONE call's tensors; returns a dict.
    Pools every valid pixel of the tensors it is given -- the reference calls it per image
    (batch size 1, main.py:40-41, 80-82)."""
    out = np.asarray(output, dtype=np.float32).reshape(-1)
    tgt = np.asarray(target, dtype=np.float32).reshape(-1)
    valid = (tgt > 0) | (out > 0)                      # metrics.py:32  ((t>0)+(o>0))>0
    o = np.float32(1e3) * out[valid]                   # metrics.py:34-35 -> millimetres
    t = np.float32(1e3) * tgt[valid]
    ad = np.abs(o - t)
    r = {}
    r['mse'] = float(np.mean(ad * ad, dtype=np.float32))
    r['rmse'] = math.sqrt(r['mse'])
    r['mae'] = float(np.mean(ad, dtype=np.float32))
    with np.errstate(divide='ignore', invalid='ignore'):
        ln10 = np.float32(math.log(10))
        r['lg10'] = float(np.mean(np.abs(np.log(o) / ln10 - np.log(t) / ln10), dtype=np.float32))
        r['absrel'] = float(np.mean(ad / t, dtype=np.float32))
        ratio = np.maximum(o / t, t / o)
        for k in (1, 2, 3):
            r['delta%d' % k] = float(np.mean((ratio < 1.25 ** k).astype(np.float32), dtype=np.float32))
        inv = np.abs(np.float32(1) / o - np.float32(1) / t)
        r['irmse'] = math.sqrt(float(np.mean(inv * inv, dtype=np.float32)))
        r['imae'] = float(np.mean(inv, dtype=np.float32))
    return r


def average_per_image(outputs, targets):
    """AverageMeter over per-image Result.evaluate calls (reference metrics.py:71-95 with
    main.py:80-82 at batch size 1): mean of per-image metrics, NOT pooled pixels."""
    sums = {k: 0.0 for k in METRIC_NAMES}
    n = 0
    for o, t in zip(outputs, targets):
        r = evaluate_one(o, t)
        for k in METRIC_NAMES:
            sums[k] += r[k]
        n += 1
    return {k: v / n for k, v in sums.items()}, n
