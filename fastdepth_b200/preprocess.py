"""NYU-Depth-v2 validation pre-processing on the GPU (the step on the other side of the
boundary, reference dataloaders/nyu.py:48-59 + dataloaders/transforms.py:311-341, 344-405 + dataloaders/dataloader.py:90-111).

The reference pipeline is three nearest-neighbour gathers and a scale:
    Resize(250/480)  ->  CenterCrop((228, 304))  ->  Resize((224, 224))  ->  rgb / 255
implemented with ``scipy.misc.imresize(..., 'nearest')`` (a thin wrapper over ``PIL.Image.resize(NEAREST)``; the function
was removed from SciPy 1.3, so the reference's own loader no longer runs on a current stack).  Because every step only
selects source pixels, the whole chain is ONE gather through a row table and a column table; the tables are produced by
pushing coordinate ramps through PIL's own nearest resize (so they are PIL's behaviour by construction) and the gather +
/255 + NHWC->NCHW + dtype conversion is one kernel launch (``fd_nyu_val_gather``).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .plan import DTYPE_CODE

_TABLES = {}


def _pil_nearest(arr_f32, size_wh):
    from PIL import Image
    return np.asarray(Image.fromarray(arr_f32, mode='F').resize(size_wh, resample=Image.NEAREST))


def nyu_val_index_maps(h_in=480, w_in=640, out_hw=(224, 224), iheight=480.0):
    """(rows[out_h], cols[out_w]) int32: source pixel of every output pixel for the reference's val transform."""
    key = (h_in, w_in, tuple(out_hw), iheight)
    if key not in _TABLES:
        frac = 250.0 / iheight                                  # transforms.Resize(250.0 / iheight), nyu.py:51
        size1 = (int(w_in * frac), int(h_in * frac))            # imresize with a float: (array(im.size) * size).astype(int)
        th, tw = 228, 304                                       # transforms.CenterCrop((228, 304)), nyu.py:52
        maps = []
        for axis in (0, 1):
            ramp = np.arange(h_in if axis == 0 else w_in, dtype=np.float32)
            img = np.repeat(ramp[:, None], w_in, 1) if axis == 0 else np.repeat(ramp[None, :], h_in, 0)
            a = _pil_nearest(np.ascontiguousarray(img), size1)
            i = int(round((a.shape[0] - th) / 2.0)); j = int(round((a.shape[1] - tw) / 2.0))   # CenterCrop.get_params
            a = a[i:i + th, j:j + tw]
            a = _pil_nearest(np.ascontiguousarray(a), (out_hw[1], out_hw[0]))                # Resize(output_size), nyu.py:53
            maps.append((a[:, 0] if axis == 0 else a[0, :]).astype(np.int32))
        _TABLES[key] = tuple(maps)
    return _TABLES[key]


def nyu_val_transform(rgb_u8, depth, out_hw=(224, 224), dtype=torch.float32):
    """rgb_u8: [n,H,W,3] uint8 CUDA, depth: [n,H,W] float32 CUDA (or None) -> (input [n,3,oh,ow] ``dtype`` in [0,1],
    target [n,1,oh,ow] float32) exactly as NYUDataset(split='val') + ToTensor produce them, one kernel launch."""
    lib = _lib.load()
    assert rgb_u8.is_cuda and rgb_u8.dtype == torch.uint8 and rgb_u8.dim() == 4 and rgb_u8.shape[3] == 3
    n, h_in, w_in, _ = rgb_u8.shape
    rows, cols = nyu_val_index_maps(h_in, w_in, out_hw)
    dev = rgb_u8.device
    rows_d = torch.from_numpy(rows).to(dev)
    cols_d = torch.from_numpy(cols).to(dev)
    x = torch.empty((n, 3, out_hw[0], out_hw[1]), dtype=dtype, device=dev)
    t = torch.empty((n, 1, out_hw[0], out_hw[1]), dtype=torch.float32, device=dev) if depth is not None else None
    rgb_u8 = rgb_u8.contiguous()
    if depth is not None:
        depth = depth.contiguous().float()
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(lib.fd_nyu_val_gather(rgb_u8.data_ptr(), depth.data_ptr() if depth is not None else None,
                                     rows_d.data_ptr(), cols_d.data_ptr(), n, h_in, w_in, out_hw[0], out_hw[1],
                                     DTYPE_CODE[dtype], x.data_ptr(), t.data_ptr() if t is not None else None,
                                     dev.index, stream))
    return x, t
