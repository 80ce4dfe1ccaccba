"""Deterministic synthetic weights and inputs for parity tests, smoke() and bench.py.

Trained FastDepth weights are download links in the reference (README.md:26-46) and there is
no network, so every measurement here uses random-init weights of the named architecture.
Plain default init collapses activations to ~1e-3 and makes relative-error parity vacuous
(SURVEY.md section 4 "synthetic-weight pitfalls"), hence this recipe:

* conv weights: encoder N(0, sqrt(2/(k*k*Cout))) as ``weights_init`` would (reference
  models.py:36-41); decoder U(-b, b), b = 1/sqrt(fan_in) (PyTorch default, which the SkipAdd
  decoder keeps because of the ``weights_init``-on-Sequential no-op, reference models.py:699-704);
* BN: running_mean ~ N(0, 0.1), running_var ~ U(0.25, 0.75), gamma ~ U(0.5, 3), beta ~ N(0.2, 0.3);
  per-block gain re-normalisation keeps activations alive through 20 stages;
* last BN (decode_conv6): gamma = 1, beta = 0.5 so the final ReLU is not identically zero.

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` in a fixed order, so the values
are identical on every machine and torch version (the golden fixtures depend on that).
"""
import numpy as np
import torch

# NetAdapt-pruned widths recovered from the reference's TVM tuning log
# (tvm_compile/tuning/tx2-gpu.mobilenet-nnconv5dw-skipadd-pruned.trials=2000.stop=600.log,
#  lines 38..1 in network order).
PRUNED_ENCODER = (16, 56, 88, 120, 144, 256, 408, 376, 272, 288, 296, 328, 480, 512)
PRUNED_DECODER = (200, 256, 120, 56, 16)
PRUNED_WIDTHS = (PRUNED_ENCODER, PRUNED_DECODER)

STOCK_ENCODER = (32, 64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 1024, 1024)
STOCK_DECODER = (512, 256, 128, 64, 32)
STOCK_WIDTHS = (STOCK_ENCODER, STOCK_DECODER)


def _bn(rng, c, out, prefix, last=False):
    if last:
        gamma = np.ones(c); beta = np.full(c, 0.5)
        mean = rng.normal(0.0, 0.05, c); var = rng.uniform(0.5, 0.75, c)
    else:
        gamma = rng.uniform(0.5, 3.0, c); beta = rng.normal(0.2, 0.3, c)
        mean = rng.normal(0.0, 0.1, c); var = rng.uniform(0.25, 0.75, c)
    out[prefix + '.weight'] = gamma.astype(np.float32)
    out[prefix + '.bias'] = beta.astype(np.float32)
    out[prefix + '.running_mean'] = mean.astype(np.float32)
    out[prefix + '.running_var'] = var.astype(np.float32)
    out[prefix + '.num_batches_tracked'] = np.asarray(0, dtype=np.int64)


def synthetic_state_dict(widths=STOCK_WIDTHS, seed=1, gain=1.0):
    """state_dict (torch fp32 CPU tensors) with the MobileNetSkipAdd key schema
    (SURVEY.md section 8a-a2)."""
    enc, dec = widths
    rng = np.random.Generator(np.random.PCG64(seed))
    sd = {}

    def gauss(shape, k, c_out, g):
        return (rng.normal(0.0, g * np.sqrt(2.0 / (k * k * c_out)), shape)).astype(np.float32)

    def unif(shape, fan_in, g):
        b = g / np.sqrt(fan_in)
        return rng.uniform(-b, b, shape).astype(np.float32)

    # stem: conv0.0 / conv0.1
    sd['conv0.0.weight'] = gauss((enc[0], 3, 3, 3), 3, enc[0], 3.0 * gain)
    _bn(rng, enc[0], sd, 'conv0.1')
    for i in range(1, 14):
        c_in, c_out = enc[i - 1], enc[i]
        # depthwise fan-in is 9: weights_init's sqrt(2/(9*C)) starves it, so lift the gain
        sd['conv%d.0.weight' % i] = gauss((c_in, 1, 3, 3), 3, c_in, gain * np.sqrt(c_in) * 0.5)
        _bn(rng, c_in, sd, 'conv%d.1' % i)
        sd['conv%d.3.weight' % i] = gauss((c_out, c_in, 1, 1), 1, c_out, 0.6 * gain * np.sqrt(c_out / c_in))
        _bn(rng, c_out, sd, 'conv%d.4' % i)
    c = enc[13]
    for j, c_out in enumerate(dec, start=1):
        sd['decode_conv%d.0.0.weight' % j] = unif((c, 1, 5, 5), 25, 0.8 * gain)
        _bn(rng, c, sd, 'decode_conv%d.0.1' % j)
        sd['decode_conv%d.1.0.weight' % j] = unif((c_out, c, 1, 1), c, 0.9 * gain)
        _bn(rng, c_out, sd, 'decode_conv%d.1.1' % j)
        c = c_out
    sd['decode_conv6.0.weight'] = np.abs(unif((1, c, 1, 1), c, 1.5 * gain))   # positive head: depth-like, not half-dead
    _bn(rng, 1, sd, 'decode_conv6.1', last=True)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


def synthetic_input(n, h, w, seed=0):
    """[n,3,h,w] fp32 in [0,1) -- the range the reference pipeline yields
    (dataloaders/transforms.py:216-224, dataloaders/nyu.py:56)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy(rng.random((n, 3, h, w), dtype=np.float32))


def synthetic_target(pred, seed=1):
    """Strictly positive pseudo ground truth around an oracle prediction:
    t = pred * (1 + 0.1*randn), clamped to >= 1e-3 so metrics.py's masks/logs stay finite."""
    rng = np.random.Generator(np.random.PCG64(seed))
    noise = torch.from_numpy(rng.standard_normal(tuple(pred.shape)).astype(np.float32))
    return (pred.float().cpu() * (1.0 + 0.1 * noise)).clamp_min(1e-3)
