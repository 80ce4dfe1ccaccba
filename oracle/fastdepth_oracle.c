/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C restatement of the ARITHMETIC of the hot path MobileNetSkipAdd.forward (reference models.py:706-732): the
 * reference delegates every operator to PyTorch (a dependency that is not vendored under /root/reference), and
 * oracle/fastdepth_oracle.py restates the path with the same PyTorch primitives.  This file restates the primitives
 * themselves from their published definitions as explicit loops -- no BLAS, no PyTorch -- so that the oracle does not
 * share an operator library with the thing it pins:
 *   fo_conv_dense     nn.Conv2d(ci, co, k, stride, pad, bias=False)            reference imagenet/mobilenet.py:24
 *   fo_conv_depthwise nn.Conv2d(c, c, k, stride, pad, groups=c, bias=False)    imagenet/mobilenet.py:31, models.py:64
 *   fo_conv_pointwise nn.Conv2d(ci, co, 1, 1, 0, bias=False)                   imagenet/mobilenet.py:35, models.py:71
 *   fo_bn_act         nn.BatchNorm2d in eval mode + ReLU / ReLU6               imagenet/mobilenet.py:25-26, models.py:66-67
 *   fo_upsample2x     F.interpolate(scale_factor=2, mode='nearest')            models.py:723
 *   fo_add            x + x1                                                   models.py:725
 * Tensors are contiguous NCHW float32 like the reference's; sums are accumulated in double (the reference's fp32 sum
 * order is an implementation detail of its BLAS; double is the order-independent answer to ~1e-7).
 * Composition into the network: oracle/c_oracle.py.  Pinned against the golden vectors produced by the live reference
 * (tests/test_oracle.py::test_c_oracle_*).  Built by oracle/build_oracle.py (gcc -O2 -shared), which __graft_entry__.build() calls. */
#include <math.h>
#include <stddef.h>

#define IDX4(n, c, y, x, C, H, W) ((((size_t)(n) * (C) + (c)) * (H) + (y)) * (W) + (x))

/* cross-correlation with symmetric zero padding, out[n][co][oy][ox] = sum_ci,ky,kx in[n][ci][oy*s+ky-p][ox*s+kx-p] * w[co][ci][ky][kx] */
void fo_conv_dense(const float* in, const float* w, float* out, int n, int ci, int h, int wd, int co, int k, int stride, int pad) {
    const int ho = (h + 2 * pad - k) / stride + 1, wo = (wd + 2 * pad - k) / stride + 1;
    for (int b = 0; b < n; ++b)
        for (int o = 0; o < co; ++o)
            for (int oy = 0; oy < ho; ++oy)
                for (int ox = 0; ox < wo; ++ox) {
                    double acc = 0.0;
                    for (int c = 0; c < ci; ++c)
                        for (int ky = 0; ky < k; ++ky) {
                            const int iy = oy * stride + ky - pad;
                            if (iy < 0 || iy >= h) continue;
                            for (int kx = 0; kx < k; ++kx) {
                                const int ix = ox * stride + kx - pad;
                                if (ix < 0 || ix >= wd) continue;
                                acc += (double)in[IDX4(b, c, iy, ix, ci, h, wd)] * (double)w[((size_t)(o * ci + c) * k + ky) * k + kx];
                            }
                        }
                    out[IDX4(b, o, oy, ox, co, ho, wo)] = (float)acc;
                }
}

/* groups == channels: every channel is convolved with its own k x k filter w[c][0][ky][kx] */
void fo_conv_depthwise(const float* in, const float* w, float* out, int n, int c, int h, int wd, int k, int stride, int pad) {
    const int ho = (h + 2 * pad - k) / stride + 1, wo = (wd + 2 * pad - k) / stride + 1;
    for (int b = 0; b < n; ++b)
        for (int ch = 0; ch < c; ++ch)
            for (int oy = 0; oy < ho; ++oy)
                for (int ox = 0; ox < wo; ++ox) {
                    double acc = 0.0;
                    for (int ky = 0; ky < k; ++ky) {
                        const int iy = oy * stride + ky - pad;
                        if (iy < 0 || iy >= h) continue;
                        for (int kx = 0; kx < k; ++kx) {
                            const int ix = ox * stride + kx - pad;
                            if (ix < 0 || ix >= wd) continue;
                            acc += (double)in[IDX4(b, ch, iy, ix, c, h, wd)] * (double)w[((size_t)ch * k + ky) * k + kx];
                        }
                    }
                    out[IDX4(b, ch, oy, ox, c, ho, wo)] = (float)acc;
                }
}

/* 1x1 convolution: a matrix product over channels at every pixel */
void fo_conv_pointwise(const float* in, const float* w, float* out, int n, int ci, int hw, int co) {
    for (int b = 0; b < n; ++b)
        for (int o = 0; o < co; ++o) {
            float* dst = out + ((size_t)b * co + o) * hw;
            for (int p = 0; p < hw; ++p) {
                double acc = 0.0;
                for (int c = 0; c < ci; ++c) acc += (double)in[((size_t)b * ci + c) * hw + p] * (double)w[(size_t)o * ci + c];
                dst[p] = (float)acc;
            }
        }
}

/* y = (x - mean) / sqrt(var + eps) * gamma + beta, then act: 0 none, 1 ReLU, 2 ReLU6.  In place. */
void fo_bn_act(float* x, const float* mean, const float* var, const float* gamma, const float* beta, float eps, int act, int n, int c, int hw) {
    for (int b = 0; b < n; ++b)
        for (int ch = 0; ch < c; ++ch) {
            const double inv = (double)gamma[ch] / sqrt((double)var[ch] + (double)eps);
            const double shift = (double)beta[ch] - (double)mean[ch] * inv;
            float* p = x + ((size_t)b * c + ch) * hw;
            for (int i = 0; i < hw; ++i) {
                double v = (double)p[i] * inv + shift;
                if (act >= 1 && v < 0.0) v = 0.0;
                if (act == 2 && v > 6.0) v = 6.0;
                p[i] = (float)v;
            }
        }
}

/* nearest neighbour, scale 2: out[y][x] = in[y / 2][x / 2] */
void fo_upsample2x(const float* in, float* out, int nc, int h, int w) {
    for (int p = 0; p < nc; ++p)
        for (int y = 0; y < 2 * h; ++y)
            for (int x = 0; x < 2 * w; ++x) out[((size_t)p * 2 * h + y) * 2 * w + x] = in[((size_t)p * h + y / 2) * w + x / 2];
}

void fo_add(float* x, const float* y, size_t count) {
    for (size_t i = 0; i < count; ++i) x[i] = x[i] + y[i];
}
