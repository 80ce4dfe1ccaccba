// Per-image depth metrics on device.
//
// Semantics follow the reference's Result.evaluate (metrics.py:31-55) applied to ONE image at a
// time, which is what its only caller does (DataLoader batch_size=1, main.py:40-41, 80-82), and
// AverageMeter's running sums (metrics.py:71-95): for every image the 10 metric values are added
// to sums[0..9] and 1 to sums[10].  The caller divides by the count after its single cross-GPU
// all-reduce.  Element math is fp32 like the reference; the per-image reductions
// use double accumulators (the reference's fp32 cascade sum agrees to ~1e-6 relative) and the per-image
// results are rounded to fp32, the precision the reference carries them in.
#include "fd_common.cuh"

namespace fd {

constexpr int MT_THREADS = 512;
constexpr int MT_NACC = 10;   // cnt, sum ad^2, sum ad, sum lg10, sum absrel, d1, d2, d3, sum inv^2, sum inv

template <typename T>
__global__ void __launch_bounds__(MT_THREADS)
metrics_kernel(const T* __restrict__ pred, const float* __restrict__ target, int hw, double* __restrict__ sums) {
    const int img = blockIdx.x;
    const T* p = pred + (size_t)img * hw;
    const float* t = target + (size_t)img * hw;
    double acc[MT_NACC];
#pragma unroll
    for (int i = 0; i < MT_NACC; ++i) acc[i] = 0.0;
    const float inv_ln10 = 1.0f / 2.302585092994046f;
    const float th1 = 1.25f, th2 = 1.25f * 1.25f, th3 = 1.25f * 1.25f * 1.25f;
    for (int i = threadIdx.x; i < hw; i += MT_THREADS) {
        const float o0 = Traits<T>::to_f(p[i]);
        const float t0 = t[i];
        if (!(t0 > 0.f || o0 > 0.f)) continue;                 // metrics.py:32
        const float o = 1e3f * o0, tt = 1e3f * t0;             // metrics.py:34-35
        const float ad = fabsf(o - tt);
        acc[0] += 1.0;
        acc[1] += (double)(ad * ad);
        acc[2] += (double)ad;
        acc[3] += (double)fabsf(logf(o) * inv_ln10 - logf(tt) * inv_ln10);
        acc[4] += (double)(ad / tt);
        const float ratio = fmaxf(o / tt, tt / o);
        acc[5] += ratio < th1 ? 1.0 : 0.0;
        acc[6] += ratio < th2 ? 1.0 : 0.0;
        acc[7] += ratio < th3 ? 1.0 : 0.0;
        const float inv = fabsf(1.0f / o - 1.0f / tt);
        acc[8] += (double)(inv * inv);
        acc[9] += (double)inv;
    }
    __shared__ double red[MT_NACC][MT_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < MT_NACC; ++i) {
        double v = acc[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[i][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[MT_NACC];
        for (int i = 0; i < MT_NACC; ++i) {
            double v = 0.0;
            for (int w = 0; w < MT_THREADS / 32; ++w) v += red[i][w];
            tot[i] = v;
        }
        const double cnt = tot[0];
        const double mse = tot[1] / cnt;
        const double vals[10] = {sqrt(tot[8] / cnt), tot[9] / cnt, mse,          sqrt(mse),     tot[2] / cnt,
                                 tot[4] / cnt,       tot[3] / cnt, tot[5] / cnt, tot[6] / cnt, tot[7] / cnt};
        // The reference's per-image values are fp32 (float(tensor.mean()), metrics.py:38-55) and AverageMeter adds them up in
        // double: round every per-image value to fp32 before the fp64 accumulation.  A sum of 24-bit values of similar
        // magnitude is then EXACT in fp64, i.e. independent of the order of the atomics and of how images are partitioned
        // over ranks -- the N-rank sums equal the 1-rank sums bit for bit.
        for (int i = 0; i < 10; ++i) atomicAdd(&sums[i], (double)(float)vals[i]);
        atomicAdd(&sums[10], 1.0);
    }
}

int launch_metrics(int dtype, const void* pred, const float* target, int n, int hw, double* sums, cudaStream_t st) {
    if (n <= 0) return FD_OK;
    switch (dtype) {
        case FD_F32: metrics_kernel<float><<<n, MT_THREADS, 0, st>>>((const float*)pred, target, hw, sums); break;
        case FD_F16: metrics_kernel<__half><<<n, MT_THREADS, 0, st>>>((const __half*)pred, target, hw, sums); break;
        case FD_BF16:
            metrics_kernel<__nv_bfloat16><<<n, MT_THREADS, 0, st>>>((const __nv_bfloat16*)pred, target, hw, sums);
            break;
        default: return fail(FD_ERR_INVALID, "bad dtype");
    }
    FD_CUDA_OK(cudaGetLastError());
    return FD_OK;
}

// ----------------------------------------------------------------------------------------------
// NYU val pre-processing as one gather (reference dataloaders/nyu.py:48-59: Resize -> CenterCrop -> Resize, all
// nearest-neighbour, then rgb / 255; dataloaders/dataloader.py:90-111: HWC -> CHW float).  rows/cols: composed source
// index tables.  Thread = output pixel; NCHW planes are written coalesced along x.
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
nyu_val_gather_kernel(const uint8_t* __restrict__ rgb, const float* __restrict__ depth, const int* __restrict__ rows,
                      const int* __restrict__ cols, int n, int h_in, int w_in, int oh, int ow, T* __restrict__ x,
                      float* __restrict__ t) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * oh * ow;
    if (idx >= total) return;
    const int ox = (int)(idx % ow);
    const int oy = (int)((idx / ow) % oh);
    const int img = (int)(idx / ((long long)ow * oh));
    const size_t src = ((size_t)img * h_in + rows[oy]) * w_in + cols[ox];
    const size_t plane = (size_t)oh * ow;
    T* xo = x + (size_t)img * 3 * plane + (size_t)oy * ow + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c)      // np.asfarray(rgb, 'float') / 255 (double), then .float()
        xo[c * plane] = Traits<T>::from_f((float)((double)rgb[src * 3 + c] / 255.0));
    if (t != nullptr) t[(size_t)img * plane + (size_t)oy * ow + ox] = depth[src];
}

int launch_nyu_val_gather(int dtype, const uint8_t* rgb, const float* depth, const int* rows, const int* cols, int n, int h_in,
                          int w_in, int oh, int ow, void* x, float* t, cudaStream_t st) {
    const long long total = (long long)n * oh * ow;
    if (total <= 0) return FD_OK;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    switch (dtype) {
        case FD_F32: nyu_val_gather_kernel<float><<<blocks, 256, 0, st>>>(rgb, depth, rows, cols, n, h_in, w_in, oh, ow, (float*)x, t); break;
        case FD_F16: nyu_val_gather_kernel<__half><<<blocks, 256, 0, st>>>(rgb, depth, rows, cols, n, h_in, w_in, oh, ow, (__half*)x, t); break;
        case FD_BF16: nyu_val_gather_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(rgb, depth, rows, cols, n, h_in, w_in, oh, ow, (__nv_bfloat16*)x, t); break;
        default: return fail(FD_ERR_INVALID, "bad dtype");
    }
    FD_CUDA_OK(cudaGetLastError());
    return FD_OK;
}

}  // namespace fd
