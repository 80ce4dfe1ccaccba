// microbenchmark: 5x5 depthwise inner product for a 4x4 pixel block and a channel PAIR per lane, under the fused kernel's
// register cap (576-thread launch bound -> 96 registers), 8 working warps per SM (two per scheduler):
//   mode 0: two FHFMA per pixel-tap, one pass (what block_tc_kernel<KS=5> does)
//   mode 1: FFMA2, two passes of 4 rows x 2 columns, taps re-loaded + widened lazily per pass   (tried in the kernel: slower)
//   mode 2: FFMA2, two passes of 2 rows x 4 columns, all 25 taps widened once and kept (50 registers)
//   mode 3: FFMA2 for kernel rows 0..2 (15 taps kept as fp32 pairs) + FHFMA for rows 3..4 (10 taps as 16-bit words), one pass
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ void fhfma2(float& lo, float& hi, unsigned a, unsigned b) {
    asm("{.reg .f16 al,ah,bl,bh; mov.b32 {al,ah}, %2; mov.b32 {bl,bh}, %3; fma.rn.f32.f16 %0, al, bl, %0; fma.rn.f32.f16 %1, ah, bh, %1;}"
        : "+f"(lo), "+f"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ void ffma2(f32x2& acc, f32x2 a, f32x2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ f32x2 widen(unsigned h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return ((f32x2)__float_as_uint(f.y) << 32) | __float_as_uint(f.x);
}
__device__ __forceinline__ float lo32(f32x2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(f32x2 v) { return __uint_as_float((unsigned)(v >> 32)); }

template <int MODE>
__global__ void __launch_bounds__(576, 1) k(const unsigned* __restrict__ src, float* out, int iters, long long* cyc) {
    __shared__ unsigned tile[64 * 32 + 64], wts[25 * 32 + 64];
    for (int i = threadIdx.x; i < 64 * 32 + 64; i += blockDim.x) tile[i] = src[i];
    for (int i = threadIdx.x; i < 25 * 32 + 64; i += blockDim.x) wts[i] = src[i + 2200];
    __syncthreads();
    if (threadIdx.x >= 256) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("" ::: "memory");                      // shared memory must be re-read: nothing is hoisted out of the loop
        const unsigned* in = tile + lane + (it & 1) * 32;       // addresses change every iteration
        const unsigned* wp = wts + lane + (it & 1) * 32;
        if (MODE == 0) {
            unsigned w[25];
#pragma unroll
            for (int i = 0; i < 25; ++i) w[i] = wp[i * 32];
            float a[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) a[i][j][0] = a[i][j][1] = 0.f;
#pragma unroll
            for (int iy = 0; iy < 8; ++iy) {
                unsigned row[8];
#pragma unroll
                for (int ix = 0; ix < 8; ++ix) row[ix] = in[(iy * 8 + ix) * 32];
#pragma unroll
                for (int oy = 0; oy < 4; ++oy) {
                    const int ky = iy - oy;
                    if (ky < 0 || ky >= 5) continue;
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                        for (int kx = 0; kx < 5; ++kx) fhfma2(a[oy][ox][0], a[oy][ox][1], row[ox + kx], w[ky * 5 + kx]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s += a[i][j][0] + a[i][j][1];
        } else if (MODE == 1) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                f32x2 a[4][2];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i][0] = a[i][1] = 0ull;
                f32x2 wq[5][5];
#pragma unroll
                for (int iy = 0; iy < 8; ++iy) {
                    if (iy < 5) {
#pragma unroll
                        for (int kx = 0; kx < 5; ++kx) wq[iy][kx] = widen(wp[(iy * 5 + kx) * 32]);
                    }
                    f32x2 row[6];
#pragma unroll
                    for (int ix = 0; ix < 6; ++ix) row[ix] = widen(in[(iy * 8 + pass * 2 + ix) * 32]);
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy) {
                        const int ky = iy - oy;
                        if (ky < 0 || ky >= 5) continue;
#pragma unroll
                        for (int ox = 0; ox < 2; ++ox)
#pragma unroll
                            for (int kx = 0; kx < 5; ++kx) ffma2(a[oy][ox], row[ox + kx], wq[ky][kx]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) s += lo32(a[i][0]) + hi32(a[i][0]) + lo32(a[i][1]) + hi32(a[i][1]);
            }
        } else if (MODE == 2) {
            f32x2 wq[25];
#pragma unroll
            for (int i = 0; i < 25; ++i) wq[i] = widen(wp[i * 32]);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {            // output rows 2*pass, 2*pass+1 <- input rows 2*pass .. 2*pass+5
                f32x2 a[2][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) a[0][j] = a[1][j] = 0ull;
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const int iy = 2 * pass + r;
#pragma unroll
                    for (int ix = 0; ix < 8; ++ix) {
                        const f32x2 v = widen(in[(iy * 8 + ix) * 32]);
#pragma unroll
                        for (int o = 0; o < 2; ++o) {
                            const int ky = r - o;
                            if (ky < 0 || ky >= 5) continue;
#pragma unroll
                            for (int ox = 0; ox < 4; ++ox) {
                                const int kx = ix - ox;
                                if (kx < 0 || kx >= 5) continue;
                                ffma2(a[o][ox], v, wq[ky * 5 + kx]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) s += lo32(a[0][j]) + hi32(a[0][j]) + lo32(a[1][j]) + hi32(a[1][j]);
            }
        } else {
            f32x2 wq[15];
            unsigned wh[10];
#pragma unroll
            for (int i = 0; i < 15; ++i) wq[i] = widen(wp[i * 32]);
#pragma unroll
            for (int i = 0; i < 10; ++i) wh[i] = wp[(15 + i) * 32];
            float a[4][4][2];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) a[i][j][0] = a[i][j][1] = 0.f;
#pragma unroll
            for (int iy = 0; iy < 8; ++iy) {
#pragma unroll
                for (int ix = 0; ix < 8; ++ix) {
                    const unsigned hw = in[(iy * 8 + ix) * 32];
                    const f32x2 v = widen(hw);
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy) {
                        const int ky = iy - oy;
                        if (ky < 0 || ky >= 5) continue;
#pragma unroll
                        for (int ox = 0; ox < 4; ++ox) {
                            const int kx = ix - ox;
                            if (kx < 0 || kx >= 5) continue;
                            if (ky < 3) {
                                f32x2 t = ((f32x2)__float_as_uint(a[oy][ox][1]) << 32) | __float_as_uint(a[oy][ox][0]);
                                ffma2(t, v, wq[ky * 5 + kx]);
                                a[oy][ox][0] = lo32(t); a[oy][ox][1] = hi32(t);
                            } else {
                                fhfma2(a[oy][ox][0], a[oy][ox][1], hw, wh[(ky - 3) * 5 + kx]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s += a[i][j][0] + a[i][j][1];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; unsigned* src;
    cudaMalloc(&out, 148 * 256 * 4); cudaMallocManaged(&cyc, 8); cudaMalloc(&src, 16384); cudaMemset(src, 0x3c, 16384);
    const int iters = 1000;
    const char* names[4] = {"FHFMA one pass", "FFMA2 2 x (4x2), lazy taps", "FFMA2 2 x (2x4), 25 taps kept", "FFMA2 rows 0-2 + FHFMA rows 3-4"};
    for (int m = 0; m < 4; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            if (m == 0) k<0><<<148, 576>>>(src, out, iters, cyc);
            if (m == 1) k<1><<<148, 576>>>(src, out, iters, cyc);
            if (m == 2) k<2><<<148, 576>>>(src, out, iters, cyc);
            if (m == 3) k<3><<<148, 576>>>(src, out, iters, cyc);
            cudaDeviceSynchronize();
        }
        printf("%-34s : %.0f cycles per 4x4 block and warp (two warps per scheduler)   %s\n", names[m], (double)*cyc / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
