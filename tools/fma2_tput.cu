// microbenchmark: the depthwise inner product  acc[px] += in[px + tap] * w[tap]  for a channel PAIR per lane, written as
//   mode 0: two FHFMA (fma.rn.f32.f16 with .H0/.H1 operands) per pixel-tap           (what block_tc_kernel does)
//   mode 1: inputs / taps converted to fp32 pairs once, one FFMA2 (fma.rn.f32x2) per pixel-tap
// 8 warps per SM (two per scheduler) like the kernel; 4 output pixels x 3 taps per input row of 6, 16 accumulator pairs.
// CAVEAT: the asm statements are `volatile`, so the compiler keeps them in program order and both variants run latency-bound;
// the 2.0x this prints (1383 vs 697 cycles) is the instruction-count ratio, not a pipe-throughput ratio.  fma2_tput5.cu has
// the schedulable (non-volatile) version under the kernel's register cap; the in-kernel A/B (FD_DW3_FHFMA build) is +1.2 %.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ void fhfma2(float& lo, float& hi, unsigned a, unsigned b) {
    asm volatile("{.reg .f16 al,ah,bl,bh; mov.b32 {al,ah}, %2; mov.b32 {bl,bh}, %3; fma.rn.f32.f16 %0, al, bl, %0; fma.rn.f32.f16 %1, ah, bh, %1;}"
                 : "+f"(lo), "+f"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long cvt2(unsigned h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return ((unsigned long long)__float_as_uint(f.y) << 32) | __float_as_uint(f.x);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(const unsigned* __restrict__ src, float* out, int iters, long long* cyc) {
    __shared__ unsigned tile[36 * 32 + 64];
    for (int i = threadIdx.x; i < 36 * 32 + 64; i += blockDim.x) tile[i] = src[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float a[4][4][2];
    unsigned long long a2[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j][0] = a[i][j][1] = 0.f; a2[i][j] = 0ull; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            unsigned w[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) w[i] = tile[36 * 32 + i] + it;
#pragma unroll
            for (int iy = 0; iy < 6; ++iy) {
                unsigned row[6];
#pragma unroll
                for (int ix = 0; ix < 6; ++ix) row[ix] = tile[(iy * 6 + ix) * 32 + lane];
#pragma unroll
                for (int oy = 0; oy < 4; ++oy) {
                    const int ky = iy - oy;
                    if (ky < 0 || ky >= 3) continue;
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) fhfma2(a[oy][ox][0], a[oy][ox][1], row[ox + kx], w[ky * 3 + kx]);
                }
            }
        } else {
            unsigned long long w[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) w[i] = cvt2(tile[36 * 32 + i] + it);
#pragma unroll
            for (int iy = 0; iy < 6; ++iy) {
                unsigned long long row[6];
#pragma unroll
                for (int ix = 0; ix < 6; ++ix) row[ix] = cvt2(tile[(iy * 6 + ix) * 32 + lane]);
#pragma unroll
                for (int oy = 0; oy < 4; ++oy) {
                    const int ky = iy - oy;
                    if (ky < 0 || ky >= 3) continue;
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) ffma2(a2[oy][ox], row[ox + kx], w[ky * 3 + kx]);
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += a[i][j][0] + a[i][j][1] + __uint_as_float((unsigned)a2[i][j]) + __uint_as_float((unsigned)(a2[i][j] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; unsigned* src;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 8); cudaMalloc(&src, 8192); cudaMemset(src, 0x3c, 8192);
    const int iters = 2000;
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int m = 0; m < 2; ++m) {
            for (int rep = 0; rep < 2; ++rep) {
                if (m == 0) k<0><<<148, warps * 32>>>(src, out, iters, cyc);
                if (m == 1) k<1><<<148, warps * 32>>>(src, out, iters, cyc);
                cudaDeviceSynchronize();
            }
            printf("%-26s warps/SM %2d : %.0f cycles per 4x4x3x3 unit (%d warps -> %.0f per scheduler-unit)\n",
                   m == 0 ? "2 x FHFMA per pixel-tap" : "cvt + FFMA2 per pixel-tap", warps, (double)*cyc / iters, warps, (double)*cyc / iters / (warps / 4.0));
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
