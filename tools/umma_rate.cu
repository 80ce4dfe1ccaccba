// Issue-rate microbenchmark of tcgen05.mma (kind::f16, K = 16 per instruction) on resident shared-memory operands:
//   * cta_group::1, M = 128, N = 256 / 128 / 64  (one CTA)
//   * cta_group::2, M = 256 over a CTA pair, N = 256 (each CTA holds half of B)
// prints cycles per instruction and MAC / clk / SM.  Operand contents are irrelevant (zeros).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdepth_b200/csrc -o tools/umma_rate tools/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <string>
#include "fd_tc_common.cuh"
namespace fd {
void set_error(const std::string&) {}
int fail(int code, const std::string& msg) { fprintf(stderr, "error: %s\n", msg.c_str()); return code; }
}
using namespace fd;

__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <bool TWO>
__global__ void __launch_bounds__(128, 1) rate_kernel(int n_cols, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) { if (TWO) tmem_alloc2(smem_u32(&tb), 512); else tmem_alloc(smem_u32(&tb), 512); }
    tc_fence_before();
    if (TWO) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t rank = TWO ? cluster_ctarank() : 0;
    if (threadIdx.x == 0 && rank == 0) {
        const uint32_t m = TWO ? 256u : 128u;
        const uint32_t idesc = (1u << 4) | ((uint32_t)(n_cols >> 3) << 17) | ((m >> 4) << 24);
        const uint32_t a_lo = sw128_desc_lo(base), b_lo = sw128_desc_lo(base + 16384);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (TWO) umma2(tb, a_lo + 2 * k, b_lo + 2 * k, kSw128DescHi, idesc, 1u);
                else umma_f16_lohi(tb, a_lo + 2 * k, b_lo + 2 * k, kSw128DescHi, idesc, 1u);
            }
        }
        if (TWO) commit2(smem_u32(&bar)); else umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        out[0] = t1 - t0;
    }
    tc_fence_before();
    if (TWO) cluster_sync_all(); else __syncthreads();
    if (warp == 0) { tc_fence_after(); if (TWO) tmem_dealloc2(tb, 512); else tmem_dealloc(tb, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    const int reps = 2000;
    cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int n : {256, 128, 64}) {
        rate_kernel<false><<<1, 128, 64 * 1024>>>(n, reps, d);
        long long c = 0; cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("cta_group::1 M128 N%-3d K16: %.1f cycles / instruction  (%.0f MAC/clk/SM)  [%s]\n", n, (double)c / (reps * 4), 128.0 * n * 16 * reps * 4 / (double)c, cudaGetErrorString(e));
    }
    for (int n : {256, 128}) {
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, rate_kernel<true>, n, reps, d);
        long long c = 0; cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("cta_group::2 M256 N%-3d K16: %.1f cycles / instruction  (%.0f MAC/clk/SM)  [%s]\n", n, (double)c / (reps * 4), 256.0 * n * 16 * reps * 4 / (double)c / 2, cudaGetErrorString(e));
    }
    return 0;
}
