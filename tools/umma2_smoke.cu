// Smoke test + microbenchmark of the mechanisms the 2-CTA chain kernel relies on (run on a B200):
//   * cluster of 2 CTAs, tcgen05.alloc / mma / commit with cta_group::2 (M = 256 over the CTA pair, N = 256 per instruction, each CTA
//     holding HALF of the B tile), accumulators in both CTAs' TMEM
//   * the A operand written by ordinary shared-memory stores of each CTA's own threads (128B-swizzled K-major rows), made visible with
//     fence.proxy.async and signalled to the LEADER's mbarrier with a remote arrive
//   * B halves loaded by each CTA's own TMA with the completion posted on the leader's mbarrier (cp.async.bulk.tensor ... cta_group::2)
//   * tcgen05.commit ... multicast::cluster to barriers in both CTAs
// and of the TMEM drain rate (tcgen05.ld 32x32b.x32 by 4 / 8 / 16 warps).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fastdepth_b200/csrc -o tools/umma2_smoke tools/umma2_smoke.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "fd_tc_common.cuh"

namespace fd {
void set_error(const std::string&) {}
int fail(int code, const std::string& msg) { fprintf(stderr, "error: %s\n", msg.c_str()); return code; }
}
using namespace fd;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAITC_DONE;\n\tbra WAITC_LOOP;\n\tWAITC_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc),
        "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

constexpr int KB = 2;                  // K-blocks of 64
struct Smem {
    uint64_t a_full[KB], b_full[KB], acc_full;
    uint32_t tmem_base, pad;
};

// A: [256][K] fp16 row-major (K = 64*KB), W: [256][K] fp16 (N x K), D: [256][256] fp32.  One cluster of 2 CTAs, 8 "producer" warps + 1 TMA
// warp + 1 MMA warp each; CTA r owns rows 128r.. of A/D and loads rows 128r.. of W.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
umma2_kernel(const __grid_constant__ CUtensorMap tm_w, const __half* __restrict__ A, float* __restrict__ D, int K, long long* cyc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    // [A kb0][A kb1][B kb0][B kb1][Smem]
    const uint32_t a_off = 0, b_off = KB * 16384;
    Smem* sm = reinterpret_cast<Smem*>(smem + b_off + KB * 16384);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    if (threadIdx.x == 0) {
        for (int i = 0; i < KB; ++i) { mbar_init(smem_u32(&sm->a_full[i]), 16); mbar_init(smem_u32(&sm->b_full[i]), 1); }
        mbar_init(smem_u32(&sm->acc_full), 1);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc2(smem_u32(&sm->tmem_base), 256);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = sm->tmem_base;
    const long long t0 = clock64();
    if (warp < 8) {
        // producers: warp w writes rows [16w, 16w+16) of this CTA's A tile for every K-block; lane = channel pair
        for (int kb = 0; kb < KB; ++kb) {
            uint8_t* a_s = smem + a_off + kb * 16384;
            for (int r = 0; r < 16; ++r) {
                const int m = warp * 16 + r;
                const uint32_t v = *reinterpret_cast<const uint32_t*>(A + (size_t)(rank * 128 + m) * K + kb * 64 + lane * 2);
                *reinterpret_cast<uint32_t*>(a_s + m * 128 + (((lane >> 2) ^ (m & 7)) << 4) + ((lane & 3) << 2)) = v;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa(smem_u32(&sm->a_full[kb]), 0));
        }
    } else if (warp == 8) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                if (rank == 0) mbar_expect_tx(smem_u32(&sm->b_full[kb]), 2 * 16384);
                tma_load_2d_2sm(base + b_off + kb * 16384, &tm_w, mapa(smem_u32(&sm->b_full[kb]), 0), kb * 64, (int)rank * 128);
            }
        }
    } else if (warp == 9 && rank == 0) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
            for (int kb = 0; kb < KB; ++kb) {
                mbar_wait_cluster(smem_u32(&sm->a_full[kb]), 0);
                mbar_wait_cluster(smem_u32(&sm->b_full[kb]), 0);
                tc_fence_after();
                const uint32_t a_lo = sw128_desc_lo(base + a_off + kb * 16384), b_lo = sw128_desc_lo(base + b_off + kb * 16384);
                for (int k = 0; k < 4; ++k) umma2_f16(tmem_base, a_lo + 2 * k, b_lo + 2 * k, kSw128DescHi, idesc, (kb | k) ? 1u : 0u);
            }
            umma2_commit_mc(smem_u32(&sm->acc_full), 3);
        }
    }
    if (warp < 4) {
        // epilogue (both CTAs): warp q reads lanes 32q.. of its own TMEM: D rows rank*128 + 32q + lane
        mbar_wait(smem_u32(&sm->acc_full), 0);
        tc_fence_after();
        const int row = rank * 128 + warp * 32 + lane;
        for (int c = 0; c < 256; c += 32) {
            uint32_t r[32];
            tmem_ld32_sync(tmem_base + ((uint32_t)(warp * 32) << 16) + c, r);
            for (int j = 0; j < 32; ++j) D[(size_t)row * 256 + c + j] = __uint_as_float(r[j]);
        }
    }
    if (threadIdx.x == 0 && cyc) cyc[rank] = clock64() - t0;
    tc_fence_before();
    cluster_sync_all();
    if (warp == 9) { tc_fence_after(); tmem_dealloc2(tmem_base, 256); }
}

// TMEM drain rate: W warps read all 512 columns of the 128 lanes (warp w: lane quarter w % 4, columns split over the W / 4 warps of a quarter)
__global__ void __launch_bounds__(512, 1) tmem_rate_kernel(int nwarps, int reps, long long* out, uint32_t* sink) {
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(smem_u32(&tb), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tb;
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < nwarps) {
        const int q = warp & 3, part = warp >> 2, parts = nwarps >> 2;
        const int c0 = part * (512 / parts), c1 = c0 + 512 / parts;
        for (int rep = 0; rep < reps; ++rep)
            for (int c = c0; c < c1; c += 32) {
                uint32_t r[32];
                tmem_ld32_sync(tmem + ((uint32_t)(q * 32) << 16) + c, r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= r[j];
            }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / reps;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    const int K = 64 * KB;
    std::vector<__half> hA(256 * K), hW(256 * K);
    srand(1);
    for (auto& v : hA) v = __float2half((rand() % 17 - 8) / 8.0f);
    for (auto& v : hW) v = __float2half((rand() % 13 - 6) / 16.0f);
    __half *dA, *dW; float* dD; long long* dc;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&dD, 256 * 256 * 4); cudaMalloc(&dc, 64);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 256 * 256 * 4);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || !fnp) { printf("no encoder\n"); return 2; }
    PFN_encodeTiled enc = (PFN_encodeTiled)fnp;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)K, 256}; cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128}; cuuint32_t estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dW, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 2; }
    const size_t smem = 4 * 16384 + sizeof(Smem) + 1024;
    cudaFuncSetAttribute(umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma2_kernel<<<2, 320, smem>>>(tm, dA, dD, K, dc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("umma2 kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hD(256 * 256);
    long long hc[2];
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int m = 0; m < 256; ++m)
        for (int n = 0; n < 256; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)__half2float(hA[m * K + k]) * (double)__half2float(hW[n * K + k]);
            const double err = fabs(ref - hD[m * 256 + n]);
            if (!(err <= 1e-3)) { if (bad < 5) printf("  mismatch D[%d][%d] = %g want %g\n", m, n, hD[m * 256 + n], ref); ++bad; }
            if (err > maxerr) maxerr = err;
        }
    printf("umma2 (cta_group::2, M256 N256 K%d): max abs err %.3g, mismatches %d, cycles cta0 %lld cta1 %lld -> %s\n", K, maxerr, bad, hc[0], hc[1],
           bad ? "FAIL" : "OK");
    // TMEM drain rate
    long long* dout; uint32_t* dsink; cudaMalloc(&dout, 8); cudaMalloc(&dsink, 4);
    for (int nw : {4, 8, 16}) {
        tmem_rate_kernel<<<148, 512>>>(nw, 20, dout, dsink);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long c = 0; cudaMemcpy(&c, dout, 8, cudaMemcpyDeviceToHost);
        printf("tmem drain 128 lanes x 512 cols fp32 (256 KB) with %2d warps: %lld cycles (%.1f B/clk/SM)  [%s]\n", nw, c, 262144.0 / (double)c, cudaGetErrorString(e2));
    }
    return bad ? 1 : 0;
}
