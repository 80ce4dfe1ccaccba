// Fused depthwise -> pointwise block kernel for sm_100a (tcgen05 + TMEM + TMA), persistent and
// warp-specialised.
//
// One kernel per conv_dw / decode_conv block (reference imagenet/mobilenet.py:29-38, models.py:61-75,
// 683-697) including, for decoder blocks, the nearest-x2 upsample and the skip add of models.py:723-729
// in the epilogue, and optionally decode_conv6 (models.py:698,731) folded below the last upsample.
//
// Work item = one tile of 128 output pixels (NI images x TH x TW) times n_cta output channels.  The grid is
// one CTA per SM; every CTA walks items blockIdx.x, +gridDim.x, ... and all four roles run concurrently on
// different items / K-blocks:
//   warp 16     TMA producer : per 64-channel K-block, one 4-D box load of the input halo tile
//                              [NI][IH][IW][64ch] (OOB zero fill == conv zero padding) and the [BN x 64]
//                              slices of the pointwise weights (128B-swizzled, K-major); when the whole weight
//                              matrix fits it is loaded once and stays resident
//   warps 0-7   depthwise    : lane = channel pair, 4x4 output pixels per warp, FHFMA (16-bit x 16-bit + fp32,
//                              exact products) -> BN affine -> act -> 16-bit, written straight into the
//                              128B-swizzled K-major A operand tile in shared memory (never touches HBM)
//   warp 17     MMA issuer   : tcgen05.mma.cta_group::1.kind::f16, M=128, N<=256, K=16, fp32 accumulators in
//                              TMEM (double-buffered so the next item's MMAs overlap this item's epilogue)
//   warps 8-15  epilogue     : tcgen05.ld 32x32b -> BN affine + act -> 16-bit -> global (x4 replicated + skip
//                              for decoder blocks; or the folded 1-channel head)
// mbarrier rings: input stages (TMA -> dw), A stages (dw -> MMA), B stages (TMA -> MMA), accumulators
// (MMA -> epilogue).
#include <cuda.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "fd_common.cuh"

namespace fd {

constexpr int TC_DW_WARPS = 8;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_WARP_EPI0 = TC_DW_WARPS;                    // 8..15 (warp % 4 == TMEM lane quarter)
constexpr int TC_WARP_TMA = TC_DW_WARPS + TC_EPI_WARPS;      // 16
constexpr int TC_WARP_MMA = TC_WARP_TMA + 1;                 // 17
constexpr int TC_THREADS = (TC_WARP_MMA + 1) * 32;           // 576
constexpr int TC_KBLK = 64;                     // channels per K-block (one 128-byte swizzle row)
constexpr int TC_A_STAGE_BYTES = 128 * 128;     // 128 rows x 64 x 2 B
constexpr int TC_MAX_IN = 6, TC_MAX_A = 4, TC_MAX_B = 16;

struct TcParams {
    int n, h_in, w_in, h_out, w_out, c_in, c_out;
    int tiles_x, tiles_y;
    int items, splits;    // work items = spatial tiles x output-channel splits
    int kblocks;          // ceil(c_in / 64)
    int cin_pad;          // kblocks * 64
    int n_cta;            // output channels per item (multiple of 16)
    int bn;               // B sub-block width (columns per tcgen05.mma), multiple of 16, <= 256
    int nb;               // sub-blocks per K-block = ceil(n_cta / bn)
    int s_in, s_a, s_b;   // pipeline depths
    int b_resident;       // 1: all kblocks*nb weight blocks are loaded once and kept (s_b == kblocks*nb)
    int nacc;             // TMEM accumulator buffers (2 when 2*n_cta <= 512)
    int in_stage_bytes, b_stage_bytes;
    int tmem_cols;        // power of two >= 32, >= nacc * n_cta
    int act, upsample;
    int head;             // 1: fold the C->1 head (writes head_out instead of out)
    int head_act;
    float head_scale, head_bias;
    const void* skip;
    void* out;
    void* head_out;
    int in_stage_stride;  // in_stage_bytes + dw parameter block, rounded to 128
    int dwp_bytes;        // bytes of one K-block's depthwise parameter block
    int cpad_all;         // n_cta * splits: padded length of the pointwise BN vectors
    const void* dwp;      // [kblocks] x { [k*k][64] 16-bit taps, [64] fp32 scale, [64] fp32 bias }
    const float* pw_scale;
    const float* pw_bias; // [cpad_all]
    const float* head_w;  // [cpad_all]
};

// ----------------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// plain (non-tensor) bulk copy global -> shared, completion on an mbarrier (SASS UBLKCP)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both operands K-major, 16-bit inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// 32 consecutive columns of this thread's TMEM lane, load + wait in ONE asm statement so that no use of the
// destination registers can be scheduled before tcgen05.wait::ld
__device__ __forceinline__ void tmem_ld32_sync(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_sync(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 UMMA): start address >> 4, LBO (unused for
// swizzled K-major) = 1, SBO = 1024 B between 8-row groups, descriptor version 1, layout type 2 (128B swizzle).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// mixed-precision FMA: exact 16-bit x 16-bit product added into fp32 (SASS FHFMA / FHFMA.BF16)
template <typename T> struct MixFma;
template <> struct MixFma<__half> {
    __device__ __forceinline__ static void fma2(float& lo, float& hi, uint32_t a, uint32_t b) {
        asm("{\n\t.reg .f16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %2;\n\tmov.b32 {bl, bh}, %3;\n\t"
            "fma.rn.f32.f16 %0, al, bl, %0;\n\tfma.rn.f32.f16 %1, ah, bh, %1;\n\t}" : "+f"(lo), "+f"(hi) : "r"(a), "r"(b));
    }
    __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static float2 unpack(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
    static constexpr uint32_t kUmmaFormat = 0;   // F16
};
template <> struct MixFma<__nv_bfloat16> {
    __device__ __forceinline__ static void fma2(float& lo, float& hi, uint32_t a, uint32_t b) {
        asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %2;\n\tmov.b32 {bl, bh}, %3;\n\t"
            "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %1, ah, bh, %1;\n\t}" : "+f"(lo), "+f"(hi) : "r"(a), "r"(b));
    }
    __device__ __forceinline__ static uint32_t pack(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static float2 unpack(uint32_t v) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v)); }
    static constexpr uint32_t kUmmaFormat = 1;   // BF16
};

// shared-memory bookkeeping block (after the operand stages)
struct TcBarriers {
    uint64_t in_full[TC_MAX_IN], in_empty[TC_MAX_IN];
    uint64_t a_full[TC_MAX_A], a_empty[TC_MAX_A];
    uint64_t b_full[TC_MAX_B], b_empty[TC_MAX_B];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

struct ItemCoord { int img0, oy0, ox0, n0; };
__device__ __forceinline__ ItemCoord decode_item(const TcParams& p, int w, int NI, int TH, int TW) {
    ItemCoord c;
    const int split = w % p.splits;
    int t = w / p.splits;
    const int tile_x = t % p.tiles_x; t /= p.tiles_x;
    const int tile_y = t % p.tiles_y; t /= p.tiles_y;
    c.img0 = t * NI; c.oy0 = tile_y * TH; c.ox0 = tile_x * TW; c.n0 = split * p.n_cta;
    return c;
}

// ----------------------------------------------------------------------------------------------
// the kernel
// ----------------------------------------------------------------------------------------------
template <typename T, int KS, int STRIDE, int NI, int TH, int TW>
__global__ void __launch_bounds__(TC_THREADS, 1)
block_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w, const TcParams p) {
    static_assert(NI * TH * TW == 128, "tile must hold 128 pixels");
    static_assert(NI * (TH / 4) * (TW / 4) == TC_DW_WARPS, "one 4x4 pixel block per depthwise warp");
    constexpr int PAD = (KS - 1) / 2;
    constexpr int IH = (TH - 1) * STRIDE + KS, IW = (TW - 1) * STRIDE + KS;     // input box
    constexpr int IBH = 3 * STRIDE + KS, IBW = 3 * STRIDE + KS;                   // per-warp input block
    using MF = MixFma<T>;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    // carve-up: [A stages][B stages][input stages (+ dw parameter block each)][pw BN vectors][barriers]
    const uint32_t a_off = 0;
    const uint32_t b_off = a_off + p.s_a * TC_A_STAGE_BYTES;
    const uint32_t in_off = b_off + p.s_b * p.b_stage_bytes;
    const uint32_t pw_off = in_off + p.s_in * p.in_stage_stride;
    const uint32_t bar_off = pw_off + 3u * (uint32_t)p.cpad_all * 4u;
    TcBarriers* bars = reinterpret_cast<TcBarriers*>(smem + bar_off);
    float* s_pw_scale = reinterpret_cast<float*>(smem + pw_off);
    float* s_pw_bias = s_pw_scale + p.cpad_all;
    float* s_head_w = s_pw_bias + p.cpad_all;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_MAX_IN; ++i) { mbar_init(smem_u32(&bars->in_full[i]), 1); mbar_init(smem_u32(&bars->in_empty[i]), TC_DW_WARPS); }
        for (int i = 0; i < TC_MAX_A; ++i) { mbar_init(smem_u32(&bars->a_full[i]), TC_DW_WARPS); mbar_init(smem_u32(&bars->a_empty[i]), 1); }
        for (int i = 0; i < TC_MAX_B; ++i) { mbar_init(smem_u32(&bars->b_full[i]), 1); mbar_init(smem_u32(&bars->b_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->acc_full[i]), 1); mbar_init(smem_u32(&bars->acc_empty[i]), TC_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == TC_WARP_MMA) tmem_alloc(smem_u32(&bars->tmem_base), (uint32_t)p.tmem_cols);     // MMA warp owns TMEM
    if (warp == TC_WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tm_in);
        tma_prefetch_desc(&tm_w);
    }
    for (int i = threadIdx.x; i < p.cpad_all; i += TC_THREADS) {          // pointwise BN affine (+ head weights) -> smem
        s_pw_scale[i] = p.pw_scale[i];
        s_pw_bias[i] = p.pw_bias[i];
        s_head_w[i] = p.head ? p.head_w[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == TC_WARP_TMA) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            uint32_t it = 0, jb = 0;
            bool first = true;
            for (int w = blockIdx.x; w < p.items; w += gridDim.x, first = false) {
                const ItemCoord c = decode_item(p, w, NI, TH, TW);
                for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                    const uint32_t s = it % (uint32_t)p.s_in, ph = (it / (uint32_t)p.s_in) & 1u;
                    mbar_wait(smem_u32(&bars->in_empty[s]), ph ^ 1u);
                    mbar_expect_tx(smem_u32(&bars->in_full[s]), (uint32_t)(p.in_stage_bytes + p.dwp_bytes));
                    tma_load_4d(smem_base + in_off + s * p.in_stage_stride, &tm_in, smem_u32(&bars->in_full[s]), kb * TC_KBLK,
                                c.ox0 * STRIDE - PAD, c.oy0 * STRIDE - PAD, c.img0);
                    bulk_load(smem_base + in_off + s * p.in_stage_stride + p.in_stage_bytes,
                              reinterpret_cast<const uint8_t*>(p.dwp) + (size_t)kb * p.dwp_bytes, (uint32_t)p.dwp_bytes,
                              smem_u32(&bars->in_full[s]));
                    if (p.b_resident && !first) continue;
                    for (int nbi = 0; nbi < p.nb; ++nbi, ++jb) {
                        uint32_t sb;
                        if (p.b_resident) {
                            sb = (uint32_t)(kb * p.nb + nbi);
                        } else {
                            sb = jb % (uint32_t)p.s_b;
                            mbar_wait(smem_u32(&bars->b_empty[sb]), ((jb / (uint32_t)p.s_b) & 1u) ^ 1u);
                        }
                        mbar_expect_tx(smem_u32(&bars->b_full[sb]), (uint32_t)p.b_stage_bytes);
                        tma_load_2d(smem_base + b_off + sb * p.b_stage_bytes, &tm_w, smem_u32(&bars->b_full[sb]), kb * TC_KBLK,
                                    c.n0 + nbi * p.bn);
                    }
                }
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B 16-bit K-major, M = 128, N filled per sub-block
            const uint32_t idesc_base = (1u << 4) | (MF::kUmmaFormat << 7) | (MF::kUmmaFormat << 10) | ((128u >> 4) << 24);
            uint32_t it = 0, jb = 0, i = 0;
            for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++i) {
                const uint32_t ab = i % (uint32_t)p.nacc, pa = (i / (uint32_t)p.nacc) & 1u;
                mbar_wait(smem_u32(&bars->acc_empty[ab]), pa ^ 1u);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + ab * (uint32_t)p.n_cta;
                for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                    const uint32_t sa = it % (uint32_t)p.s_a, pha = (it / (uint32_t)p.s_a) & 1u;
                    mbar_wait(smem_u32(&bars->a_full[sa]), pha);
                    tc_fence_after();
                    const uint64_t a_desc = make_kmajor_sw128_desc(smem_base + a_off + sa * TC_A_STAGE_BYTES);
                    for (int nbi = 0; nbi < p.nb; ++nbi, ++jb) {
                        uint32_t sb;
                        if (p.b_resident) {
                            sb = (uint32_t)(kb * p.nb + nbi);
                            if (i == 0) { mbar_wait(smem_u32(&bars->b_full[sb]), 0); tc_fence_after(); }
                        } else {
                            sb = jb % (uint32_t)p.s_b;
                            mbar_wait(smem_u32(&bars->b_full[sb]), (jb / (uint32_t)p.s_b) & 1u);
                            tc_fence_after();
                        }
                        const uint64_t b_desc = make_kmajor_sw128_desc(smem_base + b_off + sb * p.b_stage_bytes);
                        const int n_cur = min(p.bn, p.n_cta - nbi * p.bn);
                        const uint32_t idesc = idesc_base | ((uint32_t)(n_cur >> 3) << 17);
#pragma unroll
                        for (int k = 0; k < TC_KBLK / 16; ++k)     // advance 32 B (16 elements) inside the swizzle row
                            umma_f16(d_tmem + nbi * p.bn, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        if (!p.b_resident) umma_commit(smem_u32(&bars->b_empty[sb]));
                    }
                    umma_commit(smem_u32(&bars->a_empty[sa]));
                }
                umma_commit(smem_u32(&bars->acc_full[ab]));
            }
        }
    } else if (warp < TC_DW_WARPS) {
        // =========================== depthwise workers ===========================
        constexpr int BPR = TW / 4, BPI = (TH / 4) * BPR;
        const int ni = warp / BPI, rem = warp % BPI;
        const int br = rem / BPR, bc = rem % BPR;
        const uint32_t in_warp_off = (uint32_t)((ni * IH + br * 4 * STRIDE) * IW + bc * 4 * STRIDE) * 128u + lane * 4u;
        uint32_t it = 0;
        for (int w = blockIdx.x; w < p.items; w += gridDim.x) {
            for (int kb = 0; kb < p.kblocks; ++kb, ++it) {
                const uint32_t s = it % (uint32_t)p.s_in, ph = (it / (uint32_t)p.s_in) & 1u;
                const uint32_t sa = it % (uint32_t)p.s_a, pha = (it / (uint32_t)p.s_a) & 1u;
                mbar_wait(smem_u32(&bars->in_full[s]), ph);
                const uint8_t* stage = smem + in_off + s * p.in_stage_stride;
                const uint8_t* in_s = stage + in_warp_off;
                // this K-block's depthwise taps + folded BN for the lane's channel pair (landed with the tile)
                const uint8_t* prm = stage + p.in_stage_bytes;
                uint32_t wv[KS * KS];
#pragma unroll
                for (int i = 0; i < KS * KS; ++i) wv[i] = *reinterpret_cast<const uint32_t*>(prm + i * 128 + lane * 4);
                const float2 sc = *reinterpret_cast<const float2*>(prm + KS * KS * 128 + lane * 8);
                const float2 bi = *reinterpret_cast<const float2*>(prm + KS * KS * 128 + 256 + lane * 8);
                float acc[4][4][2];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.f;
#pragma unroll
                for (int iy = 0; iy < IBH; ++iy) {
                    uint32_t row[IBW];
#pragma unroll
                    for (int ix = 0; ix < IBW; ++ix) row[ix] = *reinterpret_cast<const uint32_t*>(in_s + (iy * IW + ix) * 128);
#pragma unroll
                    for (int oy = 0; oy < 4; ++oy) {
                        const int ky = iy - oy * STRIDE;
                        if (ky < 0 || ky >= KS) continue;
#pragma unroll
                        for (int ox = 0; ox < 4; ++ox)
#pragma unroll
                            for (int kx = 0; kx < KS; ++kx)
                                MF::fma2(acc[oy][ox][0], acc[oy][ox][1], row[ox * STRIDE + kx], wv[ky * KS + kx]);
                    }
                }
                // the input stage can be refilled as soon as every warp has read it
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->in_empty[s]));

                mbar_wait(smem_u32(&bars->a_empty[sa]), pha ^ 1u);
                uint8_t* a_s = smem + a_off + sa * TC_A_STAGE_BYTES;
#pragma unroll
                for (int oy = 0; oy < 4; ++oy)
#pragma unroll
                    for (int ox = 0; ox < 4; ++ox) {
                        const int m = (ni * TH + br * 4 + oy) * TW + bc * 4 + ox;
                        const float lo = apply_act(fmaf(acc[oy][ox][0], sc.x, bi.x), p.act);
                        const float hi = apply_act(fmaf(acc[oy][ox][1], sc.y, bi.y), p.act);
                        *reinterpret_cast<uint32_t*>(a_s + m * 128 + (((lane >> 2) ^ (m & 7)) << 4) + ((lane & 3) << 2)) = MF::pack(lo, hi);
                    }
                fence_proxy_async();             // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[sa]));
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int ew = warp - TC_WARP_EPI0;
        const int q = ew & 3, hsel = ew >> 2;              // TMEM lane quarter (== warp % 4), column-batch parity
        const int m = q * 32 + lane;
        const int e_ni = m / (TH * TW), e_ty = (m / TW) % TH, e_tx = m % TW;
        const int batches = (p.n_cta + 31) >> 5;           // 32 accumulator columns per batch
        T* __restrict__ outp = reinterpret_cast<T*>(p.out);
        const T* __restrict__ skipp = reinterpret_cast<const T*>(p.skip);
        uint32_t i = 0;
        for (int w = blockIdx.x; w < p.items; w += gridDim.x, ++i) {
            const ItemCoord c = decode_item(p, w, NI, TH, TW);
            const uint32_t ab = i % (uint32_t)p.nacc, pa = (i / (uint32_t)p.nacc) & 1u;
            const int img = c.img0 + e_ni, oy = c.oy0 + e_ty, ox = c.ox0 + e_tx;
            const bool valid = img < p.n && oy < p.h_out && ox < p.w_out;
            mbar_wait(smem_u32(&bars->acc_full[ab]), pa);
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + ab * (uint32_t)p.n_cta;

            if (!p.head) {
                for (int b = hsel; b < batches; b += 2) {
                    uint32_t r[32];
                    const bool full = b * 32 + 32 <= p.n_cta;          // n_cta is a multiple of 16
                    if (full) tmem_ld32_sync(t_lane + b * 32, r);
                    else tmem_ld16_sync(t_lane + b * 32, r);
#pragma unroll
                    for (int hg = 0; hg < 2; ++hg) {                   // two 16-column groups
                        if (hg == 1 && !full) break;
                        const int c0 = c.n0 + b * 32 + hg * 16;
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 s2 = *reinterpret_cast<const float2*>(s_pw_scale + c0 + 2 * j);
                            const float2 b2 = *reinterpret_cast<const float2*>(s_pw_bias + c0 + 2 * j);
                            const float lo = apply_act(fmaf(__uint_as_float(r[hg * 16 + 2 * j]), s2.x, b2.x), p.act);
                            const float hi = apply_act(fmaf(__uint_as_float(r[hg * 16 + 2 * j + 1]), s2.y, b2.y), p.act);
                            pk[j] = MF::pack(lo, hi);
                        }
                        // stores are predicated (no divergent early-out: the next tcgen05.ld is warp-collective)
                        const bool v0 = valid && c0 + 8 <= p.c_out, v1 = valid && c0 + 16 <= p.c_out;
                        if (!p.upsample) {
                            T* o = outp + (((size_t)img * p.h_out + oy) * p.w_out + ox) * p.c_out + c0;
                            if (v0) *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            if (v1) *reinterpret_cast<uint4*>(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        } else {
                            const int w2 = 2 * p.w_out;
                            const size_t off00 = (((size_t)img * 2 * p.h_out + 2 * oy) * w2 + 2 * ox) * p.c_out + c0;
                            size_t off[4];
#pragma unroll
                            for (int d = 0; d < 4; ++d) off[d] = off00 + ((size_t)(d >> 1) * w2 + (d & 1)) * p.c_out;
                            if (skipp != nullptr) {
                                // x = interpolate(x) (already rounded to the storage dtype); x = x + skip (models.py:723-729).
                                // all eight skip vectors are requested before the first is consumed
                                uint4 sv[4][2];
#pragma unroll
                                for (int d = 0; d < 4; ++d) {
                                    sv[d][0] = v0 ? __ldg(reinterpret_cast<const uint4*>(skipp + off[d])) : make_uint4(0, 0, 0, 0);
                                    sv[d][1] = v1 ? __ldg(reinterpret_cast<const uint4*>(skipp + off[d] + 8)) : make_uint4(0, 0, 0, 0);
                                }
#pragma unroll
                                for (int d = 0; d < 4; ++d) {
                                    const uint32_t sk[8] = {sv[d][0].x, sv[d][0].y, sv[d][0].z, sv[d][0].w,
                                                            sv[d][1].x, sv[d][1].y, sv[d][1].z, sv[d][1].w};
                                    uint32_t z[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float2 a = MF::unpack(pk[j]), bq = MF::unpack(sk[j]);
                                        z[j] = MF::pack(a.x + bq.x, a.y + bq.y);
                                    }
                                    if (v0) *reinterpret_cast<uint4*>(outp + off[d]) = make_uint4(z[0], z[1], z[2], z[3]);
                                    if (v1) *reinterpret_cast<uint4*>(outp + off[d] + 8) = make_uint4(z[4], z[5], z[6], z[7]);
                                }
                            } else {
#pragma unroll
                                for (int d = 0; d < 4; ++d) {
                                    if (v0) *reinterpret_cast<uint4*>(outp + off[d]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                    if (v1) *reinterpret_cast<uint4*>(outp + off[d] + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                                }
                            }
                        }
                    }
                }
            } else if (hsel == 0) {
                // decode_conv6 folded below the last upsample: dot over this block's (<= 64) output channels
                float dot = 0.f;
                for (int b = 0; b < batches; ++b) {
                    uint32_t r[32];
                    const bool full = b * 32 + 32 <= p.n_cta;
                    if (full) tmem_ld32_sync(t_lane + b * 32, r);
                    else tmem_ld16_sync(t_lane + b * 32, r);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j >= 8 && !full) break;
                        const int c0 = b * 32 + 2 * j;
                        const float2 s2 = *reinterpret_cast<const float2*>(s_pw_scale + c0);
                        const float2 b2 = *reinterpret_cast<const float2*>(s_pw_bias + c0);
                        const float2 hw = *reinterpret_cast<const float2*>(s_head_w + c0);
                        const float lo = apply_act(fmaf(__uint_as_float(r[2 * j]), s2.x, b2.x), p.act);
                        const float hi = apply_act(fmaf(__uint_as_float(r[2 * j + 1]), s2.y, b2.y), p.act);
                        const float2 rq = MF::unpack(MF::pack(lo, hi));      // decode_conv5's output is stored in 16 bits
                        dot = fmaf(rq.x, hw.x, dot);
                        dot = fmaf(rq.y, hw.y, dot);
                    }
                }
                if (valid) {
                    const float y = apply_act(fmaf(dot, p.head_scale, p.head_bias), p.head_act);
                    const uint32_t yy = MF::pack(y, y);
                    T* ho = reinterpret_cast<T*>(p.head_out) + ((size_t)img * 2 * p.h_out + 2 * oy) * (2 * p.w_out) + 2 * ox;
                    *reinterpret_cast<uint32_t*>(ho) = yy;
                    *reinterpret_cast<uint32_t*>(ho + 2 * p.w_out) = yy;
                }
            }
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    return fn;
}

struct BlockTcPlan {
    CUtensorMap tm_in, tm_w;
    TcParams p;
    dim3 grid;
    size_t smem_bytes;
    int dtype, ks, stride, tile;           // tile: 0 = (1,8,16), 1 = (2,8,8)
    void* dwp = nullptr;                   // owned device copies (packed / padded)
    float* pw_scale = nullptr;
    float* pw_bias = nullptr;
    float* head_w = nullptr;
    std::string name;
};

static int pick_tile(const StageGeom& g) {
    // 2 images x 8x8 when a whole image fits an 8x8 box (7x7 stages), else 1 image x 8 rows x 16 cols
    return (g.h_out <= 8 && g.w_out <= 8) ? 1 : 0;
}

bool block_tc_supported(int dtype, const StageGeom& g, bool head_fused) {
    if (dtype != FD_F16 && dtype != FD_BF16) return false;
    if (!((g.ksize == 3 && (g.stride == 1 || g.stride == 2)) || (g.ksize == 5 && g.stride == 1))) return false;
    if (g.c_in % 8 || g.c_out % 8) return false;
    if (head_fused && g.c_out > 64) return false;
    return get_encode() != nullptr;
}

template <typename T, int KS, int STRIDE, int NI, int TH, int TW>
static int launch_inst(BlockTcPlan* bp, cudaStream_t st) {
    auto kern = block_tc_kernel<T, KS, STRIDE, NI, TH, TW>;
    static bool attr_set = false;
    if (!attr_set) {
        FD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    kern<<<bp->grid, TC_THREADS, bp->smem_bytes, st>>>(bp->tm_in, bp->tm_w, bp->p);
    FD_CUDA_OK(cudaGetLastError());
    return FD_OK;
}

template <typename T>
static int launch_t(BlockTcPlan* bp, cudaStream_t st) {
    const int key = bp->ks * 100 + bp->stride * 10 + bp->tile;
    switch (key) {
        case 310: return launch_inst<T, 3, 1, 1, 8, 16>(bp, st);
        case 311: return launch_inst<T, 3, 1, 2, 8, 8>(bp, st);
        case 320: return launch_inst<T, 3, 2, 1, 8, 16>(bp, st);
        case 321: return launch_inst<T, 3, 2, 2, 8, 8>(bp, st);
        case 510: return launch_inst<T, 5, 1, 1, 8, 16>(bp, st);
        case 511: return launch_inst<T, 5, 1, 2, 8, 8>(bp, st);
        default: return fail(FD_ERR_UNSUPPORTED, "no fused kernel instance for this block");
    }
}

int block_tc_launch(BlockTcPlan* bp, cudaStream_t st, void* head_out) {
    if (bp->p.head) {
        if (!head_out) return fail(FD_ERR_INVALID, "head-fused block needs the output pointer");
        bp->p.head_out = head_out;
    }
    return bp->dtype == FD_F16 ? launch_t<__half>(bp, st) : launch_t<__nv_bfloat16>(bp, st);
}

const char* block_tc_name(BlockTcPlan* bp) { return bp->name.c_str(); }

void block_tc_destroy(BlockTcPlan* bp) {
    if (!bp) return;
    cudaFree(bp->dwp); cudaFree(bp->pw_scale); cudaFree(bp->pw_bias);
    cudaFree(bp->head_w);
    delete bp;
}

// per-K-block depthwise parameter block: [taps][64] 16-bit taps | [64] fp32 scale | [64] fp32 bias (zero padded)
template <typename T>
__global__ void pack_dwp_kernel(const float* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ bias,
                                uint8_t* __restrict__ dst, int taps, int c_in, int kblocks, int block_bytes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kblocks * 64) return;
    const int kb = i / 64, cl = i % 64, c = kb * 64 + cl;
    uint8_t* blk = dst + (size_t)kb * block_bytes;
    T* wt = reinterpret_cast<T*>(blk);
    for (int t = 0; t < taps; ++t) wt[t * 64 + cl] = Traits<T>::from_f(c < c_in ? w[t * c_in + c] : 0.f);
    reinterpret_cast<float*>(blk + taps * 128)[cl] = c < c_in ? scale[c] : 0.f;
    reinterpret_cast<float*>(blk + taps * 128 + 256)[cl] = c < c_in ? bias[c] : 0.f;
}
__global__ void pad_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_src, int n_dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_dst) dst[i] = i < n_src ? src[i] : 0.f;
}

static int padded_copy(const float* src, int n_src, int n_dst, float** out) {
    FD_CUDA_OK(cudaMalloc(out, (size_t)n_dst * 4));
    pad_copy_kernel<<<(n_dst + 255) / 256, 256>>>(src, *out, n_src, n_dst);
    FD_CUDA_OK(cudaGetLastError());
    return FD_OK;
}

int block_tc_prepare(int dtype, const BlockArgs& a, const float* head_w, float head_scale, float head_bias, int head_act,
                     void* head_out, BlockTcPlan** out) {
    PFN_encodeTiled encode = get_encode();
    if (!encode) return fail(FD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const StageGeom& g = a.g;
    BlockTcPlan* bp = new (std::nothrow) BlockTcPlan();
    if (!bp) return fail(FD_ERR_CUDA, "out of host memory");
    bp->dtype = dtype; bp->ks = g.ksize; bp->stride = g.stride; bp->tile = pick_tile(g);
    const int NI = bp->tile ? 2 : 1, TH = 8, TW = bp->tile ? 8 : 16;
    const int IH = (TH - 1) * g.stride + g.ksize, IW = (TW - 1) * g.stride + g.ksize;
    TcParams& p = bp->p;
    memset(&p, 0, sizeof(p));
    p.n = g.n; p.h_in = g.h_in; p.w_in = g.w_in; p.h_out = g.h_out; p.w_out = g.w_out; p.c_in = g.c_in; p.c_out = g.c_out;
    p.tiles_x = (g.w_out + TW - 1) / TW; p.tiles_y = (g.h_out + TH - 1) / TH;
    const int tiles_img = (g.n + NI - 1) / NI;
    const int n_tiles = p.tiles_x * p.tiles_y * tiles_img;
    p.kblocks = (g.c_in + TC_KBLK - 1) / TC_KBLK; p.cin_pad = p.kblocks * TC_KBLK;
    p.act = g.act; p.upsample = g.upsample;
    p.head = head_w != nullptr; p.head_act = head_act; p.head_scale = head_scale; p.head_bias = head_bias;
    p.skip = a.skip; p.out = a.out; p.head_out = head_out;

    // split the output channels into items until there are enough items for the 148 SMs (each item
    // recomputes the cheap depthwise half) and the per-item accumulator fits TMEM
    const int cout_pad = (g.c_out + 15) / 16 * 16;
    int splits = 1;
    while ((cout_pad + splits - 1) / splits > 256 || (n_tiles * splits < 148 && (cout_pad / (splits * 2)) >= 64 && !p.head)) splits *= 2;
    p.n_cta = ((cout_pad + splits - 1) / splits + 15) / 16 * 16;
    splits = (cout_pad + p.n_cta - 1) / p.n_cta;
    p.splits = splits;
    p.items = n_tiles * splits;
    p.nacc = 2;                                       // n_cta <= 256 -> two accumulators fit 512 columns
    p.tmem_cols = 32;
    while (p.tmem_cols < p.nacc * p.n_cta) p.tmem_cols *= 2;
    p.in_stage_bytes = NI * IH * IW * 128;
    const int taps = g.ksize * g.ksize;
    p.dwp_bytes = taps * 128 + 512;
    p.in_stage_stride = (p.in_stage_bytes + p.dwp_bytes + 127) / 128 * 128;
    p.cpad_all = p.n_cta * splits;
    p.s_a = 2;
    const int avail = 212 * 1024 - (int)sizeof(TcBarriers) - 1024 - p.s_a * TC_A_STAGE_BYTES - 3 * p.cpad_all * 4;
    // weights: resident when the whole [cin_pad x n_cta] matrix is small and every item uses the same one
    p.bn = p.n_cta < 256 ? p.n_cta : 256;
    p.nb = (p.n_cta + p.bn - 1) / p.bn;
    const int w_all = p.kblocks * p.nb * p.bn * 128;
    p.b_resident = (splits == 1 && p.kblocks * p.nb <= TC_MAX_B && w_all <= 64 * 1024 && w_all + 2 * p.in_stage_stride <= avail) ? 1 : 0;
    if (p.b_resident) {
        p.s_b = p.kblocks * p.nb;
    } else {
        p.s_b = 2;
        while (p.bn > 16 && p.s_b * p.bn * 128 + 2 * p.in_stage_stride > avail) p.bn = (p.bn / 2 + 15) / 16 * 16;
        p.nb = (p.n_cta + p.bn - 1) / p.bn;
        while (p.s_b < 4 && (p.s_b + 1) * p.bn * 128 + 2 * p.in_stage_stride <= avail && p.s_b < p.kblocks * p.nb) ++p.s_b;
    }
    p.b_stage_bytes = p.bn * 128;
    p.s_in = (avail - p.s_b * p.b_stage_bytes) / p.in_stage_stride;
    if (p.s_in > TC_MAX_IN) p.s_in = TC_MAX_IN;
    if (p.s_in < 1) { delete bp; return fail(FD_ERR_UNSUPPORTED, "fused block does not fit shared memory"); }
    bp->smem_bytes = (size_t)p.s_a * TC_A_STAGE_BYTES + (size_t)p.s_b * p.b_stage_bytes + (size_t)p.s_in * p.in_stage_stride +
                     3 * (size_t)p.cpad_all * 4 + sizeof(TcBarriers) + 1024;
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    bp->grid = dim3((unsigned)(p.items < sms ? p.items : sms), 1, 1);

    // packed / padded parameter copies (device -> device)
    int rc = FD_OK;
    if (cudaMalloc(&bp->dwp, (size_t)p.kblocks * p.dwp_bytes) != cudaSuccess) rc = fail(FD_ERR_CUDA, "cudaMalloc failed");
    if (rc == FD_OK) {
        const int tot = p.kblocks * 64;
        if (dtype == FD_F16) pack_dwp_kernel<__half><<<(tot + 127) / 128, 128>>>(a.dw_w, a.dw_scale, a.dw_bias, (uint8_t*)bp->dwp, taps, g.c_in, p.kblocks, p.dwp_bytes);
        else pack_dwp_kernel<__nv_bfloat16><<<(tot + 127) / 128, 128>>>(a.dw_w, a.dw_scale, a.dw_bias, (uint8_t*)bp->dwp, taps, g.c_in, p.kblocks, p.dwp_bytes);
        if (cudaGetLastError() != cudaSuccess) rc = fail(FD_ERR_CUDA, "pack_dwp_kernel launch failed");
    }
    if (rc == FD_OK) rc = padded_copy(a.pw_scale, g.c_out, p.cpad_all, &bp->pw_scale);
    if (rc == FD_OK) rc = padded_copy(a.pw_bias, g.c_out, p.cpad_all, &bp->pw_bias);
    if (rc == FD_OK) rc = padded_copy(p.head ? head_w : a.pw_scale, p.head ? g.c_out : 0, p.cpad_all, &bp->head_w);
    if (rc == FD_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(FD_ERR_CUDA, "parameter packing failed");
    if (rc != FD_OK) { block_tc_destroy(bp); return rc; }
    p.dwp = bp->dwp; p.pw_scale = bp->pw_scale; p.pw_bias = bp->pw_bias; p.head_w = bp->head_w;

    // tensor maps
    const size_t es = 2;
    const CUtensorMapDataType dt = dtype == FD_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    {   // input: NHWC viewed as (C, W, H, N); box (64, IW, IH, NI); no swizzle; OOB -> 0 (== zero padding)
        cuuint64_t dims[4] = {(cuuint64_t)g.c_in, (cuuint64_t)g.w_in, (cuuint64_t)g.h_in, (cuuint64_t)g.n};
        cuuint64_t strides[3] = {(cuuint64_t)g.c_in * es, (cuuint64_t)g.w_in * g.c_in * es, (cuuint64_t)g.h_in * g.w_in * g.c_in * es};
        cuuint32_t box[4] = {(cuuint32_t)TC_KBLK, (cuuint32_t)IW, (cuuint32_t)IH, (cuuint32_t)NI};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&bp->tm_in, dt, 4, const_cast<void*>(a.in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { block_tc_destroy(bp); return fail(FD_ERR_CUDA, "cuTensorMapEncodeTiled(input) failed: " + std::to_string((int)r)); }
    }
    {   // pointwise weights [c_out][c_in] viewed as (K = c_in, N = c_out); box (64, bn); 128B swizzle
        cuuint64_t dims[2] = {(cuuint64_t)g.c_in, (cuuint64_t)g.c_out};
        cuuint64_t strides[1] = {(cuuint64_t)g.c_in * es};
        cuuint32_t box[2] = {(cuuint32_t)TC_KBLK, (cuuint32_t)p.bn};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&bp->tm_w, dt, 2, const_cast<void*>(a.pw_w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { block_tc_destroy(bp); return fail(FD_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r)); }
    }
    char buf[160];
    snprintf(buf, sizeof(buf), "block_tc<k%d,s%d,%s>%s%s%s[n%dx%d,bn%d%s,kb%d,in%d]", g.ksize, g.stride, bp->tile ? "2x8x8" : "1x8x16",
             g.upsample ? "+up2x" : "", a.skip ? "+skip" : "", p.head ? "+head" : "", p.n_cta, p.splits, p.bn,
             p.b_resident ? "r" : "", p.kblocks, p.s_in);
    bp->name = buf;
    *out = bp;
    return FD_OK;
}

}  // namespace fd
