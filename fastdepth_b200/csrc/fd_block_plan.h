// Shared-memory / pipeline planning of the fused block kernel (pure host C++, no CUDA) so that it can be
// unit-tested on a machine without a GPU (tests/test_block_plan.py through fd_debug_block_plan).
#pragma once

namespace fd {

constexpr int kPlanKblk = 64;
constexpr int kPlanAStage = 128 * 128;       // one A operand stage: 128 rows x 64 x 2 B
constexpr int kPlanStg = 16384;              // one epilogue staging tile: 128 px x 64 ch x 2 B
constexpr int kPlanMaxIn = 6, kPlanMaxA = 4, kPlanMaxB = 16;
constexpr int kPlanSmemBudget = 224 * 1024;  // of the 227 KB a CTA may use (alignment slack is budgeted separately)

struct BlockPlanIn {
    int ksize, stride, tile;     // tile: 0 = 1 image x 8 x 16, 1 = 2 images x 8 x 8
    int c_in, c_out, n_tiles;    // n_tiles: 128-pixel tiles of the whole problem
    int head;                    // decode_conv6 folded into the epilogue
    int barrier_bytes;           // sizeof(TcBarriers)
};
struct BlockPlanOut {
    int ok;
    int kblocks, cin_pad, splits, n_cta, cpad_all, items, tmem_cols;
    int in_stage_bytes, dwp_bytes, in_stage_stride;
    int s_in, s_a, s_b, bn, nb, b_resident, b_stage_bytes;
    int epi_groups, n_stg;       // n_stg = staging tiles in total (epi_groups x 1 or 2)
    int smem_bytes;
};

inline BlockPlanOut plan_block(const BlockPlanIn& q) {
    BlockPlanOut p{};
    const int NI = q.tile ? 2 : 1, TH = 8, TW = q.tile ? 8 : 16;
    const int IH = (TH - 1) * q.stride + q.ksize, IW = (TW - 1) * q.stride + q.ksize;
    p.kblocks = (q.c_in + kPlanKblk - 1) / kPlanKblk;
    p.cin_pad = p.kblocks * kPlanKblk;
    // Split the output channels into items (each item recomputes the depthwise half for its 128 pixels, so splitting is
    // not free).  Candidates: n_cta <= 256 (two TMEM accumulators), multiples of 64 when there is more than one split
    // (the epilogue moves whole [128 px][64 ch] tiles and must not touch a neighbouring split's columns).  Pick the
    // candidate with the smallest modelled kernel time: rounds over the 148 SMs x (K-blocks x max(depthwise, MMA) cycles)
    // + the last item's exposed epilogue.  The per-K-block cycle counts are the measured ones (profiles/r01_trace_*).
    const int cout_pad = (q.c_out + 15) / 16 * 16;
    int splits = 1;
    {
        const long dw_c = q.ksize == 5 ? 2100 : (q.stride == 2 ? 1200 : 1000);
        long best_t = -1;
        for (int n_cta = 256; n_cta >= 64; n_cta -= 64) {
            int sp = (cout_pad + n_cta - 1) / n_cta, nc = n_cta;
            if (sp == 1) nc = cout_pad;                              // a single split needs no 64-alignment
            if (sp == 1 && cout_pad > 256) continue;
            if (q.head && sp > 1) continue;
            const long items = (long)q.n_tiles * sp;
            const long rounds = (items + 147) / 148;
            const long mma_c = 2L * nc;                              // 128 x nc x 64 MACs at 4096 MAC/clk
            const long kb_c = (dw_c > mma_c ? dw_c : mma_c) + 100;
            const long t = rounds * p.kblocks * kb_c + 45L * nc;
            if (best_t < 0 || t < best_t) { best_t = t; splits = sp; p.n_cta = nc; }
        }
        if (cout_pad <= 64 || q.head) { splits = 1; p.n_cta = cout_pad; }
    }
    p.splits = splits;
    p.items = q.n_tiles * splits;
    p.cpad_all = p.n_cta * splits;
    p.tmem_cols = 32;
    while (p.tmem_cols < 2 * p.n_cta) p.tmem_cols *= 2;
    p.in_stage_bytes = NI * IH * IW * 128;
    p.dwp_bytes = q.ksize * q.ksize * 128 + 512;
    p.in_stage_stride = (p.in_stage_bytes + p.dwp_bytes + 127) / 128 * 128;

    // small search over ring depths / weight sub-block width / epilogue organisation, scored by what matters for
    // the block at hand (stride-2 blocks stage 72 KB of input per K-block and leave little room; single-K-block
    // blocks want a deep A ring to hide the serial latency of the MMA issue thread)
    const int fixed = q.barrier_bytes + 2048 + (q.head ? 3 : 2) * p.cpad_all * 4;      // 2048: two 1 KB alignment slacks
    const int total = kPlanSmemBudget - fixed;
    long best = -(1L << 60);
    bool found = false;
    const int bn_top = p.n_cta < 256 ? p.n_cta : 256;
    for (int s_a = kPlanMaxA; s_a >= 2; --s_a)
        for (int eg = q.head ? 1 : 3; eg >= 0; --eg) {
            // epilogue organisation: 3 = two groups x two tiles, 2 = two groups x one tile, 1 = one group x two, 0 = one x one
            const int groups = q.head ? 2 : (eg >= 2 ? 2 : 1);
            const int n_stg = q.head ? 0 : groups * ((eg & 1) ? 2 : 1);
            if (q.head && eg != 1) continue;
            const int avail = total - s_a * kPlanAStage - n_stg * kPlanStg;
            for (int bn = bn_top;; bn = (bn / 2 + 15) / 16 * 16) {
                const int nb = (p.n_cta + bn - 1) / bn;
                const int w_all = p.kblocks * nb * bn * 128;
                const bool can_res = splits == 1 && bn == bn_top && p.kblocks * nb <= kPlanMaxB && w_all <= 64 * 1024;
                for (int res = can_res ? 1 : 0; res >= 0; --res)
                    for (int s_b = res ? p.kblocks * nb : 6; s_b >= (res ? p.kblocks * nb : 2); --s_b) {
                        if (!res && s_b > p.kblocks * nb && s_b > 2) continue;
                        const int left = avail - s_b * bn * 128;
                        if (left < 0) continue;
                        int s_in = left / p.in_stage_stride;
                        if (s_in > kPlanMaxIn) s_in = kPlanMaxIn;
                        if (s_in < 2 && !(s_in == 1 && p.kblocks == 1 && p.items <= 148)) continue;
                        const int bn_eff = bn < 128 ? bn : 128;
                        // weight ring depth in K-blocks: below 2 the MMA of K-block k+1 waits for a weight load that could
                        // only start when the MMA of K-block k had finished (measured: conv7 lost a third of its time there)
                        const int b_ahead2 = res ? 4 : (2 * s_b / nb > 4 ? 4 : 2 * s_b / nb);      // in half K-blocks, capped at 2 K-blocks
                        long score = (long)bn_eff * 100 + (bn >= 256 ? 500 : 0) + (s_in > 4 ? 4 : s_in) * 2500 +
                                     s_a * (p.kblocks <= 2 ? 1500 : 400) + groups * 2500 + n_stg * 300 + (res ? 1000 : 0) +
                                     b_ahead2 * 1800;
                        if (bn < 64 && bn < bn_top) score -= 20000;          // narrow MMAs are a last resort
                        if (score > best) {
                            found = true; best = score;
                            p.s_a = s_a; p.n_stg = n_stg; p.epi_groups = groups; p.s_in = s_in; p.s_b = s_b; p.bn = bn; p.b_resident = res;
                        }
                    }
                if (bn <= 16) break;
            }
        }
    p.ok = found ? 1 : 0;
    if (!found) return p;
    p.nb = (p.n_cta + p.bn - 1) / p.bn;
    p.b_stage_bytes = p.bn * 128;
    p.smem_bytes = p.s_a * kPlanAStage + p.s_b * p.b_stage_bytes + p.s_in * p.in_stage_stride + p.n_stg * kPlanStg + fixed;
    return p;
}

}  // namespace fd
