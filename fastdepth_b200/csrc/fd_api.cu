// C-ABI of fastdepth_b200 (see include/fastdepth_b200.h): plan construction, weight packing,
// forward dispatch, stage timing.  Host-side C++; kernels live in the sibling .cu files.
#include <algorithm>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "fd_block_plan.h"
#include "fd_common.cuh"

namespace fd {

BlockPlanOut block_tc_debug_plan(int ksize, int stride, int h_out, int w_out, int n, int c_in, int c_out, int head);

// ---- error state -----------------------------------------------------------------------
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

// ---- kernels (fd_kernels_simt.cu, fd_block_tc.cu, fd_metrics.cu) -----------------------------
int launch_stem(int dtype, const void* x, void* out, const float* w, const float* scale, const float* bias,
                const StageGeom& g, cudaStream_t st);
int launch_dw(int dtype, const BlockArgs& a, cudaStream_t st);
int launch_pw(int dtype, const BlockArgs& a, cudaStream_t st);
int launch_head(int dtype, const void* in, void* out, const float* w, float scale, float bias, long long m_total, int c,
                int in_pitch, int h, int wd, int up, int act, cudaStream_t st);
int launch_metrics(int dtype, const void* pred, const float* target, int n, int hw, double* sums, cudaStream_t st);
int launch_nyu_val_gather(int dtype, const uint8_t* rgb, const float* depth, const int* rows, const int* cols, int n, int h_in,
                          int w_in, int oh, int ow, void* x, float* t, cudaStream_t st);
// fused tcgen05 block kernel
struct BlockTcPlan;   // opaque per-stage state (tensor maps, tile config)
bool block_tc_supported(int dtype, const StageGeom& g, bool head_fused);
int block_tc_prepare(int dtype, const BlockArgs& a, const float* head_w, float head_scale, float head_bias, int head_act,
                     void* head_out, bool tma_epilogue, const TcLaunchOpts& opts, BlockTcPlan** out);
int block_tc_launch(BlockTcPlan* p, cudaStream_t st, void* head_out);
void block_tc_destroy(BlockTcPlan* p);
const char* block_tc_name(BlockTcPlan* p);
int block_tc_trace(BlockTcPlan* bp, cudaStream_t st, void* head_out, unsigned long long* out_host, int* rows, int* cols);
// fused multi-layer chain kernel (fd_chain_tc.cu): a run of 3x3 stride-1 blocks on a small map in one 2-CTA-cluster kernel
struct ChainTcPlan;
bool chain_tc_supported(int dtype, const StageGeom* g, int n_layers);
int chain_tc_prepare(int dtype, const BlockArgs* layers, int n_layers, const TcLaunchOpts& opts, ChainTcPlan** out);
int chain_tc_launch(ChainTcPlan* cp, cudaStream_t st);
void chain_tc_destroy(ChainTcPlan* cp);
const char* chain_tc_name(ChainTcPlan* cp);
int chain_tc_trace(ChainTcPlan* cp, cudaStream_t st, unsigned long long* out_host, int* rows, int* cols);
// tensor-core stem (fd_stem_tc.cu)
struct StemTcPlan;
bool stem_tc_supported(int dtype, const StageGeom& g);
int stem_tc_prepare(int dtype, const StageGeom& g, const float* w27_dev, const float* scale_dev, const float* bias_dev, void* out,
                    const TcLaunchOpts& opts, StemTcPlan** res);
int stem_tc_launch(StemTcPlan* sp, const void* x, cudaStream_t st);
void stem_tc_destroy(StemTcPlan* sp);
const char* stem_tc_name(StemTcPlan* sp);

static size_t dtype_size(int dtype) { return dtype == FD_F32 ? 4 : 2; }

struct Stage {
    fd_stage_desc d{};
    StageGeom g{};
    int out_h = 0, out_w = 0;            // spatial size of the stage's output buffer (after upsample)
    void* out = nullptr;                 // NHWC [n,out_h,out_w,c_out]   (STEM/DWPW)
    void* out_eff = nullptr;             // buffer the stage really writes (== a skip source when accumulating in place)
    void* out_alloc = nullptr;           // what this stage cudaMalloc'ed (out may be a channel slice of another stage's buffer)
    int out_c = 0;                       // channels the NEXT stage sees (c_out, or c_out + c_skip after a concat)
    int out_pitch = 0;                   // elements between pixels of `out`
    int concat_src = -1;                 // >= 0: this stage's output is a channel slice of stage concat_src's wide buffer
    void* mid = nullptr;                 // NHWC [n,h_out,w_out,c_in]     (DWPW, path 0)
    float* dw_w = nullptr;               // [k*k][c_in]
    float* dw_scale = nullptr;
    float* dw_bias = nullptr;
    void* pw_w = nullptr;                // [c_out][c_in] plan dtype (DWPW)
    float* pw_w_f32 = nullptr;           // stem: [27][c_out] tap-major ; head: [c_in]
    float* pw_scale = nullptr;
    float* pw_bias = nullptr;
    float head_scale = 0.f, head_bias = 0.f;
    bool have_weights = false;
    BlockTcPlan* tc = nullptr;
    StemTcPlan* stc = nullptr;
    ChainTcPlan* chain = nullptr;        // set on the FIRST stage of a run executed by the chain kernel
    int chained = 0;                     // 1: this stage runs inside a chain kernel (its own output buffer is only written if it is
                                         //    the run's last stage)
};

struct Step {
    int stage;
    std::string name;
    double alg_bytes, macs;
    double dw_macs = 0.0;                // of which depthwise (SIMT FMA pipe); the rest is the dense contraction (tensor pipe)
    std::function<int(cudaStream_t, const void*, void*)> run;
};

}  // namespace fd

using namespace fd;

struct fd_plan {
    int n = 0, h = 0, w = 0, dtype = 0, device = 0, n_sms = 148;
    std::vector<Stage> stages;
    std::vector<Step> steps;
    bool steps_valid = false;
    int opt_path = 1, opt_fold_head = 1, opt_graph = 1, opt_tma_epilogue = 1, opt_inplace_skip = 1, opt_pdl = 0, opt_wait_sleep_ns = 0;
    int opt_chain = 1;
    int opt_cluster = 1;
    size_t workspace_bytes = 0;
    // fd_pipeline_*: host batches flow H2D -> forward -> D2H through kPipeSlots device slots on three streams
    struct PipeSlot { void* x = nullptr; void* y = nullptr; cudaEvent_t up = nullptr, done = nullptr, down = nullptr; bool busy = false; };
    PipeSlot pipe[3];
    cudaStream_t pipe_h2d = nullptr, pipe_run = nullptr, pipe_d2h = nullptr;
    unsigned long long pipe_next = 0;    // next ticket
    void* stage_x = nullptr;             // device staging for fd_forward_host
    void* stage_y = nullptr;
    cudaEvent_t last_done = nullptr;     // recorded after every fd_forward on its stream (cross-stream ordering of one plan's buffers)
    cudaStream_t last_stream = nullptr;
    bool last_stream_set = false;
    void* l2_flush = nullptr;
    size_t l2_flush_bytes = 0;
    // CUDA graph cache keyed on the (x, y) pointer pair: callers that rotate a few buffers (or let a
    // caching allocator hand the same blocks back) replay; a new pair is captured once, LRU-evicted.
    struct GraphEntry { const void* x; void* y; cudaGraphExec_t exec; unsigned long long stamp; };
    std::vector<GraphEntry> graphs;
    unsigned long long graph_clock = 0;
    int graph_misses = 0;                // consecutive fd_forward calls that found no captured graph for their (x, y) pair
};
static const size_t kMaxGraphs = 8;

namespace fd {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int dev_alloc(fd_plan* p, void** ptr, size_t bytes) {
    bytes = (bytes + 255) & ~size_t(255);
    FD_CUDA_OK(cudaMalloc(ptr, bytes));
    FD_CUDA_OK(cudaMemset(*ptr, 0, bytes));
    p->workspace_bytes += bytes;
    return FD_OK;
}

static void invalidate(fd_plan* p) {
    p->steps_valid = false;
    p->steps.clear();
    for (auto& g : p->graphs) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
    for (auto& s : p->stages) {
        if (s.tc) { block_tc_destroy(s.tc); s.tc = nullptr; }
        if (s.stc) { stem_tc_destroy(s.stc); s.stc = nullptr; }
        if (s.chain) { chain_tc_destroy(s.chain); s.chain = nullptr; }
        s.chained = 0;
    }
}

// fp32 host array -> device array of the plan dtype (exact when the values came from that dtype)
static int upload_as_dtype(int dtype, const float* host, size_t count, void* dev) {
    if (dtype == FD_F32) {
        FD_CUDA_OK(cudaMemcpy(dev, host, count * 4, cudaMemcpyHostToDevice));
    } else if (dtype == FD_F16) {
        std::vector<__half> tmp(count);
        for (size_t i = 0; i < count; ++i) tmp[i] = __float2half_rn(host[i]);
        FD_CUDA_OK(cudaMemcpy(dev, tmp.data(), count * 2, cudaMemcpyHostToDevice));
    } else {
        std::vector<__nv_bfloat16> tmp(count);
        for (size_t i = 0; i < count; ++i) tmp[i] = __float2bfloat16_rn(host[i]);
        FD_CUDA_OK(cudaMemcpy(dev, tmp.data(), count * 2, cudaMemcpyHostToDevice));
    }
    return FD_OK;
}

static int build_steps(fd_plan* p) {
    invalidate(p);
    const int ns = (int)p->stages.size();
    const double es = (double)dtype_size(p->dtype);
    for (auto& s : p->stages)
        if (!s.have_weights) return fail(FD_ERR_STATE, "fd_plan_set_stage_weights was not called for every stage");

    TcLaunchOpts lopts;                   // every kernel plan keeps its own copy (no process-wide launch state)
    lopts.pdl = p->opt_pdl; lopts.sleep_ns = p->opt_wait_sleep_ns; lopts.n_sms = p->n_sms; lopts.cluster = p->opt_cluster;
    Stage& head = p->stages[ns - 1];
    Stage& last = p->stages[ns - 2];
    // decode_conv6 below the last upsample: exact because a 1x1 conv, a per-channel affine and ReLU
    // act pixel-wise and nearest upsampling only replicates pixels.
    const bool fold = p->opt_fold_head && last.d.kind == FD_STAGE_DWPW && last.d.upsample && last.d.skip_src < 0;
    bool head_fused = false;

    for (int i = 0; i < ns; ++i) {
        Stage& s = p->stages[i];
        const void* in = i > 0 ? p->stages[i - 1].out_eff : nullptr;
        s.out_eff = s.out;
        s.g.in_pitch = i > 0 ? p->stages[i - 1].out_pitch : 0;
        s.g.out_pitch = s.out_pitch;
        s.g.skip_pitch = (s.d.skip_src >= 0 && !s.d.skip_mode) ? p->stages[s.d.skip_src].out_pitch : s.out_pitch;
        if (s.d.kind == FD_STAGE_STEM) {
            Step st;
            st.stage = i;
            st.name = "stem_kernel";
            st.macs = (double)s.g.n * s.g.h_out * s.g.w_out * s.g.c_out * 27.0;
            st.alg_bytes = ((double)s.g.n * 3 * s.g.h_in * s.g.w_in + (double)s.g.n * s.g.h_out * s.g.w_out * s.g.c_out) * es +
                           29.0 * s.g.c_out * 4;
            Stage* sp = &s;
            const int dtype = p->dtype;
            if (p->opt_path == 1 && stem_tc_supported(dtype, s.g)) {
                int rc = stem_tc_prepare(dtype, s.g, s.pw_w_f32, s.pw_scale, s.pw_bias, s.out, lopts, &s.stc);
                if (rc != FD_OK) return rc;
                st.name = stem_tc_name(s.stc);
                StemTcPlan* stc = s.stc;
                st.run = [stc](cudaStream_t stream, const void* x, void*) { return stem_tc_launch(stc, x, stream); };
            } else {
                st.run = [sp, dtype](cudaStream_t stream, const void* x, void*) {
                    return launch_stem(dtype, x, sp->out, sp->pw_w_f32, sp->pw_scale, sp->pw_bias, sp->g, stream);
                };
            }
            p->steps.push_back(st);
        } else if (s.d.kind == FD_STAGE_DWPW && s.chained) {
            continue;                                 // executed by the chain kernel launched at the run's first stage
        } else if (s.d.kind == FD_STAGE_DWPW) {
            // ---- a run of 3x3 stride-1 blocks on a small map: ONE chain kernel (2-CTA clusters, activations stay in shared memory)
            if (p->opt_path == 1 && p->opt_chain) {
                std::vector<BlockArgs> run;
                std::vector<StageGeom> geoms;
                int j = i;
                for (; j < ns - 1 && (int)run.size() < 8; ++j) {
                    Stage& t = p->stages[j];
                    if (t.d.kind != FD_STAGE_DWPW || t.d.ksize != 3 || t.d.stride != 1 || t.d.upsample || t.d.skip_src >= 0) break;
                    if (j > i && p->stages[j - 1].concat_src >= 0) break;     // the previous output must really be written (concat slice)
                    bool is_skip_source = false;                              // ... and so must a tensor a decoder stage will add / concatenate
                    for (int k2 = j + 1; k2 < ns; ++k2) if (p->stages[k2].d.skip_src == j) is_skip_source = true;
                    BlockArgs b{};
                    b.g = t.g;
                    b.g.in_pitch = j > 0 ? p->stages[j - 1].out_pitch : 0;
                    b.g.out_pitch = t.out_pitch;
                    b.in = j > 0 ? p->stages[j - 1].out_eff : nullptr;
                    b.out = t.out;
                    b.dw_w = t.dw_w; b.dw_scale = t.dw_scale; b.dw_bias = t.dw_bias;
                    b.pw_w = t.pw_w; b.pw_scale = t.pw_scale; b.pw_bias = t.pw_bias;
                    run.push_back(b); geoms.push_back(b.g);
                    if (is_skip_source) { ++j; break; }                       // a skip source may END a run, not sit inside one
                }
                int len = (int)run.size();
                while (len >= 2 && !chain_tc_supported(p->dtype, geoms.data(), len)) --len;
                if (len >= 2) {
                    int rc = chain_tc_prepare(p->dtype, run.data(), len, lopts, &s.chain);
                    if (rc != FD_OK) return rc;
                    Step st;
                    st.stage = i + len - 1;                                   // reported under the run's last stage (the tensor it writes)
                    st.macs = 0; st.dw_macs = 0;
                    double wb = 0;
                    for (int k2 = 0; k2 < len; ++k2) {
                        const StageGeom& g = geoms[k2];
                        const double px = (double)g.n * g.h_out * g.w_out;
                        st.dw_macs += px * g.c_in * 9.0;
                        st.macs += px * g.c_in * 9.0 + px * g.c_in * g.c_out;
                        wb += (double)g.c_in * 9 * 4 + 2.0 * g.c_in * 4 + (double)g.c_in * g.c_out * es + 2.0 * g.c_out * 4;
                        p->stages[i + k2].chained = 1;
                        p->stages[i + k2].out_eff = p->stages[i + k2].out;
                    }
                    // algorithmic bytes of the MERGED stage (SURVEY.md 8d rule): external input once + external output once + weights once
                    st.alg_bytes = ((double)geoms[0].n * geoms[0].h_in * geoms[0].w_in * geoms[0].c_in +
                                    (double)geoms[len - 1].n * geoms[len - 1].h_out * geoms[len - 1].w_out * geoms[len - 1].c_out) * es + wb;
                    char nm[200];
                    snprintf(nm, sizeof(nm), "%s{stages %d-%d}", chain_tc_name(s.chain), i, i + len - 1);
                    st.name = nm;
                    ChainTcPlan* cpn = s.chain;
                    st.run = [cpn](cudaStream_t stream, const void*, void*) { return chain_tc_launch(cpn, stream); };
                    p->steps.push_back(st);
                    s.chained = 1;
                    continue;
                }
            }
            BlockArgs a{};
            a.g = s.g;
            const bool folded_here = fold && (&s == &last);
            if (folded_here) a.g.upsample = 0;
            a.in = in;
            a.mid = s.mid;
            a.out = s.out;
            a.skip = (s.d.skip_src >= 0 && !s.d.skip_mode) ? p->stages[s.d.skip_src].out_eff : nullptr;   // concat: the source wrote its slice itself
            a.dw_w = s.dw_w; a.dw_scale = s.dw_scale; a.dw_bias = s.dw_bias;
            a.pw_w = s.pw_w; a.pw_scale = s.pw_scale; a.pw_bias = s.pw_bias;
            const double px_in = (double)s.g.n * s.g.h_in * s.g.w_in, px_out = (double)s.g.n * s.g.h_out * s.g.w_out;
            const double up = a.g.upsample ? 4.0 : 1.0;
            const double dw_macs = px_out * s.g.c_in * s.g.ksize * s.g.ksize, pw_macs = px_out * s.g.c_in * s.g.c_out;
            const double w_bytes = (double)s.g.c_in * s.g.ksize * s.g.ksize * 4 + 2.0 * s.g.c_in * 4 +
                                   (double)s.g.c_in * s.g.c_out * es + 2.0 * s.g.c_out * 4;
            const double fused_bytes = (px_in * s.g.c_in + px_out * up * s.g.c_out * (a.skip ? 2.0 : 1.0)) * es + w_bytes;
            const int dtype = p->dtype;
            const bool fuse_head = folded_here && p->opt_path == 1 && block_tc_supported(dtype, a.g, true);
            bool use_tc = p->opt_path == 1 && block_tc_supported(dtype, a.g, false);
            if (use_tc) {
                // decoder blocks with a skip accumulate INTO the skip tensor (TMA reduce-add): that buffer becomes the
                // block's output and the skip never has to be read by the SM
                const bool tma_epi = p->opt_tma_epilogue != 0;
                if (tma_epi && p->opt_inplace_skip && a.skip != nullptr) { a.out = const_cast<void*>(a.skip); s.out_eff = a.out; }
                int rc = fuse_head ? block_tc_prepare(dtype, a, head.pw_w_f32, head.head_scale, head.head_bias, head.g.act, nullptr, false, lopts, &s.tc)
                                   : block_tc_prepare(dtype, a, nullptr, 0.f, 0.f, 0, nullptr, tma_epi && (a.skip == nullptr || a.skip == a.out), lopts, &s.tc);
                if (rc != FD_OK) return rc;
                Step st;
                st.stage = i;
                st.name = block_tc_name(s.tc);
                st.macs = dw_macs + pw_macs + (fuse_head ? px_out * s.g.c_out : 0.0);
                st.dw_macs = dw_macs;
                st.alg_bytes = fuse_head ? (px_in * s.g.c_in + px_out * 4.0) * es + w_bytes + s.g.c_out * 4.0 : fused_bytes;
                BlockTcPlan* tc = s.tc;
                st.run = [tc](cudaStream_t stream, const void*, void* y) { return block_tc_launch(tc, stream, y); };
                p->steps.push_back(st);
                head_fused = fuse_head;
            } else {
                Step d;
                d.stage = i;
                d.name = s.g.ksize == 3 ? "dw_kernel<3>" : "dw_kernel<5>";
                d.macs = dw_macs;
                d.dw_macs = dw_macs;
                d.alg_bytes = (px_in + px_out) * s.g.c_in * es + (double)s.g.c_in * (s.g.ksize * s.g.ksize + 2) * 4;
                d.run = [a, dtype](cudaStream_t stream, const void*, void*) { return launch_dw(dtype, a, stream); };
                p->steps.push_back(d);
                Step q;
                q.stage = i;
                q.name = "pw_kernel";
                q.macs = pw_macs;
                q.alg_bytes = (px_out * s.g.c_in + px_out * up * s.g.c_out * (a.skip ? 2.0 : 1.0)) * es +
                              (double)s.g.c_in * s.g.c_out * es + 2.0 * s.g.c_out * 4;
                q.run = [a, dtype](cudaStream_t stream, const void*, void*) { return launch_pw(dtype, a, stream); };
                p->steps.push_back(q);
            }
        } else if (!head_fused) {  // HEAD (unless decode_conv6 already ran inside the last block's epilogue)
            const bool up = fold;
            const int hh = up ? last.g.h_out : s.g.h_in, ww = up ? last.g.w_out : s.g.w_in;
            const long long m_total = (long long)s.g.n * hh * ww;
            Step st;
            st.stage = i;
            st.name = up ? "head_kernel<up2x>" : "head_kernel";
            st.macs = (double)m_total * s.g.c_in;
            st.alg_bytes = ((double)m_total * s.g.c_in + (double)s.g.n * s.g.h_in * s.g.w_in) * es + s.g.c_in * 4.0;
            Stage* sp = &s;
            const int dtype = p->dtype;
            const int c = s.g.c_in, act = s.g.act, ipitch = s.g.in_pitch;
            st.run = [sp, in, dtype, m_total, c, ipitch, hh, ww, up, act](cudaStream_t stream, const void*, void* y) {
                return launch_head(dtype, in, y, sp->pw_w_f32, sp->head_scale, sp->head_bias, m_total, c, ipitch, hh, ww, up ? 1 : 0,
                                   act, stream);
            };
            p->steps.push_back(st);
        }
    }
    p->steps_valid = true;
    return FD_OK;
}

static int run_steps(fd_plan* p, const void* x, void* y, cudaStream_t st) {
    for (auto& s : p->steps) {
        int rc = s.run(st, x, y);
        if (rc != FD_OK) return fail(rc, std::string(fd_last_error()) + " [stage " + std::to_string(s.stage) + ": " + s.name + "]");
    }
    return FD_OK;
}

}  // namespace fd

// =========================================================================================
// C-ABI
// =========================================================================================
extern "C" {

int fd_abi_version(void) { return FD_ABI_VERSION; }
const char* fd_last_error(void) { return g_last_error.c_str(); }

int fd_plan_create(const fd_stage_desc* stages, int n_stages, int n, int h, int w, int dtype, int device, fd_plan** out) {
    if (!out) return fail(FD_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!stages || n_stages < 3) return fail(FD_ERR_INVALID, "need at least stem + one block + head");
    if (n <= 0 || h <= 0 || w <= 0) return fail(FD_ERR_INVALID, "n, h, w must be positive");
    if (h % 32 || w % 32) return fail(FD_ERR_INVALID, "H and W must be multiples of 32 (skip shapes would not line up)");
    if (dtype != FD_F32 && dtype != FD_F16 && dtype != FD_BF16) return fail(FD_ERR_INVALID, "bad dtype");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FD_ERR_CUDA, "no CUDA device: fastdepth_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(FD_ERR_INVALID, "bad device index");
    cudaDeviceProp prop{};
    FD_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(FD_ERR_UNSUPPORTED, "fastdepth_b200 is built for sm_100a (B200) only");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FD_ERR_CUDA, "cudaSetDevice failed");

    fd_plan* p = new fd_plan();
    p->n = n; p->h = h; p->w = w; p->dtype = dtype; p->device = device; p->n_sms = prop.multiProcessorCount;
    p->stages.resize(n_stages);
    const size_t es = dtype_size(dtype);
    int ch = 3, hh = h, ww = w;
    int rc = FD_OK;
    for (int i = 0; i < n_stages && rc == FD_OK; ++i) {
        Stage& s = p->stages[i];
        s.d = stages[i];
        const fd_stage_desc& d = s.d;
        const bool first = i == 0, lastst = i == n_stages - 1;
        if ((d.kind == FD_STAGE_STEM) != first || (d.kind == FD_STAGE_HEAD) != lastst ||
            (!first && !lastst && d.kind != FD_STAGE_DWPW)) { rc = fail(FD_ERR_INVALID, "stage list must be STEM, DWPW..., HEAD"); break; }
        if (d.c_in != ch) { rc = fail(FD_ERR_INVALID, "stage " + std::to_string(i) + ": c_in does not match producer"); break; }
        if (d.act != FD_ACT_RELU && d.act != FD_ACT_RELU6) { rc = fail(FD_ERR_INVALID, "bad act"); break; }
        s.g.n = n; s.g.h_in = hh; s.g.w_in = ww; s.g.c_in = d.c_in; s.g.c_out = d.c_out;
        s.g.ksize = d.ksize; s.g.stride = d.stride; s.g.act = d.act; s.g.upsample = d.upsample ? 1 : 0;
        if (d.kind == FD_STAGE_STEM) {
            if (d.ksize != 3 || d.c_in != 3 || d.c_out % 8 || d.stride < 1 || d.stride > 2 || d.upsample || d.skip_src >= 0) {
                rc = fail(FD_ERR_INVALID, "stem must be 3x3, c_in 3, c_out % 8 == 0, stride 1|2"); break; }
            s.g.h_out = (hh + 2 - 3) / d.stride + 1; s.g.w_out = (ww + 2 - 3) / d.stride + 1;
        } else if (d.kind == FD_STAGE_DWPW) {
            if ((d.ksize != 3 && d.ksize != 5) || d.stride < 1 || d.stride > 2 || d.c_in % 8 || d.c_out % 8 || d.c_out <= 0) {
                rc = fail(FD_ERR_INVALID, "block stage needs k in {3,5}, stride 1|2, channels % 8 == 0"); break; }
            const int pad = (d.ksize - 1) / 2;
            s.g.h_out = (hh + 2 * pad - d.ksize) / d.stride + 1; s.g.w_out = (ww + 2 * pad - d.ksize) / d.stride + 1;
        } else {
            if (d.ksize != 1 || d.c_out != 1 || d.c_in % 8 || d.upsample || d.skip_src >= 0) {
                rc = fail(FD_ERR_INVALID, "head must be 1x1, c_out 1, c_in % 8 == 0"); break; }
            s.g.h_out = hh; s.g.w_out = ww;
        }
        s.out_h = s.g.h_out * (s.g.upsample ? 2 : 1);
        s.out_w = s.g.w_out * (s.g.upsample ? 2 : 1);
        s.out_c = d.c_out; s.out_pitch = d.c_out;
        if (d.skip_src >= 0) {
            if (d.kind != FD_STAGE_DWPW || !d.upsample || d.skip_src >= i) { rc = fail(FD_ERR_INVALID, "bad skip_src"); break; }
            Stage& src = p->stages[d.skip_src];
            if (src.out_h != s.out_h || src.out_w != s.out_w || (!d.skip_mode && src.g.c_out != d.c_out)) {
                rc = fail(FD_ERR_INVALID, "stage " + std::to_string(i) + ": skip tensor shape does not match the upsampled output");
                break; }
            if (d.skip_mode) {
                // concatenation (models.py:806-811): one wide NHWC buffer [.., c_out + c_skip]; this stage writes the channel
                // slice [0, c_out), the skip source is re-pointed to write (and be read by its consumer) in slice [c_out, ..)
                if (d.skip_mode != 1 || src.concat_src >= 0 || src.d.skip_src >= 0) { rc = fail(FD_ERR_INVALID, "bad skip_mode / skip source"); break; }
                s.out_c = d.c_out + src.g.c_out; s.out_pitch = s.out_c;
            }
        }
        if (d.kind != FD_STAGE_HEAD) {
            rc = dev_alloc(p, &s.out, (size_t)n * s.out_h * s.out_w * s.out_pitch * es);
            if (rc) break;
            s.out_alloc = s.out;
            if (d.skip_src >= 0 && d.skip_mode) {
                Stage& src = p->stages[d.skip_src];
                // the source's own dense buffer stays allocated (freed with the plan) but is no longer used
                src.out = static_cast<char*>(s.out) + (size_t)d.c_out * es;
                src.out_pitch = s.out_pitch;
                src.concat_src = i;
            }
        }
        if (d.kind == FD_STAGE_DWPW) {
            if ((rc = dev_alloc(p, &s.mid, (size_t)n * s.g.h_out * s.g.w_out * d.c_in * es))) break;
            if ((rc = dev_alloc(p, (void**)&s.dw_w, (size_t)d.ksize * d.ksize * d.c_in * 4))) break;
            if ((rc = dev_alloc(p, (void**)&s.dw_scale, (size_t)d.c_in * 4))) break;
            if ((rc = dev_alloc(p, (void**)&s.dw_bias, (size_t)d.c_in * 4))) break;
            if ((rc = dev_alloc(p, &s.pw_w, (size_t)d.c_in * d.c_out * es))) break;
        } else if (d.kind == FD_STAGE_STEM) {
            if ((rc = dev_alloc(p, (void**)&s.pw_w_f32, (size_t)27 * d.c_out * 4))) break;
        } else {
            if ((rc = dev_alloc(p, (void**)&s.pw_w_f32, (size_t)d.c_in * 4))) break;
        }
        if ((rc = dev_alloc(p, (void**)&s.pw_scale, (size_t)d.c_out * 4))) break;
        if ((rc = dev_alloc(p, (void**)&s.pw_bias, (size_t)d.c_out * 4))) break;
        ch = s.out_c; hh = s.out_h; ww = s.out_w;
    }
    if (rc == FD_OK && (hh != h || ww != w)) rc = fail(FD_ERR_INVALID, "stage list does not return to the input resolution");
    if (rc != FD_OK) {
        std::string keep = g_last_error;
        fd_plan_destroy(p);
        g_last_error = keep;
        return rc;
    }
    *out = p;
    return FD_OK;
}

int fd_plan_set_stage_weights(fd_plan* p, int stage, const float* dw_w, const float* dw_scale, const float* dw_bias,
                              const float* pw_w, const float* pw_scale, const float* pw_bias) {
    if (!p) return fail(FD_ERR_INVALID, "plan is NULL");
    if (stage < 0 || stage >= (int)p->stages.size()) return fail(FD_ERR_INVALID, "bad stage index");
    if (!pw_w || !pw_scale || !pw_bias) return fail(FD_ERR_INVALID, "pw_w / pw_scale / pw_bias are required");
    DeviceGuard guard(p->device);
    Stage& s = p->stages[stage];
    const fd_stage_desc& d = s.d;
    if (d.kind == FD_STAGE_DWPW) {
        if (!dw_w || !dw_scale || !dw_bias) return fail(FD_ERR_INVALID, "block stage needs dw_w / dw_scale / dw_bias");
        const int kk = d.ksize * d.ksize;
        std::vector<float> t((size_t)kk * d.c_in);            // [c][k][k] -> [k*k][c]
        for (int c = 0; c < d.c_in; ++c)
            for (int j = 0; j < kk; ++j) t[(size_t)j * d.c_in + c] = dw_w[(size_t)c * kk + j];
        FD_CUDA_OK(cudaMemcpy(s.dw_w, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        FD_CUDA_OK(cudaMemcpy(s.dw_scale, dw_scale, (size_t)d.c_in * 4, cudaMemcpyHostToDevice));
        FD_CUDA_OK(cudaMemcpy(s.dw_bias, dw_bias, (size_t)d.c_in * 4, cudaMemcpyHostToDevice));
        int rc = upload_as_dtype(p->dtype, pw_w, (size_t)d.c_in * d.c_out, s.pw_w);
        if (rc) return rc;
    } else if (d.kind == FD_STAGE_STEM) {
        std::vector<float> t((size_t)27 * d.c_out);            // [co][ci][ky][kx] -> [(ci,ky,kx)][co]
        for (int co = 0; co < d.c_out; ++co)
            for (int j = 0; j < 27; ++j) t[(size_t)j * d.c_out + co] = pw_w[(size_t)co * 27 + j];
        FD_CUDA_OK(cudaMemcpy(s.pw_w_f32, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    } else {
        FD_CUDA_OK(cudaMemcpy(s.pw_w_f32, pw_w, (size_t)d.c_in * 4, cudaMemcpyHostToDevice));
        s.head_scale = pw_scale[0];
        s.head_bias = pw_bias[0];
    }
    FD_CUDA_OK(cudaMemcpy(s.pw_scale, pw_scale, (size_t)d.c_out * 4, cudaMemcpyHostToDevice));
    FD_CUDA_OK(cudaMemcpy(s.pw_bias, pw_bias, (size_t)d.c_out * 4, cudaMemcpyHostToDevice));
    s.have_weights = true;
    invalidate(p);
    return FD_OK;
}

static int* option_slot(fd_plan* p, const char* name) {
    if (!p || !name) return nullptr;
    if (!strcmp(name, "path")) return &p->opt_path;
    if (!strcmp(name, "fold_head")) return &p->opt_fold_head;
    if (!strcmp(name, "graph")) return &p->opt_graph;
    if (!strcmp(name, "tma_epilogue")) return &p->opt_tma_epilogue;
    if (!strcmp(name, "inplace_skip")) return &p->opt_inplace_skip;
    if (!strcmp(name, "pdl")) return &p->opt_pdl;
    if (!strcmp(name, "wait_sleep_ns")) return &p->opt_wait_sleep_ns;
    if (!strcmp(name, "chain")) return &p->opt_chain;
    if (!strcmp(name, "cluster")) return &p->opt_cluster;
    return nullptr;
}

int fd_plan_set_option(fd_plan* p, const char* name, int value) {
    int* slot = option_slot(p, name);
    if (!slot) return fail(FD_ERR_INVALID, std::string("unknown option: ") + (name ? name : "(null)"));
    const bool is_time = !strcmp(name, "wait_sleep_ns");
    if (is_time ? (value < 0 || value > 100000) : (value != 0 && value != 1))
        return fail(FD_ERR_INVALID, is_time ? "wait_sleep_ns must be in [0, 100000]" : "option value must be 0 or 1");
    if (*slot != value) { *slot = value; invalidate(p); }
    return FD_OK;
}

int fd_plan_get_option(fd_plan* p, const char* name, int* value) {
    int* slot = option_slot(p, name);
    if (!slot || !value) return fail(FD_ERR_INVALID, "unknown option or NULL value");
    *value = *slot;
    return FD_OK;
}

static int ensure_steps(fd_plan* p) {
    if (p->steps_valid) return FD_OK;
    return build_steps(p);
}

static int forward_enqueue(fd_plan* p, const void* x_dev, void* y_dev, cudaStream_t st);

int fd_forward(fd_plan* p, const void* x_dev, void* y_dev, void* stream) {
    if (!p || !x_dev || !y_dev) return fail(FD_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    int rc = ensure_steps(p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // A plan owns ONE set of activation buffers: a forward enqueued on another stream than the previous one first waits for that
    // one to finish (stream-ordered, no host synchronisation) instead of silently running into its intermediates.  Concurrency
    // is spelled with several plans (fastdepth_b200.engine.ForwardLanes).
    if (!p->last_done) FD_CUDA_OK(cudaEventCreateWithFlags(&p->last_done, cudaEventDisableTiming));
    if (p->last_stream_set && p->last_stream != st) FD_CUDA_OK(cudaStreamWaitEvent(st, p->last_done, 0));
    rc = forward_enqueue(p, x_dev, y_dev, st);
    if (rc) return rc;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
        FD_CUDA_OK(cudaEventRecord(p->last_done, st));
        p->last_stream = st; p->last_stream_set = true;
    }
    return FD_OK;
}

static int forward_enqueue(fd_plan* p, const void* x_dev, void* y_dev, cudaStream_t st) {
    int rc = FD_OK;
    if (!p->opt_graph) return run_steps(p, x_dev, y_dev, st);

    // Replay from a CUDA graph captured for this (x, y) pair.
    for (auto& g : p->graphs)
        if (g.x == x_dev && g.y == y_dev) {
            g.stamp = ++p->graph_clock;
            p->graph_misses = 0;
            FD_CUDA_OK(cudaGraphLaunch(g.exec, st));
            return FD_OK;
        }
    // A caller that never repeats an (x, y) pair (fresh output tensors that it keeps alive, a stream of distinct inputs) would
    // pay a capture + instantiate + eviction per call: after two cache-fulls of consecutive misses launch directly until a
    // pair repeats again (19 PDL-chained launches cost far less than one capture).
    if (++p->graph_misses > 2 * (int)kMaxGraphs) {
        if (p->graph_misses > (1 << 30)) p->graph_misses = 2 * (int)kMaxGraphs + 1;
        return run_steps(p, x_dev, y_dev, st);
    }
    cudaStream_t cap;
    FD_CUDA_OK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
        rc = run_steps(p, x_dev, y_dev, cap);
        e = cudaStreamEndCapture(cap, &graph);
        if (rc == FD_OK && e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    }
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    if (rc != FD_OK) return rc;
    if (e != cudaSuccess) return fail(FD_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    if (p->graphs.size() >= kMaxGraphs) {
        size_t victim = 0;
        for (size_t i = 1; i < p->graphs.size(); ++i)
            if (p->graphs[i].stamp < p->graphs[victim].stamp) victim = i;
        cudaGraphExecDestroy(p->graphs[victim].exec);
        p->graphs.erase(p->graphs.begin() + victim);
    }
    p->graphs.push_back({x_dev, y_dev, exec, ++p->graph_clock});
    FD_CUDA_OK(cudaGraphLaunch(exec, st));
    return FD_OK;
}

int fd_forward_host(fd_plan* p, const void* x_host, void* y_host, void* stream) {
    if (!p || !x_host || !y_host) return fail(FD_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    const size_t es = dtype_size(p->dtype);
    const size_t xb = (size_t)p->n * 3 * p->h * p->w * es, yb = (size_t)p->n * p->h * p->w * es;
    if (!p->stage_x) {
        int rc = dev_alloc(p, &p->stage_x, xb);
        if (rc) return rc;
        if ((rc = dev_alloc(p, &p->stage_y, yb))) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FD_CUDA_OK(cudaMemcpyAsync(p->stage_x, x_host, xb, cudaMemcpyHostToDevice, st));
    int rc = fd_forward(p, p->stage_x, p->stage_y, stream);
    if (rc) return rc;
    FD_CUDA_OK(cudaMemcpyAsync(y_host, p->stage_y, yb, cudaMemcpyDeviceToHost, st));
    FD_CUDA_OK(cudaStreamSynchronize(st));
    return FD_OK;
}

static const int kPipeSlots = 3;

int fd_pipeline_submit(fd_plan* p, const void* x_host, void* y_host, unsigned long long* ticket) {
    if (!p || !x_host || !y_host || !ticket) return fail(FD_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    const size_t es = dtype_size(p->dtype);
    const size_t xb = (size_t)p->n * 3 * p->h * p->w * es, yb = (size_t)p->n * p->h * p->w * es;
    if (!p->pipe_h2d) {
        FD_CUDA_OK(cudaStreamCreateWithFlags(&p->pipe_h2d, cudaStreamNonBlocking));
        FD_CUDA_OK(cudaStreamCreateWithFlags(&p->pipe_run, cudaStreamNonBlocking));
        FD_CUDA_OK(cudaStreamCreateWithFlags(&p->pipe_d2h, cudaStreamNonBlocking));
        for (auto& sl : p->pipe) {
            int rc = dev_alloc(p, &sl.x, xb);
            if (rc) return rc;
            if ((rc = dev_alloc(p, &sl.y, yb))) return rc;
            FD_CUDA_OK(cudaEventCreateWithFlags(&sl.up, cudaEventDisableTiming));
            FD_CUDA_OK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
            FD_CUDA_OK(cudaEventCreateWithFlags(&sl.down, cudaEventDisableTiming));
        }
    }
    const unsigned long long t = p->pipe_next;
    fd_plan::PipeSlot& sl = p->pipe[t % kPipeSlots];
    if (sl.busy) {                                   // slot still owned by ticket t - kPipeSlots: wait for its download
        FD_CUDA_OK(cudaEventSynchronize(sl.down));
        sl.busy = false;
    }
    // upload on the copy-in stream; the forward waits for it; the download waits for the forward
    FD_CUDA_OK(cudaMemcpyAsync(sl.x, x_host, xb, cudaMemcpyHostToDevice, p->pipe_h2d));
    FD_CUDA_OK(cudaEventRecord(sl.up, p->pipe_h2d));
    FD_CUDA_OK(cudaStreamWaitEvent(p->pipe_run, sl.up, 0));
    int rc = fd_forward(p, sl.x, sl.y, p->pipe_run);
    if (rc) return rc;
    FD_CUDA_OK(cudaEventRecord(sl.done, p->pipe_run));
    FD_CUDA_OK(cudaStreamWaitEvent(p->pipe_d2h, sl.done, 0));
    FD_CUDA_OK(cudaMemcpyAsync(y_host, sl.y, yb, cudaMemcpyDeviceToHost, p->pipe_d2h));
    FD_CUDA_OK(cudaEventRecord(sl.down, p->pipe_d2h));
    // (a slot is only re-filled after the host saw its previous download complete, which implies its forward has
    //  finished reading sl.x -- no stream-level dependency is needed, and adding one would serialise the uploads)
    sl.busy = true;
    *ticket = t;
    p->pipe_next = t + 1;
    return FD_OK;
}

int fd_pipeline_wait(fd_plan* p, unsigned long long ticket) {
    if (!p) return fail(FD_ERR_INVALID, "NULL plan");
    if (ticket >= p->pipe_next) return fail(FD_ERR_INVALID, "unknown ticket");
    if (ticket + kPipeSlots < p->pipe_next) return FD_OK;        // its slot was already recycled, hence complete
    DeviceGuard guard(p->device);
    fd_plan::PipeSlot& sl = p->pipe[ticket % kPipeSlots];
    if (sl.busy) {
        FD_CUDA_OK(cudaEventSynchronize(sl.down));
        sl.busy = false;
    }
    return FD_OK;
}

int fd_stage_buffer(fd_plan* p, int stage, int which, void** dev_ptr, int* n, int* h, int* w, int* c, int* c_stride) {
    if (!p || stage < 0 || stage >= (int)p->stages.size() || !dev_ptr) return fail(FD_ERR_INVALID, "bad argument");
    Stage& s = p->stages[stage];
    int hh, ww, cc;
    void* ptr;
    if (which == 0) {
        ptr = s.out_eff ? s.out_eff : s.out; hh = s.out_h; ww = s.out_w; cc = s.g.c_out;
        // with decode_conv6 folded below the last upsample the last block writes its low-res output
        if (p->opt_fold_head && stage == (int)p->stages.size() - 2 && s.d.upsample && s.d.skip_src < 0) { hh = s.g.h_out; ww = s.g.w_out; }
    } else if (which == 1) {
        ptr = s.mid; hh = s.g.h_out; ww = s.g.w_out; cc = s.g.c_in;
    } else {
        return fail(FD_ERR_INVALID, "which must be 0 or 1");
    }
    if (!ptr) return fail(FD_ERR_STATE, "stage has no such buffer");
    *dev_ptr = ptr;
    if (n) *n = p->n;
    if (h) *h = hh;
    if (w) *w = ww;
    if (c) *c = cc;
    if (c_stride) *c_stride = (which == 0 && s.out_pitch > 0) ? s.out_pitch : cc;
    return FD_OK;
}

int fd_plan_launches_per_forward(fd_plan* p, int* n_launches) {
    if (!p || !n_launches) return fail(FD_ERR_INVALID, "NULL argument");
    DeviceGuard guard(p->device);
    int rc = ensure_steps(p);
    if (rc) return rc;
    *n_launches = (int)p->steps.size();
    return FD_OK;
}

int fd_plan_workspace_bytes(fd_plan* p, size_t* bytes) {
    if (!p || !bytes) return fail(FD_ERR_INVALID, "NULL argument");
    *bytes = p->workspace_bytes;
    return FD_OK;
}

int fd_plan_step_count(fd_plan* p, int* n_steps) { return fd_plan_launches_per_forward(p, n_steps); }

int fd_plan_step_info(fd_plan* p, int step, int* stage, double* alg_bytes, double* macs, char* kernel_name, int name_cap) {
    if (!p) return fail(FD_ERR_INVALID, "NULL plan");
    DeviceGuard guard(p->device);
    int rc = ensure_steps(p);
    if (rc) return rc;
    if (step < 0 || step >= (int)p->steps.size()) return fail(FD_ERR_INVALID, "bad step index");
    const Step& s = p->steps[step];
    if (stage) *stage = s.stage;
    if (alg_bytes) *alg_bytes = s.alg_bytes;
    if (macs) *macs = s.macs;
    if (kernel_name && name_cap > 0) {
        strncpy(kernel_name, s.name.c_str(), name_cap - 1);
        kernel_name[name_cap - 1] = 0;
    }
    return FD_OK;
}

int fd_plan_step_macs(fd_plan* p, int step, double* dw_macs, double* dense_macs) {
    if (!p) return fail(FD_ERR_INVALID, "NULL plan");
    DeviceGuard guard(p->device);
    int rc = ensure_steps(p);
    if (rc) return rc;
    if (step < 0 || step >= (int)p->steps.size()) return fail(FD_ERR_INVALID, "bad step index");
    const Step& s = p->steps[step];
    if (dw_macs) *dw_macs = s.dw_macs;
    if (dense_macs) *dense_macs = s.macs - s.dw_macs;
    return FD_OK;
}

int fd_plan_time_steps(fd_plan* p, const void* x_dev, void* y_dev, void* stream, int warmup, int iters, int flush_l2,
                       float* ms_out) {
    if (!p || !x_dev || !y_dev || !ms_out || iters <= 0) return fail(FD_ERR_INVALID, "bad argument");
    DeviceGuard guard(p->device);
    int rc = ensure_steps(p);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (flush_l2 && !p->l2_flush) {
        p->l2_flush_bytes = size_t(256) << 20;               // > 126 MB L2
        FD_CUDA_OK(cudaMalloc(&p->l2_flush, p->l2_flush_bytes));
    }
    rc = run_steps(p, x_dev, y_dev, st);                      // make every intermediate valid
    if (rc) return rc;
    cudaEvent_t e0, e1;
    FD_CUDA_OK(cudaEventCreate(&e0));
    FD_CUDA_OK(cudaEventCreate(&e1));
    for (size_t i = 0; i < p->steps.size() && rc == FD_OK; ++i) {
        for (int k = 0; k < warmup && rc == FD_OK; ++k) rc = p->steps[i].run(st, x_dev, y_dev);
        float total = 0.f;
        for (int k = 0; k < iters && rc == FD_OK; ++k) {
            if (flush_l2) cudaMemsetAsync(p->l2_flush, k & 0xff, p->l2_flush_bytes, st);
            cudaEventRecord(e0, st);
            rc = p->steps[i].run(st, x_dev, y_dev);
            cudaEventRecord(e1, st);
            cudaError_t e = cudaEventSynchronize(e1);
            if (e != cudaSuccess) { rc = fail(FD_ERR_CUDA, std::string("step ") + p->steps[i].name + ": " + cudaGetErrorString(e)); break; }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
        }
        ms_out[i] = total / iters;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return rc;
}

int fd_plan_trace_stage(fd_plan* p, int stage, void* y_dev, void* stream, unsigned long long* out_host, int cap, int* rows, int* cols) {
    if (!p || stage < 0 || stage >= (int)p->stages.size() || !out_host || !rows || !cols) return fail(FD_ERR_INVALID, "bad argument");
    DeviceGuard guard(p->device);
    int rc = ensure_steps(p);
    if (rc) return rc;
    if (cap < 12 * 256) return fail(FD_ERR_INVALID, "trace buffer too small (need 3072 entries)");
    if (p->stages[stage].chain) return chain_tc_trace(p->stages[stage].chain, (cudaStream_t)stream, out_host, rows, cols);
    if (!p->stages[stage].tc) return fail(FD_ERR_STATE, "stage does not run the fused block kernel");
    return block_tc_trace(p->stages[stage].tc, (cudaStream_t)stream, y_dev, out_host, rows, cols);
}

int fd_debug_block_plan(int ksize, int stride, int h_out, int w_out, int n, int c_in, int c_out, int head, int* out, int cap) {
    if (!out || cap < 16) return fail(FD_ERR_INVALID, "need an int[16] output");
    const BlockPlanOut q = block_tc_debug_plan(ksize, stride, h_out, w_out, n, c_in, c_out, head);
    const int v[16] = {q.ok, q.splits, q.n_cta, q.items, q.kblocks, q.s_in, q.s_a, q.s_b, q.bn, q.nb, q.b_resident, q.epi_groups,
                       q.n_stg, q.smem_bytes, q.tmem_cols, q.in_stage_stride};
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    if (cap >= 18) { out[16] = q.nacc; out[17] = q.epi_colsplit; }
    if (cap >= 19) out[18] = q.epi_wide;
    if (cap >= 20) out[19] = q.cs;
    if (cap >= 21) out[20] = q.dw_teams;
    return FD_OK;
}

int fd_metrics_accumulate(const void* pred_dev, const float* target_dev, int dtype, int n, int hw, double* sums_dev,
                          int device, void* stream) {
    if (!pred_dev || !target_dev || !sums_dev || n < 0 || hw <= 0) return fail(FD_ERR_INVALID, "bad argument");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FD_ERR_CUDA, "cudaSetDevice failed");
    return launch_metrics(dtype, pred_dev, target_dev, n, hw, sums_dev, (cudaStream_t)stream);
}

int fd_nyu_val_gather(const uint8_t* rgb_dev, const float* depth_dev, const int* rows_dev, const int* cols_dev, int n, int h_in,
                      int w_in, int out_h, int out_w, int dtype, void* x_dev, float* target_dev, int device, void* stream) {
    if (!rgb_dev || !rows_dev || !cols_dev || !x_dev || n < 0 || h_in <= 0 || w_in <= 0 || out_h <= 0 || out_w <= 0)
        return fail(FD_ERR_INVALID, "bad argument");
    if ((depth_dev == nullptr) != (target_dev == nullptr)) return fail(FD_ERR_INVALID, "depth and target must come together");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(FD_ERR_CUDA, "cudaSetDevice failed");
    return launch_nyu_val_gather(dtype, rgb_dev, depth_dev, rows_dev, cols_dev, n, h_in, w_in, out_h, out_w, x_dev, target_dev,
                                 (cudaStream_t)stream);
}

void fd_plan_destroy(fd_plan* p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    invalidate(p);
    for (auto& s : p->stages) {
        cudaFree(s.out_alloc ? s.out_alloc : s.out); cudaFree(s.mid); cudaFree(s.dw_w); cudaFree(s.dw_scale); cudaFree(s.dw_bias);
        cudaFree(s.pw_w); cudaFree(s.pw_w_f32); cudaFree(s.pw_scale); cudaFree(s.pw_bias);
    }
    for (auto& sl : p->pipe) {
        cudaFree(sl.x); cudaFree(sl.y);
        if (sl.up) cudaEventDestroy(sl.up);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.down) cudaEventDestroy(sl.down);
    }
    if (p->last_done) cudaEventDestroy(p->last_done);
    if (p->pipe_h2d) cudaStreamDestroy(p->pipe_h2d);
    if (p->pipe_run) cudaStreamDestroy(p->pipe_run);
    if (p->pipe_d2h) cudaStreamDestroy(p->pipe_d2h);
    cudaFree(p->stage_x); cudaFree(p->stage_y); cudaFree(p->l2_flush);
    delete p;
}

}  // extern "C"
