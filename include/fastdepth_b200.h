/*
 * fastdepth_b200 -- C-ABI of the B200-native FastDepth forward path.
 *
 * This is the drop-in boundary for ONE hot path of dwofk/fast-depth:
 *     MobileNetSkipAdd.forward            (reference models.py:706-732)
 * i.e. the MobileNet encoder (reference imagenet/mobilenet.py:22-54), the NNConv5
 * depthwise-separable decoder with 2x nearest upsampling and additive skips
 * (reference models.py:61-75, 683-698, 720-731), plus the per-image depth metrics the
 * only caller computes on the result (reference metrics.py:31-55 via main.py:80-82).
 *
 * The reference has no FFI of its own (it is pure Python on PyTorch); the entry points
 * below are what a binding for this path binds.  The repo's own binding is
 * fastdepth_b200/_lib.py (ctypes); INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain C: opaque handle, ints, raw pointers; no C++/torch types cross the boundary.
 *   - every function returns 0 on success, a negative fd_status otherwise; the message is
 *     retrievable (thread-local) with fd_last_error().  No C++ exception crosses the ABI.
 *   - "dev" pointers are CUDA device pointers on the plan's device, "host" pointers are
 *     ordinary (ideally pinned) host memory.  `stream` is a cudaStream_t passed as void*.
 *   - all work is stream-ordered and asynchronous unless stated otherwise.
 *   - a plan is not thread-safe; use one plan per (device, N, H, W, dtype) per thread.
 *   - there is NO CPU fallback: without a CUDA device every call fails with FD_ERR_CUDA.
 */
#ifndef FASTDEPTH_B200_H
#define FASTDEPTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FD_ABI_VERSION 2

typedef struct fd_plan fd_plan;

typedef enum {
    FD_OK = 0,
    FD_ERR_INVALID = -1,      /* bad argument / unsupported shape (H or W % 32, C % 8 ...)   */
    FD_ERR_CUDA = -2,         /* CUDA runtime / driver error, or no device                   */
    FD_ERR_STATE = -3,        /* call order violated (weights missing, plan destroyed ...)   */
    FD_ERR_UNSUPPORTED = -4   /* valid request this build has no kernel for                  */
} fd_status;

/* storage + arithmetic type of activations and of the pointwise contraction inputs;
 * accumulation, BN scale/bias and activations are always computed in fp32.
 * (reference: dtype of the tensor handed to model(input), main.py:68,74-75) */
typedef enum { FD_F32 = 0, FD_F16 = 1, FD_BF16 = 2 } fd_dtype;

typedef enum {
    FD_STAGE_STEM = 0,   /* dense 3x3 stride-s conv + BN + act, NCHW in -> NHWC out
                            (conv_bn, reference imagenet/mobilenet.py:22-27)                 */
    FD_STAGE_DWPW = 1,   /* depthwise kxk(stride) + BN + act -> pointwise 1x1 + BN + act
                            (conv_dw, reference imagenet/mobilenet.py:29-38;
                             depthwise+pointwise, reference models.py:61-75, 683-697)        */
    FD_STAGE_HEAD = 2    /* pointwise C->1 + BN + act, NHWC in -> [N,1,H,W] out
                            (decode_conv6, reference models.py:698, 731)                     */
} fd_stage_kind;

typedef enum { FD_ACT_RELU = 0, FD_ACT_RELU6 = 1 } fd_act;

typedef struct {
    int32_t kind;        /* fd_stage_kind                                                    */
    int32_t c_in;        /* input channels  (3 for the stem)                                 */
    int32_t c_out;       /* output channels (1 for the head)                                 */
    int32_t ksize;       /* spatial kernel: 3 (stem, encoder dw), 5 (decoder dw), 1 (head)   */
    int32_t stride;      /* stride of the spatial conv (1 or 2)                              */
    int32_t act;         /* fd_act applied after BOTH halves (encoder ReLU6, decoder ReLU)   */
    int32_t upsample;    /* 1: output is 2x nearest-upsampled (F.interpolate, models.py:723) */
    int32_t skip_src;    /* stage index whose output is combined with this stage's upsampled
                            output (models.py:724-729 / 806-811), or -1                      */
    int32_t skip_mode;   /* 0: ADD (MobileNetSkipAdd, models.py:724-729); 1: CONCATENATE along
                            channels, upsampled output first (MobileNetSkipConcat,
                            models.py:806-811) -- the next stage then has c_in = c_out + c_skip  */
} fd_stage_desc;

/* Build a plan for a stage list (always: 1 STEM, k DWPW, 1 HEAD) at a fixed problem size.
 * Allocates NHWC activation buffers and packed-weight storage on `device`.
 * H and W must be multiples of 32 (reference forward's skip shapes only line up then),
 * every c_in/c_out except the stem's c_in and the head's c_out a multiple of 8. */
int fd_plan_create(const fd_stage_desc* stages, int n_stages,
                   int n, int h, int w, int dtype /* fd_dtype */, int device, fd_plan** out);

/* Upload one stage's parameters (HOST pointers, fp32, BatchNorm already folded to a
 * per-channel affine y = conv * scale + bias by the caller; folding is exact in fp32,
 * reference BN eval formula, eps 1e-5).  Synchronous.  May be called again after a
 * parameter update.
 *   STEM : dw_* = NULL ; pw_w = [c_out][c_in][k][k]  ; pw_scale/pw_bias = [c_out]
 *   DWPW : dw_w = [c_in][k][k], dw_scale/dw_bias = [c_in] ; pw_w = [c_out][c_in], pw_scale/pw_bias = [c_out]
 *   HEAD : dw_* = NULL ; pw_w = [1][c_in]            ; pw_scale/pw_bias = [1]              */
int fd_plan_set_stage_weights(fd_plan* plan, int stage,
                              const float* dw_w, const float* dw_scale, const float* dw_bias,
                              const float* pw_w, const float* pw_scale, const float* pw_bias);

/* Tunables, by name.  Unknown names fail with FD_ERR_INVALID.
 *   "path"       0 = SIMT reference-quality kernels (all dtypes), 1 = fused tcgen05 block
 *                kernels where available (16-bit dtypes)            [default 1]
 *   "fold_head"  1 = apply decode_conv6 below the last upsample (exact: 1x1 conv/BN/ReLU
 *                commute with nearest upsampling) [default 1]
 *   "graph"      1 = replay fd_forward from a captured CUDA graph   [default 1]
 *   "tma_epilogue" 1 = fused blocks write their output tiles with TMA tensor stores  [default 1]
 *   "inplace_skip" 1 = decoder blocks ADD their upsampled output into the skip tensor in place
 *                (TMA reduce-add); the skip source's stage buffer then holds the decoder output
 *                after fd_forward (set 0 for stage-by-stage inspection)  [default 1]
 *   "pdl"        1 = tensor-core kernels are launched with programmatic dependent launch so that each
 *                kernel's prologue overlaps the previous kernel's tail.  With one 227 KB CTA per SM the early-launched
 *                dependents mostly hold SMs idle while they wait for the previous grid: measured 596.8 us per forward
 *                with it, 589.4 us without (round 2), and it costs more when several plans run concurrently
 *                [default 0]
 *   "chain"      1 = a run of consecutive 3x3 stride-1 blocks on a small feature map (conv7..conv11 at 14x14) executes as ONE
 *                kernel on 2-CTA clusters with every intermediate activation resident in shared memory; the intermediate
 *                stages' buffers are then not written (set 0 for stage-by-stage inspection)  [default 1]
 *   "cluster"    1 = the block planner may run a block on thread-block clusters of 2 or 4 CTAs that share one 128-pixel tile:
 *                each CTA computes the depthwise half of a quarter (half) of the K-blocks, broadcasts its operand tiles to
 *                the others through distributed shared memory and runs the MMAs of one output-channel split (chosen by
 *                the planner's cost model for the small-map, many-channel blocks)  [default 1]
 *   "wait_sleep_ns" > 0: latency-tolerant roles of the fused block kernel (epilogue warps waiting for an
 *                accumulator, TMA producer waiting for a free stage) sleep this many ns between barrier
 *                probes instead of spinning (measured: no effect on B200, the spinning waiters do not
 *                take issue slots the depthwise warps could use)  [default 0]   */
int fd_plan_set_option(fd_plan* plan, const char* name, int value);
int fd_plan_get_option(fd_plan* plan, const char* name, int* value);

/* The hot path.  x_dev: [N,3,H,W] contiguous, plan dtype.  y_dev: [N,1,H,W] contiguous,
 * plan dtype.  Enqueues on `stream`; returns without synchronising.  A plan owns one set of activation
 * buffers: calls on different streams are ordered after each other by an event (never corrupted, never
 * overlapped); to keep several forwards in flight use several plans (one per stream).
 * Replaces: pred = model(input)  (reference main.py:74-75 -> models.py:706-732). */
int fd_forward(fd_plan* plan, const void* x_dev, void* y_dev, void* stream);

/* Same, end to end from HOST buffers: H2D copy of x, forward, D2H copy of y, then waits for
 * the stream.  (reference main.py:68 input.cuda() ... main.py:85-98 pred.cpu()) */
int fd_forward_host(fd_plan* plan, const void* x_host, void* y_host, void* stream);

/* Pipelined end-to-end evaluation from HOST buffers (what the reference's DataLoader(pin_memory) + input.cuda()
 * + pred.cpu() loop does, main.py:40-41, 68, 85-98, with the copies overlapped): submit() enqueues the upload of
 * x_host (pinned), the forward and the download into y_host (pinned) on the plan's own three streams and returns
 * a ticket at once; up to 3 batches are in flight (submit blocks on the oldest when all slots are busy).
 * wait(ticket) returns when y_host holds that batch's depth maps.  Tickets complete in order. */
int fd_pipeline_submit(fd_plan* plan, const void* x_host, void* y_host, unsigned long long* ticket);
int fd_pipeline_wait(fd_plan* plan, unsigned long long ticket);

/* Introspection for stage-parity tests: the NHWC buffer stage `stage` wrote in the last
 * fd_forward (valid until the next one).  c_stride = elements between pixels.
 * which = 0: the stage output (after upsample/skip-add); 1: the depthwise intermediate
 * (only materialised on path 0). */
int fd_stage_buffer(fd_plan* plan, int stage, int which, void** dev_ptr,
                    int* n, int* h, int* w, int* c, int* c_stride);

/* Bookkeeping used by bench.py.  A "step" is one kernel launch of fd_forward under the current
 * options (a DWPW stage is one fused step on path 1, a dw + a pw step on path 0). */
int fd_plan_launches_per_forward(fd_plan* plan, int* n_launches);
int fd_plan_workspace_bytes(fd_plan* plan, size_t* bytes);
int fd_plan_step_count(fd_plan* plan, int* n_steps);
/* Stage index, ALGORITHMIC bytes (external inputs once + outputs once + weights once; SURVEY.md
 * section 8d) and MACs of step `step`, plus its kernel name. */
int fd_plan_step_info(fd_plan* plan, int step, int* stage, double* alg_bytes, double* macs,
                      char* kernel_name, int name_cap);
/* The same step's MACs split by the pipe that executes them: depthwise taps (channel-wise, SIMT FMA pipe; north_star keeps
 * them off the tensor cores) and the dense contractions (stem im2col, pointwise 1x1, head; tensor pipe on path 1).  bench.py
 * turns them into the per-stage FMA-pipe and tensor-pipe floors it prints next to the HBM floor. */
int fd_plan_step_macs(fd_plan* plan, int step, double* dw_macs, double* dense_macs);
/* Time every step's kernel alone with CUDA events on `stream` (`warmup` + `iters` launches each;
 * a 256 MB buffer is written between launches when flush_l2 != 0 so inputs come from HBM).
 * ms_out[n_steps] = mean launch duration.  Synchronous. */
int fd_plan_time_steps(fd_plan* plan, const void* x_dev, void* y_dev, void* stream,
                       int warmup, int iters, int flush_l2, float* ms_out);

/* Debug: run stage `stage`'s fused block kernel once with its in-kernel timeline enabled and return
 * the SM-clock stamps of CTA 0: rows = {TMA issue, dw start, dw math done, A tile published, MMA
 * operands ready, MMA issued, epilogue start, epilogue done}, one column per K-block / item.
 * The previous fd_forward's activations are reused as inputs. Synchronous. */
int fd_plan_trace_stage(fd_plan* plan, int stage, void* y_dev, void* stream,
                        unsigned long long* out_host, int cap, int* rows, int* cols);

/* Debug (host only, needs no GPU): the shared-memory / pipeline plan the fused block kernel would use for one
 * block.  out[0..15] = {ok, splits, n_cta, items, kblocks, s_in, s_a, s_b, bn, nb, b_resident, epi_groups, n_stg,
 * smem_bytes, tmem_cols, in_stage_stride}; with cap >= 18 also out[16..17] = {nacc (TMEM accumulators), epi_colsplit},
 * with cap >= 19 out[18] = epi_wide, with cap >= 20 out[19] = cs (cluster size: CTAs sharing one tile's depthwise half), with cap >= 21 out[20] = dw_teams.
 * cap must be at least 16. */
int fd_debug_block_plan(int ksize, int stride, int h_out, int w_out, int n, int c_in, int c_out, int head,
                        int* out, int cap);

/* Per-image depth metrics on device (reference metrics.py:31-55 applied per image, as
 * main.py:40-41,80-82 does at batch size 1).  pred: [n, hw] of `dtype`; target: [n, hw] fp32.
 * Adds, for each image, its 10 metric values into sums_dev[0..9] (order: irmse, imae, mse,
 * rmse, mae, absrel, lg10, delta1, delta2, delta3) and 1.0 into sums_dev[10] (count), all
 * double (reference AverageMeter, metrics.py:71-95).  The cross-GPU reduction of that
 * 11-vector is the caller's single all-reduce. */
int fd_metrics_accumulate(const void* pred_dev, const float* target_dev, int dtype, int n, int hw,
                          double* sums_dev, int device, void* stream);

/* NYU-Depth-v2 validation pre-processing as ONE gather launch (reference dataloaders/nyu.py:48-59: Resize(250/480) ->
 * CenterCrop(228x304) -> Resize(out), all nearest-neighbour, rgb / 255; dataloaders/dataloader.py:90-111: HWC -> CHW).
 * rgb_dev: [n, h_in, w_in, 3] uint8; depth_dev: [n, h_in, w_in] float or NULL; rows_dev[out_h] / cols_dev[out_w]: the
 * composed source-index tables (fastdepth_b200/preprocess.py builds them from PIL's own nearest resize);
 * x_dev: [n, 3, out_h, out_w] of `dtype`; target_dev: [n, 1, out_h, out_w] float or NULL (with depth_dev). */
int fd_nyu_val_gather(const uint8_t* rgb_dev, const float* depth_dev, const int* rows_dev, const int* cols_dev,
                      int n, int h_in, int w_in, int out_h, int out_w, int dtype,
                      void* x_dev, float* target_dev, int device, void* stream);

void fd_plan_destroy(fd_plan* plan);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* fd_last_error(void);
int fd_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FASTDEPTH_B200_H */
