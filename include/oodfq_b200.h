/*
 * oodfq_b200.h -- C ABI of the B200 (sm_100a) fake-quantisation hot path.
 *
 * This is the drop-in boundary.  The reference (weesunghyun/OOD-DFQ) has no FFI
 * of its own: its operator API is the Python package `quantization_utils`
 * (imported at main_direct.py:21 and trainer_direct.py:19).  Each entry point
 * below replaces the chain of ATen launches behind one reference call site; the
 * citation after "replaces:" is file:line in the reference tree.  The Python
 * mirror of the reference interface (ood_dfq_b200/quantization_utils/) binds
 * these symbols with ctypes -- see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless the name ends in
 *     `_host`; tensors are dense; `stream` is a cudaStream_t passed as void*.
 *   - all work is enqueued on `stream`; no entry point synchronises the host.
 *   - return value: 0 on success, a negative OODFQ_E* code otherwise; the
 *     message of the last failure on the calling thread is oodfq_last_error().
 *     Nothing throws across this boundary.
 *   - memory is owned by the caller (torch allocators in the Python mirror).
 *   - k is the bit-width: codes are integers in [-2^(k-1), 2^(k-1)-1].
 *   - arithmetic order is the reference's, one fp32 rounding per step
 *     (SURVEY.md section 8(a')): integer codes are bit-exact with the reference.
 */
#ifndef OODFQ_B200_H
#define OODFQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OODFQ_ABI_VERSION 4

#define OODFQ_OK 0
#define OODFQ_EINVAL (-1)  /* bad argument (null pointer, k out of range, misaligned ...) */
#define OODFQ_ECUDA (-2)   /* a CUDA runtime call or kernel launch failed                   */

/* element-wise modes of oodfq_fq_forward */
#define OODFQ_MODE_FAKEQUANT 0 /* quantise -> clamp -> dequantise                    */
#define OODFQ_MODE_QUANTIZE 1  /* round(scale*x - zp), no clamp (linear_quantize)    */
#define OODFQ_MODE_DEQUANTIZE 2/* (q + zp) / scale            (linear_dequantize)   */

/* flags */
#define OODFQ_SYMMETRIC 1      /* the *_DSG family: zero-point ignored               */
#define OODFQ_PARAMS_GIVEN 2   /* p0/p1 are (scale, zero_point) instead of (min,max) */
#define OODFQ_RELU_FIRST 4     /* x <- max(x, 0) before quantising: the nn.Sequential(ReLU, QuantAct) of
                                  main_direct.py:464-465 in one pass (scalar range, FAKEQUANT mode) */
#define OODFQ_NO_ONCHIP 8      /* act_calib_forward: always take the two-kernel path (tests, comparisons) */
#define OODFQ_ONCHIP_TMA 16    /* accepted and ignored since ABI 4: the on-chip kernel is always fed by TMA bulk copies */

/* flags of the per-channel (BatchNorm) entry points */
#define OODFQ_BN_RELU 1
#define OODFQ_BN_QUANT 2
#define OODFQ_BN_NHWC 4        /* x, y, grads are channels_last: [N*H*W rows][C]; needs C % 4 == 0 */
#define OODFQ_BN_POOL_REGISTER 16 /* bn_pool_forward: always take the register kernel (tests, comparisons) */
#define OODFQ_AUG_SRC_NHWC 8   /* crop_resize_flip: the stored image set is channels_last too */

typedef void* oodfq_stream_t;  /* cudaStream_t */

/* ---- library bookkeeping ------------------------------------------------- */
int oodfq_abi_version(void);
const char* oodfq_last_error(void);
/* number of kernels this library has launched since load / since the last reset
 * (bench.py reports it as "gpu_launches"). */
unsigned long long oodfq_launch_count(void);
void oodfq_reset_launch_count(void);
/* bytes of zero-initialised device scratch the *_ws entry points need */
size_t oodfq_workspace_bytes(void);

/* ---- a1: scale / zero-point ---------------------------------------------
 * replaces: asymmetric_linear_quantization_params, quant_utils.py:107-128
 *           symmetric_linear_quantization_params_DSG, quant_utils.py:238-259
 * lo, hi, scale, zero_point: [n]. */
int oodfq_quant_params(const float* lo, const float* hi, float* scale, float* zero_point,
                       long long n, int k, oodfq_stream_t stream);

/* ---- a2-a5: element-wise quantise / clamp / dequantise ------------------
 * replaces: AsymmetricQuantFunction.forward, quant_utils.py:138-157
 *           SymmetricQuantFunction_DSG.forward, quant_utils.py:268-286
 *           linear_quantize / linear_dequantize (+_DSG), quant_utils.py:61-104, 192-235
 * x, y: [numel]; p0, p1: [rows] (rows == 1: one range for the whole tensor,
 * else one range per leading-dimension row of numel/rows elements).
 * codes (nullable, k <= 8): int8 integer codes, FAKEQUANT mode only.
 * y may alias x (in-place variants of the reference helpers). */
int oodfq_fq_forward(const float* x, float* y, int8_t* codes, long long numel,
                     const float* p0, const float* p1, long long rows,
                     int k, int mode, int flags, oodfq_stream_t stream);

/* ---- a6: calibrating QuantAct forward -----------------------------------
 * replaces: QuantAct.forward with running_stat=True, quant_modules.py:75-96
 *           (QuantAct_DSG.forward :360-386 with OODFQ_SYMMETRIC)
 * One call = data min/max of x, the bias-corrected running-range update of
 * (x_min, x_max, beta_t) IN PLACE on the device, then the fake-quantised y
 * with the UPDATED range.  y == NULL: range update only (full_precision_flag).
 * Tensors of up to 96 MB (asymmetric, k <= 8, numel % 4 == 0) run as ONE cooperative kernel that stages x into
 * shared memory with TMA bulk copies and keeps it there across the grid-wide range reduction, so x crosses HBM
 * once; larger ones (or OODFQ_NO_ONCHIP) as a reducing kernel followed by the streaming fake-quant.  Results are
 * bit-identical either way.
 * workspace: oodfq_workspace_bytes() of device memory, zeroed once by the
 * caller and then owned by this library between calls on one stream. */
int oodfq_act_calib_forward(const float* x, float* y, int8_t* codes, long long numel,
                            float* x_min, float* x_max, const float* beta, float* beta_t,
                            int k, int flags, void* workspace, oodfq_stream_t stream);

/* ---- a6 + a11 in one pass (BASELINE.json north_star (b)) --------------------
 * replaces: QuantAct.forward with running_stat=True, quant_modules.py:80-94, AND the two per-channel
 *           reductions of the BN-statistics hook on the same tensor, trainer_direct.py:388-393
 *           (input.mean([0,2,3]), input.var([0,2,3], unbiased=False)) -- the reference never combines them.
 * x, y: [N, C, HW] NCHW, or channels_last [N*HW rows][C] with OODFQ_BN_NHWC (C % 4 == 0).  One call = data
 * min/max, the in-place running-range update, y = fakequant(x) with the UPDATED range (bit-identical to
 * oodfq_act_calib_forward) and sums[2*C] fp64 = (sum_c x, then sum_c x^2) over N and HW.
 * Up to 96 MB (k <= 8; NCHW: C <= 1024; channels_last: C a power of two in [4, 1024]) this is ONE cooperative
 * kernel / ONE read of x from HBM; otherwise (or with OODFQ_NO_ONCHIP) the range reduction followed by one pass
 * that quantises and accumulates (12 B/elem, the floor once 4*numel exceeds L2).  flags: OODFQ_BN_NHWC,
 * OODFQ_NO_ONCHIP. */
int oodfq_act_calib_stats_forward(const float* x, float* y, int N, int C, long long HW,
                                  float* x_min, float* x_max, const float* beta, float* beta_t,
                                  int k, int flags, double* sums, void* workspace, oodfq_stream_t stream);

/* data min / max only (NaN-propagating like torch.min/max): out[0]=min, out[1]=max
 * replaces: x.data.min(), x.data.max(), quant_modules.py:81-82 */
int oodfq_minmax(const float* x, long long numel, float* out2, void* workspace,
                 oodfq_stream_t stream);

/* ---- a7/a8: weight fake-quant, many tensors in one launch ----------------
 * replaces: Quant_Conv2d.forward :266-279, Quant_Linear.forward :215-230 and the
 *           *_DSG twins :420-431, :465-479 (per-output-row min/max + a5), for
 *           every layer of a model at once. */
typedef struct {
    const float* w;     /* [rows, row_len] dense                              */
    float* wq;          /* same shape: fake-quantised weight                  */
    float* lo;          /* [rows] row minimum actually used (nullable)        */
    float* hi;          /* [rows] row maximum actually used (nullable)        */
    int8_t* codes;      /* [rows, row_len] integer codes (nullable, k <= 8)   */
    long long rows;
    long long row_len;
    int k;
    int flags;          /* OODFQ_SYMMETRIC or 0                               */
} oodfq_weight_desc;

int oodfq_weight_fq_multi(const oodfq_weight_desc* descs_host, int n_tensors,
                          oodfq_stream_t stream);

/* ---- a11: per-channel statistics of a BN input ---------------------------
 * replaces: hook_fn_forward, trainer_direct.py:388-393 / distill_data.py:69-73
 * x: [N, C, HW] (NCHW).  sums: [2*C] fp64 receives S1_c = sum(x - shift_c) and
 * S2_c = sum((x - shift_c)^2) over N and HW (accumulated around a local pivot and
 * re-based in fp64, so no cancellation).  shift: [C] or NULL (zero).
 * Optional fused fake-quant of the same read (north_star (b)): when y != NULL,
 * y = fakequant(x) with the scalar range (fq_lo, fq_hi), k = fq_k <= 8.
 * flags: 0, or OODFQ_BN_NHWC when x (and y) are channels_last [N*H*W rows][C] (C % 4 == 0). */
int oodfq_bn_stats_forward(const float* x, int N, int C, long long HW, const float* shift,
                           double* sums, float* y, const float* fq_lo, const float* fq_hi,
                           int fq_k, int flags, void* workspace, oodfq_stream_t stream);

/* mean_c = shift_c + S1_c/count, var_c = S2_c/count - (S1_c/count)^2 (biased).
 * `sums` may have been all-reduced over ranks; count is the global N*HW. */
int oodfq_bn_stats_finalize(const double* sums, const float* shift, int C, double count,
                            float* mean, float* var, oodfq_stream_t stream);

/* ---- a12: BN-statistics loss over L layers, packed ------------------------
 * replaces: trainer_direct.py:473-486 and distill_data.py:252-265
 * Layer l owns channels [ch_off_host[l], ch_off_host[l+1]) of every packed [Ctot]
 * array.  sums is [2*Ctot] fp64 with each layer's output of oodfq_bn_stats_forward kept
 * together: floats [2*off_l, 2*off_l + C_l) are S1, the next C_l are S2.
 * counts_host[l] = global N*H*W of layer l.  Outputs: loss3[0] = (sum_l MSE_mean + MSE_var)/L,
 * loss3[1] = sum_l MSE_mean / L, loss3[2] = sum_l MSE_var / L; mean, var [Ctot];
 * gmean, gvar [Ctot] = d loss / d mean_c, d loss / d var_c. */
int oodfq_bns_loss(const double* sums, const float* shift, const float* run_mean,
                   const float* run_var, const int* ch_off_host, const double* counts_host,
                   int L, float* loss3, float* mean, float* var, float* gmean, float* gvar,
                   oodfq_stream_t stream);

/* ---- a12 backward ----------------------------------------------------------
 * grad_x[n,c,i] = (grad_in ? grad_in[n,c,i] : 0)
 *               + g * ( gmean_c/count + gvar_c * 2*(x[n,c,i]-mean_c)/count )
 * g = *gscale (device scalar, NULL -> 1).  grad_x may alias grad_in. */
int oodfq_bn_stats_backward(const float* x, const float* grad_in, float* grad_x,
                            int N, int C, long long HW, const float* mean,
                            const float* gmean, const float* gvar, double count,
                            const float* gscale, int flags, oodfq_stream_t stream);

/* ---- eval-mode BatchNorm fused with the ReLU + QuantAct behind it ----------
 * replaces: nn.BatchNorm2d in eval() (student and teacher always are: trainer_direct.py:411-412)
 *           followed by the nn.Sequential(ReLU, QuantAct) of main_direct.py:464-465
 * forward : y = [fakequant]( [relu]( a_c*x + b_c ) ),  a_c = w_c/sqrt(rv_c+eps), b_c = bias_c - rm_c*a_c
 * backward: g' = grad_y * [a_c*x+b_c > 0];  grad_x = g'*a_c;
 *           dwdb[c] = sum g'*(x-rm_c)/sqrt(rv_c+eps), dwdb[C+c] = sum g'   (accumulated in fp64, stored as fp32; NULL: skip)
 * flags: OODFQ_BN_RELU | OODFQ_BN_QUANT (QUANT: scalar range fq_lo/fq_hi, k = fq_k <= 8; its
 * backward is the identity STE).  z_debug (nullable): the fp32 value handed to the quantiser.
 * weight / bias may be NULL (1 / 0).
 * relu_mask (nullable; channels_last + OODFQ_BN_RELU only): forward output / backward input, one byte per 128-bit
 * column ([N*H*W*C/4] bytes in storage order), bit j = a_c*x+b_c > 0 for channel j of the column.  A backward that is
 * given the mask and no dwdb never reads x (x may be NULL): 8.25 instead of 12 B/elem. */
int oodfq_bn_eval_forward(const float* x, float* y, float* z_debug, int N, int C, long long HW,
                          const float* weight, const float* bias, const float* running_mean,
                          const float* running_var, float eps, int flags, const float* fq_lo,
                          const float* fq_hi, int fq_k, uint8_t* relu_mask, oodfq_stream_t stream);
int oodfq_bn_eval_backward(const float* x, const float* grad_y, float* grad_x, int N, int C,
                           long long HW, const float* weight, const float* bias,
                           const float* running_mean, const float* running_var, float eps,
                           int flags, float* dwdb, void* workspace, const uint8_t* relu_mask,
                           oodfq_stream_t stream);

/* ---- the reduction inside the feature-alignment loss ----------------------
 * replaces: x.pow(2).mean([2,3]) of Trainer.channel_attention, trainer_direct.py:382-383 (hooks :432-440,
 *           loss :325-330) and its autograd tape
 * forward : e[n,c] = mean_{hw} x[n,c,hw]^2          e: [N, C]
 * backward: grad_x[n,c,hw] = grad_e[n,c] * 2/HW * x[n,c,hw]
 * flags: 0 or OODFQ_BN_NHWC.  scratch (NHWC forward only): oodfq_channel_energy_scratch_floats(N, C) floats. */
size_t oodfq_channel_energy_scratch_floats(int N, int C);
int oodfq_channel_energy_forward(const float* x, float* e, int N, int C, long long HW, int flags,
                                 float* scratch, oodfq_stream_t stream);
int oodfq_channel_energy_backward(const float* x, const float* grad_e, float* grad_x, int N, int C,
                                  long long HW, int flags, oodfq_stream_t stream);

/* ---- deferred folds of the BatchNorm parameter-gradient reductions ----------------------------------------
 * replaces: nothing in the reference -- launch bookkeeping behind `loss_S.backward()` (trainer_direct.py:350-356).
 * oodfq_bn_eval_backward / oodfq_res_tail_backward / oodfq_bn_pool_backward with dwdb != NULL end in a small launch
 * that folds the per-CTA (dW, dB) partials.  Between _begin and _end those calls write their partials into regions
 * of `arena` (device memory, 256-byte aligned, caller-owned, must outlive _end) and the folds of ALL of them run
 * as ONE launch at _flush / _end on `stream` (the stream the producing kernels ran on).  Until then every dwdb
 * handed to a deferred call holds garbage and must neither be read nor freed.  A call that does not fit the arena
 * folds immediately as usual.  Results are bit-identical to the immediate fold.  Process-wide state (one process
 * per GPU), guarded by a mutex: the producing calls may come from autograd's backward thread. */
int oodfq_defer_folds_begin(void* arena, size_t arena_bytes);
int oodfq_defer_folds_flush(oodfq_stream_t stream);
int oodfq_defer_folds_end(oodfq_stream_t stream);
int oodfq_defer_folds_pending(void);

/* ---- global average pool between the last QuantAct and Quant_Linear ------------------------------------
 * replaces: features.final_pool = AvgPool2d(7 | 8) over a plane of exactly that size (pytorchcv ResNet behind
 *           ptcv_get_model, main_direct.py:380-397; reference models.py avg_pool2d(out, 4)) and its backward
 * channels_last only (flags must contain OODFQ_BN_NHWC; C % 4 == 0): x, grad_x [N, H*W, C]; y, grad_y [N, C].
 * forward : y[n,c] = (fp32 running sum of x[n,hw,c] in hw order) / HW     -- bit-identical to ATen's avg_pool2d
 * backward: grad_x[n,hw,c] = 0 + grad_y[n,c] / HW                          -- bit-identical to its backward */
int oodfq_global_avgpool_forward(const float* x, float* y, int N, int C, long long HW, int flags,
                                 oodfq_stream_t stream);
int oodfq_global_avgpool_backward(const float* grad_y, float* grad_x, int N, int C, long long HW, int flags,
                                  oodfq_stream_t stream);

/* ---- stem fusion: eval BatchNorm -> ReLU -> [QuantAct] -> MaxPool2d(3, stride 2, padding 1) -----------
 * replaces: the first QuantAct site of the ImageNet ResNets together with the max-pool that consumes it
 *           (pytorchcv ResInitBlock behind ptcv_get_model, main_direct.py:380-397; quantize_model :464-465)
 * channels_last only: x [N,H,W,C], out / idx / xhat [N,Ho,Wo,C], Ho = (H-1)/2+1.  flags must contain
 * OODFQ_BN_NHWC | OODFQ_BN_RELU, optionally OODFQ_BN_QUANT.  idx: one byte per output (window-local argmax
 * 0..8, bit 7 = ReLU active).  xhat (nullable): normalised input at the argmax, needed only for dwdb.
 * grad_out2 (nullable): a second gradient w.r.t. out, added to grad_out in registers. */
int oodfq_bn_pool_forward(const float* x, float* out, uint8_t* idx, float* xhat, int N, int C, int H, int W,
                          const float* weight, const float* bias, const float* running_mean,
                          const float* running_var, float eps, int flags, const float* fq_lo,
                          const float* fq_hi, int fq_k, oodfq_stream_t stream);
int oodfq_bn_pool_backward(const float* grad_out, const float* grad_out2, const uint8_t* idx, const float* xhat, float* grad_x,
                           int N, int C, int H, int W, const float* weight, const float* bias,
                           const float* running_mean, const float* running_var, float eps,
                           float* dwdb, void* workspace, oodfq_stream_t stream);

/* ---- QuantAct_MSE: clip-ratio range search --------------------------------------------------------------
 * replaces: the 80-iteration loop of QuantAct_MSE.forward (quant_modules.py:160-178) with find_MSESmallest
 *           (quant_utils.py:36-47) and lp_loss (quant_utils.py:26-33) inside it
 * data_minmax: [2] device floats (oodfq_minmax of x).  Candidate i in [0, ncand) uses the range
 * data_minmax * fp32(1 - i*step); score_i = mean |x - fakequant_i(x)|^p; the first strict minimum below 1e10 is
 * kept and folded into (x_min, x_max) by the plain EMA x*beta + kept*(1-beta); beta_t *= beta.
 * cur_min / cur_max (nullable) receive the data range; scores [ncand] and chosen (device int) are optional
 * outputs for tests.  scratch: oodfq_act_mse_scratch_doubles(ncand) doubles.  ncand <= 96. */
size_t oodfq_act_mse_scratch_doubles(int ncand);
int oodfq_act_mse_search(const float* x, long long numel, const float* data_minmax, int k, int ncand, double step,
                         float p, float* x_min, float* x_max, const float* beta, float* beta_t, float* cur_min,
                         float* cur_max, double* scratch, float* scores, int* chosen, oodfq_stream_t stream);

/* ---- residual-unit tail: BN1(x1) + identity -> ReLU -> [QuantAct], with the feature-alignment energy -------
 * replaces: the last BatchNorm of a residual body, the residual add of the unit's forward (pytorchcv ResUnit /
 *           reference models.py:40-47 `out += self.shortcut(x); out = self.relu2(out)`), the
 *           Sequential(ReLU, QuantAct) quantize_model puts behind it (main_direct.py:464-465) and the
 *           channel-attention reduction the trainer hooks onto the body output (trainer_direct.py:432-440,
 *           :382-383)
 * channels_last only (flags must contain OODFQ_BN_NHWC, optionally OODFQ_BN_QUANT): x1, r, y, grads [N,H,W,C].
 * rv2 == NULL: the identity is r itself; otherwise identity = BN2(r) with (w2, b2, rm2, rv2, eps2).
 * energy (nullable) [N, C] = mean_hw BN1(x1)^2; scratch: oodfq_res_tail_scratch_floats(N, C) floats.
 * relu_mask (nullable, forward output / backward input): one byte per 128-bit column, [N*H*W*C/4] bytes in storage
 *           order; bit j = the ReLU behind the add lets the gradient of channel j of that column through.  With it
 *           the backward never re-derives the mask, so `r` may be NULL unless the identity BatchNorm's parameter
 *           gradients are wanted (rv2 && dwdb), and `x1` may be NULL when neither grad_energy nor dwdb is given:
 *           16.25 instead of 20 B/elem for the common case.
 * backward: grad_y2 (nullable) is a second gradient w.r.t. y, added to grad_y in registers (the output fed two
 *           consumers); grad_energy nullable; dwdb nullable, else [2*Ct] floats (fp64 accumulation, rounded once) with Ct = C (or 2C with BN2):
 *           dW of BN1 (then BN2), followed by dB of BN1 (then BN2). */
size_t oodfq_res_tail_scratch_floats(int N, int C);
int oodfq_res_tail_forward(const float* x1, const float* r, float* y, float* energy, float* scratch,
                           uint8_t* relu_mask, int N, int C, long long HW, const float* w1, const float* b1,
                           const float* rm1, const float* rv1, float eps1, const float* w2, const float* b2,
                           const float* rm2, const float* rv2, float eps2, int flags, const float* fq_lo,
                           const float* fq_hi, int fq_k, oodfq_stream_t stream);
int oodfq_res_tail_backward(const float* grad_y, const float* grad_y2, const float* grad_energy, const float* x1, const float* r,
                            const uint8_t* relu_mask, float* grad_x1, float* grad_r, int N, int C, long long HW,
                            const float* w1, const float* b1, const float* rm1, const float* rv1, float eps1,
                            const float* w2, const float* b2, const float* rm2, const float* rv2, float eps2,
                            int flags, float* dwdb, void* workspace, oodfq_stream_t stream);

/* ---- eval-mode BatchNorm (+ReLU +QuantAct) AND the statistics of its input, one read ---------------------
 * replaces: oodfq_bn_stats_forward(x, shift) followed by oodfq_bn_eval_forward(x) for a fused BatchNorm whose input is
 *           tapped by the BN-statistics loss (distill_data.py:69-78 hooks every BatchNorm): y as oodfq_bn_eval_forward
 *           (bit-identical), sums[2*C] fp64 as oodfq_bn_stats_forward, 8 instead of 4 + 8 B/elem.
 * channels_last only (flags: OODFQ_BN_NHWC, optionally OODFQ_BN_RELU, OODFQ_BN_QUANT with a scalar range). */
int oodfq_bn_eval_stats_forward(const float* x, float* y, int N, int C, long long HW, const float* weight,
                                const float* bias, const float* running_mean, const float* running_var, float eps,
                                int flags, const float* fq_lo, const float* fq_hi, int fq_k, const float* shift,
                                double* sums, void* workspace, oodfq_stream_t stream);

/* ---- backward through an eval-mode BatchNorm whose input is tapped by the BN-statistics loss --------------
 * replaces: the chain oodfq_bn_eval_backward -> oodfq_bn_stats_backward(grad_in = its result) for a BatchNorm that is
 *           both fused and hooked by the statistics loss (distill_data.py:69-78, :252-265 put the
 *           hook on EVERY BatchNorm): grad_x = [a_c*x+b_c > 0] * grad_y * a_c + g * (gmean_c/count +
 *           gvar_c * 2*(x - mean_c)/count) in ONE pass over x and grad_y (12 B/elem instead of 8-12 + 12), same
 *           roundings in the same order as the chain.  No parameter gradients (the hooked network is frozen).
 * channels_last only (flags: OODFQ_BN_NHWC, optionally OODFQ_BN_RELU); mean, gmean, gvar, gscale as in
 * oodfq_bn_stats_backward. */
int oodfq_bn_eval_tap_backward(const float* x, const float* grad_y, float* grad_x, int N, int C, long long HW,
                               const float* weight, const float* bias, const float* running_mean,
                               const float* running_var, float eps, int flags, const float* mean,
                               const float* gmean, const float* gvar, double count, const float* gscale,
                               oodfq_stream_t stream);

/* ---- feature-alignment loss over all residual units, one kernel each way -------------------------------
 * replaces: Trainer.loss_fa (trainer_direct.py:325-330) over the maps of Trainer.channel_attention
 *           (trainer_direct.py:382-383): fa = lam * sum_l mean((F.normalize(Es_l) - F.normalize(Et_l))^2), and the
 *           autograd tape behind it (~20 element-wise / reduction launches per unit and pass in eager PyTorch).
 * e_student / e_teacher / grad_*: HOST arrays of L device pointers to [N, channels[l]] fp32 row-major tensors
 * (the per-(image, channel) energies of oodfq_res_tail_forward / oodfq_channel_energy_forward); channels: host ints.
 * forward : loss[0] (device) receives fa; row_scratch: L*N device doubles; workspace as everywhere.
 * backward: grad tables receive grad_loss[0] * d fa / d E (grad_loss: device scalar, NULL = 1); either table may be
 *           NULL, as may single entries (that side / unit gets no gradient).  L <= oodfq_fa_loss_max_layers(). */
int oodfq_fa_loss_max_layers(void);
int oodfq_fa_loss_forward(const float* const* e_student, const float* const* e_teacher, const int* channels, int L,
                          int N, float lam, float* loss, double* row_scratch, void* workspace, oodfq_stream_t stream);
int oodfq_fa_loss_backward(const float* const* e_student, const float* const* e_teacher, const int* channels, int L,
                           int N, float lam, const float* grad_loss, float* const* grad_student,
                           float* const* grad_teacher, oodfq_stream_t stream);

/* ---- space-to-depth re-layout in front of the ImageNet stem convolution ---------------------------------
 * replaces: nothing in the reference -- it is the data format on the input side of Quant_Conv2d's F.conv2d call
 *           (quant_modules.py:279-281) for the 3-channel 7x7 stride-2 stem (main_direct.py:380-397), where a
 *           stride-2 KxK convolution is run as the equal stride-1 convolution over the 2x2 space-to-depth image.
 * x [N,H,W,C] channels_last -> xs [N,(H+2*pad)/2,(W+2*pad)/2,CP], xs[n,i,j,(s,t,c)] = x[n,2i+s-pad,2j+t-pad,c]
 * (zero outside the image); CP = cpad floats per xs pixel: 0 or 4C for the plain form, or a larger multiple of 4 whose
 * extra channels are written as zeros (16 for C = 3: cuDNN converts a 12-channel tensor before every use).
 * backward is the inverse gather of the gradient (padding channels ignored).
 * C = 1, 3, 4 with 16-byte-multiple rows run as a TMA-staged ring (bulk loads and stores). */
int oodfq_s2d_stem_forward(const float* x, float* xs, int N, int H, int W, int C, int pad, int cpad,
                           oodfq_stream_t stream);
int oodfq_s2d_stem_backward(const float* grad_xs, float* grad_x, int N, int H, int W, int C, int pad, int cpad,
                            oodfq_stream_t stream);

/* ---- batch assembly: gather -> RandomResizedCrop -> grey->RGB repeat -> RandomHorizontalFlip ----------------
 * replaces: direct_dataset.__getitem__ (main_direct.py:200-204) and its torchvision pipeline
 *           RandomResizedCrop(size, scale=(0.5, 1.0)) -> Lambda(repeat to 3 channels) -> RandomHorizontalFlip
 *           (main_direct.py:158-169), run per sample on DataLoader workers, plus the collation and the per-step
 *           host -> device copy of the batch (main_direct.py:525-533)
 * images [M, C_in, H, W] NCHW fp32, the concatenated shards (main_direct.py:173-195) resident on the device;
 * index [N] int64 sample -> image; boxes [N][4] int32 (top, left, height, width) as RandomResizedCrop.get_params
 * returns them; flips [N] uint8 (non-zero = mirrored after the resize).  The draws themselves stay on the host.
 * out [N, C_out, out_h, out_w] (flags & OODFQ_BN_NHWC: channels_last).  Channels: 1->1, 1->3 (repeat), 3->3.
 * Resize = bilinear, align_corners=False, no antialiasing: identical to torchvision's antialiased filter
 * whenever the crop is not larger than the output (always, in direct_dataset: size == image size).  Entries
 * outside the image set are folded into it (never an out-of-bounds read).
 * flags & OODFQ_AUG_SRC_NHWC: images is [M, H, W, C_in] (the optimised batch of the distillation loop, which the
 * reference augments in place of a stored set: data_generate/distill_data.py:197-227).
 * backward: grad_images (layout of images) += scatter of grad_out (layout of out) with the forward's tap weights --
 * the caller zero-fills it (or passes an accumulating gradient); atomic adds, like ATen's upsample backward that
 * autograd runs for the reference's RRC(gaussian_data[j]). */
int oodfq_crop_resize_flip(const float* images, long long n_images, int C_in, int H, int W, const long long* index,
                           const int* boxes, const unsigned char* flips, float* out, int N, int C_out, int out_h,
                           int out_w, int flags, oodfq_stream_t stream);
int oodfq_crop_resize_flip_backward(const float* grad_out, float* grad_images, long long n_images, int C_in, int H, int W,
                                    const long long* index, const int* boxes, const unsigned char* flips, int N,
                                    int C_out, int out_h, int out_w, int flags, oodfq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OODFQ_B200_H */
