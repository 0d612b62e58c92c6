/*
 * clair3_b200.h — C-ABI of libclair3b200.so: the B200 (sm_100a) inference forward pass of Clair3's two
 * networks, the drop-in boundary for the one hot path this project replaces.
 *
 * The reference has no plugin/operator registry for this path; its seam is the torch module protocol used by its
 * callers (paths relative to HKU-BAL/Clair3):
 *
 *   m = Clair3_P|Clair3_F(add_indel_length, predict=True, input_channels)   clair3/CallVariantsFromCffi.py:230-243
 *   m.to(device); m.eval(); m.load_state_dict(state_dict)                    clair3/CallVariantsFromCffi.py:19-28,246-248
 *   Y = m(torch.from_numpy(X).to(device)); Y.detach().cpu().numpy()          clair3/CallVariantsFromCffi.py:48-52
 *                                                                            (same protocol: clair3/CallVariants.py:54-87,1466-1480)
 *
 * Each entry point below names the reference interface it replaces.  The Python shim that mirrors the module
 * protocol on top of this ABI is clair3_b200/model.py (bound with cffi, the way the reference binds libclair3 in
 * build.py:44-79).  Convention change vs libclair3 (which exit(1)s on failure, src/medaka_common.c:25-47): every call
 * returns an int status, 0 = ok, message via c3b_last_error(); the library never exits the process and never
 * falls back to a CPU implementation.
 *
 * Threading: a c3b_model is driven by one host thread at a time; forwards on different CUDA streams may be in
 * flight concurrently (each stream gets its own activation workspace).  Distinct models/devices are independent.
 */
#ifndef CLAIR3_B200_H
#define CLAIR3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct c3b_model c3b_model;

/* network kind                      reference class                                   */
#define C3B_PILEUP          0     /* Clair3_P  clair3/model.py:58-161                  */
#define C3B_FULL_ALIGNMENT  1     /* Clair3_F  clair3/model.py:282-416                 */

/* element type of the candidate-site batch handed to c3b_forward (reference: any dtype, x.float() model.py:131,378) */
#define C3B_DT_I8   0             /* full-alignment wire dtype (shared/param_f.py:92) and GPU-mode pileup .npy (CreateTensorPileupFromCffi.py:447) */
#define C3B_DT_I32  1             /* pileup wire dtype (CreateTensorPileupFromCffi.py:397) */
#define C3B_DT_F32  2
#define C3B_DT_I64  3             /* only for c3b_set_param (BatchNorm num_batches_tracked) */

/* arithmetic used by the kernels (c3b_set_option "precision") */
#define C3B_PREC_BF16_TC 0        /* production: bf16 operands on tcgen05 tensor cores, fp32 accumulate/state/softmax */
#define C3B_PREC_FP32    1        /* debug: the same layer graph on fp32 CUDA cores (separates layout bugs from precision) */

/* Replaces Clair3_P.__init__ / Clair3_F.__init__ (clair3/model.py:61-128, 285-368) + m.to(device) (CallVariantsFromCffi.py:246).
 * channels: 18 (pileup) | 8 | 9 with dwell (full-alignment).  Fails if the device is not compute capability 10.x. */
int c3b_create(c3b_model **out, int kind, int channels, int add_indel_length, int device_ordinal);

/* Replaces one entry of m.load_state_dict(state_dict) (clair3/CallVariantsFromCffi.py:19-28).  key is the reference
 * state_dict key (e.g. "LSTM1.weight_ih_l0_reverse", "res_block2.0.bn1.running_var"); data is the tensor as stored
 * in the .pt (dtype C3B_DT_F32, or C3B_DT_I64 for *.num_batches_tracked which is accepted and ignored).  Unknown keys
 * and shape mismatches are errors (strict, like torch). */
int c3b_set_param(c3b_model *m, const char *key, const void *host_data, int dtype, const int64_t *shape, int ndim);

/* Ends load_state_dict: checks every expected key is present (strict), folds BatchNorm into the convolutions
 * (eps 1e-3, clair3/model.py:192), sums the LSTM bias pairs, folds 1/NORMALIZE_NUM (shared/param_f.py:36) into conv1,
 * packs bf16 UMMA operand images and uploads them once. */
int c3b_finalize(c3b_model *m);

/* name: "precision" (C3B_PREC_*), "chunk_sites" (sites per internal pass), "lstm_tile" (batch columns per LSTM CTA: 16|32|64, 0 = auto), "lstm_wg" (epilogue warpgroups per LSTM sub-tile: 1|2), "lstm_trace" (debug clock stamps),
 * "lstm_mufu16" (1: gate activations with packed tanh.approx.f16x2, two sites per MUFU op; 0 default: fp32 tanh.approx, measured faster),
 * "host_async" (1: forwards with HOST buffers do not synchronise; the buffers must be pinned and the caller synchronises the
 * stream before reading y - lets a caller pipeline H2D / forward / D2H of consecutive batches over several streams),
 * "profile" (1: bracket every kernel launch with CUDA events on its stream and accumulate per-kernel time; setting it resets the totals). */
int c3b_set_option(c3b_model *m, const char *name, int value);

/* Replaces Y = m(X) (clair3/model.py:130-161 / 377-416) including the H2D/D2H of _torch_predict
 * (clair3/CallVariantsFromCffi.py:48-52) when x_on_device / y_on_device are 0.
 *   x: [batch,33,channels] (pileup) or [batch,depth,33,channels] (full-alignment, NHWC), C-contiguous, dtype x_dtype.
 *   y: [batch, c3b_out_dim()] float32 softmax probabilities, heads concatenated gt21|genotype|indel1|indel2.
 * With both buffers on the device the call is asynchronous on cuda_stream; with a host buffer on either side it
 * returns after y is complete.  batch may be any value >= 0 (ragged last batch, CallVariantsFromCffi.py:106-148). */
int c3b_forward(c3b_model *m, const void *x, int x_dtype, int x_on_device, int64_t batch, int depth,
                float *y, int y_on_device, void *cuda_stream);

/* 24 or 90 (clair3/model.py:153-159). */
int c3b_out_dim(const c3b_model *m);

/* Packed device weight image (what one rank broadcasts to the others at start-up). */
int c3b_weight_blob(c3b_model *m, void **device_ptr, size_t *bytes);

/* One-time ncclBroadcast of the packed weight image from rank `root` over NVLink (libnccl is dlopen'ed; the comm is
 * the caller's ncclComm_t).  Multi-GPU inference in the reference is N independent processes over file lists
 * (clair3/CallVariantsFromCffiGPU.py:141-199); there is no per-batch collective to replace. */
int c3b_bcast_weights(c3b_model *m, void *nccl_comm, int root, void *cuda_stream);

/* Debug tap: copy an intermediate activation of the most recent forward (default stream slot, first chunk) to the host
 * as float32.  names: pileup "lstm1"[B,33,256] "lstm2"[B,33,320] "l4_pre"[B,128]; full-alignment "conv1" "res_block1"
 * "conv3" "res_block2" "conv5" "res_block3" (NHWC) "spp"[B,3584] "l4_pre"[B,256].  *count_inout: capacity in / elements out. */
int c3b_get_tap(c3b_model *m, const char *name, float *host_out, int64_t *count_inout);

/* Per-kernel device time accumulated while option "profile" is on.  kernel names: pileup "ingest" "lstm1" "proj2" "lstm2"
 * "l4" "heads"; full-alignment "ingest" "conv0".."conv8" "spp" "l4" "heads".  Synchronises the streams it recorded on. */
int c3b_get_profile(c3b_model *m, const char *kernel, double *total_ms, int64_t *launches);

/* Number of this library's kernels launched on behalf of m so far (bench.py's gpu_launches). */
int64_t c3b_launch_count(const c3b_model *m);

/* Kernel unit-test hook (tests/test_igemm.py), not part of the drop-in surface: run the tcgen05 implicit-GEMM kernel on
 * caller matrices.  out[M][N] = a[M][K] * w[N][K]^T; swapped=0: standard orientation, +bias, optional ReLU, bf16-rounded;
 * swapped=1: weights on the TMEM lanes, split-K (ksplit) fp32 atomic accumulation, no bias.  K % 8 == 0; N % 16 == 0
 * (N % 128 == 0 when swapped). */
int c3b_debug_gemm(c3b_model *m, int swapped, int64_t M, int N, int K, const float *a, const float *w, const float *bias,
                   int relu, int ksplit, float *out);

/* Kernel timing hook: with option "lstm_trace" on, CTA (0,0) of each LSTM kernel stamps clock64 at four points of every
 * step (operands ready, MMAs issued, accumulator ready, epilogue done); copies [2 layers][33 steps][4] stamps out. */
int c3b_debug_lstm_trace(c3b_model *m, int64_t *out264);

/* Hardware probe (tools/diag.py probe): one tcgen05.mma with its A operand in TMEM (checks the assumed layout) and the
 * cycles of `reps` back-to-back MMAs with A from shared memory vs TMEM.  a[128][16], b[n][16] -> out_d[128][n]. */
int c3b_debug_ts_probe(const float *a, const float *b, int n, int reps, float *out_d, int64_t *timing10);
/* debug: cycles for back-to-back tcgen05.mma under operand / accumulator switching (tools/diag.py mmaprobe) */
int c3b_debug_mma_probe(int n, int reps, int nmodes, const int *modes, int64_t *timing);
/* debug: cycles of TMEM reads / an epilogue chunk with the tensor pipe idle and busy (tools/diag.py tmemprobe) */
int c3b_debug_tmem_probe(int reps, int64_t *timing6);

void c3b_destroy(c3b_model *m);

/* Thread-local message for the last non-zero status. */
const char *c3b_last_error(void);
const char *c3b_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CLAIR3_B200_H */
