// Implicit GEMM on 5th-gen tensor cores (tcgen05, accumulators in TMEM) with fused epilogues.
//
//   D[rows x cols] = ACT[act_rows x K] * W[w_rows x K]^T      fp16 operands, fp32 accumulate
//
// Used for (a) the 3x3 convolutions of Clair3_F (clair3/model.py:183-235; BatchNorm folded, bias+ReLU(+residual)
// epilogue, NHWC fp16 in/out), (b) the LSTM2 input projection of Clair3_P (the W_ih half of nn.LSTM's gate GEMM,
// clair3/model.py:102-107) and (c) the L4 dense layers (clair3/model.py:110,344) as split-K.
//
// Operand staging (no tensor maps): weights are pre-packed on the host into the exact shared-memory image of a
// SWIZZLE_NONE K-major UMMA operand, one contiguous piece per (k-chunk, row-block), and land in shared memory with
// cp.async.bulk (TMA engine).  When the CTA's whole weight slab fits (LSTM2 projection: 256 rows x K=256 = 128 KB; conv1,
// res_block1, conv3) it is loaded ONCE per CTA and stays resident ("W-stationary"); otherwise one piece rides in every
// pipeline stage.  Activation tiles are gathered by 128 producer threads with 16-byte cp.async (zero-fill for conv
// padding / ragged rows) straight into the same canonical layout: element (row,k) lives at (k/8)*LBO + row*16 + (k%8)*2
// bytes, i.e. LBO = rows*16, SBO = 128.
//
// Two orientations share the pipeline:
//   standard (SWAP=false): A = activation tile (M = 128 pixels -> TMEM lanes), B = weights (N = Cout <= 256 columns).
//                          A thread of the epilogue owns one pixel and writes its Cout channels contiguously.
//   swapped  (SWAP=true) : A = 128 weight rows (-> TMEM lanes), B = activation tile (N = positions / sites); a CTA tile
//                          may cover two 128-row blocks (two accumulators per tile) so every activation byte fetched
//                          from L2 feeds 256 weight rows.  A thread owns one output unit and a run of consecutive
//                          positions: the layout the persistent LSTM kernel wants its pre-gates in, and coalesced
//                          split-K atomics for L4.
//
// Roles (288 threads): warps 0-3 activation producers (+ thread 0 issues weight bulk copies), warps 4-7 epilogue
// (TMEM lane quadrant = warp % 4), warp 8 allocates TMEM and its lane 0 issues tcgen05.mma.  Persistent over tiles with
// two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "c3b_internal.h"
#include <limits.h>

#include "ptx.cuh"

namespace {

constexpr int kThreads = 320;       // warps 0-3 producers / 2nd epilogue group, 4-7 epilogue, 8 MMA, 9 bulk-copy loader
constexpr int kProducerThreads = 128;
constexpr int kLag = 3;            // cp.async groups in flight per producer thread before the oldest is published
constexpr int kMaxStages = 8;

struct IgemmDev {
    const op_t *act;
    const op_t *w_img;
    const float *bias;
    void *out;
    const op_t *residual;
    int64_t m_valid;       // valid activation rows (pixels / positions / sites)
    int64_t lda;           // plain mode row stride (elements)
    int64_t ldo;           // output row stride (elements)
    int taps;              // 9 conv gather, 1 plain row-major gather, 0 plain k-group-planar ([K/8][ld_rows][8]) via bulk copies
    int64_t ld_rows;       // planar mode: rows per k-group plane
    int hin, win, cin, hout, wout, stride;
    int cpk_shift;         // log2(cin/8) in conv mode
    int kgroups;           // real K/8
    int nchunks;           // total k-chunks
    int act_rows;          // activation rows per tile (128)
    int w_rows;            // weight rows per piece (standard: N, swapped: 128)
    int n_rowblocks;       // weight row blocks in the image (standard: 1)
    int wb;                // row blocks per CTA tile (1; 2 for the swapped resident mode)
    int w_resident;        // 1: the CTA's weight slab is loaded once and stays in shared memory
    int n_act_tiles;
    int ksplit;
    int chunks_per_split;
    int stages;
    int relu;
    int in_planar, out_planar;   // conv mode: planar padded tensors (pconv_tc.cu) on the input / output side
    PlanarGeom gin, gout;
    int64_t split_stride;  // split-K partial epilogue: elements between ks slices
    int pg_bp;             // pre-gate epilogue: padded batch (multiple of 128)
    int pg_nbl;            //                    LSTM tile (batch columns per LSTM CTA)
    long long *trace;      // optional clock stamps of CTA 0, [tile][8]
    int stage_out;         // pre-gate epilogue: 1 = stage [32 rows][NBL] blocks in shared memory and bulk-store them
};

// Tile walk shared by the three roles.  Streaming mode: tile = ((at * n_rowblocks) + rb) * ksplit + ks over a flat grid.
// Resident mode: the CTA is pinned to row-block group blockIdx.x % n_rg and strides over activation tiles.
struct TileWalk {
    int n_rg, at, at_step, tile, num_tiles;
    __device__ TileWalk(const IgemmDev &p) {
        n_rg = p.n_rowblocks / p.wb;
        num_tiles = p.n_act_tiles * p.n_rowblocks * p.ksplit;
        tile = blockIdx.x;
        at = blockIdx.x / n_rg;
        at_step = gridDim.x / n_rg;
    }
    __device__ bool valid(const IgemmDev &p) const { return p.w_resident ? at < p.n_act_tiles : tile < num_tiles; }
    __device__ void next(const IgemmDev &p) {
        if (p.w_resident) at += at_step; else tile += gridDim.x;
    }
    __device__ void decode(const IgemmDev &p, int &act_tile, int &rb0, int &c_begin, int &c_end) const {
        if (p.w_resident) {
            act_tile = at;
            rb0 = (blockIdx.x % n_rg) * p.wb;
            c_begin = 0;
            c_end = p.nchunks;
        } else {
            const int ks = tile % p.ksplit;
            rb0 = (tile / p.ksplit) % p.n_rowblocks;
            act_tile = tile / (p.ksplit * p.n_rowblocks);
            c_begin = ks * p.chunks_per_split;
            c_end = min(p.nchunks, c_begin + p.chunks_per_split);
        }
    }
};

template <bool SWAP, int EPI>
__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const IgemmDev p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[kMaxStages];
    __shared__ uint64_t empty_bar[kMaxStages];
    __shared__ uint64_t tmem_full_bar[2];
    __shared__ uint64_t tmem_empty_bar[2];
    __shared__ uint64_t w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bias_s[SWAP ? 1 : 256];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int S = p.stages;
    // gathered tiles pad the k-group pitch by one row (conflict-free cp.async writes); bulk-copied planar tiles keep it at
    // 128 rows (one contiguous 2 KB run per k-group: a single bulk copy each)
    const uint32_t act_pad = p.taps == 0 ? 0u : 1u;
    const uint32_t act_bytes = (uint32_t)(p.act_rows + act_pad) * 128u;
    const uint32_t w_bytes = (uint32_t)p.w_rows * 128u;
    const uint32_t w_res_bytes = p.w_resident ? (uint32_t)p.wb * p.nchunks * w_bytes : 0u;
    const uint32_t stage_bytes = act_bytes + (p.w_resident ? 0u : w_bytes);
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t stages_base = smem_base + w_res_bytes;
    const int ncols = SWAP ? p.act_rows : p.w_rows;     // accumulator columns per row block
    const int acc_stride = p.wb * ncols;                // TMEM columns per accumulator stage (<= 256)

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full_bar[s], p.taps == 0 ? 1 : kProducerThreads + (p.w_resident ? 0 : 1));
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full_bar[a], 1);
            ptx::mbar_init(&tmem_empty_bar[a], (SWAP && p.taps == 0) ? 256 : 128);
        }
        ptx::mbar_init(&w_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 8) ptx::tmem_alloc<512>(&tmem_base_smem);
    if (!SWAP) {
        for (int i = tid; i < p.w_rows; i += kThreads) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const bool two_groups = SWAP && p.taps == 0;       // planar operands: warps 0-3 are free -> second epilogue warpgroup
    if (warp == 9) {
        // ===================================================== bulk-copy loader (one thread): resident weight slab and, for
        // k-group-planar activations, every stage (8 contiguous 2 KB runs per chunk; TMA engine, async proxy, no fences)
        // The whole warp walks the tiles; per chunk lane 0 arms the barrier and lanes 0..7 each issue ONE 2 KB run (lane 8 the
        // streamed weight piece), so the per-copy address arithmetic runs in parallel instead of as one thread's serial
        // instruction stream (measured: the single-thread loader, not the tensor pipe, bounded the LSTM2 projection).
        if (p.w_resident && lane == 0) {
            const int rbr = (blockIdx.x % (p.n_rowblocks / p.wb)) * p.wb;
            ptx::mbar_arrive_expect_tx(&w_bar, w_res_bytes);
            for (int c = 0; c < p.nchunks; ++c)
                for (int b = 0; b < p.wb; ++b)
                    ptx::bulk_g2s(smem_base + (uint32_t)(c * p.wb + b) * w_bytes,
                                  (const char *)p.w_img + ((size_t)c * p.n_rowblocks + rbr + b) * w_bytes, w_bytes, &w_bar);
        }
        if (p.taps == 0) {
            const uint32_t lbo_a = (uint32_t)p.act_rows * 16u;
            const size_t kg_pitch = (size_t)p.ld_rows * 16;
            int s = 0;
            uint32_t ph = 0;
            for (TileWalk tw(p); tw.valid(p); tw.next(p)) {
                int at, rb0, c_begin, c_end;
                tw.decode(p, at, rb0, c_begin, c_end);
                const char *src_lane = (const char *)p.act + (size_t)at * 2048 + (size_t)(c_begin * 8 + lane) * kg_pitch;
                for (int c = c_begin; c < c_end; ++c, src_lane += 8 * kg_pitch) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
                    const uint32_t stage = stages_base + (uint32_t)s * stage_bytes;
                    const int kgs = min(8, p.kgroups - c * 8);
                    if (lane == 0) ptx::mbar_arrive_expect_tx(&full_bar[s], (uint32_t)kgs * 2048u + (p.w_resident ? 0u : w_bytes));
                    __syncwarp();
                    if (lane < kgs) ptx::bulk_g2s(stage + (uint32_t)lane * lbo_a, src_lane, 2048u, &full_bar[s]);
                    else if (lane == 8 && !p.w_resident)
                        ptx::bulk_g2s(stage + act_bytes, (const char *)p.w_img + ((size_t)c * p.n_rowblocks + rb0) * w_bytes, w_bytes, &full_bar[s]);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp < 4 && !two_groups) {
        // ===================================================== activation producers (16-byte cp.async gathers)
        {
        // Thread -> (k-group kgl = tid & 7, rows (tid >> 3) + 16 j): one warp-level cp.async covers 4 rows x 128
        // contiguous bytes of the activation matrix (fully coalesced); the padded LBO keeps the smem side conflict-free.
        const int kgl = tid & 7;
        const int rsub = tid >> 3;
        const uint32_t lbo_act = (uint32_t)(p.act_rows + 1) * 16u;
        int it = 0;                      // chunk iteration counter of this CTA (ring position)
        int pend0 = -1, pend1 = -1, pend2 = -1;   // stages issued but not yet published (oldest first)
        for (TileWalk tw(p); tw.valid(p); tw.next(p)) {
            int at, rb0, c_begin, c_end;
            tw.decode(p, at, rb0, c_begin, c_end);
            const char *row_base[8];
            int hw0[8];                  // (hi0 << 16) | (wi0 & 0xffff), or INT_MIN for an invalid row
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t g = (int64_t)at * p.act_rows + rsub + 16 * j;
                row_base[j] = (const char *)p.act;
                hw0[j] = INT_MIN;
                if (g < p.m_valid) {
                    if (p.taps == 1) {
                        row_base[j] = (const char *)(p.act + g * p.lda);
                        hw0[j] = 0;
                    } else {
                        const int wo = (int)(g % p.wout);
                        const int ho = (int)((g / p.wout) % p.hout);
                        const int64_t b = g / ((int64_t)p.wout * p.hout);
                        if (p.in_planar) {
                            // slot of padded input pixel (ho*stride + dh, wo*stride + dw) for dh = dw = 0; no bounds checks needed
                            row_base[j] = (const char *)p.act +
                                          ((size_t)p.gin.g + b * p.gin.s + (size_t)(ho * p.stride) * p.gin.wp + wo * p.stride) * 16;
                            hw0[j] = 0;
                        } else {
                            row_base[j] = (const char *)(p.act + b * (int64_t)p.hin * p.win * p.cin);
                            hw0[j] = ((ho * p.stride - 1) << 16) | ((wo * p.stride - 1) & 0xffff);
                        }
                    }
                }
            }
            for (int c = c_begin; c < c_end; ++c, ++it) {
                const int s = it % S;
                const uint32_t ph = (uint32_t)(it / S) & 1u;
                ptx::mbar_wait(&empty_bar[s], ph ^ 1u);
                const uint32_t stage = stages_base + (uint32_t)s * stage_bytes;
                if (!p.w_resident && tid == 0) {
                    ptx::mbar_arrive_expect_tx(&full_bar[s], w_bytes);
                    const char *src = (const char *)p.w_img + ((size_t)c * p.n_rowblocks + rb0) * w_bytes;
                    ptx::bulk_g2s(stage + act_bytes, src, w_bytes, &full_bar[s]);
                }
                const int gk = c * 8 + kgl;
                const bool k_ok = gk < p.kgroups;
                int dh = 0, dw = 0;
                size_t koff = (size_t)gk * 16;                   // plain mode: byte offset inside the row
                if (p.taps != 1) {
                    const int tap = gk >> p.cpk_shift;
                    const int c8 = gk - (tap << p.cpk_shift);
                    dh = tap / 3;
                    dw = tap - dh * 3;
                    koff = p.in_planar ? (size_t)c8 * p.gin.p * 16 : (size_t)c8 * 16;   // planar: channel group = plane
                }
                const uint32_t dst0 = stage + (uint32_t)kgl * lbo_act + (uint32_t)rsub * 16u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const char *src = (const char *)p.act;
                    uint32_t nbytes = 0;
                    if (k_ok && hw0[j] != INT_MIN) {
                        if (p.taps == 1) {
                            src = row_base[j] + koff;
                            nbytes = 16;
                        } else if (p.in_planar) {
                            src = row_base[j] + ((size_t)dh * p.gin.wp + dw) * 16 + koff;
                            nbytes = 16;
                        } else {
                            const int hi = (hw0[j] >> 16) + dh;
                            const int wi = (int)(short)(hw0[j] & 0xffff) + dw;
                            if (hi >= 0 && hi < p.hin && wi >= 0 && wi < p.win) {
                                src = row_base[j] + ((size_t)(hi * p.win + wi) * p.cin) * 2 + koff;
                                nbytes = 16;
                            }
                        }
                    }
                    ptx::cp_async16(dst0 + (uint32_t)j * 256u, src, nbytes);
                }
                ptx::cp_async_commit();
                if (pend2 >= 0) {                                // kLag groups already in flight: publish the oldest
                    ptx::cp_async_wait<kLag>();
                    ptx::fence_proxy_async_smem();
                    ptx::mbar_arrive(&full_bar[pend2]);
                }
                pend2 = pend1;
                pend1 = pend0;
                pend0 = s;
            }
        }
        ptx::cp_async_wait<0>();
        ptx::fence_proxy_async_smem();
        if (pend2 >= 0) ptx::mbar_arrive(&full_bar[pend2]);
        if (pend1 >= 0) ptx::mbar_arrive(&full_bar[pend1]);
        if (pend0 >= 0) ptx::mbar_arrive(&full_bar[pend0]);
        }   // gather producers
    } else if (warp == 8) {
        // ===================================================== MMA issuer (whole warp walks the tiles, one elected lane issues)
        const uint32_t idesc = ptx::umma_idesc_f16(128, (uint32_t)ncols);
        const uint32_t lbo_act = (uint32_t)(p.act_rows + act_pad) * 16u;
        const uint32_t lbo_w = (uint32_t)p.w_rows * 16u;
        // ONE elected thread runs the whole loop (barrier waits included).  tcgen05.mma issue does not run ahead of the
        // tensor pipe, so every instruction between two MMAs is pipe idle time: descriptors are advanced with 32-bit adds on
        // their start-address field (shared memory < 256 KB: no carry out of the 14 bits), ring position kept as counters.
        if (ptx::elect_one()) {
            const uint64_t act_d0 = ptx::umma_desc_nosw(0, lbo_act, 128u), w_d0 = ptx::umma_desc_nosw(0, lbo_w, 128u);
            const uint32_t act_lo0 = (uint32_t)act_d0, act_hi = (uint32_t)(act_d0 >> 32);
            const uint32_t w_lo0 = (uint32_t)w_d0, w_hi = (uint32_t)(w_d0 >> 32);
            const uint32_t act_kstep = (2u * lbo_act) >> 4, w_kstep = (2u * lbo_w) >> 4;
            const uint32_t w_bstep = w_bytes >> 4;          // next resident row block of the same chunk
            int s = 0;
            uint32_t ph = 0;
            int tcount = 0;
            if (p.w_resident) ptx::mbar_wait(&w_bar, 0);
            for (TileWalk tw(p); tw.valid(p); tw.next(p), ++tcount) {
                int at, rb0, c_begin, c_end;
                tw.decode(p, at, rb0, c_begin, c_end);
                const int acc = tcount & 1;
                const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
                const bool tr = p.trace != nullptr && blockIdx.x == 0 && tcount < 8;
                if (tr) p.trace[tcount * 8 + 0] = clock64();
                ptx::mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1u);
                ptx::tc_fence_after();
                if (tr) p.trace[tcount * 8 + 1] = clock64();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * acc_stride);
                for (int c = c_begin; c < c_end; ++c) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    if (tr && c == c_begin) p.trace[tcount * 8 + 2] = clock64();
                    const uint32_t stage = stages_base + (uint32_t)s * stage_bytes;
                    const int kgs = min(8, p.kgroups - c * 8);
                    const int ksteps = (kgs + 1) >> 1;
                    const uint32_t act_lo = act_lo0 + (stage >> 4);
                    const uint32_t w_lo = w_lo0 + ((p.w_resident ? smem_base + (uint32_t)(c * p.wb) * w_bytes : stage + act_bytes) >> 4);
                    if (ksteps == 4 && p.wb == 2) {             // the hot shape (LSTM2 input projection): fully unrolled
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = ((uint64_t)act_hi << 32) | (uint64_t)(act_lo + (uint32_t)k * act_kstep);
                            const uint32_t accum = (c > c_begin || k > 0) ? 1u : 0u;
#pragma unroll
                            for (int b = 0; b < 2; ++b) {
                                const uint64_t wd = ((uint64_t)w_hi << 32) | (uint64_t)(w_lo + (uint32_t)b * w_bstep + (uint32_t)k * w_kstep);
                                if (SWAP) ptx::umma_f16(d_tmem + (uint32_t)(b * ncols), wd, ad, idesc, accum);
                                else ptx::umma_f16(d_tmem, ad, wd, idesc, accum);
                            }
                        }
                    } else if (ksteps == 4 && p.wb == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t ad = ((uint64_t)act_hi << 32) | (uint64_t)(act_lo + (uint32_t)k * act_kstep);
                            const uint64_t wd = ((uint64_t)w_hi << 32) | (uint64_t)(w_lo + (uint32_t)k * w_kstep);
                            const uint32_t accum = (c > c_begin || k > 0) ? 1u : 0u;
                            if (SWAP) ptx::umma_f16(d_tmem, wd, ad, idesc, accum);
                            else ptx::umma_f16(d_tmem, ad, wd, idesc, accum);
                        }
                    } else {
                        for (int k = 0; k < ksteps; ++k) {
                            const uint64_t ad = ((uint64_t)act_hi << 32) | (uint64_t)(act_lo + (uint32_t)k * act_kstep);
                            const uint32_t accum = (c > c_begin || k > 0) ? 1u : 0u;
                            for (int b = 0; b < p.wb; ++b) {
                                const uint64_t wd = ((uint64_t)w_hi << 32) | (uint64_t)(w_lo + (uint32_t)b * w_bstep + (uint32_t)k * w_kstep);
                                if (SWAP) ptx::umma_f16(d_tmem + (uint32_t)(b * ncols), wd, ad, idesc, accum);
                                else ptx::umma_f16(d_tmem, ad, wd, idesc, accum);
                            }
                        }
                    }
                    ptx::umma_commit(&empty_bar[s]);
                    if (c + 1 == c_end) ptx::umma_commit(&tmem_full_bar[acc]);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                if (tr) p.trace[tcount * 8 + 3] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue warps 4..7 (+ warps 0..3 as a second group when the
        // operands arrive by bulk copy): with two row blocks per tile each group takes one, otherwise half the columns
        const int eg = warp < 4 ? 1 : 0;
        const int q = warp & 3;                       // TMEM lane quadrant
        const int r = q * 32 + lane;                  // lane / row within the tile
        int tcount = 0;
        for (TileWalk tw(p); tw.valid(p); tw.next(p), ++tcount) {
            int at, rb0, c_begin, c_end;
            tw.decode(p, at, rb0, c_begin, c_end);
            const int acc = tcount & 1;
            const uint32_t acc_ph = (uint32_t)(tcount >> 1) & 1u;
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 128 && tcount < 8;
            if (tr) p.trace[tcount * 8 + 4] = clock64();
            ptx::mbar_wait(&tmem_full_bar[acc], acc_ph);
            ptx::tc_fence_after();
            if (tr) p.trace[tcount * 8 + 5] = clock64();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * acc_stride);
            if (!SWAP) {
                // thread = pixel row, columns = output channels
                const int64_t pix = (int64_t)at * 128 + r;
                const bool ok = pix < p.m_valid;
                size_t pslot = 0;                                  // planar padded output: slot of this pixel
                if (p.out_planar && ok) {
                    const int wo = (int)(pix % p.wout);
                    const int ho = (int)((pix / p.wout) % p.hout);
                    const int64_t b = pix / ((int64_t)p.wout * p.hout);
                    pslot = (size_t)p.gout.g + b * p.gout.s + (size_t)(ho + 1) * p.gout.wp + (wo + 1);
                }
                op_t *orow = (op_t *)p.out + pix * p.ldo;
                const op_t *rrow = p.residual ? p.residual + pix * p.ldo : nullptr;
                for (int j0 = 0; j0 < ncols; j0 += 16) {
                    float v[16];
                    ptx::tmem_ld16(taddr + (uint32_t)j0, v);
                    ptx::tmem_ld_wait();
                    if (ok) {
                        uint4 res[2];
                        if (rrow) {
                            res[0] = *reinterpret_cast<const uint4 *>(rrow + j0);
                            res[1] = *reinterpret_cast<const uint4 *>(rrow + j0 + 8);
                        }
                        uint4 pk[2];
                        uint32_t *pw = reinterpret_cast<uint32_t *>(pk);
                        const op2_t *rp = reinterpret_cast<const op2_t *>(res);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float a = v[2 * i] + bias_s[j0 + 2 * i];
                            float b = v[2 * i + 1] + bias_s[j0 + 2 * i + 1];
                            if (rrow) {
                                const float2 rf = op22f2(rp[i]);
                                a += rf.x;
                                b += rf.y;
                            }
                            if (p.relu) {
                                a = fmaxf(a, 0.f);
                                b = fmaxf(b, 0.f);
                            }
                            pw[i] = f2op2_sat(a, b);
                        }
                        if (p.out_planar) {
                            op_t *o0 = (op_t *)p.out + ((size_t)(j0 >> 3) * p.gout.p + pslot) * 8;
                            *reinterpret_cast<uint4 *>(o0) = pk[0];
                            *reinterpret_cast<uint4 *>(o0 + (size_t)p.gout.p * 8) = pk[1];
                        } else {
                            *reinterpret_cast<uint4 *>(orow + j0) = pk[0];
                            *reinterpret_cast<uint4 *>(orow + j0 + 8) = pk[1];
                        }
                    }
                }
            } else {
                // thread = weight row (output unit), columns = positions
                const int64_t pos0 = (int64_t)at * p.act_rows;
                const int b_lo = (two_groups && p.wb == 2) ? eg : 0, b_hi = (two_groups && p.wb == 2) ? eg + 1 : p.wb;
                const int j_lo = (two_groups && p.wb == 1) ? eg * (ncols / 2) : 0;
                const int j_hi = (two_groups && p.wb == 1) ? j_lo + ncols / 2 : ncols;
                for (int b = b_lo; b < b_hi; ++b) {
                    const int R = (rb0 + b) * 128 + r;
                    const uint32_t tb = taddr + (uint32_t)(b * ncols);
                    if (EPI == IGEMM_EPI_F16_BIAS) {
                        // pre-gate layout pgT[dir][t][b/NBL][blk][r][b%NBL]  (see lstm_tc.cu)
                        const float bias = p.bias ? p.bias[R] : 0.f;
                        const int dir = R / 640, blk = (R % 640) >> 7;
                        const int t = (int)(pos0 / p.pg_bp);
                        const int bb0 = (int)(pos0 % p.pg_bp);
                        const int ntl = p.pg_bp / p.pg_nbl;
                        // the tile's 128 positions span 128/NBL consecutive sub-tiles of one time step
                        const size_t row_off = ((size_t)(dir * C3B_T + t) * ntl + bb0 / p.pg_nbl) * 5 * 128 * p.pg_nbl +
                                               ((size_t)blk * 128 + r) * p.pg_nbl;
                        const size_t st_stride = (size_t)5 * 128 * p.pg_nbl;
                        int st_i = j_lo / p.pg_nbl, in_st = j_lo % p.pg_nbl;
                        // A thread owns a ROW of the output (its sites are 128 B apart from the next lane's): direct 16-byte
                        // stores touch 32 lines per instruction, and LSU wavefronts are what the concurrent MMA operand fetches
                        // starve (measured: an epilogue chunk costs 2.7x more under a running MMA stream).  So each warp
                        // transposes its [32 rows][NBL] block through shared memory (XOR-swizzled 16-byte chunks: both phases
                        // conflict-free) and writes it back with fully coalesced 16-byte stores: 96 wavefronts per 4 KB block
                        // instead of 256.  (Handing the block to the TMA engine was no faster, and a swizzled GLOBAL layout
                        // cost the LSTM2 kernel more than it saved here.)
                        {
                            const uint32_t row_bytes = (uint32_t)p.pg_nbl * 2u;
                            const uint32_t nc = (uint32_t)p.pg_nbl >> 3;                        // 16-byte chunks per row: 2, 4 or 8
                            const uint32_t rsh = nc == 8 ? 0u : nc == 4 ? 1u : 2u;              // rows sharing a 128-byte bank row
                            const uint32_t ncs = nc == 8 ? 3u : nc == 4 ? 2u : 1u;              // log2(nc)
                            uint8_t *stg = smem + w_res_bytes + (uint32_t)S * stage_bytes + (uint32_t)warp * 32u * row_bytes;
                            uint8_t *my_row = stg + (uint32_t)lane * row_bytes;
                            const uint32_t sw = ((uint32_t)lane >> rsh) & (nc - 1u);
                            const size_t blk_off = row_off - (size_t)lane * p.pg_nbl;          // row q*32 of this block
                            for (int j0 = j_lo; j0 < j_hi; j0 += 16) {
                                float v[16];
                                ptx::tmem_ld16(tb + (uint32_t)j0, v);
                                if (in_st == 0) __syncwarp();           // the previous block has been read back
                                ptx::tmem_ld_wait();
                                uint4 pk[2];
                                uint32_t *pw = reinterpret_cast<uint32_t *>(pk);
#pragma unroll
                                for (int i = 0; i < 8; ++i) pw[i] = f2op2_sat(v[2 * i] + bias, v[2 * i + 1] + bias);
                                const uint32_t c0 = (uint32_t)in_st >> 3;
                                *reinterpret_cast<uint4 *>(my_row + ((c0 ^ sw) << 4)) = pk[0];
                                *reinterpret_cast<uint4 *>(my_row + (((c0 + 1u) ^ sw) << 4)) = pk[1];
                                in_st += 16;
                                if (in_st == p.pg_nbl) {
                                    __syncwarp();
                                    uint4 *gdst = reinterpret_cast<uint4 *>((__half *)p.out + blk_off + (size_t)st_i * st_stride);
                                    for (uint32_t u = (uint32_t)lane; u < 32u * nc; u += 32u) {   // 16-byte unit u of the contiguous block
                                        const uint32_t row = u >> ncs, c = u & (nc - 1u);
                                        gdst[u] = *reinterpret_cast<const uint4 *>(stg + row * row_bytes + ((c ^ ((row >> rsh) & (nc - 1u))) << 4));
                                    }
                                    in_st = 0;
                                    ++st_i;
                                }
                            }
                        }
                    } else {
                        // split-K partial sums: partial[ks][pos][R], plain stores (consecutive lanes = consecutive R: 128-byte
                        // warp stores); the consumer (heads kernel) adds the ks slices - no atomics, no memset
                        const int ks = p.w_resident ? 0 : (tw.tile % p.ksplit);
                        float *outp = (float *)p.out + (size_t)ks * p.split_stride;
                        for (int j0 = j_lo; j0 < j_hi; j0 += 16) {
                            float v[16];
                            ptx::tmem_ld16(tb + (uint32_t)j0, v);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int64_t pos = pos0 + j0 + i;
                                if (pos < p.m_valid) outp[pos * p.ldo + R] = v[i];
                            }
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty_bar[acc]);
            if (tr) p.trace[tcount * 8 + 6] = clock64();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

int ilog2(int v) {
    int s = 0;
    while ((1 << s) < v) ++s;
    return s;
}

template <bool SWAP, int EPI>
int launch(const IgemmDev &p, int grid, size_t smem, cudaStream_t s) {
    auto kern = igemm_kernel<SWAP, EPI>;
    C3B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
    kern<<<grid, kThreads, smem, s>>>(p);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int c3b_launch_igemm(const c3b_model *m, const IgemmArgs &a, cudaStream_t s) {
    if (a.m <= 0) return 0;
    IgemmDev p = {};
    p.act = a.a;
    p.w_img = a.w.w_img;
    p.bias = a.w.bias;
    p.out = a.out;
    p.residual = a.residual;
    p.m_valid = a.m;
    p.lda = a.lda;
    p.ldo = a.ldo;
    p.taps = a.taps;
    p.trace = a.trace;
    p.ld_rows = a.ld_rows;
    if (a.taps == 0 && (a.w.kgroups & 1)) { c3b_set_error("igemm(planar): K/8 must be even"); return 1; }
    p.hin = a.hin; p.win = a.win; p.cin = a.cin; p.hout = a.hout; p.wout = a.wout; p.stride = a.stride;
    p.cpk_shift = (a.taps == 9) ? ilog2(a.cin / 8) : 0;
    p.kgroups = a.w.kgroups;
    p.nchunks = a.w.nchunks;
    p.relu = a.relu;
    p.split_stride = a.split_stride;
    p.in_planar = a.in_planar; p.out_planar = a.out_planar; p.gin = a.gin; p.gout = a.gout;
    p.act_rows = 128;
    p.wb = 1;
    const bool swap = (a.epilogue != IGEMM_EPI_BF16_BIAS_RELU);
    if (!swap) {
        if (a.w.n % 16 || a.w.n > 256 || a.w.n < 16) { c3b_set_error("igemm: bad N %d", a.w.n); return 1; }
        p.w_rows = a.w.n;
        p.n_rowblocks = 1;
        p.ksplit = 1;
    } else {
        if (a.w.n % 128) { c3b_set_error("igemm(swap): weight rows %d not a multiple of 128", a.w.n); return 1; }
        p.w_rows = 128;
        p.n_rowblocks = a.w.n / 128;
        p.ksplit = a.ksplit > 0 ? a.ksplit : 1;
    }
    p.n_act_tiles = (int)((a.m + p.act_rows - 1) / p.act_rows);
    p.chunks_per_split = (p.nchunks + p.ksplit - 1) / p.ksplit;
    p.ksplit = (p.nchunks + p.chunks_per_split - 1) / p.chunks_per_split;   // drop empty splits

    // pre-gate epilogue: each of the 8 epilogue warps transposes a [32 rows][NBL] fp16 block through shared memory
    const bool stage_out = a.epilogue == IGEMM_EPI_F16_BIAS;
    if (stage_out && a.taps != 0) { c3b_set_error("igemm: the pre-gate epilogue needs k-group-planar activations"); return 1; }
    p.stage_out = stage_out ? 1 : 0;
    const size_t stage_out_bytes = stage_out ? (size_t)8 * 32 * (size_t)a.win * 2 : 0;
    const size_t budget = (stage_out ? 224 : 216) * 1024 - stage_out_bytes;
    const size_t act_bytes = (size_t)(p.act_rows + (a.taps == 0 ? 0 : 1)) * 128, w_bytes = (size_t)p.w_rows * 128;
    // W-stationary when the slab fits beside >= 4 activation stages (and there are enough tiles to amortise the load)
    if (p.ksplit == 1) {
        int wb = (swap && p.n_rowblocks % 2 == 0 && 2 * p.act_rows <= 256) ? 2 : 1;
        for (; wb >= 1; --wb) {
            const size_t slab = (size_t)wb * p.nchunks * w_bytes;
            if (slab + (stage_out ? 4 : 5) * act_bytes <= budget && p.n_act_tiles >= 2 * (m->sm_count / (p.n_rowblocks / wb))) {
                p.w_resident = 1;
                p.wb = wb;
                break;
            }
        }
    }
    int grid;
    size_t smem;
    if (p.w_resident) {
        const size_t slab = (size_t)p.wb * p.nchunks * w_bytes;
        int stages = (int)((budget - slab) / act_bytes);
        p.stages = stages > kMaxStages ? kMaxStages : stages;
        smem = slab + act_bytes * p.stages + stage_out_bytes + 256;
        const int n_rg = p.n_rowblocks / p.wb;
        int per = m->sm_count / n_rg;
        if (per > p.n_act_tiles) per = p.n_act_tiles;
        grid = n_rg * per;
    } else {
        const size_t stage_bytes = act_bytes + w_bytes;
        int stages = (int)(budget / stage_bytes);
        p.stages = stages > kMaxStages ? kMaxStages : stages;
        smem = stage_bytes * p.stages + stage_out_bytes + 256;
        const int num_tiles = p.n_act_tiles * p.n_rowblocks * p.ksplit;
        grid = num_tiles < m->sm_count ? num_tiles : m->sm_count;
    }
    if (p.stages < (a.taps == 0 ? 2 : kLag + 1)) { c3b_set_error("igemm: tile too large for the shared-memory pipeline"); return 1; }
    const_cast<c3b_model *>(m)->launches++;
    switch (a.epilogue) {
        case IGEMM_EPI_BF16_BIAS_RELU: return launch<false, IGEMM_EPI_BF16_BIAS_RELU>(p, grid, smem, s);
        case IGEMM_EPI_F16_BIAS: {
            p.pg_bp = (int)a.hin;     // caller passes padded batch / LSTM tile through hin / win in plain mode
            p.pg_nbl = (int)a.win;
            if (p.pg_bp % 128 || p.pg_nbl % 16) { c3b_set_error("igemm: bad pre-gate geometry"); return 1; }
            return launch<true, IGEMM_EPI_F16_BIAS>(p, grid, smem, s);
        }
        case IGEMM_EPI_F32_ATOMIC: return launch<true, IGEMM_EPI_F32_ATOMIC>(p, grid, smem, s);
    }
    c3b_set_error("igemm: unknown epilogue %d", a.epilogue);
    return 1;
}
