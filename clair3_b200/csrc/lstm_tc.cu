// Persistent-weight bidirectional LSTM recurrence on tcgen05 (Clair3_P's LSTM1 / LSTM2,
// clair3/model.py:96-107,132-133; torch nn.LSTM semantics: gate rows i,f,g,o, h0 = c0 = 0, the reverse direction
// walks t = 32..0 and both directions are concatenated per time step).
//
// One CTA owns (a tile of NB candidate sites) x (one direction) for all 33 steps.  The gate GEMM is issued "swapped":
// the recurrent weight matrix is the UMMA A operand (gate rows -> the 128 TMEM lanes, one 128-row block per gate), the
// activations [x_t ; h_{t-1}] of the NB sites are the B operand (sites -> TMEM columns).  So
//   * the whole weight matrix stays resident in shared memory for the 33 steps (160 KB LSTM1, 200 KB LSTM2), loaded
//     once per CTA by cp.async.bulk (TMA engine) from a host-packed SWIZZLE_NONE K-major image;
//   * epilogue thread r owns hidden unit r: it reads its i,f,g,o pre-activations for every site of the tile from four
//     TMEM column ranges of its own lane, keeps the cell state c[NB] in fp32 registers for the whole sequence, and writes
//     h_t (bf16) back into the B-operand buffer for step t+1 - no cross-thread exchange, no grid-wide sync;
//   * small NB (16/32/64) gives 2*B/NB CTAs, enough to fill 148 SMs at a 1024-site batch.
//
// LSTM1 (H=128, x has 18 channels zero-padded to 32): K = 32 + 128, the x projection is fused into the same MMAs, bias
// added in the epilogue.  LSTM2 (H=160, input 256): the input projection W_ih*h1 (+bias) is a separate big GEMM
// (igemm_tc.cu) that leaves fp16 pre-gates in the thread-friendly layout pgT[dir][t][tile][blk][row][NB]; this kernel
// keeps only W_hh resident (K = 160).  Units 0..127 are lane-aligned in row blocks 0..3; units 128..159 live in a fifth
// block laid out [i(32) f(32) g(32) o(32)] whose activated gates cross warps through a small shared-memory exchange.
//
// h_t leaves the CTA as 16-byte chunks copied from the operand buffer while the next step's MMAs run.
#include "c3b_internal.h"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 128;            // epilogue threads (one per TMEM lane); warp 4 issues the MMAs
constexpr int kBlockThreads = 160;
constexpr int kBlkBytes = 20 * 128 * 16;       // one 128-row x K=160 weight block: [K/8][128][8] bf16

struct LstmDev {
    const op_t *w_img;    // [dir][NBLK][20][128][8]
    const float *bias;             // [dir][NBLK*128] (LSTM1)
    const op_t *xs;       // LSTM1 input  [33][Bp][32]
    const __half *pg;              // LSTM2 pre-gates pgT[dir][33][Bp/NB][5][128][NB]
    op_t *hout;           // LSTM1: h1[33][Bp][256]; LSTM2: h2[Bp][33][320]
    int bp;                        // padded batch
    long long *trace;              // optional [33][4] clock64 stamps of CTA (0,0) thread 0 (debug option "lstm_trace")
};

template <int NB>
__device__ __forceinline__ void lstm_cell8(const float *gi, const float *gf, const float *gg, const float *go, float *c,
                                           float *h) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float iv = ptx::sigmoid_prehalved(gi[i]);
        const float fv = ptx::sigmoid_prehalved(gf[i]);
        const float gv = ptx::tanh_approx(gg[i]);
        const float ov = ptx::sigmoid_prehalved(go[i]);
        c[i] = fmaf(fv, c[i], iv * gv);
        h[i] = ov * ptx::tanh_approx(c[i]);
    }
}

__device__ __forceinline__ void unpack_half8(const uint4 &v, float *f) {
    const __half2 *hp = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(hp[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

template <int NB, bool LAYER2>
__global__ void __launch_bounds__(kBlockThreads, 1) lstm_tc_kernel(const LstmDev p) {
    constexpr int H = LAYER2 ? 160 : 128;
    constexpr int KX = LAYER2 ? 0 : 32;
    constexpr int K = KX + H;                       // 160 for both layers
    static_assert(K == 160, "weight block image assumes K = 160");
    constexpr int NBLK = LAYER2 ? 5 : 4;
    constexpr uint32_t LBO_B = (NB + 1) * 16;        // padded: conflict-free h stores
    constexpr uint32_t B_BYTES = (K / 8) * LBO_B;
    constexpr uint32_t TCOLS = (NBLK * NB <= 32) ? 32 : (NBLK * NB <= 64) ? 64 : (NBLK * NB <= 128) ? 128
                               : (NBLK * NB <= 256) ? 256 : 512;

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t w_bar, acc_bar, ready_bar;
    __shared__ uint32_t tmem_base_smem;
    uint8_t *w_smem = smem;
    uint8_t *b_smem = smem + NBLK * kBlkBytes;
    float *xch = reinterpret_cast<float *>(b_smem + B_BYTES);     // LAYER2 only: [NB][128]

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int tile = blockIdx.x;
    const int dir = blockIdx.y;
    const int b0 = tile * NB;
    const uint32_t w_addr = ptx::smem_u32(w_smem);
    const uint32_t b_addr = ptx::smem_u32(b_smem);

    if (tid == 0) {
        ptx::mbar_init(&w_bar, 1);
        ptx::mbar_init(&acc_bar, 1);
        ptx::mbar_init(&ready_bar, kThreads);
        ptx::fence_barrier_init();
    }
    if (warp == 4) ptx::tmem_alloc<TCOLS>(&tmem_base_smem);
    for (uint32_t i = tid * 16; i < B_BYTES; i += kBlockThreads * 16) *reinterpret_cast<uint4 *>(b_smem + i) = make_uint4(0, 0, 0, 0);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (tid == 0) {
        ptx::mbar_arrive_expect_tx(&w_bar, NBLK * kBlkBytes);
        const char *src = reinterpret_cast<const char *>(p.w_img) + (size_t)dir * NBLK * kBlkBytes;
#pragma unroll
        for (int m = 0; m < NBLK; ++m) ptx::bulk_g2s(w_addr + m * kBlkBytes, src + (size_t)m * kBlkBytes, kBlkBytes, &w_bar);
    }

    const uint32_t idesc = ptx::umma_idesc_f16(128, NB);
    if (warp == 4) {
        // ===================================================== MMA warp: wait for [x_t ; h_{t-1}], issue the step's gate GEMM
        ptx::mbar_wait(&w_bar, 0);
        for (int step = 0; step < C3B_T; ++step) {
            ptx::mbar_wait(&ready_bar, (uint32_t)step & 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                // rolled k loop: descriptors are re-derived from the loop counter on the uniform datapath every
                // iteration (a fully unrolled loop makes ptxas hoist 60 descriptors into vector registers and pay
                // R2UR moves in front of every UTCHMMA)
#pragma unroll 1
                for (int ks = 0; ks < K / 16; ++ks) {
                    const uint64_t b_desc = ptx::umma_desc_nosw(b_addr + ks * 2 * LBO_B, LBO_B, 128);
                    const uint32_t acc = ks > 0 ? 1u : 0u;
#pragma unroll
                    for (int m = 0; m < NBLK; ++m) {
                        const uint64_t a_desc = ptx::umma_desc_nosw(w_addr + m * kBlkBytes + ks * 2 * 2048, 2048, 128);
                        ptx::umma_f16(tmem_base + m * NB, a_desc, b_desc, idesc, acc);
                    }
                }
                ptx::umma_commit(&acc_bar);
            }
            __syncwarp();
        }
        ptx::tc_fence_before();
        __syncthreads();                      // matches the epilogue threads' final barrier
        ptx::tc_fence_after();
        ptx::tmem_dealloc<TCOLS>(tmem_base);
        return;
    }

    float bias_i = 0.f, bias_f = 0.f, bias_g = 0.f, bias_o = 0.f;
    if (!LAYER2) {
        const float *bp = p.bias + dir * (NBLK * 128);
        bias_i = bp[tid];
        bias_f = bp[128 + tid];
        bias_g = bp[256 + tid];
        bias_o = bp[384 + tid];
    }

    float c[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) c[i] = 0.f;
    float c_tail[LAYER2 ? NB / 4 : 1];
#pragma unroll
    for (int i = 0; i < (LAYER2 ? NB / 4 : 1); ++i) c_tail[i] = 0.f;

    const uint32_t lane_taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int ntl = p.bp / NB;

    int t_prev = 0;
    for (int step = 0; step < C3B_T; ++step) {
        const int t = dir ? (C3B_T - 1 - step) : step;

        if (!LAYER2) {
            // stage x_t: xs[t][b0+n][0..31] -> operand k-groups 0..3
            for (int idx = tid; idx < NB * 4; idx += kThreads) {
                const int n = idx >> 2, kg = idx & 3;
                const uint4 v = *reinterpret_cast<const uint4 *>(p.xs + ((size_t)t * p.bp + b0 + n) * 32 + kg * 8);
                *reinterpret_cast<uint4 *>(b_smem + kg * LBO_B + n * 16) = v;
            }
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&ready_bar);                                       // S1: this thread's operands / TMEM reads are done
        const bool tr = p.trace != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
        if (tr) p.trace[step * 4 + 0] = clock64();
        if (tr) p.trace[step * 4 + 1] = clock64();

        // while the MMAs run: ship h_{t_prev} (still in the operand buffer) to global memory
        if (step > 0) {
            for (int idx = tid; idx < NB * (H / 8); idx += kThreads) {
                const int n = idx / (H / 8), kgh = idx % (H / 8);
                const uint4 v = *reinterpret_cast<const uint4 *>(b_smem + (KX / 8 + kgh) * LBO_B + n * 16);
                op_t *dst = LAYER2 ? p.hout + ((size_t)(b0 + n) * C3B_T + t_prev) * 320 + dir * 160 + kgh * 8
                                            : p.hout + ((size_t)t_prev * p.bp + b0 + n) * 256 + dir * 128 + kgh * 8;
                *reinterpret_cast<uint4 *>(dst) = v;
            }
        }
        // LSTM2: prefetch this step's pre-gates (fp16, NB contiguous values per (block,row))
        uint4 pgv[LAYER2 ? 5 : 1][LAYER2 ? NB / 8 : 1];
        if (LAYER2) {
            const __half *pgp = p.pg + ((((size_t)(dir * C3B_T + t) * ntl + tile) * 5) * 128 + tid) * NB;
#pragma unroll
            for (int m = 0; m < 5; ++m)
#pragma unroll
                for (int j = 0; j < NB / 8; ++j)
                    pgv[m][j] = *reinterpret_cast<const uint4 *>(pgp + (size_t)m * 128 * NB + j * 8);
        }

        ptx::mbar_wait(&acc_bar, (uint32_t)step & 1u);
        ptx::tc_fence_after();
        if (tr) p.trace[step * 4 + 2] = clock64();
        ptx::named_bar_sync(1, kThreads);                                   // S2: everyone is done reading h_{t_prev}

        if (LAYER2) {
            // tail block (units 128..159): warp w holds gate w; activate and publish to the exchange buffer
#pragma unroll
            for (int j = 0; j < NB / 8; ++j) {
                float v[8], pgf[8];
                ptx::tmem_ld8(lane_taddr + 4 * NB + j * 8, v);
                ptx::tmem_ld_wait();
                unpack_half8(pgv[4][j], pgf);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float x = v[i] + pgf[i];
                    const float a = (warp == 2) ? ptx::tanh_approx(x) : ptx::sigmoid_prehalved(x);
                    xch[(j * 8 + i) * 128 + tid] = a;
                }
            }
            ptx::named_bar_sync(1, kThreads);
        }

        // main blocks: thread = hidden unit `tid`, 8 sites at a time
#pragma unroll
        for (int j = 0; j < NB / 8; ++j) {
            float gi[8], gf[8], gg[8], go[8], h[8];
            ptx::tmem_ld8(lane_taddr + 0 * NB + j * 8, gi);
            ptx::tmem_ld8(lane_taddr + 1 * NB + j * 8, gf);
            ptx::tmem_ld8(lane_taddr + 2 * NB + j * 8, gg);
            ptx::tmem_ld8(lane_taddr + 3 * NB + j * 8, go);
            ptx::tmem_ld_wait();
            if (LAYER2) {
                float pf[8];
                unpack_half8(pgv[0][j], pf);
#pragma unroll
                for (int i = 0; i < 8; ++i) gi[i] += pf[i];
                unpack_half8(pgv[1][j], pf);
#pragma unroll
                for (int i = 0; i < 8; ++i) gf[i] += pf[i];
                unpack_half8(pgv[2][j], pf);
#pragma unroll
                for (int i = 0; i < 8; ++i) gg[i] += pf[i];
                unpack_half8(pgv[3][j], pf);
#pragma unroll
                for (int i = 0; i < 8; ++i) go[i] += pf[i];
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    gi[i] += bias_i;
                    gf[i] += bias_f;
                    gg[i] += bias_g;
                    go[i] += bias_o;
                }
            }
            lstm_cell8<NB>(gi, gf, gg, go, &c[j * 8], h);
            // h[n][unit] -> operand buffer (bf16), element (n, k = KX + tid)
            const uint32_t kcol = KX + tid;
            uint8_t *dst = b_smem + (kcol >> 3) * LBO_B + (kcol & 7) * 2;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<op_t *>(dst + (j * 8 + i) * 16) = f2op(h[i]);
        }

        if (LAYER2) {
            // tail cells: unit 128 + lane, sites n = warp + 4*k
#pragma unroll
            for (int k = 0; k < NB / 4; ++k) {
                const int n = warp + 4 * k;
                const float iv = xch[n * 128 + lane];
                const float fv = xch[n * 128 + 32 + lane];
                const float gv = xch[n * 128 + 64 + lane];
                const float ov = xch[n * 128 + 96 + lane];
                c_tail[k] = fmaf(fv, c_tail[k], iv * gv);
                const float hv = ov * ptx::tanh_approx(c_tail[k]);
                const uint32_t kcol = 128 + lane;
                *reinterpret_cast<op_t *>(b_smem + (kcol >> 3) * LBO_B + n * 16 + (kcol & 7) * 2) =
                    f2op(hv);
            }
        }
        t_prev = t;
        if (tr) p.trace[step * 4 + 3] = clock64();
    }

    // last h
    ptx::named_bar_sync(1, kThreads);
    for (int idx = tid; idx < NB * (H / 8); idx += kThreads) {
        const int n = idx / (H / 8), kgh = idx % (H / 8);
        const uint4 v = *reinterpret_cast<const uint4 *>(b_smem + (KX / 8 + kgh) * LBO_B + n * 16);
        op_t *dst = LAYER2 ? p.hout + ((size_t)(b0 + n) * C3B_T + t_prev) * 320 + dir * 160 + kgh * 8
                                    : p.hout + ((size_t)t_prev * p.bp + b0 + n) * 256 + dir * 128 + kgh * 8;
        *reinterpret_cast<uint4 *>(dst) = v;
    }
    ptx::tc_fence_before();
    __syncthreads();
}

template <typename T>
__global__ void ingest_pileup_tc_kernel(const T *__restrict__ x, op_t *__restrict__ xs, int64_t batch, int bp,
                                        int channels) {
    // one thread per 8-channel group of xs[t][b][32]
    const int64_t total = (int64_t)C3B_T * bp * 4;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int kg = (int)(idx & 3);
        const int64_t tb = idx >> 2;
        const int b = (int)(tb % bp);
        const int t = (int)(tb / bp);
        __align__(16) op_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ch = kg * 8 + i;
            float f = 0.f;
            if (b < batch && ch < channels) f = (float)x[((int64_t)b * C3B_T + t) * channels + ch];
            v[i] = f2op(f);
        }
        *reinterpret_cast<uint4 *>(xs + idx * 8) = *reinterpret_cast<const uint4 *>(v);
    }
}

template <int NB, bool LAYER2>
int launch_lstm(const LstmDev &p, cudaStream_t s) {
    constexpr int NBLK = LAYER2 ? 5 : 4;
    const size_t smem = (size_t)NBLK * kBlkBytes + 20 * (NB + 1) * 16 + (LAYER2 ? (size_t)NB * 128 * 4 : 0);
    auto kern = lstm_tc_kernel<NB, LAYER2>;
    C3B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(p.bp / NB, 2);
    kern<<<grid, kBlockThreads, smem, s>>>(p);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int c3b_launch_ingest_pileup_tc(const void *x, int dtype, int channels, op_t *xs, int64_t batch, cudaStream_t s) {
    const int bp = (int)((batch + 127) / 128 * 128);
    const int64_t total = (int64_t)C3B_T * bp * 4;
    const int blocks = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
    switch (dtype) {
        case C3B_DT_I8: ingest_pileup_tc_kernel<int8_t><<<blocks, 256, 0, s>>>((const int8_t *)x, xs, batch, bp, channels); break;
        case C3B_DT_I32: ingest_pileup_tc_kernel<int32_t><<<blocks, 256, 0, s>>>((const int32_t *)x, xs, batch, bp, channels); break;
        case C3B_DT_F32: ingest_pileup_tc_kernel<float><<<blocks, 256, 0, s>>>((const float *)x, xs, batch, bp, channels); break;
        default: c3b_set_error("unsupported input dtype %d", dtype); return 1;
    }
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_lstm1_tc(const c3b_model *m, const TcPileupBuffers &b, int64_t batch, int tile, cudaStream_t s) {
    LstmDev p = {};
    p.w_img = m->lstm_tc[0][0].w_img;     // both directions are contiguous
    p.bias = m->lstm_tc[0][0].bias;
    p.xs = b.xs;
    p.hout = b.h1;
    p.trace = m->lstm_trace ? m->lstm_trace : nullptr;
    p.bp = (int)((batch + 127) / 128 * 128);
    const_cast<c3b_model *>(m)->launches++;
    switch (tile) {
        case 16: return launch_lstm<16, false>(p, s);
        case 32: return launch_lstm<32, false>(p, s);
        case 64: return launch_lstm<64, false>(p, s);
    }
    c3b_set_error("lstm1: unsupported tile %d", tile);
    return 1;
}

int c3b_launch_lstm2_tc(const c3b_model *m, const TcPileupBuffers &b, int64_t batch, int tile, cudaStream_t s) {
    LstmDev p = {};
    p.w_img = m->lstm_tc[1][0].w_img;
    p.pg = b.pg;
    p.hout = b.h2;
    p.trace = m->lstm_trace ? m->lstm_trace + C3B_T * 4 : nullptr;
    p.bp = (int)((batch + 127) / 128 * 128);
    const_cast<c3b_model *>(m)->launches++;
    switch (tile) {
        case 16: return launch_lstm<16, true>(p, s);
        case 32: return launch_lstm<32, true>(p, s);
    }
    c3b_set_error("lstm2: unsupported tile %d", tile);
    return 1;
}
