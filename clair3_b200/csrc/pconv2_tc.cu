// The shifted-view 3x3 convolution of pconv_tc.cu (same planar padded layouts, parity planes, weight images and epilogue; see
// that file for the scheme) in its BLOCK-PIPELINED / CTA-PAIR form: option "pconv_impl" = 1.  What differs:
//   * the image of a macro-tile is split into blocks (one parity plane x 64 channels), each with its own full / free barrier:
//     the MMAs start when the first block has landed, and with the stride-2 taps issued plane by plane a block is refilled for
//     the next macro-tile as soon as its 1-4 taps are done (the 140 KB four-plane images of conv4 / conv7 cannot be double
//     buffered: their load used to serialise with the MMAs - CTA time -23 % / -15 % in the cycle traces);
//   * the issue loop is ROLLED over a host-built piece schedule (weight chunk, image block, A-view offset): the tap-unrolled form
//     is ~9 KB of code and costs each CTA 4-5 k cycles of instruction-cache misses on its first macro-tile;
//   * a separate warp loads the image blocks, so weights and images are requested independently;
//   * STREAMED weights (Cout >= 128) run on a CTA PAIR (cluster of two, tcgen05 cta_group::2, M = 256): see pconv_kernel.
// Measured (tools/round2_ab.sh, same box, 12 streams): 1.08 M (no pairs) / 1.10 M (pairs) sites/s against 1.14 M for the
// pconv_tc.cu form, although the per-CTA cycle traces of conv4, conv7-9 are 10-25 % shorter - which is why this is an option and
// not the default (DESIGN.md 3.6 has the traces and what they say about the issue thread, barriers and shared-memory bandwidth).
//
// Roles (352 threads): warp 10 lane 0 loads the image blocks, warp 9 lane 0 the weight pieces (bulk copies), warp 8 issues
// tcgen05.mma (one elected lane; the peer CTA's warp 8 forwards its arrivals to the leader), warps 0-7 run the epilogue as two
// groups of four (one TMEM lane quadrant per warp, alternating 16-column chunks per group).
#include <cstdio>
#include <cstdlib>

#include "c3b_internal.h"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 352;
constexpr int kMaxWStages = 16;
constexpr int kMaxBlocks = 8;       // image blocks per buffer (64 channels of one parity plane each)
constexpr int kMaxChunks = 36;      // weight pieces per macro-tile (9 taps x C/64)
constexpr uint32_t kSchedFirst = 1u << 16, kSchedLast = 1u << 17;
constexpr int kMetaSlots = 1024;

struct PconvDev {
    const op_t *in;
    const op_t *w_img;       // pair form: the [chunk][rank][8 kg][N/2][8] image (each CTA's half of a piece is one contiguous run)
    const float *bias;
    const op_t *residual;
    op_t *out;
    int C, N;
    int H, W, Wp, S, G;
    long long T, P;
    int MT, n_macro, n_in;
    int nchunks, cpt;
    int w_resident, w_stages, img_bufs, acc_stages;
    int relu;
    // stride-2 stem convs read FOUR parity planes of the previous level (each a planar padded tensor in THIS conv's
    // geometry): tap (dh,dw) = plane (dh&1, dw&1) viewed (dh>>1)*Wp + (dw>>1) slots later - shifted views again
    int nplanes;           // 1 (stride 1) or 4 (stride 2)
    long long plane_elems; // elements between input planes
    int kpt_shift;         // log2(C/16): k-steps per tap
    int nksteps;           // 9 * C/16
    int halo_lo;           // slots loaded before the macro-tile (Wp+1 for stride 1, 0 for stride 2)
    // output: 0 = planar padded (same geometry); 1 = scatter real pixels into the four parity planes of the NEXT level
    int out_parity;
    long long out_plane_elems;
    int nS, nWp, nG;       // next level: slots per site, padded width, guard
    long long nP;          // next level: plane pitch
    long long *trace;      // optional: CTA 0 stamps [macro][8] (debug option "lstm_trace")
    // issue order of the weight pieces of a macro-tile, grouped by image block (block = (parity plane, 64-channel group); stride
    // 2: the taps that read a plane are issued together, so a plane's buffer is free - and refilled for the next tile - as soon
    // as its 1-4 taps are done, and the MMAs start when the first block has landed instead of the whole image)
    int nb, blk_kg;                  // blocks per image buffer; k-groups (bulk copies) per block
    int wg;                          // weight pieces per ring stage (one barrier wait / release per stage)
    uint32_t sched[kMaxChunks];      // weight chunk index | block << 8 | kSchedFirst (first piece of its block) | kSchedLast
    uint32_t a_off[kMaxChunks];      // C >= 64: A-view start of the piece (16-byte units from the image buffer)
    int dbg_skip_w;        // timing experiments only (env C3B_PCONV_SKIPW): never load / wait for streamed weights (wrong results)
};

// Border mask and parity-scatter target of slot l of a site (the epilogue looks this up instead of dividing per tile).
__device__ __forceinline__ uint32_t slot_meta(const PconvDev &p, int l) {
    const int hh = l / p.Wp, ww = l - hh * p.Wp;
    if (!(hh >= 1 && hh <= p.H && ww >= 1 && ww <= p.W)) return 0u;
    uint32_t m = 0x80000000u;
    if (p.out_parity) m |= (uint32_t)((hh & 1) * 2 + (ww & 1)) << 20 | (uint32_t)(((hh >> 1) + 1) * p.nWp + ((ww >> 1) + 1));
    return m;
}

// Epilogue of one macro-tile for one thread (TMEM lane = output slot): bias + residual + ReLU + border mask + fp16 pack +
// 16-byte stores, over the 16-column chunks of this thread's warp group.  TMEM reads are software-pipelined (the load of the
// next chunk - of this or the next tile - is in flight while the current one is finished); residuals are fetched one chunk
// ahead.  RES / PAR are compile-time so the no-residual and planar-output cases carry no dead work.
template <int MT, bool RES, bool PAR>
__device__ __forceinline__ void epilogue_tiles(const PconvDev &p, uint32_t tbase, long long g0, int eg, const float *bias_s,
                                               const uint32_t *meta_s) {
    const int cpg = p.N >> 5;                                            // chunks per group per tile (2, 4 or 8)
    const uint32_t in_pitch = (uint32_t)p.P * 16u;                       // bytes between k-group planes (< 4 GB, host-checked)
    const uint32_t out_pitch = PAR ? (uint32_t)p.nP * 16u : in_pitch;
    float v0[16], v1[16];
    ptx::tmem_ld16(tbase, v0);
#pragma unroll
    for (int ti = 0; ti < MT; ++ti) {
        const long long g = g0 + 128LL * ti;                             // output slot
        const bool in_data = g < p.T;
        bool real = false;
        size_t par_off = 0;                                              // parity-scatter destination (plane + slot of the next level)
        if (in_data) {
            const uint32_t b = (uint32_t)g / (uint32_t)p.S;              // T < 2^31 (checked on the host)
            const int l = (int)((uint32_t)g - b * (uint32_t)p.S);
            const uint32_t meta = p.S <= kMetaSlots ? meta_s[l] : slot_meta(p, l);
            real = (meta >> 31) != 0;
            if (PAR && real)
                par_off = (size_t)((meta >> 20) & 3u) * p.out_plane_elems + ((size_t)p.nG + (size_t)b * p.nS + (meta & 0xFFFFFu)) * 8;
        }
        const bool wr = real;       // border slots are never written: they stay zero from the workspace clear (one clear per geometry)
        const size_t slot_off = ((size_t)p.G + (size_t)g) * 8;
        const uint32_t taddr = tbase + (uint32_t)(ti * p.N);
        char *const obase = reinterpret_cast<char *>(PAR ? p.out + par_off : p.out + slot_off);
        const char *const rbase = reinterpret_cast<const char *>(p.residual + slot_off);
        uint4 res[2], rn[2];
        if (RES && real) {
            res[0] = *reinterpret_cast<const uint4 *>(rbase + (uint32_t)(2 * eg) * in_pitch);
            res[1] = *reinterpret_cast<const uint4 *>(rbase + (uint32_t)(2 * eg + 1) * in_pitch);
        }
        auto finish = [&](const float *v, int j0) {       // one chunk: 16 channels of this slot
            if (!wr) return;
            uint4 pk[2];
            {
                uint32_t *pw = reinterpret_cast<uint32_t *>(pk);
                const float4 *b4 = reinterpret_cast<const float4 *>(bias_s + j0);
                const op2_t *rp = reinterpret_cast<const op2_t *>(res);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bb = b4[i];
                    float x0 = v[4 * i] + bb.x, x1 = v[4 * i + 1] + bb.y, x2 = v[4 * i + 2] + bb.z, x3 = v[4 * i + 3] + bb.w;
                    if (RES) {
                        const float2 ra = op22f2(rp[2 * i]), rb = op22f2(rp[2 * i + 1]);
                        x0 += ra.x; x1 += ra.y; x2 += rb.x; x3 += rb.y;
                    }
                    if (p.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
                    pw[2 * i] = f2op2_sat(x0, x1);
                    pw[2 * i + 1] = f2op2_sat(x2, x3);
                }
            }
            char *d0p = obase + (uint32_t)(j0 >> 3) * out_pitch;
            *reinterpret_cast<uint4 *>(d0p) = pk[0];
            *reinterpret_cast<uint4 *>(d0p + out_pitch) = pk[1];
        };
        auto next_res = [&](int j0n) {                     // residual of the chunk after the current one
            if (RES && real && j0n < p.N) {
                rn[0] = *reinterpret_cast<const uint4 *>(rbase + (uint32_t)(j0n >> 3) * in_pitch);
                rn[1] = *reinterpret_cast<const uint4 *>(rbase + (uint32_t)((j0n >> 3) + 1) * in_pitch);
            }
        };
        for (int jc = 0; jc < cpg; jc += 2) {
            const int j0 = 16 * eg + 32 * jc;
            ptx::tmem_ld_wait();                                         // v0 = chunk jc
            ptx::tmem_ld16(taddr + (uint32_t)(32 * (jc + 1)), v1);
            next_res(j0 + 32);
            finish(v0, j0);
            if (RES) { res[0] = rn[0]; res[1] = rn[1]; }
            ptx::tmem_ld_wait();                                         // v1 = chunk jc + 1
            if (jc + 2 < cpg) ptx::tmem_ld16(taddr + (uint32_t)(32 * (jc + 2)), v0);
            else if (ti + 1 < MT) ptx::tmem_ld16(taddr + (uint32_t)p.N, v0);      // first chunk of the next tile
            next_res(j0 + 64);
            finish(v1, j0 + 32);
            if (RES) { res[0] = rn[0]; res[1] = rn[1]; }
        }
    }
}

// PAIR = the CTA-pair form for STREAMED weights (cluster of two, tcgen05 cta_group::2, M = 256): every MMA covers tile ti of
// BOTH CTAs' macro-tiles (each CTA's own image as the A operand) against a B operand of which each CTA holds HALF the output
// channels, so each SM streams half the weight bytes per MMA cycle - the per-SM L2 -> shared-memory stream (about 35 B/clk),
// not the tensor pipe, bounds the non-resident convs (a 32 KB piece feeds 2 x 4 MMAs of 129 cycles at N = 256).  The leader's
// elected thread issues every MMA and commits (multicast) to both CTAs' barriers; the peer's otherwise idle MMA warp forwards
// "my image / my weight half has landed" to the leader; both CTAs' epilogue warps return accumulator stages to the leader.
template <int MT, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) pconv_kernel(const PconvDev p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
    __shared__ uint64_t blk_full[2][kMaxBlocks], blk_empty[2][kMaxBlocks], tmem_full[2], tmem_empty[2], w_res_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[256];
    __shared__ uint32_t a_off_s[144];
    __shared__ uint32_t meta_s[kMetaSlots];   // per slot-in-site: bit 31 = real pixel, bits 20..21 = parity plane, low 20 = slot in the next level's site

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const uint32_t w_bytes = PAIR ? (uint32_t)p.N * 64u : (uint32_t)p.N * 128u;     // this CTA's part of a weight piece (4 k-steps)
    const uint32_t st_bytes = (uint32_t)p.wg * w_bytes;                              // one ring stage = wg pieces
    const uint32_t img_bytes = (uint32_t)p.nplanes * (uint32_t)(p.C / 8) * (uint32_t)p.n_in * 16u;
    const uint32_t lbo_img = (uint32_t)p.n_in * 16u;
    const uint32_t lbo_w = PAIR ? (uint32_t)p.N * 8u : (uint32_t)p.N * 16u;         // k-group pitch of the B image in shared memory
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t w_region = p.w_resident ? (uint32_t)p.nchunks * w_bytes : (uint32_t)p.w_stages * st_bytes;
    const uint32_t img_base = smem_base + w_region;

    if (tid == 0) {
        // pair, leader: a "full" barrier completes on its own bytes AND the peer's forwarded arrival (one wait per stage / block:
        // every barrier test costs the issuing thread ~90 cycles of tensor-pipe idle time)
        const uint32_t full_count = (PAIR && rank == 0) ? 2u : 1u;
        for (int s = 0; s < kMaxWStages; ++s) { ptx::mbar_init(&w_full[s], full_count); ptx::mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            for (int bk = 0; bk < kMaxBlocks; ++bk) {
                ptx::mbar_init(&blk_full[s][bk], full_count);
                ptx::mbar_init(&blk_empty[s][bk], 1);
            }
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], PAIR ? 16 : 256);     // pair: one arrival per epilogue warp of both CTAs, at the leader
        }
        ptx::mbar_init(&w_res_bar, 1);
        ptx::fence_barrier_init();
    }
    if (!PAIR && warp == 8) ptx::tmem_alloc<512>(&tmem_base_smem);
    for (int i = tid; i < p.N; i += kThreads) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    for (int l = tid; l < p.S && l < kMetaSlots; l += kThreads) meta_s[l] = slot_meta(p, l);
    if (PAIR) {
        __syncthreads();                            // the pair allocation's shared-memory write is the only access between two barriers
        if (warp == 8) ptx::tmem_alloc_pair<512>(&tmem_base_smem);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (PAIR) ptx::cluster_sync_all();            // the peer's barriers are initialised before any multicast commit / remote arrival
    // macro-tiles of this CTA: mb + rank for mb = first, first + gridDim.x, ... (pair: both CTAs run the leader's trip count; a
    // peer tile past the end is computed on the last valid image and never stored)
    const int mb0 = (int)blockIdx.x - (int)rank;

    if (warp == 10) {
        // ===================================================== image loader (one thread): the blocks of every macro-tile of this
        // CTA, each gated only by its own "block free" barrier - with two buffers it runs a whole tile ahead, with one it refills
        // block b for the next tile while the MMAs of blocks b+1.. of this tile still run
        if (lane == 0) {
            int li = 0;
            for (int mb = mb0; mb < p.n_macro; mb += gridDim.x, ++li) {
                int macro = mb + (int)rank;
                if (PAIR && macro >= p.n_macro) macro = p.n_macro - 1;
                const int buf = li % p.img_bufs;
                const uint32_t ph = (uint32_t)(li / p.img_bufs) & 1u;
                const long long slot0 = (long long)p.G + 128LL * MT * macro - p.halo_lo;
                const uint32_t dst = img_base + (uint32_t)buf * img_bytes;
                const int kg_per_plane = p.C / 8;
                for (int bk = 0; bk < p.nb; ++bk) {
                    ptx::mbar_wait(&blk_empty[buf][bk], ph ^ 1u);
                    ptx::mbar_arrive_expect_tx(&blk_full[buf][bk], (uint32_t)p.blk_kg * lbo_img);
                    for (int kg = 0; kg < p.blk_kg; ++kg) {
                        const int flat = bk * p.blk_kg + kg;                 // k-group index over [plane][C/8]
                        const int pl = flat / kg_per_plane, kgp = flat - pl * kg_per_plane;
                        ptx::bulk_g2s(dst + (uint32_t)flat * lbo_img,
                                      (const char *)(p.in + (size_t)pl * p.plane_elems) + ((size_t)kgp * p.P + slot0) * 16, lbo_img,
                                      &blk_full[buf][bk]);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ===================================================== weight loader (one thread): resident image once, or the pieces of
        // every macro-tile in schedule order through the ring
        if (lane == 0) {
            if (p.w_resident) {
                ptx::mbar_arrive_expect_tx(&w_res_bar, (uint32_t)p.nchunks * w_bytes);
                for (int c = 0; c < p.nchunks; ++c)
                    ptx::bulk_g2s(smem_base + (uint32_t)c * w_bytes, (const char *)p.w_img + (size_t)c * w_bytes, w_bytes, &w_res_bar);
            } else if (!p.dbg_skip_w) {
                int w_stage = 0;
                uint32_t w_phase = 0;
                int issued = 0;
                for (int mb = mb0; mb < p.n_macro; mb += gridDim.x) {
                    for (int j0 = 0; j0 < p.nchunks; j0 += p.wg, ++issued) {
                        const int n = (p.nchunks - j0 < p.wg) ? p.nchunks - j0 : p.wg;       // pieces in this stage
                        ptx::mbar_wait(&w_empty[w_stage], w_phase ^ 1u);
                        if (p.trace != nullptr && blockIdx.x == 0 && issued == p.w_stages) p.trace[65] = clock64();   // first refill released
                        ptx::mbar_arrive_expect_tx(&w_full[w_stage], (uint32_t)n * w_bytes);
                        for (int i = 0; i < n; ++i) {
                            const uint32_t c = p.sched[j0 + i] & 0xFFu;
                            // pair: this CTA's half of the piece's output channels (rows [rank*N/2, (rank+1)*N/2) of each k-group), packed
                            // contiguously on the host - eight 1-2 KB copies per piece instead of one cost 9-13 k cycles per macro-tile
                            ptx::bulk_g2s(smem_base + (uint32_t)w_stage * st_bytes + (uint32_t)i * w_bytes,
                                          (const char *)p.w_img + ((size_t)c * (PAIR ? 2 : 1) + rank) * w_bytes, w_bytes, &w_full[w_stage]);
                        }
                        if (++w_stage == p.w_stages) { w_stage = 0; w_phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 8) {
        // ===================================================== MMA issuer
        const uint32_t idesc = ptx::umma_idesc_f16(PAIR ? 256 : 128, (uint32_t)p.N);
        auto mma = [&](uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
            if (PAIR) ptx::umma_f16_pair(d, a_desc, b_desc, idesc, accumulate);
            else ptx::umma_f16(d, a_desc, b_desc, idesc, accumulate);
        };
        auto commit = [&](uint64_t *bar) {          // pair: the arrival is delivered to the barrier at this offset in BOTH CTAs
            if (PAIR) ptx::umma_commit_pair(bar);
            else ptx::umma_commit(bar);
        };
        // C = 16 (conv1) only: per-k-step A-view offsets (descriptor start-address units of 16 B): one tap per k-step -> plane,
        // slot shift.  (Every other conv takes its per-piece offsets from the host-built schedule.)
        for (int q = lane; q < p.nksteps && p.kpt_shift < 2; q += 32) {
            const int tap = q >> p.kpt_shift, kk = q - (tap << p.kpt_shift);
            const int dh = tap / 3, dw = tap - dh * 3;
            const int plane = p.nplanes == 4 ? ((dh & 1) * 2 + (dw & 1)) : 0;
            const int shift = p.nplanes == 4 ? (dh >> 1) * p.Wp + (dw >> 1) : dh * p.Wp + dw;   // slots into the chunk
            a_off_s[q] = ((uint32_t)(plane * (p.C / 8) + 2 * kk) * lbo_img + (uint32_t)shift * 16u) >> 4;
        }
        __syncwarp();
        const uint64_t a_desc0 = ptx::umma_desc_nosw(0, lbo_img, 128u), w_desc0 = ptx::umma_desc_nosw(0, lbo_w, 128u);
        const uint32_t a_desc_lo = (uint32_t)a_desc0, a_desc_hi = (uint32_t)(a_desc0 >> 32);
        const uint32_t w_desc_lo = (uint32_t)w_desc0, w_desc_hi = (uint32_t)(w_desc0 >> 32);
        const uint32_t b_step = (2u * lbo_w) >> 4;
        const uint32_t kstep_a = (2u * lbo_img) >> 4;
        int li = 0, w_stage = 0, w_sub = 0;        // ring stage and piece-in-stage of the next weight piece
        uint32_t w_phase = 0;
        // ONE elected thread runs the whole loop (barrier waits included): no per-chunk elect / reconvergence / warp sync.  The
        // loop is ROLLED over the schedule (a piece = 4 k-steps x MT MMAs; its weight-chunk index, image block and A-view offset
        // come from the host-built table): the tap-unrolled form was ~9 KB of code and cost every CTA 4-5 k cycles of
        // instruction-cache misses on its first macro-tile; the few integer instructions per piece hide behind the previous MMA.
        const bool elected = ptx::elect_one();
        if (elected && rank == 0) {
        if (p.w_resident) ptx::mbar_wait(&w_res_bar, 0);
        for (int mb = mb0; mb < p.n_macro; mb += gridDim.x, ++li) {
            const int buf = li % p.img_bufs;
            const uint32_t iph = (uint32_t)(li / p.img_bufs) & 1u;
            const int acc = li % p.acc_stages;
            const uint32_t aph = (uint32_t)(li / p.acc_stages) & 1u;
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && li < 8;
            if (tr) p.trace[li * 8 + 0] = clock64();
            if (PAIR) ptx::mbar_wait_cluster(&tmem_empty[acc], aph ^ 1u);
            else ptx::mbar_wait(&tmem_empty[acc], aph ^ 1u);
            ptx::tc_fence_after();
            if (tr) p.trace[li * 8 + 2] = clock64();
            const uint32_t img = img_base + (uint32_t)buf * img_bytes;
            const uint32_t d0 = tmem_base + (uint32_t)(acc * MT * p.N);
            const uint32_t a_lo = a_desc_lo + (img >> 4);
#pragma unroll 1
            for (int j = 0; j < p.nchunks; ++j) {
                const uint32_t e = p.sched[j];
                const uint32_t bk = (e >> 8) & 0xFFu;
                if (e & kSchedFirst) {                    // first piece that reads image block bk: wait until it has landed (both CTAs)
                    if (PAIR) ptx::mbar_wait_cluster(&blk_full[buf][bk], iph);
                    else ptx::mbar_wait(&blk_full[buf][bk], iph);
                    if (tr && j == 0) p.trace[li * 8 + 1] = clock64();
                }
                uint32_t w_addr;
                if (p.w_resident) {
                    w_addr = smem_base + (e & 0xFFu) * w_bytes;
                } else if (p.dbg_skip_w) {
                    w_addr = smem_base + (uint32_t)w_stage * st_bytes + (uint32_t)w_sub * w_bytes;
                } else {
                    if (w_sub == 0) {                     // first piece of a ring stage: wait for the stage (pair: both halves)
                        const bool trw = tr && li == 0 && j == p.w_stages * p.wg;      // the first refilled ring stage
                        if (trw) p.trace[69] = clock64();
                        if (PAIR) ptx::mbar_wait_cluster(&w_full[w_stage], w_phase);
                        else ptx::mbar_wait(&w_full[w_stage], w_phase);
                        if (trw) p.trace[67] = clock64();
                    }
                    w_addr = smem_base + (uint32_t)w_stage * st_bytes + (uint32_t)w_sub * w_bytes;
                }
                if ((e & kSchedFirst) || (!p.w_resident && w_sub == 0)) ptx::tc_fence_after();      // only after a barrier wait
                const uint32_t b_lo = w_desc_lo + (w_addr >> 4);
                if (p.kpt_shift >= 2) {
                    // C >= 64: the piece's 4 k-steps lie inside one tap: consecutive k-group pairs of one shifted view
                    const uint32_t a_c = a_lo + p.a_off[j];
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t b_desc = ((uint64_t)w_desc_hi << 32) | (uint64_t)(b_lo + (uint32_t)ks * b_step);
#pragma unroll
                        for (int ti = 0; ti < MT; ++ti)
                            mma(d0 + (uint32_t)(ti * p.N), ((uint64_t)a_desc_hi << 32) | (uint64_t)(a_c + (uint32_t)ks * kstep_a + (uint32_t)(ti * 128)),
                                b_desc, (j > 0 || ks > 0) ? 1u : 0u);
                    }
                } else {
                    // C = 16 (conv1): one k-step per tap, A-view offsets from the table
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const int q = 4 * j + ks;
                        if (q < p.nksteps) {
                            const uint32_t ao = a_lo + a_off_s[q];
                            const uint64_t b_desc = ((uint64_t)w_desc_hi << 32) | (uint64_t)(b_lo + (uint32_t)ks * b_step);
#pragma unroll
                            for (int ti = 0; ti < MT; ++ti)
                                mma(d0 + (uint32_t)(ti * p.N), ((uint64_t)a_desc_hi << 32) | (uint64_t)(ao + (uint32_t)(ti * 128)), b_desc, q > 0 ? 1u : 0u);
                        }
                    }
                }
                if (!p.w_resident) {
                    if (++w_sub == p.wg || j + 1 == p.nchunks) {              // last piece of the stage: release it when its MMAs finish
                        w_sub = 0;
                        if (!p.dbg_skip_w) commit(&w_empty[w_stage]);
                        if (tr && li == 0 && j + 1 == p.wg) p.trace[64] = clock64();     // stage 0 issued and its release queued
                        if (++w_stage == p.w_stages) { w_stage = 0; w_phase ^= 1u; }
                    }
                }
                if (e & kSchedLast) commit(&blk_empty[buf][bk]);      // every MMA that reads block bk has been issued: refill when they finish
            }
            commit(&tmem_full[acc]);
            if (tr) p.trace[li * 8 + 3] = clock64();
        }
        } else if (PAIR && elected) {
            // peer CTA: forward "my image block / my half of weight piece c has landed" to the leader, in the order the leader
            // consumes them.  Producer, forwarder and consumer advance in lockstep: a block / ring stage is refilled only after the
            // leader's commit for its previous use.  The arrivals are RELAXED: the bytes were written by the TMA engine (complete
            // before the local barrier flips) and are read by the tensor core straight from this CTA's shared memory - nothing of
            // this thread's needs publishing, and a release at cluster scope costs about 1000 cycles per arrival (measured: 18
            // forwards per macro-tile made the pair form's MMA phase 17-20 k cycles instead of 8.6 k).
            for (int mb = mb0; mb < p.n_macro; mb += gridDim.x, ++li) {
                const int buf = li % p.img_bufs;
                const uint32_t iph = (uint32_t)(li / p.img_bufs) & 1u;
                for (int j = 0; j < p.nchunks; ++j) {
                    const uint32_t e = p.sched[j];
                    if (e & kSchedFirst) {
                        const uint32_t bk = (e >> 8) & 0xFFu;
                        ptx::mbar_wait(&blk_full[buf][bk], iph);
                        ptx::mbar_arrive_cluster_relaxed(&blk_full[buf][bk], 0);       // the second arrival of the leader's barrier
                    }
                    if (!p.w_resident && !p.dbg_skip_w) {
                        if (w_sub == 0) {
                            ptx::mbar_wait(&w_full[w_stage], w_phase);
                            ptx::mbar_arrive_cluster_relaxed(&w_full[w_stage], 0);
                        }
                        if (++w_sub == p.wg || j + 1 == p.nchunks) {
                            w_sub = 0;
                            if (++w_stage == p.w_stages) { w_stage = 0; w_phase ^= 1u; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp < 8) {
        // ===================================================== epilogue: two groups of four warps (one TMEM lane quadrant per
        // warp), group e takes the 16-column chunks with (chunk & 1) == e; residuals are fetched one chunk ahead
        const int q = warp & 3;
        const int eg = warp >> 2;
        const int r = q * 32 + lane;
        int li = 0;
        for (int mb = mb0; mb < p.n_macro; mb += gridDim.x, ++li) {
            const int macro = mb + (int)rank;         // pair: the peer's tile past the end has no slot < T, nothing is stored
            const int acc = li % p.acc_stages;
            const uint32_t aph = (uint32_t)(li / p.acc_stages) & 1u;
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 128 && li < 8;
            if (tr) p.trace[li * 8 + 4] = clock64();
            ptx::mbar_wait(&tmem_full[acc], aph);
            ptx::tc_fence_after();
            if (tr) p.trace[li * 8 + 5] = clock64();
            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MT * p.N) + (uint32_t)(16 * eg);
            const long long g0 = (128LL * MT) * macro + r;
            if (p.residual) {
                if (p.out_parity) epilogue_tiles<MT, true, true>(p, tbase, g0, eg, bias_s, meta_s);
                else epilogue_tiles<MT, true, false>(p, tbase, g0, eg, bias_s, meta_s);
            } else {
                if (p.out_parity) epilogue_tiles<MT, false, true>(p, tbase, g0, eg, bias_s, meta_s);
                else epilogue_tiles<MT, false, false>(p, tbase, g0, eg, bias_s, meta_s);
            }
            ptx::tc_fence_before();
            if (PAIR) {
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster_relaxed(&tmem_empty[acc], 0);
            } else {
                ptx::mbar_arrive(&tmem_empty[acc]);
            }
            if (tr) p.trace[li * 8 + 6] = clock64();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (PAIR) ptx::cluster_sync_all();           // both CTAs are done with TMEM, each other's barriers and operand halves
    if (warp == 8) {
        ptx::tc_fence_after();
        if (PAIR) ptx::tmem_dealloc_pair<512>(tmem_base);
        else ptx::tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

static int launch_pconv2(void (*kern)(const PconvDev), int grid, int cluster, size_t smem, const PconvDev &p, cudaStream_t s) {
    C3B_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cluster > 1 ? 1 : 0;
    C3B_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_pconv2(const c3b_model *m, const PconvArgs &a, cudaStream_t s) {
    const PlanarGeom &g = a.geom;
    const bool c_ok = a.c == 16 || a.c == 64 || a.c == 128 || a.c == 256;
    if (!c_ok || a.n % 16 || a.n > 256 || a.n < 16) { c3b_set_error("pconv: unsupported channels %d -> %d", a.c, a.n); return 1; }
    if (g.p * 2 * a.n >= (1LL << 32) || (a.out_parity && a.next.p * 2 * a.n >= (1LL << 32)) || g.t >= (1LL << 31) || (a.out_parity && a.next.s >= (1 << 20))) { c3b_set_error("pconv: batch too large for one launch"); return 1; }
    PconvDev p = {};
    p.in = a.in; p.w_img = a.w.w_img; p.bias = a.w.bias; p.residual = a.residual; p.out = a.out;
    p.C = a.c; p.N = a.n;
    p.H = g.h; p.W = g.w; p.Wp = g.wp; p.S = g.s; p.G = g.g; p.T = g.t; p.P = g.p;
    p.relu = a.relu;
    p.trace = a.trace;
    static const int skip_w = getenv("C3B_PCONV_SKIPW") ? atoi(getenv("C3B_PCONV_SKIPW")) : 0;
    p.dbg_skip_w = skip_w;
    p.nksteps = 9 * a.c / 16;
    p.nchunks = (p.nksteps + 3) / 4;
    p.cpt = a.c / 64;
    p.kpt_shift = a.c == 16 ? 0 : a.c == 64 ? 2 : a.c == 128 ? 3 : 4;
    p.nplanes = a.stride2 ? 4 : 1;
    p.plane_elems = (long long)(a.c / 8) * g.p * 8;
    p.halo_lo = a.stride2 ? 0 : g.wp + 1;
    p.out_parity = a.out_parity;
    if (a.out_parity) {
        p.out_plane_elems = (long long)(a.n / 8) * a.next.p * 8;
        p.nS = a.next.s; p.nWp = a.next.wp; p.nG = a.next.g; p.nP = a.next.p;
    }
    if (p.nchunks != a.w.nchunks) { c3b_set_error("pconv: weight image has %d chunks, expected %d", a.w.nchunks, p.nchunks); return 1; }
    const size_t budget = 220 * 1024 - 256;      // 227 KB per CTA minus the static barriers, bias and slot tables
    const size_t w_bytes = (size_t)a.n * 128;
    const size_t w_all = (size_t)p.nchunks * w_bytes;
    // Configuration search over MT in {4,2,1}: resident weights when they fit (then small MT only costs halo re-reads and
    // balances the tile count over the SMs); streamed weights want MT >= 2 (every piece feeds MT accumulators) and a deep
    // ring, so the image is single-buffered there.  Cost model = rounds of macro-tiles x MMAs per macro-tile.
    long long best_cost = -1;
    int pair = 0;
    // streamed weights run on the CTA pair (each SM streams half of every piece); C3B_PCONV_PAIR=0 keeps one CTA per tile (A/B runs)
    static const bool allow_pair = !(getenv("C3B_PCONV_PAIR") && atoi(getenv("C3B_PCONV_PAIR")) == 0);
    static const int force_wg = getenv("C3B_PCONV_WG") ? atoi(getenv("C3B_PCONV_WG")) : 0;                 // tuning sweeps only
    static const int force_mt = getenv("C3B_PCONV_MT") ? atoi(getenv("C3B_PCONV_MT")) : 0;   // tuning sweeps only
    for (int mt = ((a.n <= 64 || force_mt == 4) ? 4 : 2); mt >= 1; mt >>= 1) {
        if (mt * a.n > 512) continue;
        const int n_in = a.stride2 ? (128 * mt + g.wp + 1 + 7) / 8 * 8 : 128 * mt + 2 * (g.wp + 1);
        const size_t img_bytes = (size_t)p.nplanes * (a.c / 8) * n_in * 16;
        int resident = 0, bufs = 0, stages = 0, pr = 0, wgc = 1;
        const long long n_macro = (g.t + 128LL * mt - 1) / (128LL * mt);
        if (w_all + 2 * img_bytes <= budget) { resident = 1; bufs = 2; }
        else if (w_all + img_bytes <= budget) { resident = 1; bufs = 1; }
        else if (mt >= 2 || a.n <= 64 || a.stride2) {
            pr = (allow_pair && n_macro >= 2 && a.w.w_img_pair != nullptr) ? 1 : 0;
            // pieces per ring stage: every stage costs the MMA-issuing thread a barrier test and a commit (~250-350 cycles that
            // the tensor pipe idles: it runs at most a few MMAs ahead of the issue), so a stage should hold >= ~700 cycles of MMAs
            const size_t piece = pr ? w_bytes / 2 : w_bytes;
            const int mma_cyc = a.n <= 64 ? 48 : a.n <= 128 ? 62 : 125;
            wgc = 1;
            if (force_wg > 0) wgc = force_wg;
            else if (4 * mt * mma_cyc < 700 && img_bytes + 4 * 2 * piece <= budget) wgc = 2;
            const size_t st_bytes = (size_t)wgc * piece;                 // one ring stage
            if (img_bytes + 2 * st_bytes <= budget) { bufs = 1; stages = (int)((budget - img_bytes) / st_bytes); }
            // pair: the refill loop of a ring stage crosses the pair twice (commit -> peer loader, peer forwarder -> leader) and
            // takes 5-6 k cycles (measured), so the ring has to cover that: a second image buffer only if >= 12 stages remain
            // (one buffer is enough since the image blocks are recycled one by one)
            const size_t min_stages2 = pr ? 12 : 6;
            if (2 * img_bytes + min_stages2 * st_bytes <= budget) { bufs = 2; stages = (int)((budget - 2 * img_bytes) / st_bytes); }
            if (!bufs) continue;
        } else continue;
        const long long rounds = (n_macro + m->sm_count - 1) / m->sm_count;
        // + a fixed per-macro-tile cost: measured, MT = 1 is 1.2-1.5x slower than MT = 2/4 on every level even with resident
        // weights (each macro-tile pays an image-chunk round trip that only several tiles of MMAs hide)
        long long cost = rounds * mt * 100 + rounds * 60 + (bufs == 1 ? rounds * 12 : 0) + (resident ? 0 : 5);
        if (force_mt > 0) cost = (mt == force_mt) ? 1 : 1000000 + cost;
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            p.MT = mt; p.n_in = n_in; p.w_resident = resident; p.img_bufs = bufs; p.w_stages = stages; pair = pr; p.wg = wgc;
            p.acc_stages = (mt * a.n * 2 <= 512) ? 2 : 1;
        }
    }
    if (best_cost < 0) { c3b_set_error("pconv: feature map does not fit shared memory"); return 1; }
    static const bool dbg = getenv("C3B_DEBUG_PCONV") != nullptr;
    if (dbg)
        fprintf(stderr, "[pconv] C=%d N=%d stride2=%d T=%lld: MT=%d resident=%d pair=%d img_bufs=%d w_stages=%d x %d pieces cost=%lld\n", a.c, a.n, a.stride2,
                (long long)g.t, p.MT, p.w_resident, pair, p.img_bufs, p.w_stages, p.wg, best_cost);
    if (p.w_stages > kMaxWStages) p.w_stages = kMaxWStages;
    // ---- the piece schedule (see PconvDev::sched)
    if (p.nchunks > kMaxChunks) { c3b_set_error("pconv: %d weight pieces per tile", p.nchunks); return 1; }
    if (a.c == 16) {
        p.nb = 1;
        p.blk_kg = p.nplanes * 2;
        for (int j = 0; j < p.nchunks; ++j) {
            p.sched[j] = (uint32_t)j | (j == 0 ? kSchedFirst : 0u) | (j + 1 == p.nchunks ? kSchedLast : 0u);
            p.a_off[j] = 0;
        }
    } else {
        static const int plane_taps[4][4] = {{0, 2, 6, 8}, {1, 7, -1, -1}, {3, 5, -1, -1}, {4, -1, -1, -1}};   // plane = (dh&1)*2 + (dw&1)
        static const int plane_ntaps[4] = {4, 2, 2, 1};
        p.nb = p.nplanes * p.cpt;
        p.blk_kg = 8;
        if (p.nb > kMaxBlocks) { c3b_set_error("pconv: %d image blocks", p.nb); return 1; }
        int j = 0;
        static const int tap_outer = getenv("C3B_PCONV_TAPOUTER") ? atoi(getenv("C3B_PCONV_TAPOUTER")) : 0;
        if (tap_outer && !a.stride2) {
            // stride 1, experiment: taps outer / channel groups inner = the weight pieces in memory order (blocks are all in use
            // from tap 0 to tap 8: no early recycling)
            for (int tap = 0; tap < 9; ++tap)
                for (int kc = 0; kc < p.cpt; ++kc, ++j) {
                    const int dh = tap / 3, dw = tap % 3;
                    p.sched[j] = (uint32_t)(tap * p.cpt + kc) | (uint32_t)kc << 8 | (tap == 0 ? kSchedFirst : 0u) | (tap == 8 ? kSchedLast : 0u);
                    p.a_off[j] = (uint32_t)(8 * kc * p.n_in + dh * g.wp + dw);
                }
        } else
        for (int pl = 0; pl < p.nplanes; ++pl)
            for (int kc = 0; kc < p.cpt; ++kc) {
                const int bk = pl * p.cpt + kc;
                const int nt = a.stride2 ? plane_ntaps[pl] : 9;
                for (int i = 0; i < nt; ++i, ++j) {
                    const int tap = a.stride2 ? plane_taps[pl][i] : i;
                    const int dh = tap / 3, dw = tap % 3;
                    const int shift = a.stride2 ? (dh >> 1) * g.wp + (dw >> 1) : dh * g.wp + dw;       // slots into the image
                    p.sched[j] = (uint32_t)(tap * p.cpt + kc) | (uint32_t)bk << 8 | (i == 0 ? kSchedFirst : 0u) | (i + 1 == nt ? kSchedLast : 0u);
                    p.a_off[j] = (uint32_t)((pl * (a.c / 8) + 8 * kc) * p.n_in + shift);
                }
            }
        if (j != p.nchunks) { c3b_set_error("pconv: schedule has %d pieces, expected %d", j, p.nchunks); return 1; }
    }
    const long long per_macro = 128LL * p.MT;
    p.n_macro = (int)((g.t + per_macro - 1) / per_macro);
    if ((long long)p.n_macro * per_macro + g.g > g.p - g.g + per_macro) { /* plane pitch covers the rounded-up slot range by construction */ }
    const size_t img_bytes = (size_t)p.nplanes * (a.c / 8) * p.n_in * 16;
    const size_t st_bytes = (size_t)p.wg * (pair ? w_bytes / 2 : w_bytes);
    const size_t smem = (p.w_resident ? (size_t)p.nchunks * w_bytes : (size_t)p.w_stages * st_bytes) + p.img_bufs * img_bytes + 256;
    int grid = p.n_macro < m->sm_count ? p.n_macro : m->sm_count;
    if (pair) grid = (grid + 1) & ~1;            // whole pairs (sm_count is even; an odd tile count gets one never-stored peer tile)
    if (pair && grid > m->sm_count) grid = m->sm_count & ~1;
    const_cast<c3b_model *>(m)->launches++;
    c3b_note_grid(grid);
    if (pair) p.w_img = a.w.w_img_pair;
    switch (p.MT * 2 + pair) {
        case 2: return launch_pconv2(pconv_kernel<1, false>, grid, 1, smem, p, s);
        case 3: return launch_pconv2(pconv_kernel<1, true>, grid, 2, smem, p, s);
        case 4: return launch_pconv2(pconv_kernel<2, false>, grid, 1, smem, p, s);
        case 5: return launch_pconv2(pconv_kernel<2, true>, grid, 2, smem, p, s);
        case 8: return launch_pconv2(pconv_kernel<4, false>, grid, 1, smem, p, s);
        case 9: return launch_pconv2(pconv_kernel<4, true>, grid, 2, smem, p, s);
        default: c3b_set_error("pconv: unsupported MT %d", p.MT); return 1;
    }
}
