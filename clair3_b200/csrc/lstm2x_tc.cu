// Both recurrent layers of Clair3_P (nn.LSTM(18, 128) and nn.LSTM(256, 160), bidirectional, clair3/model.py:96-107,132-133; torch
// semantics: gate rows i,f,g,o, h0 = c0 = 0, the reverse direction walks t = 32..0) on a CTA PAIR (tcgen05 cta_group::2).
// The text below describes LSTM2; LSTM1 is the same kernel with H = 128 (four phases), K = 48 + 128 (the 48 hi | 1 | lo input
// columns of x_t are bulk-copied into the head of the A operand buffer every step, so the x projection and the bias ride in the
// same MMAs) and no pre-gates.
//
// Why a second kernel.  The round-1 kernel (lstm_tc.cu) puts the gate rows on the TMEM lanes, so its MMAs are 128 x 32 x 16:
// TMEM (512 columns) holds the ten accumulator blocks of two ping-pong sub-tiles only up to 32 sites each, and a
// 128 x N x 16 MMA costs >= 48 cycles whatever N is - a third of the tensor rate at N = 32 - while every epilogue thread
// writes its h values as scattered 2-byte shared-memory stores.  Here the SITES sit on the lanes:
//   * one MMA is 256 x 128 x 16 over the pair (each CTA: its own 128 sites as A rows, half of the 128 gate columns of the
//     phase as B rows), i.e. 128 x 128 x 16 per SM in 64 cycles = the full tensor rate;
//   * W_hh (205 KB fp16) is split over the two CTAs' shared memory as B operand halves (102 KB each), resident for all 33 steps;
//   * the 640 gate columns of a step are issued as five PHASES of 32 hidden units x 4 gates = 128 columns through a ring of
//     three TMEM accumulator stages, so the sigma/tanh/cell epilogue of phase p runs while the MMAs of phases p+1, p+2 run;
//   * an epilogue thread owns one site and, per phase, 8 consecutive hidden units: its h values are one 16-byte shared-memory
//     store (the A operand of the next step, k-group-planar) and one 16-byte global store (coalesced over the warp); the gate
//     activations run on packed fp16 pairs (tanh.approx.f16x2: 2.5 MUFU ops per cell instead of 5), cell state in fp32;
//   * the two CTAs never exchange data (each keeps its own sites' h in its own shared memory) - only barriers cross the pair.
//
// Layouts.  B image w2x[dir][rank][phase 5][20 k-groups][64 rows][8]: rank 0 holds gates (i, f), rank 1 gates (g, o) of units
// 32 p .. 32 p + 31 (sigmoid rows pre-halved: sigma(x) = 0.5 tanh(x/2) + 0.5).  Accumulator stage: columns [0,32) i, [32,64) f,
// [64,96) g, [96,128) o.  Pre-gates (W_ih h1 + b, written by proj_tc.cu in its second layout):
// pg2[dir][t][128-site tile][80 column groups][128 sites][8] fp16, column group = 16 phase + 4 gate + (unit % 32) / 8.
// Output h2: tile-major k-group-planar [bp/128][1320][128][8], k = t*320 + dir*160 + unit (the flatten order of clair3/model.py:135).
//
// Roles (576 threads per CTA): warps 0-15 epilogue (warpgroup w = 8-unit group of each phase, warp & 3 = TMEM lane quadrant),
// warp 16: one elected thread of the LEADER CTA issues every MMA, warp 17: weights loader + TMEM allocation.
// Barriers: acc_full[3] is multicast to both CTAs by tcgen05.commit; acc_empty[3] lives in the leader (a TMEM stage has been read
// by both CTAs: 32 relaxed warp arrivals); "h_t operands complete" is collected per CTA on a LOCAL barrier (a_ready[2], 16 warp
// arrivals with cheap CTA-scope release) and the peer's completion is forwarded to the leader (a_peer[2]) by one otherwise idle
// thread - a cluster-scope release by every epilogue warp would wait for that warp's outstanding global loads / stores every
// step (measured: ~3 k cycles of the ~11 k cycle step).
#include <cstdlib>

#include "c3b_internal.h"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 576;
constexpr int kEpiWarps = 16;
constexpr int kStages = 3;

template <bool L2>
struct Shape {
    static constexpr int H = L2 ? 160 : 128;
    static constexpr int KX = L2 ? 0 : C3B_X1_COLS;                 // x columns at the head of the A operand (LSTM1)
    static constexpr int KG = (KX + H) / 8;                           // k-groups of the A / B operands: 20 | 22
    static constexpr int kPhases = H / 32;                            // 5 | 4
    static constexpr uint32_t kWPhaseBytes = KG * 64 * 16;            // one phase of one rank: [KG][64 rows][8]
    static constexpr uint32_t kWBytes = kPhases * kWPhaseBytes;       // 102,400 | 90,112
    static constexpr uint32_t kABytes = KG * 128 * 16;                // one operand buffer: [KG][128 sites][8]
    static constexpr uint32_t kXBytes = (KX / 8) * 128 * 16;          // its x part (LSTM1): 12,288
};

struct Lstm2xDev {
    const op_t *w_img;      // [dir][rank][phase][KG][64][8]
    const __half *pg;       // LSTM2: pg2
    const op_t *xs;         // LSTM1: xs2 [33][bp/128][6 k-groups][128 sites][8]
    op_t *hout;             // tile-major: LSTM2 h2 [bp/128][1320][128][8]; LSTM1 h1 [33*bp/128][32][128][8] (row = t*bp + site)
    int bp;                 // padded batch (multiple of 256)
    long long *trace;
};

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t tanh_h2u(uint32_t x) {
    uint32_t r;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}
__device__ __forceinline__ uint32_t hadd2u(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t hmul2u(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t hfma2u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// packed pre-activation pair: accumulator (+ fp16 pre-gate pair for LSTM2)
template <bool L2>
__device__ __forceinline__ uint32_t preact(float lo, float hi, uint32_t pg) {
    return L2 ? hadd2u(pack_h2(lo, hi), pg) : pack_h2(lo, hi);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    const __half2 h = *reinterpret_cast<const __half2 *>(&v);
    return __half22float2(h);
}

template <bool L2>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) lstm_pair_kernel(const Lstm2xDev p) {
    using S = Shape<L2>;
    constexpr int kPhases = S::kPhases;
    constexpr uint32_t kWPhaseBytes = S::kWPhaseBytes, kWBytes = S::kWBytes, kABytes = S::kABytes;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t w_bar, a_ready[2], a_peer[2], acc_full[kStages], acc_empty[kStages], x_full[2], x_ready[2], x_free[2];
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const int dir = blockIdx.y;
    const int tile128 = (int)blockIdx.x;                   // cluster = two consecutive 128-site tiles: blockIdx.x = 2*pair + rank
    const uint32_t w_addr = ptx::smem_u32(smem);
    const uint32_t a_addr = w_addr + kWBytes;
    uint8_t *a_smem = smem + kWBytes;

    if (tid == 0) {
        ptx::mbar_init(&w_bar, 1);
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&a_ready[b], kEpiWarps);     // this CTA's 16 epilogue warps have written their h_t values
            ptx::mbar_init(&a_peer[b], 1);              // leader: the peer CTA's a_ready has completed
            ptx::mbar_init(&x_full[b], 1);        // this CTA's x_t bulk copy has landed (LSTM1)
            ptx::mbar_init(&x_ready[b], 2);       // leader: both CTAs' x_t have landed
            ptx::mbar_init(&x_free[b], 1);        // every MMA that read operand buffer b has completed (multicast commit, once per step)
        }
        for (int s = 0; s < kStages; ++s) {
            ptx::mbar_init(&acc_full[s], 1);
            ptx::mbar_init(&acc_empty[s], 2 * kEpiWarps);
        }
        ptx::fence_barrier_init();
    }
    // h_{-1} = 0 (buffer 0); the h part of buffer 1 is completely written by the epilogue of step 0 before step 1 reads it, the
    // x parts (LSTM1) are bulk-copied per step
    for (uint32_t i = tid * 16; i < kABytes; i += kThreads * 16) *reinterpret_cast<uint4 *>(a_smem + i) = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async_smem();
    __syncthreads();
    if (warp == 17) ptx::tmem_alloc_pair<512>(&tmem_base_smem);     // after the barrier: the allocation's shared-memory write is the
    ptx::tc_fence_before();                                         // only access between the two barriers (racecheck-clean)
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 17 && lane == 0) {
        ptx::mbar_arrive_expect_tx(&w_bar, kWBytes);
        const char *src = reinterpret_cast<const char *>(p.w_img) + ((size_t)dir * 2 + rank) * kWBytes;
#pragma unroll
        for (int ph = 0; ph < kPhases; ++ph) ptx::bulk_g2s(w_addr + ph * kWPhaseBytes, src + (size_t)ph * kWPhaseBytes, kWPhaseBytes, &w_bar);
    }
    ptx::mbar_wait(&w_bar, 0);                 // every thread: this CTA's weights have landed ...
    ptx::cluster_sync_all();                   // ... and so have the peer's; its barriers are initialised, its h buffer zeroed

    if (warp == 16) {
        // ===================================================== MMA issuer (leader CTA only)
        if (rank == 0 && ptx::elect_one()) {
            const uint32_t idesc = ptx::umma_idesc_f16(256, 128);
            const uint64_t a_d0 = ptx::umma_desc_nosw(a_addr, 2048u, 128u), b_d0 = ptx::umma_desc_nosw(w_addr, 1024u, 128u);
            const uint32_t a_lo0 = (uint32_t)a_d0, a_hi = (uint32_t)(a_d0 >> 32);
            const uint32_t b_lo0 = (uint32_t)b_d0, b_hi = (uint32_t)(b_d0 >> 32);
            constexpr uint32_t a_kstep = (2u * 2048u) >> 4, a_bstep = kABytes >> 4, b_kstep = (2u * 1024u) >> 4, b_pstep = kWPhaseBytes >> 4;
            int st = 0;
            uint32_t use_par = 0;               // parity to wait on acc_empty: (use - 1) & 1; the first use of a stage needs no wait
            bool first_round = true;
            for (int step = 0; step < C3B_T; ++step) {
                const int ab = step & 1;
                if (step > 0) {
                    ptx::mbar_wait(&a_ready[ab], (uint32_t)((step - 1) >> 1) & 1u);
                    ptx::mbar_wait_cluster(&a_peer[ab], (uint32_t)((step - 1) >> 1) & 1u);
                }
                if (!L2) ptx::mbar_wait_cluster(&x_ready[ab], (uint32_t)(step >> 1) & 1u);
                ptx::tc_fence_after();
                const uint32_t a_lo = a_lo0 + (uint32_t)ab * a_bstep;
#pragma unroll 1
                for (int ph = 0; ph < kPhases; ++ph) {
                    if (!first_round) {
                        ptx::mbar_wait_cluster(&acc_empty[st], use_par);
                        ptx::tc_fence_after();
                    }
                    const uint32_t d = tmem_base + (uint32_t)(st * 128);
                    uint32_t al = a_lo, bl = b_lo0 + (uint32_t)ph * b_pstep;
#pragma unroll 1
                    for (int ks = 0; ks < S::KG / 2; ++ks) {
                        ptx::umma_f16_pair(d, ((uint64_t)a_hi << 32) | (uint64_t)al, ((uint64_t)b_hi << 32) | (uint64_t)bl, idesc, ks > 0 ? 1u : 0u);
                        al += a_kstep;
                        bl += b_kstep;
                    }
                    ptx::umma_commit_pair(&acc_full[st]);
                    if (++st == kStages) {
                        st = 0;
                        if (first_round) first_round = false; else use_par ^= 1u;
                    }
                }
                if (!L2) ptx::umma_commit_pair(&x_free[ab]);      // LSTM1: the loaders may refill this buffer's x columns (step + 2)
            }
        } else if (rank == 1 && ptx::elect_one()) {
            // peer CTA: forward "my epilogue warps have all written h_t" to the leader.  The arrival is RELAXED: the h values are in
            // this CTA's shared memory (written by the epilogue warps, made visible to the async proxy by their fence, published
            // to this thread by a_ready) and the tensor core reads them from there; a release at cluster scope costs this thread
            // several hundred cycles per step on the recurrence's critical path (measured: 261 -> 242 us per launch without it)
            for (int step = 1; step < C3B_T; ++step) {
                ptx::mbar_wait(&a_ready[step & 1], (uint32_t)((step - 1) >> 1) & 1u);
                ptx::mbar_arrive_cluster_relaxed(&a_peer[step & 1], 0);
            }
        }
        __syncwarp();
    } else if (warp == 17) {
        // ===================================================== LSTM1: x_t -> head of the operand buffer of step t (one 12 KB bulk
        // copy per step and CTA; xs2 is tile-major per time step), then tell the leader that this CTA's copy has landed
        if (!L2 && lane == 0) {
            const int ntile = p.bp >> 7;
            for (int step = 0; step < C3B_T; ++step) {
                const int t = dir ? (C3B_T - 1 - step) : step;
                const int ab = step & 1;
                // buffer ab was last read by the MMAs of step-2: their completion arrives on x_free[ab] exactly once per step, and
                // the issuer cannot start step `step` before this thread has delivered x_step - producer and consumer advance in
                // lockstep, so the parity wait is unambiguous (steps 0, 1: the buffers are free)
                if (step >= 2) ptx::mbar_wait(&x_free[ab], (uint32_t)((step - 2) >> 1) & 1u);
                ptx::mbar_arrive_expect_tx(&x_full[ab], S::kXBytes);
                ptx::bulk_g2s(a_addr + (uint32_t)ab * kABytes, (const char *)p.xs + ((size_t)t * ntile + tile128) * S::kXBytes, S::kXBytes, &x_full[ab]);
                ptx::mbar_wait(&x_full[ab], (uint32_t)(step >> 1) & 1u);
                ptx::mbar_arrive_cluster_relaxed(&x_ready[ab], 0);
            }
        }
        __syncwarp();
    } else if (warp < kEpiWarps) {
        // ===================================================== epilogue: thread = site, warpgroup = 8-unit group of every phase
        const int wg = warp >> 2, q = warp & 3;
        const int site = q * 32 + lane;
        const int ntile = p.bp >> 7;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(8 * wg);
        const size_t gsite = (size_t)tile128 * 128 + site;
        const uint32_t half2_half = 0x38003800u;      // (0.5, 0.5)
        float c[kPhases * 8];
#pragma unroll
        for (int i = 0; i < kPhases * 8; ++i) c[i] = 0.f;
        // pre-gate addressing: pg2[dir][t][tile][80][128][8]; this thread's site, first column group
        const size_t pg_tstride = (size_t)ntile * 80 * 128 * 8;
        const __half *pg_site = L2 ? p.pg + ((size_t)dir * C3B_T * ntile + tile128) * 80 * 128 * 8 + (size_t)site * 8 : nullptr;
        uint4 pgv[4];           // this thread's pre-gates of the phase about to be computed (refilled in place one phase ahead)
#pragma unroll
        for (int g = 0; g < 4; ++g)
            pgv[g] = L2 ? *reinterpret_cast<const uint4 *>(pg_site + (size_t)(dir ? C3B_T - 1 : 0) * pg_tstride + (size_t)(4 * g + wg) * 128 * 8)
                        : make_uint4(0, 0, 0, 0);
        int st = 0;
        uint32_t full_par = 0;
        const bool tr = p.trace != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
        for (int step = 0; step < C3B_T; ++step) {
            const int t = dir ? (C3B_T - 1 - step) : step;
            uint8_t *a_next = a_smem + (uint32_t)((step + 1) & 1) * kABytes + (uint32_t)(S::KX / 8) * 2048u + (uint32_t)site * 16u;
            const __half *pg_t = L2 ? pg_site + (size_t)t * pg_tstride : nullptr;
            const __half *pg_tn = L2 ? pg_site + (size_t)(dir ? t - 1 : t + 1) * pg_tstride : nullptr;     // next step (if any)
            op_t *h_t = p.hout + (L2 ? c3b_tile_major_offset(gsite, t * 40 + dir * 20, 1320)
                                     : c3b_tile_major_offset((size_t)t * p.bp + gsite, dir * 16, 32));
            if (L2 && step + 1 < C3B_T && (lane & 7) == 0) {
                // the pre-gates stream from HBM (86 MB per 1024 sites): pull the NEXT step's lines into L2 now (one prefetch per
                // 128-byte line), so the register loads one phase ahead below find them there
#pragma unroll
                for (int i = 0; i < kPhases * 4; ++i)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pg_tn + (size_t)(16 * (i >> 2) + 4 * (i & 3) + wg) * 128 * 8));
            }
#pragma unroll
            for (int ph = 0; ph < kPhases; ++ph) {
                // this phase's pre-gates were loaded one phase ago (registers refilled in place as soon as a gate has been consumed)
                const bool wrap = ph + 1 == kPhases;
                const bool more = L2 && (!wrap || step + 1 < C3B_T);
                const __half *pg_nx = (wrap ? pg_tn : pg_t) + (size_t)(16 * (wrap ? 0 : ph + 1) + wg) * 128 * 8;
                const bool trp = tr && step >= 8 && step < 8 + 6;
                if (trp) p.trace[((step - 8) * kPhases + ph) * 4 + 0] = clock64();
                ptx::mbar_wait(&acc_full[st], full_par);
                ptx::tc_fence_after();
                if (trp) p.trace[((step - 8) * kPhases + ph) * 4 + 1] = clock64();
                const uint32_t ta = lane_taddr + (uint32_t)(st * 128);
                float a[8];
                uint32_t si[4], ig[4], sf[4], so[4];
                // gate i, then g (tcgen05.ld is a 12-cycle operation: four short round trips cost less than the registers of wider loads)
                ptx::tmem_ld8(ta + 0, a);
                ptx::tmem_ld_wait();
                {
                    const uint32_t *pp = reinterpret_cast<const uint32_t *>(&pgv[0]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) si[j] = hfma2u(tanh_h2u(preact<L2>(a[2 * j], a[2 * j + 1], pp[j])), half2_half, half2_half);
                }
                ptx::tmem_ld8(ta + 64, a);
                ptx::tmem_ld_wait();
                {
                    const uint32_t *pp = reinterpret_cast<const uint32_t *>(&pgv[2]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) ig[j] = hmul2u(si[j], tanh_h2u(preact<L2>(a[2 * j], a[2 * j + 1], pp[j])));
                }
                if (more) {          // gates i and g of the next phase (four 16-byte loads per phase, coalesced over the warp)
                    pgv[0] = *reinterpret_cast<const uint4 *>(pg_nx);
                    pgv[2] = *reinterpret_cast<const uint4 *>(pg_nx + (size_t)8 * 128 * 8);
                }
                // gate f
                ptx::tmem_ld8(ta + 32, a);
                ptx::tmem_ld_wait();
                {
                    const uint32_t *pp = reinterpret_cast<const uint32_t *>(&pgv[1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) sf[j] = hfma2u(tanh_h2u(preact<L2>(a[2 * j], a[2 * j + 1], pp[j])), half2_half, half2_half);
                }
                // gate o
                ptx::tmem_ld8(ta + 96, a);
                ptx::tmem_ld_wait();
                // the stage has been read: hand it back to the MMA issuer (one arrival per warp, to the leader CTA)
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster_relaxed(&acc_empty[st], 0);
                if (trp) p.trace[((step - 8) * kPhases + ph) * 4 + 2] = clock64();
                {
                    const uint32_t *pp = reinterpret_cast<const uint32_t *>(&pgv[3]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) so[j] = hfma2u(tanh_h2u(preact<L2>(a[2 * j], a[2 * j + 1], pp[j])), half2_half, half2_half);
                }
                if (more) {          // gates f and o of the next phase
                    pgv[1] = *reinterpret_cast<const uint4 *>(pg_nx + (size_t)4 * 128 * 8);
                    pgv[3] = *reinterpret_cast<const uint4 *>(pg_nx + (size_t)12 * 128 * 8);
                }
                // cell update in fp32, h = o * tanh(c) on packed pairs
                uint4 hv;
                uint32_t *hp = reinterpret_cast<uint32_t *>(&hv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f2 = unpack_h2(sf[j]), g2 = unpack_h2(ig[j]);
                    float &c0 = c[ph * 8 + 2 * j], &c1 = c[ph * 8 + 2 * j + 1];
                    c0 = fmaf(f2.x, c0, g2.x);
                    c1 = fmaf(f2.y, c1, g2.y);
                    hp[j] = hmul2u(so[j], tanh_h2u(pack_h2(c0, c1)));
                }
                // h_t[site][units 32 ph + 8 wg .. + 8): next step's A operand (k-group 4 ph + wg) and the layer output
                *reinterpret_cast<uint4 *>(a_next + (uint32_t)(4 * ph + wg) * 2048u) = hv;
                if (trp) p.trace[((step - 8) * kPhases + ph) * 4 + 3] = clock64();
                if (++st == kStages) { st = 0; full_par ^= 1u; }
            }
            // every h_t value of this warp's sites and unit groups is in the operand buffer: tell the MMA issuer
            if (step + 1 < C3B_T) {
                ptx::fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&a_ready[(step + 1) & 1]);
            }
            // the layer output leaves AFTER the release above (a cluster-scope release waits for the arriving thread's earlier
            // global stores: issued before it, they would put one HBM write latency on the recurrence's critical path every
            // step): each thread copies its own 16-byte chunks back out of the operand buffer while the next step's MMAs run
#pragma unroll
            for (int ph = 0; ph < kPhases; ++ph)
                *reinterpret_cast<uint4 *>(h_t + (size_t)(4 * ph + wg) * 128 * 8) = *reinterpret_cast<const uint4 *>(a_next + (uint32_t)(4 * ph + wg) * 2048u);
        }
    }
    // teardown: both CTAs have finished every TMEM read / MMA before the pair's allocation is returned
    ptx::tc_fence_before();
    ptx::cluster_sync_all();
    if (warp == 17) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_pair<512>(tmem_base);
    }
}

}  // namespace

template <bool L2>
static int launch_pair(const c3b_model *m, const Lstm2xDev &p, cudaStream_t s) {
    const size_t smem = (size_t)Shape<L2>::kWBytes + 2 * Shape<L2>::kABytes;
    C3B_CUDA(cudaFuncSetAttribute(lstm_pair_kernel<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(p.bp / 128, 2);
    const_cast<c3b_model *>(m)->launches++;
    c3b_note_grid((long long)grid.x * grid.y);
    lstm_pair_kernel<L2><<<grid, kThreads, smem, s>>>(p);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_lstm2x(const c3b_model *m, const op_t *w_img, const __half *pg2, op_t *h2, int bp, long long *trace, cudaStream_t s) {
    if (bp % 256) { c3b_set_error("lstm2x: padded batch %d is not a multiple of 256", bp); return 1; }
    Lstm2xDev p = {};
    p.w_img = w_img; p.pg = pg2; p.hout = h2; p.bp = bp; p.trace = trace;
    return launch_pair<true>(m, p, s);
}

// LSTM1 on the CTA pair: xs2 = [33][bp/128][6][128][8] (ingest layout 1), h1 tile-major with rows t*bp + site
int c3b_launch_lstm1x(const c3b_model *m, const op_t *w_img, const op_t *xs2, op_t *h1, int bp, long long *trace, cudaStream_t s) {
    if (bp % 256) { c3b_set_error("lstm1x: padded batch %d is not a multiple of 256", bp); return 1; }
    Lstm2xDev p = {};
    p.w_img = w_img; p.xs = xs2; p.hout = h1; p.bp = bp; p.trace = trace;
    return launch_pair<false>(m, p, s);
}
