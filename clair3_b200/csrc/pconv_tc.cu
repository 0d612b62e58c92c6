// Stride-1 3x3 convolution (the six residual-block convs of Clair3_F, 83 % of its FLOPs; clair3/model.py:200-235) as an
// implicit GEMM over SHIFTED VIEWS of a shared-memory-resident, zero-padded, channel-group-planar feature map.
//
// Activation layout ("planar padded"): [C/8][P][8] fp16.  A site's H x W feature map is stored as its (H+2) x (W+2)
// zero-bordered raster, sites back to back: slot g = b*S + (h+1)*Wp + (w+1), S = (H+2)*Wp, Wp = W+2, at plane offset G + g
// (G guard slots of zeros at both ends).  With that layout
//   * the input of a macro-tile (MT x 128 consecutive output slots plus a (Wp+1)-slot halo on each side) is C/8 CONTIGUOUS
//     runs of memory: it lands in shared memory with C/8 cp.async.bulk (TMA engine, async proxy; no per-thread gathers,
//     no im2col expansion, no proxy fences) in exactly the SWIZZLE_NONE K-major UMMA layout [k-group][slot][8];
//   * the A operand of tap (dh,dw) is the SAME image viewed (dh-1)*Wp + (dw-1) slots later: only the descriptor's start
//     address changes, so every input byte is fetched once per macro-tile and used by all nine taps;
//   * border slots are computed like any other row and overwritten with zeros in the epilogue, so the output is again a
//     valid planar padded tensor for the next convolution (and the residual add reads the same slot of its own input).
// Weights: the host-packed per-chunk images of igemm_tc.cu ([tap*C/64 + kc][8 k-groups][N][8]); resident when they fit
// (res_block1: 72 KB), otherwise streamed through a ring with each piece applied to MT accumulators.
//
// Roles (288 threads): warp 0 lane 0 issues all bulk copies, warp 8 issues tcgen05.mma (one elected lane), warps 4-7 run
// the epilogue on their TMEM lane quadrant (bias + residual + ReLU + border mask, 16-byte coalesced planar stores).
#include "c3b_internal.h"
#include "ptx.cuh"

namespace {

constexpr int kThreads = 288;
constexpr int kMaxWStages = 8;

struct PconvDev {
    const op_t *in;
    const op_t *w_img;
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
};

template <int MT>
__global__ void __launch_bounds__(kThreads, 1) pconv_kernel(const PconvDev p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t w_full[kMaxWStages], w_empty[kMaxWStages];
    __shared__ uint64_t img_full[2], img_empty[2], tmem_full[2], tmem_empty[2], w_res_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ float bias_s[256];

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const uint32_t w_bytes = (uint32_t)p.N * 128u;
    const uint32_t img_bytes = (uint32_t)p.nplanes * (uint32_t)(p.C / 8) * (uint32_t)p.n_in * 16u;
    const uint32_t lbo_img = (uint32_t)p.n_in * 16u;
    const uint32_t lbo_w = (uint32_t)p.N * 16u;
    const uint32_t smem_base = ptx::smem_u32(smem);
    const uint32_t w_region = p.w_resident ? (uint32_t)p.nchunks * w_bytes : (uint32_t)p.w_stages * w_bytes;
    const uint32_t img_base = smem_base + w_region;

    if (tid == 0) {
        for (int s = 0; s < kMaxWStages; ++s) { ptx::mbar_init(&w_full[s], 1); ptx::mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&img_full[s], 1);
            ptx::mbar_init(&img_empty[s], 1);
            ptx::mbar_init(&tmem_full[s], 1);
            ptx::mbar_init(&tmem_empty[s], 128);
        }
        ptx::mbar_init(&w_res_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 8) ptx::tmem_alloc<512>(&tmem_base_smem);
    for (int i = tid; i < p.N; i += kThreads) bias_s[i] = p.bias ? p.bias[i] : 0.f;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================================================== loader (one thread): image chunks + weight pieces
        if (lane == 0) {
            if (p.w_resident) {
                ptx::mbar_arrive_expect_tx(&w_res_bar, (uint32_t)p.nchunks * w_bytes);
                for (int c = 0; c < p.nchunks; ++c)
                    ptx::bulk_g2s(smem_base + (uint32_t)c * w_bytes, (const char *)p.w_img + (size_t)c * w_bytes, w_bytes, &w_res_bar);
            }
            auto load_img = [&](int macro, int li) {
                const int buf = li % p.img_bufs;
                const uint32_t ph = (uint32_t)(li / p.img_bufs) & 1u;
                ptx::mbar_wait(&img_empty[buf], ph ^ 1u);
                ptx::mbar_arrive_expect_tx(&img_full[buf], img_bytes);
                const long long slot0 = (long long)p.G + 128LL * MT * macro - p.halo_lo;
                const uint32_t dst = img_base + (uint32_t)buf * img_bytes;
                for (int pl = 0; pl < p.nplanes; ++pl)
                    for (int kg = 0; kg < p.C / 8; ++kg)
                        ptx::bulk_g2s(dst + (uint32_t)(pl * (p.C / 8) + kg) * lbo_img,
                                      (const char *)(p.in + (size_t)pl * p.plane_elems) + ((size_t)kg * p.P + slot0) * 16, lbo_img,
                                      &img_full[buf]);
            };
            int li = 0, wit = 0;
            if ((int)blockIdx.x < p.n_macro) load_img(blockIdx.x, 0);
            for (int macro = blockIdx.x; macro < p.n_macro; macro += gridDim.x, ++li) {
                const int next = macro + gridDim.x;
                if (p.img_bufs == 2 && next < p.n_macro) load_img(next, li + 1);      // prefetch while this tile computes
                if (!p.w_resident) {
                    for (int c = 0; c < p.nchunks; ++c, ++wit) {
                        const int s = wit % p.w_stages;
                        const uint32_t ph = (uint32_t)(wit / p.w_stages) & 1u;
                        ptx::mbar_wait(&w_empty[s], ph ^ 1u);
                        ptx::mbar_arrive_expect_tx(&w_full[s], w_bytes);
                        ptx::bulk_g2s(smem_base + (uint32_t)s * w_bytes, (const char *)p.w_img + (size_t)c * w_bytes, w_bytes, &w_full[s]);
                    }
                }
                if (p.img_bufs == 1 && next < p.n_macro) load_img(next, li + 1);      // single buffer: after this tile's MMAs
            }
        }
    } else if (warp == 8) {
        // ===================================================== MMA issuer
        const uint32_t idesc = ptx::umma_idesc_f16(128, (uint32_t)p.N);
        int li = 0, wit = 0;
        if (p.w_resident) ptx::mbar_wait(&w_res_bar, 0);
        for (int macro = blockIdx.x; macro < p.n_macro; macro += gridDim.x, ++li) {
            const int buf = li % p.img_bufs;
            const uint32_t iph = (uint32_t)(li / p.img_bufs) & 1u;
            const int acc = li % p.acc_stages;
            const uint32_t aph = (uint32_t)(li / p.acc_stages) & 1u;
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && li < 8;
            if (tr) p.trace[li * 8 + 0] = clock64();
            ptx::mbar_wait(&img_full[buf], iph);
            if (tr) p.trace[li * 8 + 1] = clock64();
            ptx::mbar_wait(&tmem_empty[acc], aph ^ 1u);
            ptx::tc_fence_after();
            if (tr) p.trace[li * 8 + 2] = clock64();
            const uint32_t img = img_base + (uint32_t)buf * img_bytes;
            const uint32_t d0 = tmem_base + (uint32_t)(acc * MT * p.N);
            for (int c = 0; c < p.nchunks; ++c) {
                uint32_t w_addr = smem_base + (uint32_t)c * w_bytes;
                int s = 0;
                if (!p.w_resident) {
                    s = wit % p.w_stages;
                    ptx::mbar_wait(&w_full[s], (uint32_t)(wit / p.w_stages) & 1u);
                    ptx::tc_fence_after();
                    w_addr = smem_base + (uint32_t)s * w_bytes;
                    ++wit;
                }
                if (ptx::elect_one()) {
                    // descriptors differ only in their 14-bit start-address field: build one per operand and add constants
                    // (rolled k loop + unrolled tile loop keeps the issue loop on the uniform datapath)
                    const uint64_t b_base = ptx::umma_desc_nosw(w_addr, lbo_w, 128u);
                    const uint64_t a_img = ptx::umma_desc_nosw(img, lbo_img, 128u);
                    const uint32_t b_step = (2u * lbo_w) >> 4;
                    const int ks_n = min(4, p.nksteps - 4 * c);
#pragma unroll 1
                    for (int ks = 0; ks < ks_n; ++ks) {
                        const int q = 4 * c + ks;                         // global k-step: tap = q / (C/16), kk = q % (C/16)
                        const int tap = q >> p.kpt_shift, kk = q - (tap << p.kpt_shift);
                        const int dh = tap / 3, dw = tap - dh * 3;
                        const int plane = p.nplanes == 4 ? ((dh & 1) * 2 + (dw & 1)) : 0;
                        const int shift = p.nplanes == 4 ? (dh >> 1) * p.Wp + (dw >> 1) : dh * p.Wp + dw;   // slots into the chunk
                        const uint64_t a_ks = a_img + (uint64_t)(((uint32_t)(plane * (p.C / 8) + 2 * kk) * lbo_img + (uint32_t)shift * 16u) >> 4);
                        const uint64_t b_desc = b_base + (uint64_t)(ks * b_step);
                        const uint32_t accum = (c > 0 || ks > 0) ? 1u : 0u;
#pragma unroll
                        for (int ti = 0; ti < MT; ++ti)
                            ptx::umma_f16(d0 + (uint32_t)(ti * p.N), a_ks + (uint64_t)(ti * 128), b_desc, idesc, accum);
                    }
                    if (!p.w_resident) ptx::umma_commit(&w_empty[s]);
                    if (c + 1 == p.nchunks) {
                        ptx::umma_commit(&tmem_full[acc]);
                        ptx::umma_commit(&img_empty[buf]);
                    }
                }
                __syncwarp();
            }
            if (tr) p.trace[li * 8 + 3] = clock64();
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue warps 4..7
        const int q = warp & 3;
        const int r = q * 32 + lane;
        int li = 0;
        for (int macro = blockIdx.x; macro < p.n_macro; macro += gridDim.x, ++li) {
            const int acc = li % p.acc_stages;
            const uint32_t aph = (uint32_t)(li / p.acc_stages) & 1u;
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && tid == 128 && li < 8;
            if (tr) p.trace[li * 8 + 4] = clock64();
            ptx::mbar_wait(&tmem_full[acc], aph);
            ptx::tc_fence_after();
            if (tr) p.trace[li * 8 + 5] = clock64();
            for (int ti = 0; ti < MT; ++ti) {
                const long long g = (128LL * MT) * macro + 128LL * ti + r;       // output slot
                const bool in_data = g < p.T;
                bool real = false;
                size_t par_off = 0;                 // parity-scatter destination (plane + slot of the next level)
                if (in_data) {
                    const long long b = g / p.S;
                    const int l = (int)(g - b * p.S);
                    const int hh = l / p.Wp, ww = l - hh * p.Wp;
                    real = hh >= 1 && hh <= p.H && ww >= 1 && ww <= p.W;
                    if (p.out_parity && real)
                        par_off = (size_t)((hh & 1) * 2 + (ww & 1)) * p.out_plane_elems +
                                  ((size_t)p.nG + b * p.nS + (size_t)((hh >> 1) + 1) * p.nWp + ((ww >> 1) + 1)) * 8;
                }
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * MT * p.N + ti * p.N);
                for (int j0 = 0; j0 < p.N; j0 += 16) {
                    float v[16];
                    ptx::tmem_ld16(taddr + (uint32_t)j0, v);
                    ptx::tmem_ld_wait();
                    if (in_data && (real || !p.out_parity)) {
                        const size_t o0 = ((size_t)(j0 >> 3) * p.P + p.G + g) * 8;       // this conv's own slot (residual / plain output)
                        const size_t o1 = o0 + (size_t)p.P * 8;
                        uint4 pk[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                        if (real) {
                            uint4 res[2];
                            if (p.residual) {
                                res[0] = *reinterpret_cast<const uint4 *>(p.residual + o0);
                                res[1] = *reinterpret_cast<const uint4 *>(p.residual + o1);
                            }
                            uint32_t *pw = reinterpret_cast<uint32_t *>(pk);
                            const op2_t *rp = reinterpret_cast<const op2_t *>(res);
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float a = v[2 * i] + bias_s[j0 + 2 * i];
                                float b = v[2 * i + 1] + bias_s[j0 + 2 * i + 1];
                                if (p.residual) {
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
                        }
                        if (p.out_parity) {
                            op_t *d0p = p.out + par_off + ((size_t)(j0 >> 3) * p.nP) * 8;
                            *reinterpret_cast<uint4 *>(d0p) = pk[0];
                            *reinterpret_cast<uint4 *>(d0p + (size_t)p.nP * 8) = pk[1];
                        } else {
                            *reinterpret_cast<uint4 *>(p.out + o0) = pk[0];
                            *reinterpret_cast<uint4 *>(p.out + o1) = pk[1];
                        }
                    }
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tmem_empty[acc]);
            if (tr) p.trace[li * 8 + 6] = clock64();
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace

int c3b_launch_pconv(const c3b_model *m, const PconvArgs &a, cudaStream_t s) {
    const PlanarGeom &g = a.geom;
    const bool c_ok = a.c == 16 || a.c == 64 || a.c == 128 || a.c == 256;
    if (!c_ok || a.n % 16 || a.n > 256 || a.n < 16) { c3b_set_error("pconv: unsupported channels %d -> %d", a.c, a.n); return 1; }
    PconvDev p = {};
    p.in = a.in; p.w_img = a.w.w_img; p.bias = a.w.bias; p.residual = a.residual; p.out = a.out;
    p.C = a.c; p.N = a.n;
    p.H = g.h; p.W = g.w; p.Wp = g.wp; p.S = g.s; p.G = g.g; p.T = g.t; p.P = g.p;
    p.relu = a.relu;
    p.trace = a.trace;
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
    const size_t budget = 220 * 1024;
    const size_t w_bytes = (size_t)a.n * 128;
    const size_t w_all = (size_t)p.nchunks * w_bytes;
    // Configuration search over MT in {4,2,1}: resident weights when they fit (then small MT only costs halo re-reads and
    // balances the tile count over the SMs); streamed weights want MT >= 2 (every piece feeds MT accumulators) and a deep
    // ring, so the image is single-buffered there.  Cost model = rounds of macro-tiles x MMAs per macro-tile.
    long long best_cost = -1;
    for (int mt = (a.n <= 64 ? 4 : 2); mt >= 1; mt >>= 1) {
        if (mt * a.n > 512) continue;
        const int n_in = a.stride2 ? (128 * mt + g.wp + 1 + 7) / 8 * 8 : 128 * mt + 2 * (g.wp + 1);
        const size_t img_bytes = (size_t)p.nplanes * (a.c / 8) * n_in * 16;
        int resident = 0, bufs = 0, stages = 0;
        if (w_all + 2 * img_bytes <= budget) { resident = 1; bufs = 2; }
        else if (w_all + img_bytes <= budget) { resident = 1; bufs = 1; }
        else if (mt >= 2 || a.n <= 64 || a.stride2) {
            if (img_bytes + 2 * w_bytes <= budget) { bufs = 1; stages = (int)((budget - img_bytes) / w_bytes); }
            if (2 * img_bytes + 6 * w_bytes <= budget) { bufs = 2; stages = (int)((budget - 2 * img_bytes) / w_bytes); }
            if (!bufs) continue;
        } else continue;
        const long long n_macro = (g.t + 128LL * mt - 1) / (128LL * mt);
        const long long rounds = (n_macro + m->sm_count - 1) / m->sm_count;
        // + a fixed per-macro-tile cost (image latency is only hidden behind several tiles of MMAs: MT = 1 measured 1.5x slower)
        long long cost = rounds * mt * 100 + rounds * 60 + (bufs == 1 ? rounds * 12 : 0) + (resident ? 0 : 5);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            p.MT = mt; p.n_in = n_in; p.w_resident = resident; p.img_bufs = bufs; p.w_stages = stages;
            p.acc_stages = (mt * a.n * 2 <= 512) ? 2 : 1;
        }
    }
    if (best_cost < 0) { c3b_set_error("pconv: feature map does not fit shared memory"); return 1; }
    if (p.w_stages > kMaxWStages) p.w_stages = kMaxWStages;
    const long long per_macro = 128LL * p.MT;
    p.n_macro = (int)((g.t + per_macro - 1) / per_macro);
    if ((long long)p.n_macro * per_macro + g.g > g.p - g.g + per_macro) { /* plane pitch covers the rounded-up slot range by construction */ }
    const size_t img_bytes = (size_t)p.nplanes * (a.c / 8) * p.n_in * 16;
    const size_t smem = (p.w_resident ? (size_t)p.nchunks * w_bytes : (size_t)p.w_stages * w_bytes) + p.img_bufs * img_bytes + 256;
    const int grid = p.n_macro < m->sm_count ? p.n_macro : m->sm_count;
    const_cast<c3b_model *>(m)->launches++;
    switch (p.MT) {
        case 1:
            C3B_CUDA(cudaFuncSetAttribute(pconv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
            pconv_kernel<1><<<grid, kThreads, smem, s>>>(p);
            break;
        case 2:
            C3B_CUDA(cudaFuncSetAttribute(pconv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
            pconv_kernel<2><<<grid, kThreads, smem, s>>>(p);
            break;
        case 4:
            C3B_CUDA(cudaFuncSetAttribute(pconv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
            pconv_kernel<4><<<grid, kThreads, smem, s>>>(p);
            break;
        default: c3b_set_error("pconv: unsupported MT %d", p.MT); return 1;
    }
    C3B_CUDA(cudaGetLastError());
    return 0;
}
