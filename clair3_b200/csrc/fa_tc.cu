// Small memory-bound kernels around the tensor-core convolution stack of Clair3_F:
//  * ingest: int8 NHWC read image [B,D,33,C] -> fp16 parity planes (Cpad = 8 or 16); the 1/100 normalisation of
//    clair3/model.py:378 is folded into conv1's weights, and int8 values are exact in fp16.
//  * spp: 3-level spatial pyramid max pool (clair3/model.py:250-279) on the planar padded res_block3 output.
#include "c3b_internal.h"

namespace {

// int8/int32/f32 NHWC [B][D][33][channels] -> fp16, written straight into the four PARITY PLANES the stride-2 stem conv reads
// (pconv_tc.cu): padded pixel (hp, wp) = (h+1, w+1) goes to plane (hp&1, wp&1), slot (hp>>1, wp>>1) of the level-1 geometry.
// Threads are ordered like the OUTPUT (site, plane, row, column fastest) so a warp's 16-byte stores are consecutive slots of
// one plane; the reads are 8-byte pieces two pixels apart.  Channel groups with no real channel (conv1's input is padded to
// one UMMA k-step = 16 channels) stay zero from the workspace clear and are never written.  The 1/100 scale is folded into
// conv1's weights.
template <typename T>
__global__ void ingest_fa_tc_kernel(const T *__restrict__ x, op_t *__restrict__ out, int64_t batch, int channels,
                                    int cpad, int depth, PlanarGeom g1) {
    const int groups = (channels + 7) / 8;                 // groups holding real channels
    const int rows = g1.h + 1, cols = g1.w + 1;            // plane cells that can hold a real pixel
    const int per_site = 4 * rows * cols;
    const int64_t total = batch * per_site * groups;
    const size_t plane_elems = (size_t)(cpad / 8) * g1.p * 8;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cell = idx / groups;
        const int g = (int)(idx - cell * groups);
        const int64_t b = cell / per_site;
        int rem = (int)(cell - b * per_site);
        const int plane = rem / (rows * cols);
        rem -= plane * rows * cols;
        const int i = rem / cols, j = rem - i * cols;
        const int hp = 2 * i + (plane >> 1), wp = 2 * j + (plane & 1);
        if (hp < 1 || hp > depth || wp < 1 || wp > 33) continue;
        const T *src = x + ((b * depth + (hp - 1)) * 33 + (wp - 1)) * channels;
        __align__(16) op_t v[8];
        if (sizeof(T) == 1 && channels == 8 && (reinterpret_cast<uintptr_t>(x) & 7) == 0) {   // the common case: one aligned 8-byte load per pixel
            const uint2 q = *reinterpret_cast<const uint2 *>(src);
            const int8_t *b8 = reinterpret_cast<const int8_t *>(&q);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = f2op((float)b8[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int ch = g * 8 + e;
                v[e] = f2op_sat(ch < channels ? (float)src[ch] : 0.f);
            }
        }
        const size_t off = (size_t)plane * plane_elems + ((size_t)g * g1.p + g1.g + b * g1.s + (size_t)(i + 1) * g1.wp + (j + 1)) * 8;
        *reinterpret_cast<uint4 *>(out + off) = *reinterpret_cast<const uint4 *>(v);
    }
}

// One thread = one (pyramid cell, channel group of 8): 16-byte loads from the planar padded map, 8 running maxima,
// one 16-byte store into the k-group-planar L4 operand.  grid = B, block = 14 * C/8 threads (448 for C = 256).
// The loads of one window row are issued together (a rolled loop with a running max serialises one L2 round trip per
// pixel), and the 1x1 cell is the max of the four 2x2 cells (their windows tile the map), taken from shared memory.
__global__ void spp_tc_kernel(const op_t *__restrict__ x, PlanarGeom pg, op_t *__restrict__ out, int c, int bp) {
    __shared__ float p2max[4][32][8];
    const int h = pg.h, w = pg.w;
    const int64_t b = blockIdx.x;
    const int ncg = c >> 3;                                  // <= 32
    auto store = [&](int cell, int cg, const float *m) {
        // feature index f = cell*c + cg*8 + k (the reference's flatten order) -> k-group (f >> 3), tile-major [bp/128][3584/8][128][8]
        uint4 o;
        op2_t *oh2 = reinterpret_cast<op2_t *>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) oh2[k] = f2op2(m[2 * k], m[2 * k + 1]);
        *reinterpret_cast<uint4 *>(out + c3b_tile_major_offset((size_t)b, cell * ncg + cg, 14 * ncg)) = o;
    };
    for (int i = threadIdx.x; i < 13 * ncg; i += blockDim.x) {
        const int cell = i / ncg, cg = i - cell * ncg;
        int p, idx;
        if (cell < 9) { p = 3; idx = cell; }
        else { p = 2; idx = cell - 9; }
        const int wh = (h + p - 1) / p, ww = (w + p - 1) / p;
        const int oh = (h + wh - 1) / wh, ow = (w + ww - 1) / ww;
        const int ph = max((oh - 1) * wh + wh - h, 0), pw = max((ow - 1) * ww + ww - w, 0);
        const int pt = ph / 2, pl = pw / 2;
        const int oi = idx / ow, oj = idx - oi * ow;
        const int h0 = max(oi * wh - pt, 0), h1 = min(oi * wh - pt + wh, h);
        const int w0 = max(oj * ww - pl, 0), w1 = min(oj * ww - pl + ww, w);
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = 0.f;              // inputs are post-ReLU: zero padding == floor at 0
        const op_t *plane = x + ((size_t)cg * pg.p + pg.g + b * pg.s) * 8;
        for (int hh = h0; hh < h1; ++hh)
            for (int wb = w0; wb < w1; wb += 4) {
                uint4 v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[j] = wb + j < w1 ? *reinterpret_cast<const uint4 *>(plane + ((size_t)(hh + 1) * pg.wp + (wb + j + 1)) * 8)
                                       : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const op2_t *hp = reinterpret_cast<const op2_t *>(&v[j]);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = op22f2(hp[k]);
                        m[2 * k] = fmaxf(m[2 * k], f.x);
                        m[2 * k + 1] = fmaxf(m[2 * k + 1], f.y);
                    }
                }
            }
        store(cell, cg, m);
        if (cell >= 9) {
#pragma unroll
            for (int k = 0; k < 8; ++k) p2max[cell - 9][cg][k] = m[k];
        }
    }
    __syncthreads();
    for (int cg = threadIdx.x; cg < ncg; cg += blockDim.x) {
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
            m[k] = fmaxf(fmaxf(p2max[0][cg][k], p2max[1][cg][k]), fmaxf(p2max[2][cg][k], p2max[3][cg][k]));
        store(13, cg, m);
    }
}

}  // namespace

int c3b_launch_ingest_fa_tc(const void *x, int dtype, int channels, int cpad, op_t *out, int64_t batch, int depth, const PlanarGeom &g1,
                            cudaStream_t s) {
    if (batch == 0) return 0;
    if (channels > cpad) { c3b_set_error("full-alignment input has %d channels, at most %d supported", channels, cpad); return 1; }
    const int64_t total = batch * 4 * (g1.h + 1) * (g1.w + 1) * ((channels + 7) / 8);
    const int blocks = (int)((total + 255) / 256 < 8192 ? (total + 255) / 256 : 8192);
    c3b_note_grid(blocks);
    switch (dtype) {
        case C3B_DT_I8: ingest_fa_tc_kernel<int8_t><<<blocks, 256, 0, s>>>((const int8_t *)x, out, batch, channels, cpad, depth, g1); break;
        case C3B_DT_I32: ingest_fa_tc_kernel<int32_t><<<blocks, 256, 0, s>>>((const int32_t *)x, out, batch, channels, cpad, depth, g1); break;
        case C3B_DT_F32: ingest_fa_tc_kernel<float><<<blocks, 256, 0, s>>>((const float *)x, out, batch, channels, cpad, depth, g1); break;
        default: c3b_set_error("unsupported input dtype %d", dtype); return 1;
    }
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_spp_tc(const op_t *x, const PlanarGeom &g, op_t *out, int64_t batch, int c, int bp, cudaStream_t s) {
    if (batch == 0) return 0;
    c3b_note_grid(batch);
    spp_tc_kernel<<<(unsigned)batch, 448, 0, s>>>(x, g, out, c, bp);
    C3B_CUDA(cudaGetLastError());
    return 0;
}
