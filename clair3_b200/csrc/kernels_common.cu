// Kernels shared by both precisions: input ingest (dtype -> float) and the dense multi-task heads
// (SELU(L4) -> SELU(L5_k) -> SELU(Y_k) -> softmax, clair3/model.py:136-159 and 391-411), all fp32 on CUDA cores.
#include "c3b_internal.h"

namespace {

constexpr float kSeluAlpha = 1.6732632423543772f;
constexpr float kSeluScale = 1.0507009873554805f;

__device__ __forceinline__ float selu(float x) {
    return kSeluScale * (x > 0.f ? x : kSeluAlpha * expm1f(x));
}

template <typename T>
__global__ void ingest_f32_kernel(const T *__restrict__ x, float *__restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (float)x[i];
}

constexpr int HEADS_G = 8;        // sites per block
constexpr int HEADS_THREADS = 256;

// z4: [B][D4] L4 pre-activation WITHOUT bias.  One block = HEADS_G sites x all heads; thread (t/128, t%128) owns one L5
// unit of one head (two heads in flight, four heads in two passes), weights are read once per block through the
// read-only path, activations broadcast from shared memory.
__global__ void __launch_bounds__(HEADS_THREADS) heads_kernel(const float *__restrict__ z4, HeadsParams hp,
                                                              float *__restrict__ out, int64_t batch) {
    extern __shared__ float smem[];
    const int d4 = hp.d4;
    float *a = smem;                                   // [G][d4]
    float *l5 = a + HEADS_G * d4;                      // [nheads][G][128]
    float *yv = l5 + C3B_MAX_HEADS * HEADS_G * 128;    // [G][96]
    const int tid = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * HEADS_G;
    const int g_n = (int)min((int64_t)HEADS_G, batch - b0);

    // a is stored [k][G] so the L5 loop reads the 8 sites of one k with two 16-byte broadcast loads
    for (int i = tid; i < HEADS_G * d4; i += HEADS_THREADS) {
        const int g = i / d4, k = i - g * d4;
        a[k * HEADS_G + g] = (g < g_n) ? selu(z4[(b0 + g) * d4 + k] + __ldg(hp.b4 + k)) : 0.f;
    }
    __syncthreads();

    const int j = tid & 127;
    for (int h = tid >> 7; h < hp.nheads; h += 2) {
        const float *__restrict__ w = hp.h[h].w5t + j;
        float acc[HEADS_G];
        const float bias = __ldg(hp.h[h].b5 + j);
#pragma unroll
        for (int g = 0; g < HEADS_G; ++g) acc[g] = bias;
        static_assert(HEADS_G == 8, "two float4 per k");
#pragma unroll 8
        for (int k = 0; k < d4; ++k) {
            const float wv = __ldg(w + (size_t)k * 128);
            const float4 a0 = *reinterpret_cast<const float4 *>(a + k * HEADS_G);
            const float4 a1 = *reinterpret_cast<const float4 *>(a + k * HEADS_G + 4);
            acc[0] = fmaf(a0.x, wv, acc[0]);
            acc[1] = fmaf(a0.y, wv, acc[1]);
            acc[2] = fmaf(a0.z, wv, acc[2]);
            acc[3] = fmaf(a0.w, wv, acc[3]);
            acc[4] = fmaf(a1.x, wv, acc[4]);
            acc[5] = fmaf(a1.y, wv, acc[5]);
            acc[6] = fmaf(a1.z, wv, acc[6]);
            acc[7] = fmaf(a1.w, wv, acc[7]);
        }
#pragma unroll
        for (int g = 0; g < HEADS_G; ++g) l5[(h * HEADS_G + g) * 128 + j] = selu(acc[g]);
    }
    __syncthreads();

    for (int i = tid; i < HEADS_G * hp.out_dim; i += HEADS_THREADS) {
        const int g = i / hp.out_dim, o = i - g * hp.out_dim;
        int h = 0;
        while (h + 1 < hp.nheads && o >= hp.h[h + 1].out_off) ++h;
        const int n = hp.h[h].n, oo = o - hp.h[h].out_off;
        const float *__restrict__ wy = hp.h[h].wyt + oo;
        const float *lv = l5 + (h * HEADS_G + g) * 128;
        float s = __ldg(hp.h[h].by + oo);
#pragma unroll 8
        for (int jj = 0; jj < 128; ++jj) s = fmaf(lv[jj], __ldg(wy + jj * n), s);
        yv[g * 96 + o] = selu(s);
    }
    __syncthreads();

    if (tid < g_n * hp.nheads) {
        const int g = tid / hp.nheads, h = tid - g * hp.nheads;
        const int n = hp.h[h].n, off = hp.h[h].out_off;
        const float *v = yv + g * 96 + off;
        float mx = v[0];
        for (int o = 1; o < n; ++o) mx = fmaxf(mx, v[o]);
        float sum = 0.f;
        for (int o = 0; o < n; ++o) sum += expf(v[o] - mx);
        const float inv = 1.f / sum;
        float *dst = out + (b0 + g) * hp.out_dim + off;
        for (int o = 0; o < n; ++o) dst[o] = expf(v[o] - mx) * inv;
    }
}

template <typename T>
int launch_ingest(const void *x, float *out, int64_t n, cudaStream_t s) {
    if (n == 0) return 0;
    int blocks = (int)min((int64_t)4096, (n + 255) / 256);
    ingest_f32_kernel<T><<<blocks, 256, 0, s>>>((const T *)x, out, n);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int ingest_any(const void *x, int dtype, float *out, int64_t n, cudaStream_t s) {
    switch (dtype) {
        case C3B_DT_I8: return launch_ingest<int8_t>(x, out, n, s);
        case C3B_DT_I32: return launch_ingest<int32_t>(x, out, n, s);
        case C3B_DT_F32: return launch_ingest<float>(x, out, n, s);
        default: c3b_set_error("unsupported input dtype %d", dtype); return 1;
    }
}

}  // namespace

int c3b_launch_ingest_pileup_f32(const void *x, int dtype, float *out, int64_t n, cudaStream_t s) {
    return ingest_any(x, dtype, out, n, s);
}
int c3b_launch_ingest_fa_f32(const void *x, int dtype, float *out, int64_t n, cudaStream_t s) {
    return ingest_any(x, dtype, out, n, s);
}

int c3b_launch_heads(const float *z4, const HeadsParams &hp, float *out, int64_t batch, cudaStream_t s) {
    if (batch == 0) return 0;
    size_t smem = sizeof(float) * (HEADS_G * hp.d4 + C3B_MAX_HEADS * HEADS_G * 128 + HEADS_G * 96);
    int blocks = (int)((batch + HEADS_G - 1) / HEADS_G);
    heads_kernel<<<blocks, HEADS_THREADS, smem, s>>>(z4, hp, out, batch);
    C3B_CUDA(cudaGetLastError());
    return 0;
}
