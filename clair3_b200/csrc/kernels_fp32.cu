// fp32 CUDA-core debug path (C3B_PREC_FP32): the same layer graph as the tensor-core path, written for
// obviousness, used to separate layout bugs from fp16 precision loss.  Not the product path.
#include "c3b_internal.h"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// One block per (site, direction); thread j owns hidden unit j.  torch nn.LSTM semantics
// (gate rows i,f,g,o; h0=c0=0; reverse runs t = T-1..0), clair3/model.py:96-107,132-133.
__global__ void lstm_f32_kernel(const float *__restrict__ x, LstmF32 wf, LstmF32 wb, float *__restrict__ out,
                                int in_dim, int hidden) {
    extern __shared__ float sm[];
    float *xs = sm;                 // [in_dim]
    float *hs = sm + in_dim;        // [hidden]
    const int64_t b = blockIdx.x;
    const int dir = blockIdx.y;
    const LstmF32 w = dir ? wb : wf;
    const int j = threadIdx.x;
    const int g4 = 4 * hidden;
    float c = 0.f;
    hs[j] = 0.f;
    const float bi = w.bias[j], bf = w.bias[hidden + j], bg = w.bias[2 * hidden + j], bo = w.bias[3 * hidden + j];
    for (int step = 0; step < C3B_T; ++step) {
        const int t = dir ? (C3B_T - 1 - step) : step;
        __syncthreads();
        for (int k = j; k < in_dim; k += hidden) xs[k] = x[(b * C3B_T + t) * in_dim + k];
        __syncthreads();
        float ai = bi, af = bf, ag = bg, ao = bo;
        for (int k = 0; k < in_dim; ++k) {
            const float v = xs[k];
            const float *wr = w.wih_t + (size_t)k * g4 + j;
            ai = fmaf(v, wr[0], ai);
            af = fmaf(v, wr[hidden], af);
            ag = fmaf(v, wr[2 * hidden], ag);
            ao = fmaf(v, wr[3 * hidden], ao);
        }
        for (int k = 0; k < hidden; ++k) {
            const float v = hs[k];
            const float *wr = w.whh_t + (size_t)k * g4 + j;
            ai = fmaf(v, wr[0], ai);
            af = fmaf(v, wr[hidden], af);
            ag = fmaf(v, wr[2 * hidden], ag);
            ao = fmaf(v, wr[3 * hidden], ao);
        }
        c = sigmoidf_(af) * c + sigmoidf_(ai) * tanhf(ag);
        const float h = sigmoidf_(ao) * tanhf(c);
        __syncthreads();
        hs[j] = h;
        out[(b * C3B_T + t) * (2 * hidden) + dir * hidden + j] = h;
    }
}

constexpr int DENSE_G = 8;
// out[b][o] = sum_k x[b][k] * w_t[k][o]   (no bias / activation: the heads kernel applies them)
__global__ void dense_f32_kernel(const float *__restrict__ x, const float *__restrict__ w_t, float *__restrict__ out,
                                 int64_t batch, int k_dim, int n) {
    __shared__ float xs[DENSE_G][256];
    const int o = threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * DENSE_G;
    float acc[DENSE_G];
#pragma unroll
    for (int g = 0; g < DENSE_G; ++g) acc[g] = 0.f;
    for (int k0 = 0; k0 < k_dim; k0 += 256) {
        const int kc = min(256, k_dim - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < DENSE_G * 256; i += blockDim.x) {
            int g = i >> 8, kk = i & 255;
            xs[g][kk] = (b0 + g < batch && kk < kc) ? x[(b0 + g) * k_dim + k0 + kk] : 0.f;
        }
        __syncthreads();
        if (o < n) {
            for (int kk = 0; kk < kc; ++kk) {
                const float w = w_t[(size_t)(k0 + kk) * n + o];
#pragma unroll
                for (int g = 0; g < DENSE_G; ++g) acc[g] = fmaf(xs[g][kk], w, acc[g]);
            }
        }
    }
    if (o < n) {
#pragma unroll
        for (int g = 0; g < DENSE_G; ++g)
            if (b0 + g < batch) out[(b0 + g) * n + o] = acc[g];
    }
}

constexpr int CONV_P = 4;   // output pixels per block
// NHWC 3x3 conv, pad 1, folded BN bias, optional residual, ReLU.  Block = Cout threads, CONV_P pixels.
__global__ void conv3x3_f32_kernel(const float *__restrict__ x, ConvF32 w, const float *__restrict__ residual,
                                   float *__restrict__ out, int64_t npix, int hin, int win, int hout, int wout) {
    extern __shared__ float patch[];   // [CONV_P][9*cin]
    const int cin = w.cin, cout = w.cout;
    const int64_t p0 = (int64_t)blockIdx.x * CONV_P;
    const int kk = 9 * cin;
    for (int i = threadIdx.x; i < CONV_P * kk; i += blockDim.x) {
        const int p = i / kk, r = i - p * kk;
        const int tap = r / cin, c = r - tap * cin;
        const int64_t pix = p0 + p;
        float v = 0.f;
        if (pix < npix) {
            const int wo = (int)(pix % wout);
            const int ho = (int)((pix / wout) % hout);
            const int64_t b = pix / ((int64_t)wout * hout);
            const int hi = ho * w.stride - 1 + tap / 3;
            const int wi = wo * w.stride - 1 + tap % 3;
            if (hi >= 0 && hi < hin && wi >= 0 && wi < win) v = x[((b * hin + hi) * win + wi) * cin + c];
        }
        patch[i] = v;
    }
    __syncthreads();
    const int co = threadIdx.x;
    float acc[CONV_P];
    const float bias = w.bias[co];
#pragma unroll
    for (int p = 0; p < CONV_P; ++p) acc[p] = bias;
    for (int r = 0; r < kk; ++r) {
        const float wv = w.w[(size_t)r * cout + co];
#pragma unroll
        for (int p = 0; p < CONV_P; ++p) acc[p] = fmaf(patch[p * kk + r], wv, acc[p]);
    }
#pragma unroll
    for (int p = 0; p < CONV_P; ++p) {
        const int64_t pix = p0 + p;
        if (pix < npix) {
            float v = acc[p];
            if (residual) v += residual[pix * cout + co];
            out[pix * cout + co] = fmaxf(v, 0.f);
        }
    }
}

// 3-level spatial pyramid max pool with TF-'SAME' zero padding (clair3/model.py:250-279); input is post-ReLU (>= 0)
// so the zero pad is equivalent to clamping the window and taking max with 0.  out: [B][(9+4+1)*C], NHWC order.
__global__ void spp_f32_kernel(const float *__restrict__ x, float *__restrict__ out, int64_t batch, int h, int w, int c) {
    const int64_t b = blockIdx.x;
    const int cells = 14;
    for (int i = threadIdx.x; i < cells * c; i += blockDim.x) {
        const int cell = i / c, ch = i - cell * c;
        int p, idx;
        if (cell < 9) { p = 3; idx = cell; }
        else if (cell < 13) { p = 2; idx = cell - 9; }
        else { p = 1; idx = 0; }
        const int wh = (h + p - 1) / p, ww = (w + p - 1) / p;
        const int oh = (h + wh - 1) / wh, ow = (w + ww - 1) / ww;
        const int ph = max((oh - 1) * wh + wh - h, 0), pw = max((ow - 1) * ww + ww - w, 0);
        const int pt = ph / 2, pl = pw / 2;
        const int oi = idx / ow, oj = idx - oi * ow;
        const int h0 = oi * wh - pt, w0 = oj * ww - pl;
        bool padded = false;
        float m = -1e30f;
        for (int hh = h0; hh < h0 + wh; ++hh)
            for (int wv = w0; wv < w0 + ww; ++wv) {
                if (hh < 0 || hh >= h || wv < 0 || wv >= w) { padded = true; continue; }
                m = fmaxf(m, x[((b * h + hh) * w + wv) * c + ch]);
            }
        if (padded) m = fmaxf(m, 0.f);
        out[b * (cells * c) + i] = m;
    }
}

}  // namespace

int c3b_launch_lstm_f32(const float *x, const LstmF32 &fwd, const LstmF32 &bwd, float *out, int64_t batch, int in_dim,
                        int hidden, cudaStream_t s) {
    if (batch == 0) return 0;
    dim3 grid((unsigned)batch, 2);
    size_t smem = sizeof(float) * (in_dim + hidden);
    lstm_f32_kernel<<<grid, hidden, smem, s>>>(x, fwd, bwd, out, in_dim, hidden);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_dense_f32(const float *x, const float *w_t, float *out, int64_t batch, int k, int n, cudaStream_t s) {
    if (batch == 0) return 0;
    int blocks = (int)((batch + DENSE_G - 1) / DENSE_G);
    dense_f32_kernel<<<blocks, 256, 0, s>>>(x, w_t, out, batch, k, n);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_conv_f32(const float *x, const ConvF32 &w, const float *residual, float *out, int64_t batch, int hin,
                        int win, int hout, int wout, cudaStream_t s) {
    if (batch == 0) return 0;
    const int64_t npix = batch * hout * wout;
    size_t smem = sizeof(float) * CONV_P * 9 * w.cin;
    int blocks = (int)((npix + CONV_P - 1) / CONV_P);
    conv3x3_f32_kernel<<<blocks, w.cout, smem, s>>>(x, w, residual, out, npix, hin, win, hout, wout);
    C3B_CUDA(cudaGetLastError());
    return 0;
}

int c3b_launch_spp_f32(const float *x, float *out, int64_t batch, int h, int w, int c, cudaStream_t s) {
    if (batch == 0) return 0;
    spp_f32_kernel<<<(unsigned)batch, 256, 0, s>>>(x, out, batch, h, w, c);
    C3B_CUDA(cudaGetLastError());
    return 0;
}
