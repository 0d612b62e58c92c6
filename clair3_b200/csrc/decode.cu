// First, fully data-parallel stage of the reference's per-site decoder (`batch_output` -> `output_with` -> `output_from` ->
// `possible_outcome_probabilites_from`, clair3/CallVariants.py:1069-1116, 676-700, 510-576) on the GPU, so that only the sites
// that are NOT an early-out homozygous-reference call travel back to the (pure-Python, per-site) decoder:
//   * head slicing gt21 | genotype | indel_1 | indel_2 (param.label_shape_cum, CallVariants.py:1072,1082)
//   * the early-out test  homo_reference >= 0.5 and gt21[ref_base+ref_base] >= 0.5 (and, with indel heads, both
//     variant_length[0 + index_offset] >= 0.5)                                   CallVariants.py:532-534, 573-576
//   * homo_Ref_probability, the product the early-out returns, in the reference's float32 evaluation order   :527, 569-572
//   * per-head arg-max (first maximum, like numpy) and max probability
//   * QUAL of the early-out call, quality_score_from before its round(.., 2)       CallVariants.py:375-381
//   * the stable (ascending) compaction of the remaining site indices.
// Integer outputs (flags, arg-max, indices, count) are bit-exact vs the numpy restatement in oracle/decode_oracle.py; the
// float32 product is IEEE-exact (no contraction: multiplies only).
#include "c3b_internal.h"

namespace {

constexpr int kDecodeThreads = 1024;

struct DecodeDev {
    const float *y;
    const uint8_t *ref_gt21;
    int64_t batch;
    int out_dim, nheads;
    uint8_t *is_ref;
    float *ref_prob;
    int32_t *argmax;      // [batch][nheads]
    float *maxprob;       // [batch][nheads]
    double *qual;         // [batch]
    int32_t *nonref_idx;  // [batch]
    int32_t *n_nonref;    // [1]
};

__global__ void __launch_bounds__(kDecodeThreads) decode_stage1_kernel(const DecodeDev p) {
    __shared__ int warp_cnt[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int off[5] = {0, 21, 24, 57, 90};
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < p.batch; b0 += kDecodeThreads) {
        const int64_t b = b0 + tid;
        bool nonref = false;
        if (b < p.batch) {
            const float *row = p.y + b * p.out_dim;
            for (int h = 0; h < p.nheads; ++h) {
                int am = 0;
                float mx = row[off[h]];
                for (int o = off[h] + 1; o < off[h + 1]; ++o) {
                    const float v = row[o];
                    if (v > mx) { mx = v; am = o - off[h]; }
                }
                p.argmax[b * p.nheads + h] = am;
                p.maxprob[b * p.nheads + h] = mx;
            }
            const float homo_ref = row[21 + 0];                       // Genotype.homo_reference = 0 (clair3/task/genotype.py:7)
            const float gt_ref = row[p.ref_gt21[b]];
            bool early = homo_ref >= 0.5f && gt_ref >= 0.5f;
            float prob;
            if (p.nheads == 4) {
                const float v1 = row[24 + 16], v2 = row[57 + 16];     // variant_length index_offset = 16 (task/variant_length.py:6)
                early = early && v1 >= 0.5f && v2 >= 0.5f;
                prob = __fmul_rn(__fmul_rn(__fmul_rn(v1, v2), homo_ref), gt_ref);
            } else {
                prob = __fmul_rn(homo_ref, gt_ref);
            }
            p.is_ref[b] = early ? 1 : 0;
            p.ref_prob[b] = prob;
            // quality_score_from: max(Phred_Trans * log(((1.0 - p) + 1e-10) / (p + 1e-10)) + 10, 0); the ratio is float32 arithmetic
            // on a numpy float32 scalar (NumPy >= 2 promotion), the log is math.log of that value in double
            const float ratio = __fdiv_rn(__fadd_rn(__fsub_rn(1.0f, prob), 1e-10f), __fadd_rn(prob, 1e-10f));
            const double q = -4.342944819032518 * log((double)ratio) + 10.0;
            p.qual[b] = q > 0.0 ? q : 0.0;
            nonref = !early;
        }
        // stable compaction of the non-reference sites of this 1024-site slab
        const unsigned bal = __ballot_sync(0xffffffffu, nonref);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int c = warp_cnt[w];
            if (w < warp) before += c;
            total += c;
        }
        if (nonref) p.nonref_idx[base_s + before + __popc(bal & ((1u << lane) - 1u))] = (int32_t)b;
        __syncthreads();
        if (tid == 0) base_s += total;
        __syncthreads();
    }
    if (tid == 0) *p.n_nonref = base_s;
}

}  // namespace

int c3b_launch_decode_stage1(const float *y, const uint8_t *ref_gt21, int64_t batch, int out_dim, uint8_t *is_ref, float *ref_prob,
                             int32_t *argmax, float *maxprob, double *qual, int32_t *nonref_idx, int32_t *n_nonref, cudaStream_t s) {
    DecodeDev p;
    p.y = y; p.ref_gt21 = ref_gt21; p.batch = batch; p.out_dim = out_dim; p.nheads = out_dim == 90 ? 4 : 2;
    p.is_ref = is_ref; p.ref_prob = ref_prob; p.argmax = argmax; p.maxprob = maxprob; p.qual = qual;
    p.nonref_idx = nonref_idx; p.n_nonref = n_nonref;
    decode_stage1_kernel<<<1, kDecodeThreads, 0, s>>>(p);
    C3B_CUDA(cudaGetLastError());
    return 0;
}
