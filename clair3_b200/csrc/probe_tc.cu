// Micro-probes of tcgen05 behaviour on this part (diagnostics, tools/diag.py probe): (1) numerics of an MMA whose A
// operand lives in TMEM ("TS" form) under the assumed layout lane = row, one 32-bit column = two consecutive K elements;
// (2) cycles per MMA for back-to-back issue with A from shared memory (SS) and from TMEM (TS) at several N.
#include "c3b_internal.h"
#include "ptx.cuh"

namespace {

// a: [128][16] fp16 row-major, b: [n][16] fp16 row-major, d: [128][n] fp32.  timing[0..]: cycles for `reps` back-to-back MMAs.
__global__ void __launch_bounds__(128, 1) probe_kernel(const __half *a, const __half *b, float *d, int n, int reps,
                                                       long long *timing) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint8_t *a_smem = smem;                 // [2 kgroups][128][8]  (LBO = 2048)
    uint8_t *b_smem = smem + 4096;          // [2 kgroups][n][8]    (LBO = n*16)
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_smem);
    // stage A (smem image) and B
    for (int i = tid; i < 128 * 2; i += 128) {
        const int row = i >> 1, kg = i & 1;
        *reinterpret_cast<uint4 *>(a_smem + kg * 2048 + row * 16) = *reinterpret_cast<const uint4 *>(a + row * 16 + kg * 8);
    }
    for (int i = tid; i < n * 2; i += 128) {
        const int row = i >> 1, kg = i & 1;
        *reinterpret_cast<uint4 *>(b_smem + kg * n * 16 + row * 16) = *reinterpret_cast<const uint4 *>(b + row * 16 + kg * 8);
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t lane_t = tmem_base + ((uint32_t)(warp * 32) << 16);
    // A into TMEM columns [256, 264): thread = row, 8 x 32-bit = 16 fp16 of its row
    {
        uint32_t r[8];
        const uint4 v0 = *reinterpret_cast<const uint4 *>(a + tid * 16);
        const uint4 v1 = *reinterpret_cast<const uint4 *>(a + tid * 16 + 8);
        r[0] = v0.x; r[1] = v0.y; r[2] = v0.z; r[3] = v0.w; r[4] = v1.x; r[5] = v1.y; r[6] = v1.z; r[7] = v1.w;
        ptx::tmem_st8(lane_t + 256, r);
        ptx::tmem_st_wait();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::umma_idesc_f16(128, (uint32_t)n);
    const uint64_t a_desc = ptx::umma_desc_nosw(ptx::smem_u32(a_smem), 2048, 128);
    const uint64_t b_desc = ptx::umma_desc_nosw(ptx::smem_u32(b_smem), (uint32_t)n * 16u, 128);
    uint32_t phase = 0;
    // ---- numerics: D_ts at columns [0,n), D_ss at columns [n, 2n) is skipped (SS is what every production kernel uses)
    if (tid == 0) {
        ptx::umma_f16_ts(tmem_base, tmem_base + 256, b_desc, idesc, 0);
        ptx::umma_commit(&bar);
    }
    ptx::mbar_wait(&bar, phase); phase ^= 1;
    ptx::tc_fence_after();
    for (int j0 = 0; j0 < n; j0 += 8) {
        float v[8];
        ptx::tmem_ld8(lane_t + j0, v);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 8; ++i) d[tid * n + j0 + i] = v[i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    // ---- timing: reps back-to-back MMAs (accumulating), SS then TS
    // modes: 0 SS same accumulator, 1 TS same accumulator, 2 SS round-robin over 4 accumulators (n <= 64),
    //        3 SS round-robin over 2 accumulators (n <= 128), 4 TS round-robin over 4 accumulators (n <= 32)
    for (int mode = 0; mode < 5; ++mode) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                if (mode == 0) ptx::umma_f16(tmem_base, a_desc, b_desc, idesc, 1);
                else if (mode == 1) ptx::umma_f16_ts(tmem_base, tmem_base + 256, b_desc, idesc, 1);
                else if (mode == 2) ptx::umma_f16(tmem_base + (uint32_t)((i & 3) * (n <= 64 ? n : 0)), a_desc, b_desc, idesc, 1);
                else if (mode == 3) ptx::umma_f16(tmem_base + (uint32_t)((i & 1) * (n <= 128 ? n : 0)), a_desc, b_desc, idesc, 1);
                else ptx::umma_f16_ts(tmem_base + (uint32_t)((i & 3) * (n <= 32 ? n : 0)), tmem_base + 256, b_desc, idesc, 1);
            }
            ptx::umma_commit(&bar);
            timing[mode * 2 + 0] = clock64() - t0;          // issue time
        }
        ptx::mbar_wait(&bar, phase); phase ^= 1;
        if (tid == 0) timing[mode * 2 + 1] = clock64() - t0;  // completion time
        ptx::tc_fence_after();
        __syncthreads();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

// Timing-only probe of operand / accumulator switching: `reps` back-to-back 128 x n x 16 MMAs where MMA i uses
//   A tile (i % 8) at a_step bytes apart (+ a_mis bytes), B tile ((i / b_div) % 8) at b_step bytes apart, accumulator (i % d_cnt).
struct MmaProbeMode { int a_step, a_mis, b_step, b_div, d_cnt; };
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int reps, int nmodes, const MmaProbeMode *modes, long long *timing) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_smem);
    for (int i = tid; i < (64 + 80) * 1024 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t idesc = ptx::umma_idesc_f16(128, (uint32_t)n);
    const uint32_t a0 = ptx::smem_u32(smem), b0 = a0 + 64 * 1024;
    uint32_t phase = 0;
    for (int mode = 0; mode < nmodes; ++mode) {
        const MmaProbeMode md = modes[mode];
        long long t0 = 0;
        if (tid == 0) {
            const uint64_t a_desc = ptx::umma_desc_nosw(a0 + (uint32_t)md.a_mis, 2048 + 512, 128);
            const uint64_t b_desc = ptx::umma_desc_nosw(b0, (uint32_t)n * 16u, 128);
            t0 = clock64();
            // descriptors / accumulator addresses of one period of 8 MMAs are computed up front so the timed loop is only
            // the MMA issue (b_div and d_cnt are powers of two <= 8)
            uint64_t av[8], bv[8];
            uint32_t dv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                av[j] = a_desc + (uint64_t)((j * md.a_step) >> 4);
                bv[j] = b_desc + (uint64_t)(((j / md.b_div) * md.b_step) >> 4);
                dv[j] = tmem_base + (uint32_t)((j % md.d_cnt) * n);
            }
            t0 = clock64();
            for (int i = 0; i < reps; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) ptx::umma_f16(dv[j], av[j], bv[j], idesc, 1);
            }
            ptx::umma_commit(&bar);
            timing[mode * 2 + 0] = clock64() - t0;
        }
        ptx::mbar_wait(&bar, phase); phase ^= 1;
        if (tid == 0) timing[mode * 2 + 1] = clock64() - t0;
        ptx::tc_fence_after();
        __syncthreads();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

// TMEM read probe: warps 0-3 time `reps` x (tcgen05.ld 32x32b.x16 + wait::ld), singly and in batches of four loads per
// wait, first with the tensor pipe idle and then while warp 4 streams 128 x 128 x 16 MMAs into other TMEM columns.
__global__ void __launch_bounds__(160, 1) tmem_probe_kernel(int reps, long long *timing, float *sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ volatile int stop_flag;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); stop_flag = 0; }
    if (warp == 0) ptx::tmem_alloc<512>(&tmem_base_smem);
    for (int i = tid; i < 32 * 1024 / 16; i += 160) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0, 0);
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t lane_t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    for (int phase = 0; phase < 2; ++phase) {          // 0: tensor pipe idle, 1: MMA stream running
        if (warp == 4) {
            if (phase == 1 && ptx::elect_one()) {
                const uint32_t idesc = ptx::umma_idesc_f16(128, 128);
                const uint64_t a_desc = ptx::umma_desc_nosw(ptx::smem_u32(smem), 2048, 128);
                const uint64_t b_desc = ptx::umma_desc_nosw(ptx::smem_u32(smem) + 8192, 2048, 128);
                while (!stop_flag) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) ptx::umma_f16(tmem_base + 256 + (uint32_t)((i & 1) * 128), a_desc, b_desc, idesc, 1);
                }
                ptx::umma_commit(&bar);
                ptx::mbar_wait(&bar, 0);
            }
            __syncwarp();
        } else {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {                              // one load per wait
                float v[16];
                ptx::tmem_ld16(lane_t + (uint32_t)((i & 7) * 16), v);
                ptx::tmem_ld_wait();
                acc += v[0] + v[15];
            }
            long long t1 = clock64();
            for (int i = 0; i < reps; i += 4) {                           // four loads per wait
                float v0[16], v1[16], v2[16], v3[16];
                ptx::tmem_ld16(lane_t + 0, v0);
                ptx::tmem_ld16(lane_t + 16, v1);
                ptx::tmem_ld16(lane_t + 32, v2);
                ptx::tmem_ld16(lane_t + 48, v3);
                ptx::tmem_ld_wait();
                acc += v0[0] + v1[1] + v2[2] + v3[3];
            }
            long long t2 = clock64();
            // a chunk as the conv epilogue does it: load, wait, 16 adds, 8 packs, two 16-byte stores (coalesced)
            for (int i = 0; i < reps; ++i) {
                float v[16];
                ptx::tmem_ld16(lane_t + (uint32_t)((i & 7) * 16), v);
                ptx::tmem_ld_wait();
                uint4 pk[2];
                uint32_t *pw = reinterpret_cast<uint32_t *>(pk);
#pragma unroll
                for (int k = 0; k < 8; ++k) pw[k] = f2op2_sat(v[2 * k] + 1.f, v[2 * k + 1] + 1.f);
                uint4 *dst = reinterpret_cast<uint4 *>(sink) + ((size_t)(i & 63) * 2 * 128 + tid);
                dst[0] = pk[0];
                dst[128] = pk[1];
            }
            long long t3 = clock64();
            if (tid == 0) {
                timing[phase * 3 + 0] = t1 - t0;
                timing[phase * 3 + 1] = t2 - t1;
                timing[phase * 3 + 2] = t3 - t2;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (phase == 1 && tid == 0) stop_flag = 1;
        }
        if (phase == 0) __syncthreads();
    }
    if (acc == 123.456f) sink[0] = acc;
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<512>(tmem_base); }
}

}  // namespace

// timing[6]: {1 ld/wait, 4 ld/wait, full chunk} x {tensor idle, MMA stream running}, cycles for `reps` chunks
extern "C" int c3b_debug_tmem_probe(int reps, int64_t *timing) {
    void *dt, *ds;
    C3B_CUDA(cudaMalloc(&dt, 6 * 8));
    C3B_CUDA(cudaMalloc(&ds, 64 * 2 * 128 * 16 + 64));
    C3B_CUDA(cudaFuncSetAttribute(tmem_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    tmem_probe_kernel<<<1, 160, 64 * 1024>>>(reps, (long long *)dt, (float *)ds);
    C3B_CUDA(cudaGetLastError());
    C3B_CUDA(cudaDeviceSynchronize());
    C3B_CUDA(cudaMemcpy(timing, dt, 48, cudaMemcpyDeviceToHost));
    cudaFree(dt); cudaFree(ds);
    return 0;
}

// modes: nmodes x 5 ints {a_step, a_mis, b_step, b_div, d_cnt}; timing: nmodes x {issue, done} cycles
extern "C" int c3b_debug_mma_probe(int n, int reps, int nmodes, const int *modes, int64_t *timing) {
    if (n % 16 || n < 16 || n > 256 || nmodes < 1 || nmodes > 32) { c3b_set_error("mma probe: bad arguments"); return 1; }
    void *dm, *dt;
    C3B_CUDA(cudaMalloc(&dm, (size_t)nmodes * sizeof(MmaProbeMode)));
    C3B_CUDA(cudaMalloc(&dt, (size_t)nmodes * 16));
    C3B_CUDA(cudaMemcpy(dm, modes, (size_t)nmodes * sizeof(MmaProbeMode), cudaMemcpyHostToDevice));
    C3B_CUDA(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 144 * 1024));
    mma_probe_kernel<<<1, 128, 144 * 1024>>>(n, reps, nmodes, (const MmaProbeMode *)dm, (long long *)dt);
    C3B_CUDA(cudaGetLastError());
    C3B_CUDA(cudaDeviceSynchronize());
    C3B_CUDA(cudaMemcpy(timing, dt, (size_t)nmodes * 16, cudaMemcpyDeviceToHost));
    cudaFree(dm); cudaFree(dt);
    return 0;
}

// out_d: [128][n] fp32 (TS-form result), timing4: 5 modes x {issue, done} cycles for `reps` MMAs
extern "C" int c3b_debug_ts_probe(const float *a, const float *b, int n, int reps, float *out_d, int64_t *timing4) {
    if (n % 16 || n < 16 || n > 256) { c3b_set_error("probe: bad n"); return 1; }
    std::vector<uint16_t> ah(128 * 16), bh((size_t)n * 16);
    for (size_t i = 0; i < ah.size(); ++i) ah[i] = c3b_f2op(a[i]);
    for (size_t i = 0; i < bh.size(); ++i) bh[i] = c3b_f2op(b[i]);
    void *da, *db, *dd, *dt;
    C3B_CUDA(cudaMalloc(&da, ah.size() * 2));
    C3B_CUDA(cudaMalloc(&db, bh.size() * 2));
    C3B_CUDA(cudaMalloc(&dd, (size_t)128 * n * 4));
    C3B_CUDA(cudaMalloc(&dt, 10 * 8));
    C3B_CUDA(cudaMemcpy(da, ah.data(), ah.size() * 2, cudaMemcpyHostToDevice));
    C3B_CUDA(cudaMemcpy(db, bh.data(), bh.size() * 2, cudaMemcpyHostToDevice));
    probe_kernel<<<1, 128, 4096 + n * 32 + 128>>>((const __half *)da, (const __half *)db, (float *)dd, n, reps, (long long *)dt);
    C3B_CUDA(cudaGetLastError());
    C3B_CUDA(cudaDeviceSynchronize());
    C3B_CUDA(cudaMemcpy(out_d, dd, (size_t)128 * n * 4, cudaMemcpyDeviceToHost));
    C3B_CUDA(cudaMemcpy(timing4, dt, 80, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dt);
    return 0;
}
