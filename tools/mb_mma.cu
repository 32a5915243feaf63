// mb_mma.cu -- microbenchmark of tcgen05.mma issue / execution for the shapes the fused kernels use.
// One CTA per SM (or one CTA).  Measures, with clock64 in the issuing thread:
//   (a) throughput: cycles per MMA of a long back-to-back stream (M = 128, N, K = 16, kind::f16), A from shared memory
//       (SS) or from tensor memory (TS);
//   (b) latency of one "job": 13 MMAs + commit until the mbarrier completes (what an epilogue warp waits for);
//   (c) the same with `busy` other warps hammering shared memory / the FMA pipe (the epilogue's footprint).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mb_mma_bin tools/mb_mma.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../flowchain_iccv2023_b200/csrc/tc_common.cuh"

using namespace fc::tc;

__global__ void __launch_bounds__(640, 1) k_mb(int N, int ts_mode, int n_stream, int n_job, int reps, int busy, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 64);
    uint8_t* sA = smem + 1024;                 // [128 x 64] fp16 canonical K-major (16 KB)
    uint8_t* sB = sA + 128 * 64 * 2;           // [256 x 64] fp16 (32 KB)
    float* scratch = reinterpret_cast<float*>(sB + 256 * 64 * 2);   // busy-warp playground (16 KB)
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 64 * 2 + 256 * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sA)[i] = 0x3C003C00u;
    fence_async_smem();
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(tslot);
    __shared__ volatile int stop;
    if (tid == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
        if (leader) {      // same issue pattern as the fused kernels: warp-uniform operands, one elected lane
            const uint32_t idesc = umma_idesc_f16(128, N);
            const uint64_t da = umma_desc(smem_u32(sA), 128, 8 * 128), db = umma_desc(smem_u32(sB), 128, 8 * 128);
            const uint32_t ta = tm + 384;      // A operand columns (garbage is fine)
            uint32_t phase = 0;
            // (a) throughput: straight-line groups of 16 MMAs with compile-time descriptor offsets (no per-MMA set-up)
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                for (int i = 0; i < n_stream; i += 16) {
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int ks = q & 3;
                        if (ts_mode) umma_ts(tm, ta + ks * 8, db + (uint64_t)(ks * 16), idesc, 1u);
                        else umma_bf16(tm, da + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc, 1u);
                    }
                }
                umma_commit(bar);
                mbar_wait(bar, phase); phase ^= 1u;
            }
            long long t1 = clock64();
            out[blockIdx.x * 4 + 0] = (t1 - t0);
            // (b) latency of one job
            long long acc_issue = 0, acc_done = 0;
            for (int r = 0; r < reps; ++r) {
                long long a0 = clock64();
                if (n_job == 13) {
#pragma unroll
                    for (int i = 0; i < 13; ++i) {
                        const int ks = i & 3;
                        if (ts_mode) umma_ts(tm, ta + ks * 8, db + (uint64_t)(ks * 16), idesc, i > 0 ? 1u : 0u);
                        else umma_bf16(tm, da + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc, i > 0 ? 1u : 0u);
                    }
                } else {
                    for (int i = 0; i < n_job; ++i) {
                        const int ks = i & 3;
                        if (ts_mode) umma_ts(tm, ta + ks * 8, db + (uint64_t)(ks * 16), idesc, i > 0 ? 1u : 0u);
                        else umma_bf16(tm, da + (uint64_t)(ks * 16), db + (uint64_t)(ks * 16), idesc, i > 0 ? 1u : 0u);
                    }
                }
                umma_commit(bar);
                long long a1 = clock64();
                mbar_wait(bar, phase); phase ^= 1u;
                long long a2 = clock64();
                acc_issue += a1 - a0; acc_done += a2 - a0;
            }
            out[blockIdx.x * 4 + 1] = acc_issue;
            out[blockIdx.x * 4 + 2] = acc_done;
            stop = 1;
        }
        __syncwarp();
    } else if (warp <= busy) {
        // epilogue-like load: tensor-memory loads + FMA/MUFU + 16-byte shared stores
        float acc = 0.f;
        const uint32_t tb = tm + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t sa = smem_u32(scratch) + (uint32_t)(tid & 255) * 16;
        while (!stop) {
            float v[8];
            tmem_ld8(tb + 256 + (warp & 7) * 8, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 8; ++i) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v[i])); acc = fmaf(e, 1.0001f, acc); }
            asm volatile("st.shared.v4.f32 [%0], {%1,%1,%1,%1};" ::"r"(sa), "f"(acc) : "memory");
        }
        if (acc == 123.456f) out[0] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main(int argc, char** argv) {
    const int grid = argc > 1 ? atoi(argv[1]) : 1;
    long long* d;
    cudaMalloc(&d, sizeof(long long) * 4 * grid);
    const int smem = 1024 + 16384 + 32768 + 16384;
    cudaFuncSetAttribute(k_mb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("# grid %d; cycles per MMA (stream of 256, reps 20) | job of 13 MMAs: issue cycles, issue+complete cycles\n", grid);
    for (int busy : {0, 8, 16}) for (int ts = 0; ts < 2; ++ts) for (int N : {48, 64, 80, 128, 256}) {
        const int reps = 20, n_stream = 256, n_job = 13;
        cudaMemset(d, 0, sizeof(long long) * 4 * grid);
        k_mb<<<grid, 640, smem>>>(N, ts, n_stream, n_job, reps, busy, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(4 * grid);
        cudaMemcpy(h.data(), d, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
        printf("busy_warps %2d  %s  N %3d : %7.1f cyc/MMA | job issue %6.0f  job done %6.0f\n", busy, ts ? "TS" : "SS", N,
               (double)h[0] / (reps * n_stream), (double)h[1] / reps, (double)h[2] / reps);
    }
    return 0;
}
