// mb_tmem.cu -- microbenchmark (development aid): what an epilogue warp pays for its tensor-memory traffic.
// Per warp, averaged over `reps`: cycles of  tcgen05.ld.x16 + wait::ld,  4 x tcgen05.st.x4 + wait::st,  and the same store
// followed by tcgen05.fence::before_thread_sync + mbarrier.arrive -- with 1 / 8 / 16 warps doing it concurrently and with
// or without a back-to-back TS-mode MMA stream (M = 128, N = 64) on the tensor pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mb_tmem_bin tools/mb_tmem.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../flowchain_iccv2023_b200/csrc/tc_common.cuh"
#include "../flowchain_iccv2023_b200/csrc/tc_f16x3.cuh"

using namespace fc::tc;
using namespace fc::t2;

__global__ void __launch_bounds__(640, 1) k_mb(int nw, int mma_on, int reps, long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar2 = bar + 1;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 64);
    uint8_t* sB = smem + 1024;           // [64 x 64] fp16
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (64 * 64 * 2) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sB)[i] = 0x3C003C00u;
    fence_async_smem();
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar2, 1 << 19); mbar_fence_init(); }
    if (warp == 16) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(tslot);
    __shared__ volatile int stop;
    __shared__ int done_cnt;
    if (tid == 0) { stop = 0; done_cnt = 0; }
    __syncthreads();
    if (warp == 17) {
        uint32_t leader;
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
        if (leader && mma_on) {
            const uint32_t idesc = umma_idesc_f16(128, 64);
            const uint64_t db = umma_desc(smem_u32(sB), 128, 8 * 128);
            uint32_t phase = 0;
            long long f_after = 0, f_before = 0, n = 0;
            while (!stop) {
#pragma unroll
                for (int q = 0; q < 16; ++q) umma_ts(tm + 448, tm + 384 + (q & 3) * 8, db + (uint64_t)((q & 3) * 16), idesc, 1u);
                umma_commit(bar);
                // what do the tcgen05 fences cost the issuing thread while its own MMAs are still in flight?
                long long c0 = clock64();
                tc_fence_after();
                long long c1 = clock64();
                tc_fence_before();
                long long c2 = clock64();
                f_after += c1 - c0; f_before += c2 - c1; ++n;
                mbar_wait(bar, phase); phase ^= 1u;
            }
            out[8] = f_after / (n ? n : 1); out[9] = f_before / (n ? n : 1);
        }
        __syncwarp();
    } else if (warp < nw) {
        const uint32_t tb = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2) * 64;   // 16 warps -> 4 x 64 columns
        long long t_ld = 0, t_st = 0, t_sta = 0, t_stc = 0;
        float keep = 0.f;
        for (int r = 0; r < reps; ++r) {
            float v[16];
            long long a0 = clock64();
            tmem_ld16(tb, v);
            tmem_wait_ld();
            long long a1 = clock64();
            uint32_t h[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) h[i] = __float_as_uint(v[i] + v[i + 4]);
            tmem_st4(tb, h); tmem_st4(tb + 4, h); tmem_st4(tb + 8, h); tmem_st4(tb + 12, h);
            tmem_wait_st();
            long long a2 = clock64();
            tmem_st4(tb + 16, h); tmem_st4(tb + 20, h); tmem_st4(tb + 24, h); tmem_st4(tb + 28, h);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar2);
            long long a3 = clock64();
            // store, then ~200 independent FMAs, then wait::st: is the store latency hidden by the math?
            tmem_st4(tb + 32, h); tmem_st4(tb + 36, h); tmem_st4(tb + 40, h); tmem_st4(tb + 44, h);
            float f = v[0];
#pragma unroll
            for (int i = 0; i < 200; ++i) f = fmaf(f, 1.0001f, 0.5f);
            keep += f;
            long long a4 = clock64();
            tmem_wait_st();
            long long a5 = clock64();
            t_ld += a1 - a0; t_st += a2 - a1; t_sta += a3 - a2; t_stc += a5 - a4;
        }
        if (lane == 0 && warp == 0) { out[0] = t_ld; out[1] = t_st; out[2] = t_sta; out[3] = t_stc; }
        if (keep == 123.f) out[4] = 1;
        __syncwarp();
        if (lane == 0 && atomicAdd(&done_cnt, 1) == nw - 1) stop = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tm, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 128);
    const int smem = 1024 + 8192;
    cudaFuncSetAttribute(k_mb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("# cycles per op, one CTA: ld.x16+wait | 4 x st.x4+wait | 4 x st.x4+wait+fence+arrive | wait::st after 200 FMAs of cover\n");
    for (int mma = 0; mma < 2; ++mma) for (int nw : {1, 4, 8, 16}) {
        const int reps = 2000;
        cudaMemset(d, 0, 128);
        k_mb<<<1, 640, smem>>>(nw, mma, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        long long h[10];
        cudaMemcpy(h, d, 80, cudaMemcpyDeviceToHost);
        printf("mma %d  warps %2d : ld %6.1f | st %6.1f | st+arrive %6.1f | covered wait %6.1f\n", mma, nw, (double)h[0] / reps, (double)h[1] / reps,
               (double)h[2] / reps, (double)h[3] / reps);
        if (mma) printf("        issuer thread, 16 MMAs in flight: tcgen05.fence::after_thread_sync %lld cycles, ::before_thread_sync %lld cycles\n", h[8], h[9]);
    }
    return 0;
}
