// kernels_tc3.cuh -- f16x3 / f16 tensor-core per-step density with TWO resident 128-row tiles per CTA.
//
// Split-fp16 operands, tcgen05.mma, fp32 accumulate (arithmetic helpers and the weight image: tc_f16x3.cuh), organised so
// that two independent tiles share one SM and 256 tensor-memory columns each.  Two groups of 8 epilogue warps own a tile
// each; the service warpgroup is four MMA issuers, one per (net, tile), and the issuers of the last tile also feed the
// TMA weight ring (one chunk serves both tiles: half the L2 traffic per row).  The groups' epilogues overlap each other
// and the other tile's MMAs, which hides most of the MMA -> epilogue -> MMA latency chain that bounds a single tile.
// Two tensor-memory layouts (see the column map below): hidden operands in tensor memory, converted in place over
// ping-pong accumulator regions (layers <= 64 wide), or hidden operands in shared memory (layers up to 80 wide, chain
// mode, sampler).
#pragma once
#include "kernels_fp32.cuh"
#include "tc_f16x3.cuh"


namespace fc {

// tensor-memory column map of one tile (256 columns per tile), two layouts:
//  HTS = false (layers up to 80 wide, chain mode, sampler): accumulators + the narrow FIRST-layer operands in tensor
//        memory (TS-mode MMAs), hidden-layer operands in shared memory (SS-mode MMAs);
//  HTS = true  (per-step launches whose layers are all <= 64 wide): hidden-layer operands in tensor memory, first-layer
//        operands (K <= 32) in shared memory.  An SS-mode MMA pays ~44 cycles for reading its [128 x 16] A tile from
//        shared memory on top of the N/2 cycles of math (measured, tools/mb_mma.cu: N = 64: 76 cycles SS vs 43 TS); the
//        hidden layers are 80 % of the MMAs, so they get the TS slot.  Each net owns TWO 64-column regions that its
//        jobs use in turn (job k accumulates into region k & 1): the epilogue converts the 16 accumulator columns of a
//        K-step IN PLACE into that K-step's fp16 operand (hi: 8 columns, lo: the next 8), and the next job reads it
//        from there while accumulating into the other region.
constexpr uint32_t kT3Acc0 = 0, kT3Acc1 = 80;            // HTS = false: fp32 accumulators of net 0 / 1 (N <= 80)
constexpr uint32_t kT3InHi = 160, kT3InLo = 184;         // [z|w, cond, prefix] operand, K <= 48
constexpr uint32_t kT3UHi = 208, kT3ULo = 232;           // [u, mx] operand, K <= 48
constexpr uint32_t kT3NetCols = 128, kT3RegCols = 64;   // HTS = true: net n, job parity r -> columns [128 n + 64 r, + 64)
constexpr int kT3InuK = 32;                              // HTS = true: widest first-layer operand, [128 x 32] fp16 tiles in smem
constexpr uint32_t kT3TileCols = 256;
constexpr int kT3Threads = 16 * 32 + 128;                 // 2 x 8 compute warps + a service warpgroup: 4 MMA issuers (net x tile)
constexpr int kT3ComputeRegs = 104, kT3ServiceRegs = 56;   // setmaxnreg split of the register file (see the kernel)
constexpr int kT3OnesOff = 12288;                         // constant [128 x 16] fp16 tile, columns 0..2 = 1 (bias fold)
constexpr int kT3Header = 16384;                          // barriers, step table, row-partial buffers, bias ring, ones tile
constexpr int kT3BiasEntry = 768;                         // bytes per (net, job & 3) entry of the bias ring

struct T3Dims {
    int32_t slot_bytes;      // ring slot (>= largest chunk of the launched steps), multiple of 128
    int32_t nslots_log2;     // ring depth (1 -> 2 slots, 2 -> 4 slots)
    int32_t Wmax;            // widest hidden operand of the launched steps (64 or 80): shared-memory operand stride
    int32_t T, C0;
    int32_t step_lo, step_hi;   // inclusive range of steps (per-step mode) / result indices (chain mode) this launch covers
    int32_t seq;                // 1: chain mode (flow.log_prob_sequential): item (tiles, j) applies net[j], ..., net[klo];
                                // 2: sampler (flow.sample_with_log_prob): item = tiles, inverse steps 0 .. T-1
    int32_t klo, jlo;           // chain mode: lowest chain step applied, first result index written
    int32_t no_fold;            // 1: ignore the chunks' bias tiles (launches whose steps mix folded and plain widths)
    int32_t single;             // 1: one tile per item (group 1 and its issuers idle): small launches that would leave SMs empty otherwise
};

// development aid (-DFC_T3_TRACE=1): lane 0 of every warp of CTA 0 logs (tag << 48 | clock) events of one launch
#ifndef FC_T3_TRACE
#define FC_T3_TRACE 0
#endif
#if FC_T3_TRACE
__device__ unsigned long long g_t3trace[20][1024];
#endif
#if FC_T3_TRACE
#define T3_TR(tag)                                                                                                      \
    if (FC_T3_TRACE && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trace_n < 1024) {                                 \
        long long now_;                                                                                                 \
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(now_));                                                            \
        g_t3trace[threadIdx.x >> 5][trace_n++] = ((unsigned long long)(tag) << 48) | ((unsigned long long)now_ & 0xFFFFFFFFFFFFull); \
    }
#else
#define T3_TR(tag)
#endif

namespace t3 {
using namespace tc;
using namespace t2;

struct Bars3 {
    uint64_t full[4], empty[4];
    uint64_t mma_done[2][2];     // [tile][net]
    uint64_t opnd[2][2];         // [tile][net]
    uint64_t fin[2];             // [tile]
};

__device__ __forceinline__ void sts128(uint32_t saddr, const uint32_t* v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// 8 activated values -> hi / lo fp16 rows of the next layer's shared-memory operand tile
template <int NPASS, int ACT, bool DOT, bool WRITE, bool BIAS, bool HTS = false>
__device__ __forceinline__ void epi_chunk3(float* u, int c0, const float4 b0, const float4 b1, uint32_t wlast, uint32_t ohi, uint32_t olo, uint32_t koff, float& dot) {
    if (BIAS) { u[0] += b0.x; u[1] += b0.y; u[2] += b0.z; u[3] += b0.w; u[4] += b1.x; u[5] += b1.y; u[6] += b1.z; u[7] += b1.w; }
    if (ACT == ACT_T) {
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
            if (NPASS == 3) tanh_acc2(u[q], u[q + 1]);
            else { asm("tanh.approx.f32 %0, %0;" : "+f"(u[q])); asm("tanh.approx.f32 %0, %0;" : "+f"(u[q + 1])); }
        }
    } else if (DOT) {
#pragma unroll
        for (int q = 0; q < 8; ++q) u[q] = fmaxf(u[q], 0.0f);
    }
    if (DOT) {
        const float4 w0 = lds128(wlast + (uint32_t)c0 * 4), w1 = lds128(wlast + (uint32_t)c0 * 4 + 16);
        dot = fmaf(u[0], w0.x, dot); dot = fmaf(u[1], w0.y, dot); dot = fmaf(u[2], w0.z, dot); dot = fmaf(u[3], w0.w, dot);
        dot = fmaf(u[4], w1.x, dot); dot = fmaf(u[5], w1.y, dot); dot = fmaf(u[6], w1.z, dot); dot = fmaf(u[7], w1.w, dot);
    }
    if (WRITE) {
        uint32_t hi[4], lo[4];
        split8<NPASS, ACT == ACT_R>(u, hi, lo);
        if (HTS) {      // next layer's A operand in tensor memory: 8 fp16 = 4 packed columns of this thread's lane
            tmem_st4(ohi + (uint32_t)(c0 >> 1), hi);
            if (NPASS == 3) tmem_st4(olo + (uint32_t)(c0 >> 1), lo);
        } else {        // canonical K-major tile in shared memory: 8 consecutive K elements of one row = one 16-byte store
            const uint32_t off = koff + (uint32_t)(c0 >> 3) * 128u;
            sts128(ohi + off, hi);
            if (NPASS == 3) sts128(olo + off, lo);
        }
    }
}

// koff = (row/8) * (W/8) * 128 + (row%8) * 16 : byte offset of this thread's row inside a [128 x W] canonical tile
// BIAS = false: the bias is already inside the accumulator (folded chunk, see the issuer)
template <int NPASS, int ACT, bool DOT, bool WRITE, bool BIAS, bool HTS = false>
__device__ __noinline__ float epi_hidden3(uint32_t acc, uint32_t ohi, uint32_t olo, uint32_t bias, uint32_t wlast, int part, int W, uint32_t koff) {
    constexpr int CSTEP = 16;
    float dot = 0.0f;
#pragma unroll 1
    for (int c0 = part * 8; c0 < W; c0 += 2 * CSTEP) {
        const int c1 = c0 + CSTEP;
        float v0[8], v1[8];
        tmem_ld8(acc + (uint32_t)c0, v0);
        if (c1 < W) tmem_ld8(acc + (uint32_t)c1, v1);
        // the bias loads ride under the tensor-memory load latency (the ring entry has 16 floats of slack past W)
        float4 b00 = make_float4(0.f, 0.f, 0.f, 0.f), b01 = b00, b10 = b00, b11 = b00;
        if (BIAS) {
            b00 = lds128(bias + (uint32_t)c0 * 4); b01 = lds128(bias + (uint32_t)c0 * 4 + 16);
            b10 = lds128(bias + (uint32_t)c1 * 4); b11 = lds128(bias + (uint32_t)c1 * 4 + 16);
        }
        tmem_wait_ld();
        epi_chunk3<NPASS, ACT, DOT, WRITE, BIAS, HTS>(v0, c0, b00, b01, wlast, ohi, olo, koff, dot);
        if (c1 < W) epi_chunk3<NPASS, ACT, DOT, WRITE, BIAS, HTS>(v1, c1, b10, b11, wlast, ohi, olo, koff, dot);
    }
    return dot;
}

// Epilogue of a layer that is exactly 16 * NB columns wide: this thread's NB chunks are processed in unguarded pairs,
// so the compiler interleaves the 16 independent value chains of a pair (the guarded loop above keeps every chunk in
// its own basic block); all NB at once measured slower than pairs.
template <int NPASS, int ACT, bool DOT, bool WRITE, bool BIAS, int NB, bool HTS = false>
__device__ __noinline__ float epi_hidden3_nb(uint32_t acc, uint32_t ohi, uint32_t olo, uint32_t bias, uint32_t wlast, int part, uint32_t koff) {
    constexpr int CSTEP = 16;
    float dot = 0.0f;
    const int c0 = part * 8;
#pragma unroll
    for (int j = 0; j < NB; j += 2) {
        const int ca = c0 + CSTEP * j, cb = ca + CSTEP;
        float v0[8], v1[8];
        tmem_ld8(acc + (uint32_t)ca, v0);
        if (j + 1 < NB) tmem_ld8(acc + (uint32_t)cb, v1);
        float4 b00 = make_float4(0.f, 0.f, 0.f, 0.f), b01 = b00, b10 = b00, b11 = b00;
        if (BIAS) {
            b00 = lds128(bias + (uint32_t)ca * 4); b01 = lds128(bias + (uint32_t)ca * 4 + 16);
            if (j + 1 < NB) { b10 = lds128(bias + (uint32_t)cb * 4); b11 = lds128(bias + (uint32_t)cb * 4 + 16); }
        }
        tmem_wait_ld();
        epi_chunk3<NPASS, ACT, DOT, WRITE, BIAS, HTS>(v0, ca, b00, b01, wlast, ohi, olo, koff, dot);
        if (j + 1 < NB) epi_chunk3<NPASS, ACT, DOT, WRITE, BIAS, HTS>(v1, cb, b10, b11, wlast, ohi, olo, koff, dot);
    }
    return dot;
}
template <int NPASS, int ACT, bool DOT, bool WRITE, bool BIAS, bool HTS = false>
__device__ __forceinline__ float epi_layer3(uint32_t acc, uint32_t ohi, uint32_t olo, uint32_t bias, uint32_t wlast, int part, int W, uint32_t koff) {
    if (W == 64) return epi_hidden3_nb<NPASS, ACT, DOT, WRITE, BIAS, 4, HTS>(acc, ohi, olo, bias, wlast, part, koff);
    if (W == 48) return epi_hidden3_nb<NPASS, ACT, DOT, WRITE, BIAS, 3, HTS>(acc, ohi, olo, bias, wlast, part, koff);
    if (!HTS && W == 80) return epi_hidden3_nb<NPASS, ACT, DOT, WRITE, BIAS, 5, HTS>(acc, ohi, olo, bias, wlast, part, koff);
    return epi_hidden3<NPASS, ACT, DOT, WRITE, BIAS, HTS>(acc, ohi, olo, bias, wlast, part, W, koff);
}
// e^x = 2^(x log2 e) on the MUFU (2 ulp; denormal results flush to zero, which only ever multiplies the noise)
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ void lds128u(uint32_t saddr, uint32_t* v) {
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(saddr));
}

// per-thread view of its group (tile)
struct X3 {
    int row, part, lane, g;
    uint32_t tb;             // tensor-memory base of the tile incl. this warp's lane quadrant
    Bars3* bars;
    uint8_t* slots;
    int slot_bytes;
    uint32_t nslots_mask, nslots_log2;
    uint32_t chunk_it;
    uint32_t mma_cnt0, mma_cnt1;
    float* part_buf;         // this group's [2][2][128] row-partial exchange
    uint32_t bias_ring;      // shared address of the bias ring: [net][job & 3] x kT3BiasEntry
    uint32_t h_base;         // shared address of this tile's hidden operand tiles: [net][part(hi,lo)] x h_stride
    uint32_t h_stride;
};

__device__ __forceinline__ const uint8_t* slot_ptr3(const X3& x, uint32_t it) { return x.slots + (size_t)(it & x.nslots_mask) * x.slot_bytes; }
// waits for the next job of `net` and returns the tensor-memory address of its accumulator (HTS: the region alternates)
template <bool HTS>
__device__ __forceinline__ uint32_t wait_mma3(X3& x, int net) {
    uint32_t cnt;
    if (net == 0) { cnt = x.mma_cnt0++; mbar_wait(&x.bars->mma_done[x.g][0], cnt & 1u); }
    else { cnt = x.mma_cnt1++; mbar_wait(&x.bars->mma_done[x.g][1], cnt & 1u); }
    tc_fence_after();
    if (HTS) return x.tb + (uint32_t)net * kT3NetCols + (cnt & 1u) * kT3RegCols;
    return x.tb + (net ? kT3Acc1 : kT3Acc0);
}
// this warp's tensor-memory accesses and shared-memory operand stores are complete -> tell the issuers
__device__ __forceinline__ void warp_arrive3(const X3& x, uint64_t* bar) {
    tmem_wait_ld();
    tmem_wait_st();
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (x.lane == 0) mbar_arrive(bar);
}
// HTS epilogue of one net-layer.  The thread (row, column part p) owns the K-steps p, p + 2, ... of the NEXT layer: it
// loads the 16 accumulator columns of a K-step, applies the activation and writes the hi / lo fp16 operand of that
// K-step over them (hi: columns [16 ks, +8), lo: [16 ks + 8, +8) -- exactly what a TS-mode MMA reads), then arrives on
// `done`.  No shared memory is written, so the arrival needs no fence.proxy.async (a MEMBAR.ALL.CTA in the SASS).
// (Letting the next job start on the first two K-steps while the last ones are still being converted was tried -- an
// extra arrival per layer and a two-stage issue -- and measured slower, 48.7 vs 46.4 ms: the tensor pipe is busy with
// the other three job streams at that moment anyway.)
template <int NPASS, int ACT, bool DOT, bool WRITE, bool BIAS>
__device__ __noinline__ float epi_inplace3(const X3& x, uint32_t acc, uint32_t bias, uint32_t wlast, int W, uint64_t* done) {
    float dot = 0.0f;
    const int nk = W >> 4;
#pragma unroll 1
    for (int ks = x.part; ks < nk; ks += 2) {
        const int c0 = ks * 16;
        float v[16];
        tmem_ld16(acc + (uint32_t)c0, v);
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0, b2 = b0, b3 = b0;
        if (BIAS) {
            b0 = lds128(bias + (uint32_t)c0 * 4); b1 = lds128(bias + (uint32_t)c0 * 4 + 16);
            b2 = lds128(bias + (uint32_t)c0 * 4 + 32); b3 = lds128(bias + (uint32_t)c0 * 4 + 48);
        }
        tmem_wait_ld();
        // epi_chunk3 stores 4 packed columns at ohi + col / 2: ohi = acc + c0 / 2 puts the two chunks at acc + c0 + {0, 4}
        const uint32_t ohi = acc + (uint32_t)(c0 >> 1), olo = ohi + 8u;
        epi_chunk3<NPASS, ACT, DOT, WRITE, BIAS, true>(v, c0, b0, b1, wlast, ohi, olo, 0u, dot);
        epi_chunk3<NPASS, ACT, DOT, WRITE, BIAS, true>(v + 8, c0 + 8, b2, b3, wlast, ohi, olo, 0u, dot);
    }
    if (WRITE) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (x.lane == 0) mbar_arrive(done);
    }
    return dot;
}
__device__ __forceinline__ float row_sum3(const X3& x, int slot, float v) {      // slot in {0, 1}
    x.part_buf[(slot * 2 + x.part) * 128 + x.row] = v;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + x.g) : "memory");
    return x.part_buf[(slot * 2) * 128 + x.row] + x.part_buf[(slot * 2 + 1) * 128 + x.row];
}

}  // namespace t3

namespace t3 {
using namespace tc;
using namespace t2;
// Shared-memory map of the kernel and the per-launch constants every role needs.
struct T3Smem {
    Bars3* bars;
    T2Step* cur;
    uint8_t *bias_ring, *hbuf, *slots, *ones;
    float* part_all;
    uint32_t h_stride, smask;
};
__device__ __forceinline__ T3Smem t3_smem(uint8_t* smem, const T3Dims& dm) {
    T3Smem m;
    m.bars = reinterpret_cast<Bars3*>(smem);
    m.cur = reinterpret_cast<T2Step*>(smem + 256);
    m.part_all = reinterpret_cast<float*>(smem + 2048);          // 2 groups x (2 x 2 x 128 floats = 2 KB)
    m.bias_ring = smem + 6144;                                     // [net][job & 3] x 768 B: fp32 tails of the weight chunks
    m.ones = smem + kT3OnesOff;
    m.h_stride = (uint32_t)(128 * dm.Wmax * 2);                    // one [128 x Wmax] fp16 tile
    m.hbuf = smem + kT3Header;                                     // [tile][net][hi,lo]
    m.slots = m.hbuf + (size_t)8 * m.h_stride;
    m.smask = (1u << dm.nslots_log2) - 1u;
    return m;
}
// All roles walk the same item sequence; when the step changes the whole CTA reloads the step table.
__device__ __forceinline__ void t3_load_step(const T2Step* steps, int step, T2Step* cur, int tid) {
    __syncthreads();
    const int nw = sizeof(T2Step) / 4;
    const int* s = reinterpret_cast<const int*>(steps + step);
    int* d = reinterpret_cast<int*>(cur);
    for (int i = tid; i < nw; i += kT3Threads) d[i] = __ldg(s + i);
    __syncthreads();
}
// Item = (tile pair, jstep).  Per-step mode: one step (= jstep).  Chain mode: steps jstep, jstep-1, ..., klo with the
// point carried in registers.  Every role runs exactly these loops so that the step-table reloads line up.
#define T3_ITEM_BEGIN(SEQ)                                                                            \
    int cur_step = -1;                                                                              \
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {                      \
        const int jstep = dm.step_hi - (int)(item / pairs_per_step); /* heavy (late) first */       \
        const long long pair = item % pairs_per_step;                                               \
        const int nsteps = (SEQ) == 0 ? 1 : ((SEQ) == 1 ? jstep - dm.klo + 1 : dm.T);           \
        (void)pair;
#define T3_STEP_BEGIN(SEQ)                                                                            \
        for (int si = 0; si < nsteps; ++si) {                                                       \
            const int step = (SEQ) == 2 ? si : jstep - si;                                       \
            if (step != cur_step) { t3_load_step(steps, step, m.cur, (int)threadIdx.x); cur_step = step; } \
            const T2Step& ts = *m.cur;
#define T3_STEP_END }
#define T3_ITEM_END }

// The image holds a step's chunks in forward order [density r][blocks 0 .. n-1][density q]; the sampler runs the inverse
// [density q][blocks n-1 .. 0][density r] with an identical job table (same shapes), so only the chunk index is mapped.
__device__ __forceinline__ int t3_chunk_of_job(const T2Step& ts, int ci, bool inverse) {
    if (!inverse) return ci;
    const int L2 = 2 * (ts.n_hidden + 1), nb6 = 6 + ts.n_blocks * L2;
    if (ci < 6) return nb6 + ci;
    if (ci < nb6) { const int bi = (ci - 6) / L2; return 6 + (ts.n_blocks - 1 - bi) * L2 + (ci - 6 - bi * L2); }
    return ci - nb6;
}

// All MMAs of one job (= net-layer of one tile): D = A_hi W_hi + A_lo W_hi + A_hi W_lo over NK K-steps of 16.
// TS: A operand in tensor memory (columns ta_hi / ta_lo, 8 columns per K-step); otherwise shared-memory descriptors.
template <int NPASS, int NK, bool TS>
__device__ __forceinline__ void issue_job(uint32_t acc, uint32_t ta_hi, uint32_t ta_lo, uint64_t ahi, uint64_t alo, uint64_t dhi, uint64_t dlo,
                                          uint32_t idesc, uint32_t acc0) {
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) {
        if (TS) umma_ts(acc, ta_hi + ks * 8, dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
        else umma_bf16(acc, ahi + (uint64_t)(ks * 16), dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
    }
    if (NPASS == 3) {
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            if (TS) umma_ts(acc, ta_lo + ks * 8, dhi + (uint64_t)(ks * 16), idesc, 1u);
            else umma_bf16(acc, alo + (uint64_t)(ks * 16), dhi + (uint64_t)(ks * 16), idesc, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            if (TS) umma_ts(acc, ta_hi + ks * 8, dlo + (uint64_t)(ks * 16), idesc, 1u);
            else umma_bf16(acc, ahi + (uint64_t)(ks * 16), dlo + (uint64_t)(ks * 16), idesc, 1u);
        }
    }
}
template <int NPASS, bool TS>
__device__ __forceinline__ void issue_job_loop(int nk, uint32_t acc, uint32_t ta_hi, uint32_t ta_lo, uint64_t ahi, uint64_t alo, uint64_t dhi,
                                               uint64_t dlo, uint32_t idesc, uint32_t acc0) {
#pragma unroll 1
    for (int ks = 0; ks < nk; ++ks) {
        if (TS) umma_ts(acc, ta_hi + ks * 8, dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
        else umma_bf16(acc, ahi + (uint64_t)(ks * 16), dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
    }
    if (NPASS == 3) {
#pragma unroll 1
        for (int ks = 0; ks < nk; ++ks) {
            if (TS) umma_ts(acc, ta_lo + ks * 8, dhi + (uint64_t)(ks * 16), idesc, 1u);
            else umma_bf16(acc, alo + (uint64_t)(ks * 16), dhi + (uint64_t)(ks * 16), idesc, 1u);
        }
#pragma unroll 1
        for (int ks = 0; ks < nk; ++ks) {
            if (TS) umma_ts(acc, ta_hi + ks * 8, dlo + (uint64_t)(ks * 16), idesc, 1u);
            else umma_bf16(acc, ahi + (uint64_t)(ks * 16), dlo + (uint64_t)(ks * 16), idesc, 1u);
        }
    }
}

// K-steps [KLO, KHI) of a job whose A operand sits in the in-place HTS layout (K-step ks: hi at ta + 16 ks, lo 8 further)
template <int NPASS, int KLO, int KHI>
__device__ __forceinline__ void issue_inplace(uint32_t acc, uint32_t ta, uint64_t dhi, uint64_t dlo, uint32_t idesc, uint32_t acc0) {
#pragma unroll
    for (int ks = KLO; ks < KHI; ++ks) umma_ts(acc, ta + ks * 16, dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
    if (NPASS == 3) {
#pragma unroll
        for (int ks = KLO; ks < KHI; ++ks) umma_ts(acc, ta + ks * 16 + 8, dhi + (uint64_t)(ks * 16), idesc, 1u);
#pragma unroll
        for (int ks = KLO; ks < KHI; ++ks) umma_ts(acc, ta + ks * 16, dlo + (uint64_t)(ks * 16), idesc, 1u);
    }
}

// ===== MMA issuers, one warp per (net, tile): four independent job streams.  Measured with the event trace
// (tools/t3_trace.py): a service warp retires roughly one instruction per 8-10 cycles (a dependent scalar stream next to
// four busy epilogue warps), so with ONE issuer per net the per-chunk bookkeeping plus the issue of two tiles' jobs took
// 3.7 k cycles and paced the whole pipeline; per (net, tile) the same work has a full layer period to itself.
// The issuer of a net's LAST tile also feeds that net's half of the weight ring: at the top of iteration k it requests
// the chunks below k + D (D = ring slots per net), each as two bulk copies on the slot's `full` barrier -- the chunk into
// the slot and its fp32 tail (biases, last-row vector) straight into the bias ring entry [net][job & 3], so that the
// epilogues keep the tail after the slot has been handed back and no warp copies it.  A slot is handed back by the
// tcgen05.commit of the MMAs that read it (`empty` counts the tiles).  With one step per item the requests run on into
// the next item of the same step.
// First layers read their operand from tensor memory (TS) and hidden layers from shared memory (SS), or the other way
// round under HTS. =====
template <int NPASS, bool TS>
__device__ __forceinline__ void issue_dispatch(int nk, uint32_t acc, uint32_t ta_hi, uint32_t ta_lo, uint64_t ahi, uint64_t alo, uint64_t dhi,
                                               uint64_t dlo, uint32_t idesc, uint32_t acc0) {
    switch (nk) {
        case 1: issue_job<NPASS, 1, TS>(acc, ta_hi, ta_lo, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 2: issue_job<NPASS, 2, TS>(acc, ta_hi, ta_lo, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 3: issue_job<NPASS, 3, TS>(acc, ta_hi, ta_lo, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 4: issue_job<NPASS, 4, TS>(acc, ta_hi, ta_lo, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 5: issue_job<NPASS, 5, TS>(acc, ta_hi, ta_lo, ahi, alo, dhi, dlo, idesc, acc0); break;
        default: issue_job_loop<NPASS, TS>(nk, acc, ta_hi, ta_lo, ahi, alo, dhi, dlo, idesc, acc0); break;
    }
}

template <int NPASS, bool HTS>
__device__ __forceinline__ void t3_issuer(const T2Step* __restrict__ steps, const uint8_t* __restrict__ img, const T3Dims& dm,
                                       long long pairs_per_step, long long total_items, uint8_t* smem, uint32_t tmem_base, int net, int t) {
    const T3Smem m = t3_smem(smem, dm);
    Bars3* bars = m.bars;
    uint8_t *slots = m.slots, *hbuf = m.hbuf;
    const uint32_t smask = m.smask, h_stride = m.h_stride;
    uint32_t ring_it = 0;
    uint32_t opnd_c = 0, fin_c = 0, jobc = 0;      // phase counters of this tile; jobs of this (net, tile) so far
    int trace_n = 0; (void)trace_n;
    const int ntile = dm.single ? 1 : 2;
    const bool active = t < ntile;                                // one tile per item: the tile-1 issuers only follow the step reloads
    const bool producer = t == ntile - 1;
    const int depth = 1 << (dm.nslots_log2 - 1);                  // ring slots of this net
    int carried = 0;                                              // chunks of the coming step already requested (previous item)
    const uint32_t tb0 = tmem_base + (uint32_t)t * kT3TileCols;
    const uint64_t d_ones = umma_desc(smem_u32(m.ones), 128, 256u);
    uint32_t leader;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
    T3_ITEM_BEGIN(dm.seq)
    T3_STEP_BEGIN(dm.seq)
        const uint32_t base = ring_it;
        ring_it += (uint32_t)ts.nchunks;
        if (!active) continue;
        const int nq = ts.nchunks >> 1;                           // chunks of this net in the step
        // may the requests run on into the next item?  (same step table: one step per item, the next item exists and is of this step)
        const bool wrap_ok = nsteps == 1 && item + gridDim.x < total_items && (item + gridDim.x) / pairs_per_step == item / pairs_per_step;
        int k_req = carried;
        carried = 0;
#pragma unroll 1
        for (int ci = net, k = 0; ci < ts.nchunks; ci += 2, ++k, ++jobc) {
            if (producer) {
                int target = k + depth;
                if (!wrap_ok && target > nq) target = nq;
#pragma unroll 1
                for (; k_req < target; ++k_req) {
                    const int rci = net + 2 * (k_req < nq ? k_req : k_req - nq);
                    const uint32_t rit = base + (k_req < nq ? 0u : (uint32_t)ts.nchunks) + (uint32_t)rci;
                    const uint32_t slot = rit & smask;
                    T3_TR(0x50 + net)       // wants the slot for chunk rci
                    mbar_wait(&bars->empty[slot], ((rit >> dm.nslots_log2) & 1u) ^ 1u);
                    T3_TR(0x52 + net)       // slot free: bulk copies issued
                    if (leader) {
                        const int pc = t3_chunk_of_job(ts, rci, dm.seq == 2);
                        const int rn = ts.job_n[rci], rk = ts.job_kp[rci];
                        const uint32_t tiles_b = (uint32_t)(NPASS == 3 ? 2 : 1) * (uint32_t)(rn * rk * 2);
                        const uint32_t tail_b = (uint32_t)ts.chunk_bytes[pc] - tiles_b - (ts.job_fold[rci] ? (uint32_t)rn * 32u : 0u);
                        mbar_expect_tx(&bars->full[slot], (uint32_t)ts.chunk_bytes[pc] + tail_b);
                        bulk_load(slots + (size_t)slot * dm.slot_bytes, img + ts.chunk_off[pc], (uint32_t)ts.chunk_bytes[pc], &bars->full[slot]);
                        bulk_load(m.bias_ring + (size_t)(net * 4 + ((rit >> 1) & 3)) * kT3BiasEntry, img + ts.chunk_off[pc] + tiles_b, tail_b,
                                  &bars->full[slot]);
                    }
                    __syncwarp();
                }
            }
            const uint32_t it = base + (uint32_t)ci;
            const int w = ts.job_wait[ci], src = ts.job_src[ci], Kp = ts.job_kp[ci], N = ts.job_n[ci];
            const bool fold = ts.job_fold[ci] != 0 && !dm.no_fold;
            const int chunk_bytes = ts.chunk_bytes[t3_chunk_of_job(ts, ci, dm.seq == 2)];
            const uint32_t b_hi = smem_u32(slots + (size_t)(it & smask) * dm.slot_bytes);
            const uint32_t idesc = umma_idesc_f16(128, N), sbo = (uint32_t)(Kp >> 3) * 128u;
            const uint64_t dhi = umma_desc(b_hi, 128, sbo), dlo = umma_desc(b_hi + (uint32_t)(N * Kp * 2), 128, sbo);
            const int nk = Kp >> 4;
            // folded bias: acc = ones[128 x 16] * bias_tile[N x 16]^T first, everything else accumulates on top
            const uint64_t d_bias = umma_desc(b_hi + (uint32_t)(chunk_bytes - N * 32), 128, 256u);
            const uint32_t acc0 = fold ? 1u : 0u;
            // HTS: job k of this net accumulates into region k & 1 and reads the operand its predecessor's epilogue left in
            // the other one; first-layer jobs read the [IN | U] tiles from shared memory.  Otherwise: accumulators and the
            // narrow first-layer operands in tensor memory, hidden operands in shared memory.
            const uint32_t acc = HTS ? tb0 + (uint32_t)net * kT3NetCols + (jobc & 1u) * kT3RegCols : tb0 + (net ? kT3Acc1 : kT3Acc0);
            const uint32_t ta = tb0 + (uint32_t)net * kT3NetCols + ((jobc & 1u) ^ 1u) * kT3RegCols;
            const uint32_t ta_hi = tb0 + (src == 0 ? kT3InHi : kT3UHi), ta_lo = tb0 + (src == 0 ? kT3InLo : kT3ULo);
            const uint32_t sa = smem_u32(hbuf) + (uint32_t)(t * 4 + (HTS ? src : net) * 2) * h_stride;
            const uint64_t ahi = umma_desc(sa, 128, sbo), alo = umma_desc(sa + h_stride, 128, sbo);
            T3_TR(0x22)              // descriptors built: waiting for the weight chunk, then the operand
            mbar_wait(&bars->full[it & smask], (it >> dm.nslots_log2) & 1u);
            T3_TR(0x20)              // weight chunk landed
            // ---- critical path: operand of this tile ready -> MMAs -> commit ----
            if (w == 3) { mbar_wait(&bars->fin[t], fin_c & 1u); fin_c++; }
            else { mbar_wait(&bars->opnd[t][net], opnd_c & 1u); opnd_c++; }
            tc_fence_after();
            T3_TR(0x30 + t)          // operand of the tile ready
            // straight-line issue for the K widths that occur (16 * nk): with compile-time operand offsets the MMAs of a job
            // go out back to back (UTCHMMA after UTCHMMA); a counted loop costs ~10 set-up instructions per MMA.
            if (leader) {
                if (fold) umma_bf16(acc, d_ones, d_bias, idesc, 0u);
                if (HTS) {
                    if (src == 2) {
                        switch (nk) {
                            case 1: issue_inplace<NPASS, 0, 1>(acc, ta, dhi, dlo, idesc, acc0); break;
                            case 2: issue_inplace<NPASS, 0, 2>(acc, ta, dhi, dlo, idesc, acc0); break;
                            case 3: issue_inplace<NPASS, 0, 3>(acc, ta, dhi, dlo, idesc, acc0); break;
                            default: issue_inplace<NPASS, 0, 4>(acc, ta, dhi, dlo, idesc, acc0); break;
                        }
                    } else issue_dispatch<NPASS, false>(nk, acc, 0, 0, ahi, alo, dhi, dlo, idesc, acc0);
                } else {
                    if (src < 2) issue_dispatch<NPASS, true>(nk, acc, ta_hi, ta_lo, 0, 0, dhi, dlo, idesc, acc0);
                    else issue_dispatch<NPASS, false>(nk, acc, 0, 0, ahi, alo, dhi, dlo, idesc, acc0);
                }
            }
            if (leader) {
                umma_commit(&bars->mma_done[t][net]);
                umma_commit(&bars->empty[it & smask]);      // this tile's MMAs on the chunk done (the barrier counts the tiles)
            }
            __syncwarp();
            T3_TR(0x40 + t)          // MMAs issued + committed
        }
        carried = k_req > nq ? k_req - nq : 0;
        // consume the final `fin` phase of the tile (the q-density epilogue) to stay in step with its group
        mbar_wait(&bars->fin[t], fin_c & 1u); fin_c++;
    T3_STEP_END
    T3_ITEM_END
}

// ===== compute groups: group g owns tile 2*pair + g; thread = (row, column part) =====
// FOLD_D / FOLD_C: the density / coupling hidden layers of every step of this launch carry folded biases
// INV: the sampler (flow.sample_with_log_prob, fastpredNF.py:586-607): every step runs CIF_step.inverse -- sample u from
// q, undo the blocks last to first, score u under r -- and the point / prefix come from the row's own samples.
// MODE (compile-time copy of T3Dims::seq for this role): 0 per-step density, 1 chain, 2 sampler
template <int NPASS, bool FOLD_D, bool FOLD_C, int MODE, bool HTS>
__device__ __forceinline__ void t3_compute(const T2Step* __restrict__ steps, const RowArgs& a, const SampleOut& o, const T3Dims& dm,
                                        long long pairs_per_step, long long total_items, uint8_t* smem, uint32_t tmem_base) {
    constexpr int NPARTS = 2;
    constexpr bool INV = MODE == 2;
    const T3Smem m = t3_smem(smem, dm);
    Bars3* bars = m.bars;
    uint8_t* bias_ring = m.bias_ring;
    const uint32_t h_stride = m.h_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int trace_n = 0; (void)trace_n;
    X3 x;
    x.lane = lane; x.g = (warp >> 3) & 1; x.part = (warp >> 2) & 1; x.row = (warp & 3) * 32 + lane;
    x.tb = tmem_base + (uint32_t)x.g * kT3TileCols + ((uint32_t)((warp & 3) * 32) << 16);
    x.bars = bars; x.slots = m.slots; x.slot_bytes = dm.slot_bytes; x.nslots_mask = m.smask; x.nslots_log2 = (uint32_t)dm.nslots_log2;
    x.chunk_it = 0; x.mma_cnt0 = 0; x.mma_cnt1 = 0;
    x.part_buf = m.part_all + x.g * 512;
    x.bias_ring = smem_u32(bias_ring);
    x.h_stride = h_stride;
    x.h_base = smem_u32(m.hbuf) + (uint32_t)x.g * 4u * h_stride;
    if (dm.single && x.g == 1) {      // one tile per item: the second group only takes part in the step-table reloads
        T3_ITEM_BEGIN(MODE)
        T3_STEP_BEGIN(MODE)
            (void)ts;
        T3_STEP_END
        T3_ITEM_END
        return;
    }
    T3_ITEM_BEGIN(MODE)
        const int g = x.g;
        long long n = (dm.single ? pair : pair * 2 + g) * 128 + x.row;
        const bool valid = n < a.N;
        if (!valid) n = a.N - 1;
        const long long ag = n / a.PPA, rem = n % a.PPA, pt = rem / a.K;
        float z0, z1, prev0, prev1, lacc = 0.0f, base_lp = 0.0f, cum = 0.0f;
        if (INV) {      // u0 ~ Normal(base_pos, std[0]) (Normal.sample = normal_() * std + mean), its log-density
            const float* pp = src_at(a.base, ag, pt, 0);
            prev0 = __ldg(pp); prev1 = __ldg(pp + 1);
            const float s0 = __ldg(&steps[0].std[0]), s1 = __ldg(&steps[0].std[1]);
            float e0, e1;
            if (a.z0 != nullptr) { e0 = __ldg(a.z0 + n * 2); e1 = __ldg(a.z0 + n * 2 + 1); }
            else { const float4 gq = philox_normal4(a.seed, n, 0xFFFFu, 0u); e0 = gq.x; e1 = gq.y; }
            z0 = __fadd_rn(__fmul_rn(e0, s0), prev0); z1 = __fadd_rn(__fmul_rn(e1, s1), prev1);
            base_lp = normal_lp2(z0, z1, prev0, prev1, s0, s1);
            if (valid && x.part == 0) o.base_lp[n] = base_lp;
        } else {
            load_query(a, ag, pt, jstep, z0, z1);
            // base mean: per-step mode = previous trajectory point; chain mode = base_pos (fastpredNF.py:454-458)
            const float* pp = (MODE == 1 || jstep == 0) ? src_at(a.base, ag, pt, 0) : src_at(a.prefix, ag, pt, jstep - 1);
            prev0 = __ldg(pp); prev1 = __ldg(pp + 1);
        }
    T3_STEP_BEGIN(MODE)
        const int C = ts.C, c = ts.c, Kin = ts.Kin, Ku = ts.Ku, Wd = ts.Wd, CPd = ts.CPd, Hc = ts.Hc;
        if (INV) lacc = 0.0f;       // the sampler reports every step's own log-det term
        // byte offset of this row inside a canonical [128 x W] operand tile, for the two hidden widths of the step
        const uint32_t koff_d = (uint32_t)((x.row >> 3) * (Wd >> 3) * 128 + (x.row & 7) * 16);
        const uint32_t koff_c = (uint32_t)((x.row >> 3) * (Hc >> 3) * 128 + (x.row & 7) * 16);
        // HTS: shared addresses of this row inside the first-layer operand tiles [IN hi, IN lo, U hi, U lo] of its tile
        const uint32_t in_row = x.h_base + (uint32_t)((x.row >> 3) * (Kin >> 3) * 128 + (x.row & 7) * 16);
        const uint32_t u_row = x.h_base + 2u * h_stride + (uint32_t)((x.row >> 3) * (Ku >> 3) * 128 + (x.row & 7) * 16);
        // ---- input operand [z (2), cond (C0), prefix (2*step), 0...] ----
        {
            const float* cp = src_at(a.cond, ag, pt, INV ? step : jstep) - 2;     // chain mode: cond[:, j] for every chain step
            // sampler: the prefix is the row's own earlier samples, read back from the output (written by this group's
            // part-0 threads before the CTA-wide barrier of the step-table reload; coherent loads, not the read-only path)
            const float* pf = INV ? o.seq + n * (long long)(2 * dm.T) : src_at(a.prefix, ag, pt, 0);
            const long long pst = INV ? 2 : a.prefix.st;
            const int c0e = 2 + dm.C0;
            for (int k0 = x.part * 8; k0 < Kin; k0 += 8 * NPARTS) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int col = k0 + i;
                    float val = 0.0f;
                    if (col < 2) val = col == 0 ? z0 : z1;
                    else if (col < c0e) val = __ldg(cp + col);
                    else if (col < c) { const int e = col - c0e; const float* q = pf + (long long)(e >> 1) * pst + (e & 1); val = INV ? __ldcg(q) : __ldg(q); }
                    v[i] = clamp_h(val);
                }
                uint32_t hi[4], lo[4];
                split8<NPASS, false>(v, hi, lo);
                if (HTS) {
                    sts128(in_row + (uint32_t)(k0 >> 3) * 128u, hi);
                    if (NPASS == 3) sts128(in_row + h_stride + (uint32_t)(k0 >> 3) * 128u, lo);
                } else {
                    tmem_st4(x.tb + kT3InHi + (uint32_t)(k0 >> 1), hi);
                    if (NPASS == 3) tmem_st4(x.tb + kT3InLo + (uint32_t)(k0 >> 1), lo);
                }
            }
        }
        warp_arrive3(x, &bars->fin[g]);

#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            // ============ density nets (shift | log-scale): two relu layers ============
#pragma unroll 1
            for (int l = 0; l < 2; ++l) {
#pragma unroll 1
                for (int net = 0; net < 2; ++net) {
                    const uint32_t bias = x.bias_ring + (uint32_t)(net * 4 + ((x.chunk_it >> 1) & 3)) * kT3BiasEntry;
                    T3_TR(0x10 + net)        // waiting for the MMAs of this net
                    const uint32_t acc = wait_mma3<HTS>(x, net);
                    T3_TR(0x12 + net)        // accumulator ready: epilogue starts
                    if (HTS) {
                        epi_inplace3<NPASS, ACT_R, false, true, !FOLD_D>(x, acc, bias, 0u, Wd, &bars->opnd[g][net]);
                    } else {
                        const uint32_t ohi = x.h_base + (uint32_t)(net * 2) * h_stride, olo = ohi + h_stride;
                        epi_layer3<NPASS, ACT_R, false, true, !FOLD_D, false>(acc, ohi, olo, bias, 0u, x.part, Wd, koff_d);
                        warp_arrive3(x, &bars->opnd[g][net]);
                    }
                    x.chunk_it++;
                }
            }
            // ---- L3 outputs: means in acc0[0,CPd), log-stddevs in acc1[0,CPd) ----
            const float* bias_mu = reinterpret_cast<const float*>(bias_ring + (size_t)(0 * 4 + ((x.chunk_it >> 1) & 3)) * kT3BiasEntry);
            const float* bias_ls = reinterpret_cast<const float*>(bias_ring + (size_t)(1 * 4 + ((x.chunk_it >> 1) & 3)) * kT3BiasEntry);
            x.chunk_it += 2;
            T3_TR(0x14)              // waiting for both L3 outputs
            const uint32_t acc_mu = wait_mma3<HTS>(x, 0), acc_ls = wait_mma3<HTS>(x, 1);
            T3_TR(0x15)              // E2 / E3 phase starts
            if (which == 0) {
                // sample u = exp(ls) eps + mu (fastpredNF.py:727-733) -> [u, mx] operand; log r = -C/2 log 2pi - sum ls - 1/2 sum eps^2
                if (INV) {      // modules reversed: the last block's BatchNorm.inverse comes first (TFCondARFlow.py:264-269)
                    const int bl = ts.n_blocks - 1;
                    z0 = fmaf(ts.bn_iscale[bl][0], z0, ts.bn_ishift[bl][0]);
                    z1 = fmaf(ts.bn_iscale[bl][1], z1, ts.bn_ishift[bl][1]);
                    lacc -= ts.bn_ldj[bl];
                }
                const int m0 = ts.masked_dim[INV ? ts.n_blocks - 1 : 0];
                float p_ls = 0.0f, p_e2 = 0.0f;
                for (int d0 = x.part * 8; d0 < Ku; d0 += 8 * NPARTS) {
                    float u[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) u[i] = 0.0f;
                    if (d0 < CPd) {
                        float mu[8], ls[8], e[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) e[i] = 0.0f;
                        tmem_ld8(acc_mu + (uint32_t)d0, mu);
                        tmem_ld8(acc_ls + (uint32_t)d0, ls);
                        if (a.eps != nullptr) {
                            const float* ep = a.eps + n * a.eps_rs + (MODE == 1 ? ts.seq_eps_off + (jstep - step) * C : ts.eps_off) + d0;
#pragma unroll
                            for (int i = 0; i < 8; ++i) if (d0 + i < C) e[i] = __ldg(ep + i);
                        } else {
                            const Normal8 gn = philox_normal8(a.seed, n, MODE == 1 ? (unsigned)(step * 64 + jstep + 4096) : (unsigned)step, (unsigned)(d0 >> 2));
                            e[0] = gn.a.x; e[1] = gn.a.y; e[2] = gn.a.z; e[3] = gn.a.w; e[4] = gn.b.x; e[5] = gn.b.y; e[6] = gn.b.z; e[7] = gn.b.w;
                        }
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            // branch-free: padded columns (d >= C) have zero weights and biases, so mu = ls = 0 there and
                            // masking the noise alone gives u = 0, no contribution to either sum
                            const float m = mu[i] + bias_mu[d0 + i], l = ls[i] + bias_ls[d0 + i];
                            const float en = d0 + i < C ? e[i] : 0.0f;
                            u[i] = clamp_h(fmaf(fast_exp(l), en, m));
                            p_ls += l;
                            p_e2 = fmaf(en, en, p_e2);
                        }
                    }
                    if (d0 == (C & ~7)) {        // mx = z * mask of the first coupling sits at columns C, C+1
                        const int cm = C & 7;
                        const float a0 = m0 == 0 ? clamp_h(z0) : 0.0f, a1 = m0 == 1 ? clamp_h(z1) : 0.0f;
#pragma unroll
                        for (int i = 0; i < 8; i += 2) { u[i] = i == cm ? a0 : u[i]; u[i + 1] = i == cm ? a1 : u[i + 1]; }   // selects: keeps u[] in registers
                    }
                    uint32_t hi[4], lo[4];
                    split8<3, false>(u, hi, lo);     // u is always kept as a hi+lo pair: log q needs it back at full precision
                    if (HTS) {
                        sts128(u_row + (uint32_t)(d0 >> 3) * 128u, hi);
                        sts128(u_row + h_stride + (uint32_t)(d0 >> 3) * 128u, lo);
                    } else {
                        tmem_st4(x.tb + kT3UHi + (uint32_t)(d0 >> 1), hi);
                        tmem_st4(x.tb + kT3ULo + (uint32_t)(d0 >> 1), lo);
                    }
                }
                const float s_ls = row_sum3(x, 0, p_ls), s_e2 = row_sum3(x, 1, p_e2);
                const float lp_u = (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_e2;      // log density of the drawn u
                lacc += INV ? lp_u : -lp_u;                                                // forward: -log r, inverse: +log q
            } else {
                // log q(u | w, y) (fastpredNF.py:713-724), u reconstructed from its hi+lo operand
                float p_ls = 0.0f, p_q = 0.0f;
                for (int d0 = x.part * 8; d0 < CPd; d0 += 8 * NPARTS) {
                    float mu[8], ls[8];
                    uint32_t uh[4], ul[4];
                    tmem_ld8(acc_mu + (uint32_t)d0, mu);
                    tmem_ld8(acc_ls + (uint32_t)d0, ls);
                    if (HTS) {
                        lds128u(u_row + (uint32_t)(d0 >> 3) * 128u, uh);
                        lds128u(u_row + h_stride + (uint32_t)(d0 >> 3) * 128u, ul);
                    } else {
                        tmem_ld4(x.tb + kT3UHi + (uint32_t)(d0 >> 1), uh);
                        tmem_ld4(x.tb + kT3ULo + (uint32_t)(d0 >> 1), ul);
                    }
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {      // branch-free: padded columns carry mu = ls = 0 and u is masked
                        const __half2 h2 = *reinterpret_cast<const __half2*>(&uh[i >> 1]), l2 = *reinterpret_cast<const __half2*>(&ul[i >> 1]);
                        const float2 hf = __half22float2(h2), lf = __half22float2(l2);
                        const float u = d0 + i < C ? ((i & 1) ? hf.y + lf.y : hf.x + lf.x) : 0.0f;     // columns C, C+1 of the operand hold mx, not u
                        const float m = mu[i] + bias_mu[d0 + i], l = ls[i] + bias_ls[d0 + i];
                        const float df = (u - m) * fast_exp(-l);
                        p_q = fmaf(df, df, p_q);
                        p_ls += l;
                    }
                }
                const float s_ls = row_sum3(x, 0, p_ls), s_q = row_sum3(x, 1, p_q);
                const float lp_u = (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_q;
                lacc += INV ? -lp_u : lp_u;                                                // forward: +log q, inverse: -log r
            }
            warp_arrive3(x, &bars->fin[g]);
            T3_TR(0x16)              // E2 / E3 phase done
            if (which == 1) break;

            // ============ n_blocks x (masked affine coupling + BatchNorm), forward ============
#pragma unroll 1
            for (int bi = 0; bi < ts.n_blocks; ++bi) {
                const int b = INV ? ts.n_blocks - 1 - bi : bi;
                const int j = 1 - ts.masked_dim[b];
                float ps = 0.0f, pt_ = 0.0f, bs = 0.0f, bt = 0.0f;
#pragma unroll 1
                for (int l = 0; l <= ts.n_hidden; ++l) {
                    const bool last = l == ts.n_hidden;
#pragma unroll 1
                    for (int net = 0; net < 2; ++net) {
                        const uint32_t tail = x.bias_ring + (uint32_t)(net * 4 + ((x.chunk_it >> 1) & 3)) * kT3BiasEntry;
                        x.chunk_it++;
                        T3_TR(0x10 + net)
                        const uint32_t acc = wait_mma3<HTS>(x, net);
                        T3_TR(0x12 + net)
                        if (HTS) {
                            if (!last) {
                                if (net == 0) epi_inplace3<NPASS, ACT_T, false, true, !FOLD_C>(x, acc, tail, 0u, Hc, &bars->opnd[g][net]);
                                else epi_inplace3<NPASS, ACT_R, false, true, !FOLD_C>(x, acc, tail, 0u, Hc, &bars->opnd[g][net]);
                            } else {
                                float pd;
                                if (net == 0) pd = epi_inplace3<NPASS, ACT_T, true, false, !FOLD_C>(x, acc, tail, tail + Hc * 4, Hc, nullptr);
                                else pd = epi_inplace3<NPASS, ACT_R, true, false, !FOLD_C>(x, acc, tail, tail + Hc * 4, Hc, nullptr);
                                if (net == 0) { ps = pd; bs = lds32(tail + Hc * 8); } else { pt_ = pd; bt = lds32(tail + Hc * 8); }
                            }
                        } else {
                            const uint32_t ohi = x.h_base + (uint32_t)(net * 2) * h_stride, olo = ohi + h_stride;
                            if (!last) {
                                if (net == 0) epi_layer3<NPASS, ACT_T, false, true, !FOLD_C, false>(acc, ohi, olo, tail, 0u, x.part, Hc, koff_c);
                                else epi_layer3<NPASS, ACT_R, false, true, !FOLD_C, false>(acc, ohi, olo, tail, 0u, x.part, Hc, koff_c);
                                warp_arrive3(x, &bars->opnd[g][net]);
                            } else {
                                float pd;
                                if (net == 0) pd = epi_layer3<NPASS, ACT_T, true, false, !FOLD_C>(acc, ohi, olo, tail, tail + Hc * 4, x.part, Hc, koff_c);
                                else pd = epi_layer3<NPASS, ACT_R, true, false, !FOLD_C>(acc, ohi, olo, tail, tail + Hc * 4, x.part, Hc, koff_c);
                                if (net == 0) { ps = pd; bs = lds32(tail + Hc * 8); } else { pt_ = pd; bt = lds32(tail + Hc * 8); }
                            }
                        }
                    }
                }
                // both nets done: affine update of the unmasked coordinate + BatchNorm (TFCondARFlow.py:360-381, 303-315)
                const float s = row_sum3(x, 0, ps) + bs, t = row_sum3(x, 1, pt_) + bt;
                const float log_s = tanhf(s);
                const bool more = bi + 1 < ts.n_blocks;
                const int bn = INV ? b - 1 : b + 1;                  // the block processed next
                if (!INV) {
                    if (j == 0) z0 = fmaf(z0, expf(log_s), t); else z1 = fmaf(z1, expf(log_s), t);
                    z0 = fmaf(ts.bn_scale[b][0], z0, ts.bn_shift[b][0]);
                    z1 = fmaf(ts.bn_scale[b][1], z1, ts.bn_shift[b][1]);
                    lacc += log_s + ts.bn_ldj[b];
                } else {            // x = (u - t) exp(-log_s) (TFCondARFlow.py:383-402), then the next block's BatchNorm.inverse
                    if (j == 0) z0 = (z0 - t) * expf(-log_s); else z1 = (z1 - t) * expf(-log_s);
                    lacc -= log_s;
                    if (more) {
                        z0 = fmaf(ts.bn_iscale[bn][0], z0, ts.bn_ishift[bn][0]);
                        z1 = fmaf(ts.bn_iscale[bn][1], z1, ts.bn_ishift[bn][1]);
                        lacc -= ts.bn_ldj[bn];
                    }
                }
                float p0, p1;
                uint32_t dst_hi, dst_lo;      // HTS: shared addresses of the (hi, lo) pair; otherwise tensor-memory columns
                int owner;
                if (more) {
                    const int mn = ts.masked_dim[bn];
                    p0 = mn == 0 ? clamp_h(z0) : 0.0f; p1 = mn == 1 ? clamp_h(z1) : 0.0f;
                    if (HTS) { dst_hi = u_row + (uint32_t)(C >> 3) * 128u + (uint32_t)(C & 7) * 2u; dst_lo = dst_hi + h_stride; }
                    else { dst_hi = kT3UHi + (uint32_t)(C >> 1); dst_lo = kT3ULo + (uint32_t)(C >> 1); }
                    owner = (C >> 3) % NPARTS;
                } else {
                    p0 = clamp_h(z0); p1 = clamp_h(z1);
                    if (HTS) { dst_hi = in_row; dst_lo = in_row + h_stride; }
                    else { dst_hi = kT3InHi; dst_lo = kT3InLo; }
                    owner = 0;
                }
                if (x.part == owner) {
                    const __half2 h = __floats2half2_rn(p0, p1);
                    const float2 f = __half22float2(h);
                    const __half2 lo2 = __floats2half2_rn(p0 - f.x, p1 - f.y);
                    if (HTS) {
                        sts32(dst_hi, *reinterpret_cast<const uint32_t*>(&h));
                        sts32(dst_lo, *reinterpret_cast<const uint32_t*>(&lo2));
                    } else {
                        tmem_st1(x.tb + dst_hi, *reinterpret_cast<const uint32_t*>(&h));
                        tmem_st1(x.tb + dst_lo, *reinterpret_cast<const uint32_t*>(&lo2));
                    }
                }
                warp_arrive3(x, &bars->fin[g]);
            }
        }
        if (INV) {      // z = the sample of this step; log_prob = base + running mean of the step terms (fastpredNF.py:598-607)
            cum += lacc;
            if (valid && x.part == 0) {
                const long long q = n * dm.T + step;
                o.seq[q * 2] = z0; o.seq[q * 2 + 1] = z1;
                o.ldjs[q] = lacc;
                o.log_prob[q] = base_lp + cum / (float)(step + 1);
            }
        }
    T3_STEP_END
        if (!INV && x.part == 0) {      // whole warps: store_result reduces K importance samples across lanes
            if (MODE == 0) {
                const float lp = normal_lp2(z0, z1, prev0, prev1, m.cur->std[0], m.cur->std[1]) + lacc;
                store_result(a, ag, rem, jstep, lp, valid);
            } else {      // base = Normal(base_pos, std[0]); ldj = mean over the chain steps of this result index
                const float lp = normal_lp2(z0, z1, prev0, prev1, __ldg(&steps[0].std[0]), __ldg(&steps[0].std[1])) + lacc / (float)(jstep - dm.jlo + 1);
                store_result(a, ag, rem, jstep - dm.jlo, lp, valid);
            }
        }
    T3_ITEM_END
}

}  // namespace t3

template <int NPASS, bool FOLD_D, bool FOLD_C, int MODE = 0, bool HTS = false>
__global__ void __launch_bounds__(kT3Threads, 1)
k_separate_tc3(const T2Step* __restrict__ steps, const uint8_t* __restrict__ img, const __grid_constant__ RowArgs a,
               const __grid_constant__ T3Dims dm, long long pairs_per_step, long long total_items, const __grid_constant__ SampleOut o) {
    using namespace tc;
    using namespace t2;
    using namespace t3;
    extern __shared__ __align__(128) uint8_t smem_t3[];
    uint8_t* smem = smem_t3;
    Bars3* bars = reinterpret_cast<Bars3*>(smem);
    static_assert(sizeof(Bars3) <= 240, "barrier block");
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 240);
    static_assert(sizeof(T2Step) <= 1792, "T2Step must fit the header");
    const uint32_t nslots = 1u << dm.nslots_log2;

    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        // a weight slot is handed back once the issuers of BOTH tiles have committed the MMAs that read it
        for (uint32_t s = 0; s < nslots; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], dm.single ? 1 : 2); }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&bars->mma_done[t][0], 1);
            mbar_init(&bars->mma_done[t][1], 1);
            mbar_init(&bars->opnd[t][0], 8);      // warps that run this net's hidden epilogues
            mbar_init(&bars->opnd[t][1], 8);
            mbar_init(&bars->fin[t], 8);
        }
        mbar_fence_init();
    }
    {   // constant A operand of the bias fold: row r = [1, 1, 1, 0 x 13] in the canonical K-major layout (Kp = 16)
        uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kT3OnesOff);
        for (int i = tid; i < 1024; i += kT3Threads) {            // 128 rows x 2 core columns x 16 B = 4 KB
            const int w = i & 3, kc = (i >> 5) & 1;               // word inside the 16-byte row piece, core column (k / 8)
            ones[i] = kc == 0 ? (w == 0 ? 0x3C003C00u : (w == 1 ? 0x00003C00u : 0u)) : 0u;
        }
        fence_async_smem();
    }
    if (warp == 16) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tslot);

    // two roles (inlined: separate noinline role functions measured 3 ms slower)
    // Warp-specialised register split (setmaxnreg, per warpgroup): the kernel launches with 96 registers per thread
    // (640 threads); the service warpgroup (warps 16-19) gives back all but 56, the four epilogue warpgroups grow to 104.
    // setmaxnreg.inc only draws on what .dec released: 4 x 128 x (104 - 96) = 4096 <= 128 x (96 - 56) = 5120.
    if (warp < 16) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kT3ComputeRegs));
        t3_compute<NPASS, FOLD_D, FOLD_C, MODE, HTS>(steps, a, o, dm, pairs_per_step, total_items, smem, tmem_base);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kT3ServiceRegs));
        t3_issuer<NPASS, HTS>(steps, img, dm, pairs_per_step, total_items, smem, tmem_base, (warp - 16) & 1, (warp - 16) >> 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// Shared-memory feasibility of the two-tile kernel for a packed model (same arithmetic as the launchers below): every
// step's operand tiles + at least a 2-slot ring of its largest weight chunk must fit, for the per-step launches (width
// classes) and for the chain / sampler launches (one class for the whole model).  capi.cu routes the model to the wide
// kernel when this says no (e.g. hidden_size 80: two 26.3 KB slots + 160 KB of operand tiles exceed 227 KB by 512 B).
inline bool tc3_fits(const std::vector<T2Step>& hsteps, int smem_optin) {
    int wmax = 64, slot_max = 0;
    for (const T2Step& st : hsteps) {
        const int w = st.Wd > st.Hc ? st.Wd : st.Hc;
        if (w > 80) return false;
        const int wc = w <= 64 ? 64 : 80;
        int slot = 0;
        for (int ci = 0; ci < st.nchunks; ++ci) slot = st.chunk_bytes[ci] > slot ? st.chunk_bytes[ci] : slot;
        slot = (slot + 127) / 128 * 128;
        if ((size_t)kT3Header + (size_t)8 * 128 * wc * 2 + (size_t)2 * slot > (size_t)smem_optin) return false;
        wmax = wc > wmax ? wc : wmax;
        slot_max = slot > slot_max ? slot : slot_max;
    }
    return (size_t)kT3Header + (size_t)8 * 128 * wmax * 2 + (size_t)2 * slot_max <= (size_t)smem_optin;
}

// Launches the steps in groups of equal operand-width class (hidden width <= 64: 128 KB of operand tiles + a 4-slot weight
// ring; 80: 160 KB of operand tiles + 2 slots).  Models with wider layers never get here: tc2_pack refuses them and
// capi.cu routes them to the single-tile wide kernel (kernels_tcw.cuh).
template <int NPASS>
inline int tc3_launch_separate(const T2Dims& dm2, const std::vector<T2Step>& hsteps, const T2Step* d_steps, const unsigned char* d_img,
                               const RowArgs& a, int num_sms, int smem_optin, cudaStream_t stream, uint64_t* launches, std::string* msg) {
    if (!dm2.supported || d_img == nullptr) { *msg = "the tensor-core path does not support this model shape; use FC_PREC_FP32"; return 1; }
    const int T = dm2.T;
    const long long tiles = (a.N + 127) / 128;
    // small launches (fewer tile pairs x steps than SMs) run one tile per item so that twice as many SMs work
    const bool single = ((tiles + 1) / 2) * T < num_sms;
    const long long pairs = single ? tiles : (tiles + 1) / 2;
    int lo = 0;
    while (lo < T) {
        // group consecutive steps with the same operand-width class
        // (the class also fixes which hidden layers carry folded biases: tc2_pack folds widths <= kFoldMaxN = 64)
        auto wclass = [&](int s) { const int w = hsteps[s].Wd > hsteps[s].Hc ? hsteps[s].Wd : hsteps[s].Hc; return (w <= 64 ? 64 : (w <= 80 ? 80 : 96)) + (hsteps[s].Wd <= kFoldMaxN ? 1 : 0); };
        int hi = lo;
        const int wc = wclass(lo) & ~1;
        const bool fold_d = hsteps[lo].Wd <= kFoldMaxN, fold_c = hsteps[lo].Hc <= kFoldMaxN;
        while (hi + 1 < T && wclass(hi + 1) == wclass(lo)) ++hi;
        int slot = 0;
        for (int s = lo; s <= hi; ++s)
            for (int ci = 0; ci < hsteps[s].nchunks; ++ci) slot = hsteps[s].chunk_bytes[ci] > slot ? hsteps[s].chunk_bytes[ci] : slot;
        slot = (slot + 127) / 128 * 128;
        // steps whose layers are all <= 64 wide: hidden operands in tensor memory, first-layer operands (K <= 32) in smem
        bool hts = wc == 64 && fold_d && fold_c;
        for (int s = lo; s <= hi; ++s) hts = hts && hsteps[s].Kin <= kT3InuK && hsteps[s].Ku <= kT3InuK;
        T3Dims dm{};
        dm.slot_bytes = slot; dm.Wmax = hts ? kT3InuK : wc; dm.T = T; dm.C0 = dm2.C0; dm.step_lo = lo; dm.step_hi = hi; dm.single = single ? 1 : 0;
        const size_t fixed = (size_t)kT3Header + (size_t)8 * 128 * dm.Wmax * 2;
        // a weight slot is handed back by the tcgen05.commit of the MMAs that read it, so 2 slots (one per net) already
        // pipeline; 4 when they fit
        dm.nslots_log2 = (fixed + (size_t)4 * slot <= (size_t)smem_optin) ? 2 : 1;
        const size_t smem = fixed + ((size_t)1 << dm.nslots_log2) * slot;
        if (wc > 80 || smem > (size_t)smem_optin) { *msg = "the two-tile tensor-core kernel does not fit this model shape"; return 1; }
        const long long items = pairs * (hi - lo + 1);
        const int grid = (int)(items < num_sms ? items : num_sms);
        cudaError_t e = cudaSuccess;
        auto go = [&](auto kern) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) kern<<<grid, kT3Threads, smem, stream>>>(d_steps, d_img, a, dm, pairs, items, SampleOut{});
        };
#ifdef FC_T3_HTS_NOFOLD
        if (hts) { dm.no_fold = 1; go(k_separate_tc3<NPASS, false, false, 0, true>); }
#else
        if (hts) go(k_separate_tc3<NPASS, true, true, 0, true>);
#endif
        else if (fold_d && fold_c) go(k_separate_tc3<NPASS, true, true>);
        else if (fold_c) go(k_separate_tc3<NPASS, false, true>);
        else if (fold_d) go(k_separate_tc3<NPASS, true, false>);
        else go(k_separate_tc3<NPASS, false, false>);
        if (e != cudaSuccess) { *msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
        (*launches)++;
        e = cudaGetLastError();
        if (e != cudaSuccess) { *msg = std::string("k_separate_tc3 launch: ") + cudaGetErrorString(e); return 1; }
        lo = hi + 1;
    }
    return 0;
}

// Chain mode (flow.log_prob_sequential, fastpredNF.py:548-569): one launch, item = (tile pair, result index j) walks
// net[j] ... net[klo]; the steps of an item mix operand widths, so the operand stride is the widest one and the bias
// tiles are ignored (no_fold).
template <int NPASS>
inline int tc3_launch_sequential(const T2Dims& dm2, const std::vector<T2Step>& hsteps, const T2Step* d_steps, const unsigned char* d_img,
                                 const RowArgs& a, int num_sms, int smem_optin, cudaStream_t stream, uint64_t* launches, std::string* msg) {
    if (!dm2.supported || d_img == nullptr) { *msg = "the tensor-core path does not support this model shape; use FC_PREC_FP32"; return 1; }
    const int T = dm2.T;
    const long long tiles = (a.N + 127) / 128;
    const bool single = ((tiles + 1) / 2) * (T - a.seq_jlo) < num_sms;
    const long long pairs = single ? tiles : (tiles + 1) / 2;
    int wc = 64, slot = 0;
    for (int s = 0; s < T; ++s) {
        const int w = hsteps[s].Wd > hsteps[s].Hc ? hsteps[s].Wd : hsteps[s].Hc;
        wc = w > wc ? w : wc;
        for (int ci = 0; ci < hsteps[s].nchunks; ++ci) slot = hsteps[s].chunk_bytes[ci] > slot ? hsteps[s].chunk_bytes[ci] : slot;
    }
    slot = (slot + 127) / 128 * 128;
    T3Dims dm{};
    dm.slot_bytes = slot; dm.Wmax = wc; dm.T = T; dm.C0 = dm2.C0; dm.step_lo = a.seq_jlo; dm.step_hi = T - 1;
    dm.seq = 1; dm.klo = a.seq_klo; dm.jlo = a.seq_jlo; dm.no_fold = 1; dm.single = single ? 1 : 0;
    const size_t fixed = (size_t)kT3Header + (size_t)8 * 128 * wc * 2;
    dm.nslots_log2 = (fixed + (size_t)4 * slot <= (size_t)smem_optin) ? 2 : 1;
    const size_t smem = fixed + ((size_t)1 << dm.nslots_log2) * slot;
    if (wc > 80 || smem > (size_t)smem_optin) { *msg = "the tensor-core chain kernel does not fit this model shape; use FC_PREC_FP32"; return 1; }
    auto kern = k_separate_tc3<NPASS, false, false, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    const long long items = pairs * (T - a.seq_jlo);
    const int grid = (int)(items < num_sms ? items : num_sms);
    kern<<<grid, kT3Threads, smem, stream>>>(d_steps, d_img, a, dm, pairs, items, SampleOut{});
    (*launches)++;
    e = cudaGetLastError();
    if (e != cudaSuccess) { *msg = std::string("k_separate_tc3 (chain) launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

// Sampler (flow.sample_with_log_prob): one launch, item = tile pair, steps 0 .. T-1 in the inverse direction.
template <int NPASS>
inline int tc3_launch_sample(const T2Dims& dm2, const std::vector<T2Step>& hsteps, const T2Step* d_steps, const unsigned char* d_img,
                             const RowArgs& a, const SampleOut& o, int num_sms, int smem_optin, cudaStream_t stream, uint64_t* launches,
                             std::string* msg) {
    if (!dm2.supported || d_img == nullptr) { *msg = "the tensor-core path does not support this model shape; use FC_PREC_FP32"; return 1; }
    const int T = dm2.T;
    const long long tiles = (a.N + 127) / 128;
    const bool single = (tiles + 1) / 2 < num_sms;
    const long long pairs = single ? tiles : (tiles + 1) / 2;
    int wc = 64, slot = 0;
    for (int s = 0; s < T; ++s) {
        const int w = hsteps[s].Wd > hsteps[s].Hc ? hsteps[s].Wd : hsteps[s].Hc;
        wc = w > wc ? w : wc;
        for (int ci = 0; ci < hsteps[s].nchunks; ++ci) slot = hsteps[s].chunk_bytes[ci] > slot ? hsteps[s].chunk_bytes[ci] : slot;
    }
    slot = (slot + 127) / 128 * 128;
    T3Dims dm{};
    dm.slot_bytes = slot; dm.Wmax = wc; dm.T = T; dm.C0 = dm2.C0; dm.step_lo = 0; dm.step_hi = T - 1;
    dm.seq = 2; dm.no_fold = 1; dm.single = single ? 1 : 0;
    const size_t fixed = (size_t)kT3Header + (size_t)8 * 128 * wc * 2;
    dm.nslots_log2 = (fixed + (size_t)4 * slot <= (size_t)smem_optin) ? 2 : 1;
    const size_t smem = fixed + ((size_t)1 << dm.nslots_log2) * slot;
    if (wc > 80 || smem > (size_t)smem_optin) { *msg = "the tensor-core sampler does not fit this model shape; use FC_PREC_FP32"; return 1; }
    auto kern = k_separate_tc3<NPASS, false, false, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    const int grid = (int)(pairs < num_sms ? pairs : num_sms);
    kern<<<grid, kT3Threads, smem, stream>>>(d_steps, d_img, a, dm, pairs, pairs, o);
    (*launches)++;
    e = cudaGetLastError();
    if (e != cudaSuccess) { *msg = std::string("k_separate_tc3 (sampler) launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

}  // namespace fc
