// kernels_tcw.cuh -- f16x3 / f16 tensor-core path for WIDE models: hidden sizes up to 256 (BASELINE config 5) and
// conditioning lengths up to 64 (the reference's tuning range, src/main.py:222-229), i.e. every shape the two-tile
// kernel (kernels_tc3.cuh, layers <= 80 wide) refuses.
//
// Same arithmetic as the two-tile kernel (split-fp16 operands, tcgen05.mma kind::f16, fp32 accumulate in tensor
// memory, tc_f16x3.cuh), different resource plan: a 256-wide layer needs 256 accumulator columns, and its hi + lo
// operand would be 128 KB of shared memory read at the SS-mode cost of 172 cycles per MMA (tools/mb_mma.cu), so
//   * ONE 128-row tile per CTA, 16 epilogue warps (thread = (row = TMEM lane, one of 4 column parts)) + a service
//     warpgroup (TMA weight producer, MMA issuer); the two nets of a sub-network run ONE AFTER THE OTHER;
//   * tensor memory is two 256-column REGIONS that swap roles every layer: layer l accumulates into region l & 1 and
//     reads its A operand (TS mode: 11 + N/2 cycles per MMA instead of 44 + N/2) from the other one.  The epilogue
//     converts an accumulator IN PLACE into the next layer's operand: a thread reads 16 fp32 columns of its row and
//     writes 8 packed-fp16 hi columns + 8 lo columns back into the same 16 columns, which is exactly the K-step layout
//     a TS-mode MMA reads (hi at 16 k, lo at 16 k + 8);
//   * every wide layer is issued as two N-HALVES with their own completion barriers, so the epilogue of the first half
//     overlaps the MMAs of the second (they touch disjoint columns of the accumulator region and the operand region is
//     read-only until both halves are done);
//   * the narrow first-layer operands [z|w, y] and [u, mx] live in shared memory (SS mode, K <= 96), the density means
//     wait in a shared-memory stash (each thread parks and later re-reads its own chunks) while the log-scale net runs;
//   * the weights are packed per half as K-SLABS ([N_half x ks] hi + lo tile, <= 16 KB, canonical K-major layout each)
//     that stream from L2 through a TMA ring; a slab is released by the tcgen05.commit of the MMAs that read it;
//   * biases / last-row vectors are read by the epilogue warps straight from L2/L1 (warp-uniform float4 loads).
// One job = one net-layer.  Protocol: the compute warps arrive on `opnd` once before every job (operand converted,
// accumulator region drained); the issuer waits for it, issues half 0 then half 1 slab by slab and commits `done[h]`
// after each; the compute warps wait for `done[0]` / `done[1]` before the respective half of the epilogue.
#pragma once
#include "kernels_tc3.cuh"

namespace fc {

constexpr int kTwMaxJobs = 12 + 2 * (kMaxHidden + 1) * kMaxBlocks;
constexpr int kTwThreads = 16 * 32 + 128;       // 16 compute warps + a service warpgroup (producer, issuer, 2 idle)
constexpr int kTwComputeRegs = 104, kTwServiceRegs = 56;
constexpr int kTwHeader = 8192;                 // barriers, step table, row-partial buffers
constexpr int kTwStepOff = 256, kTwPartOff = 3584;
constexpr int kTwSlotBytes = 16384;
constexpr int kTwMaxSlots = 8;
constexpr uint32_t kTwRegion = 256;             // tensor-memory columns of one region

struct TwStep {
    int32_t C, c, Kin, Ku, Wd, CPd, Hc;
    int32_t n_blocks, n_hidden, njobs, eps_off, seq_eps_off;
    int32_t masked_dim[kMaxBlocks];
    float bn_scale[kMaxBlocks][2], bn_shift[kMaxBlocks][2], bn_ldj[kMaxBlocks];
    float bn_iscale[kMaxBlocks][2], bn_ishift[kMaxBlocks][2];
    float std[2];
    // jobs in forward order: [r density: net 0 L1 L2 L3, net 1 L1 L2 L3][block b: s-net L0..Ln, t-net L0..Ln] x n_blocks[q density]
    uint32_t job_img[kTwMaxJobs];      // byte offset / 16 of the job's first slab (half 0, then half 1) in the image
    uint32_t job_tail[kTwMaxJobs];     // float offset of the job's fp32 tail (bias[N] (+ last-row vector[N], its bias))
    uint16_t job_kp[kTwMaxJobs], job_n[kTwMaxJobs];      // K (padded), N (padded)
    uint16_t job_n0[kTwMaxJobs];       // N of the first half (multiple of 16); the second half has N - n0 (0: one half only)
    uint16_t job_ksplit[kTwMaxJobs];   // K columns of the A operand that are ready after `opnd[0]` (the producing layer's first half); 0: wait for all
    uint8_t job_src[kTwMaxJobs];       // A operand: 0 = IN (shared memory), 1 = U (shared memory), 2 = the other tensor-memory region
    uint8_t job_reg[kTwMaxJobs];       // accumulator region of the job (layer index & 1)
};
static_assert(sizeof(TwStep) <= kTwPartOff - kTwStepOff, "TwStep must fit the header");

struct TwDims {
    int32_t supported;
    int32_t T, C0;
    int32_t Kmax;                          // widest first-layer operand: [128 x Kmax] fp16 tiles in shared memory
    int32_t CPmax;                         // widest density output: stash of [CPmax/8][128][8] floats
    int32_t slot_bytes, nslots;
    int32_t step_lo, step_hi;              // per-step mode: steps covered; chain mode: result indices
    int32_t seq;                           // 0 per-step density, 1 chain, 2 sampler
    int32_t klo, jlo;
};

// slab width of a half-job: as many 16-wide K steps as fit a ring slot
inline int tw_slab_k(int N, int Kp, int parts, int slot_bytes) {
    int ks = (slot_bytes / (N * 2 * parts)) & ~15;
    if (ks < 16) ks = 16;
    return ks < Kp ? ks : Kp;
}
// first half of an N-wide layer: layers of 128 columns and more are issued as two halves
inline int tw_first_half(int N) { return N >= 128 ? (N / 32) * 16 : N; }

// Packs the wide image.  dm.supported = 0 for shapes this kernel does not cover.
inline void tcw_pack(const fc_model_desc& d, const std::vector<StepPlan>& plans, const std::vector<TcStepSrc>& src, int npass,
                     std::vector<unsigned char>& img, std::vector<float>& tails, std::vector<TwStep>& steps, TwDims& dm,
                     int smem_optin) {
    img.clear(); tails.clear(); steps.clear();
    memset(&dm, 0, sizeof dm);
    const int T = d.pred_len, H = d.hidden_size, parts = npass == 3 ? 2 : 1;
    const int Cmax = d.cond_size + (T - 1) * 2, cmax = Cmax + 2;
    const int Wd_max = pad16i(2 * cmax), Kmax = pad16i(cmax), CPmax = pad16i(Cmax);
    if (H % 16 != 0 || H > 256 || (d.cond_size & 1) || Wd_max > 256) return;      // two 256-column regions
    dm.T = T; dm.C0 = d.cond_size; dm.Kmax = Kmax; dm.CPmax = CPmax;
    dm.slot_bytes = kTwSlotBytes;
    const long long ring = (long long)smem_optin - kTwHeader - (long long)4 * 128 * Kmax * 2 - (long long)128 * CPmax * 4;
    dm.nslots = (int)(ring / dm.slot_bytes);
    if (dm.nslots > kTwMaxSlots) dm.nslots = kTwMaxSlots;
    if (dm.nslots < 2) return;                                    // shared memory
    steps.resize(T);
    std::vector<float> sub, Ws, bs;
    for (int i = 0; i < T; ++i) {
        TwStep& ts = steps[i];
        memset(&ts, 0, sizeof ts);
        const StepPlan& sp = plans[i];
        const TcStepSrc& s = src[i];
        const int C = sp.C, c = sp.c;
        ts.C = C; ts.c = c; ts.Kin = pad16i(c); ts.Ku = pad16i(C + 2); ts.Wd = pad16i(2 * c); ts.CPd = pad16i(C); ts.Hc = H;
        ts.n_blocks = sp.n_blocks; ts.n_hidden = sp.n_hidden; ts.eps_off = sp.eps_off; ts.seq_eps_off = sp.seq_eps_off;
        for (int b = 0; b < sp.n_blocks; ++b) {
            ts.masked_dim[b] = sp.masked_dim[b];
            for (int q = 0; q < 2; ++q) {
                ts.bn_scale[b][q] = sp.bn_scale[b][q]; ts.bn_shift[b][q] = sp.bn_shift[b][q];
                ts.bn_iscale[b][q] = sp.bn_iscale[b][q]; ts.bn_ishift[b][q] = sp.bn_ishift[b][q];
            }
            ts.bn_ldj[b] = sp.bn_ldj[b];
        }
        ts.std[0] = sp.std[0]; ts.std[1] = sp.std[1];
        int nj = 0;
        // one job: W [rows x K] (row-major, nn.Linear.weight) -> per N-half, K-slabs of canonical K-major tiles; tail = bias (+ extra)
        auto job = [&](const float* W, const float* bias, int rows, int K, int N, int Kp, int src_, int layer, const float* extra, int n_extra) {
            while (img.size() % 128) img.push_back(0);
            ts.job_img[nj] = (uint32_t)(img.size() / 16);
            const int n0 = tw_first_half(N);
            for (int half = 0; half < 2; ++half) {
                const int r0 = half == 0 ? 0 : n0, Nh = half == 0 ? n0 : N - n0;
                if (Nh == 0) continue;
                const int Ks = tw_slab_k(Nh, Kp, parts, dm.slot_bytes);
                const int rv = rows - r0 < 0 ? 0 : (rows - r0 < Nh ? rows - r0 : Nh);       // valid (unpadded) rows of this half
                for (int k0 = 0; k0 < Kp; k0 += Ks) {
                    const int ksl = Ks < Kp - k0 ? Ks : Kp - k0;
                    const int kv = K - k0 < 0 ? 0 : (K - k0 < ksl ? K - k0 : ksl);  // valid (unpadded) columns of this slab
                    sub.assign((size_t)(rv > 0 ? rv : 1) * (kv > 0 ? kv : 1), 0.0f);
                    for (int r = 0; r < rv; ++r)
                        for (int k = 0; k < kv; ++k) sub[(size_t)r * kv + k] = W[(size_t)(r0 + r) * K + k0 + k];
                    for (int part = 0; part < parts; ++part) {
                        const size_t o = img.size();
                        img.resize(o + (size_t)Nh * ksl * 2);
                        pack_kmajor_f16(img.data() + o, sub.data(), rv, kv, Nh, ksl, part);
                    }
                }
            }
            while (tails.size() % 4) tails.push_back(0.0f);
            ts.job_tail[nj] = (uint32_t)tails.size();
            const size_t o = tails.size();
            tails.resize(o + (size_t)N + n_extra, 0.0f);
            memcpy(tails.data() + o, bias, (size_t)rows * 4);
            if (n_extra) memcpy(tails.data() + o + N, extra, (size_t)n_extra * 4);
            ts.job_kp[nj] = (uint16_t)Kp; ts.job_n[nj] = (uint16_t)N; ts.job_n0[nj] = (uint16_t)n0;
            ts.job_src[nj] = (uint8_t)src_; ts.job_reg[nj] = (uint8_t)(layer & 1);
            // a tensor-memory operand produced by a two-half layer becomes available half by half
            ts.job_ksplit[nj] = (uint16_t)((src_ == 2 && nj > 0 && ts.job_n0[nj - 1] < ts.job_n[nj - 1] && ts.job_n[nj - 1] == Kp) ? ts.job_n0[nj - 1] : 0);
            ++nj;
        };
        auto dens = [&](int which) {
            const int h2 = 2 * c;
            for (int net = 0; net < 2; ++net) {
                job(s.dens_W[which][net][0], s.dens_b[which][net][0], h2, c, ts.Wd, ts.Kin, 0, 0, nullptr, 0);
                job(s.dens_W[which][net][1], s.dens_b[which][net][1], h2, h2, ts.Wd, ts.Wd, 2, 1, nullptr, 0);
                job(s.dens_W[which][net][2], s.dens_b[which][net][2], C, h2, ts.CPd, ts.Wd, 2, 2, nullptr, 0);
            }
        };
        dens(0);
        for (int b = 0; b < sp.n_blocks; ++b) {
            const int j = 1 - sp.masked_dim[b];
            for (int net = 0; net < 2; ++net)
                for (int l = 0; l <= sp.n_hidden; ++l) {
                    std::vector<float> extra;
                    if (l == sp.n_hidden) {      // surviving row (1 - mask) of the final H -> 2 Linear + its bias
                        extra.assign(H + 4, 0.0f);
                        memcpy(extra.data(), s.cpl_W[b][net][sp.n_hidden + 1] + (size_t)j * H, (size_t)H * 4);
                        extra[H] = s.cpl_b[b][net][sp.n_hidden + 1][j];
                    }
                    const float* Wl = s.cpl_W[b][net][l];
                    const float* bl = s.cpl_b[b][net][l];
                    const int Kl = l == 0 ? c : H;
                    if (npass == 3 && net == 0) {   // s-net layers feed tanh = 1 - 2/(1 + 2^(2 log2(e) x)): fold the factor into W, b
                        Ws.assign(Wl, Wl + (size_t)H * Kl);
                        bs.assign(bl, bl + H);
                        for (float& v : Ws) v *= kTanhPrescale;
                        for (float& v : bs) v *= kTanhPrescale;
                        Wl = Ws.data(); bl = bs.data();
                    }
                    job(Wl, bl, H, Kl, H, l == 0 ? ts.Ku : H, l == 0 ? 1 : 2, l, extra.empty() ? nullptr : extra.data(), (int)extra.size());
                }
        }
        dens(1);
        ts.njobs = nj;
    }
    while (img.size() % 128) img.push_back(0);
    while (tails.size() % 4) tails.push_back(0.0f);
    dm.supported = 1;
}

inline size_t tcw_smem_bytes(const TwDims& dm) {
    return (size_t)kTwHeader + (size_t)4 * 128 * dm.Kmax * 2 + (size_t)128 * dm.CPmax * 4 + (size_t)dm.nslots * dm.slot_bytes;
}

namespace tw {
using namespace tc;
using namespace t2;

struct BarsW {
    uint64_t full[kTwMaxSlots], empty[kTwMaxSlots];
    uint64_t done[2], opnd[2];
};
static_assert(sizeof(BarsW) <= 240, "barrier block");

struct SmemW {
    BarsW* bars;
    TwStep* cur;
    float* part_buf;          // [2 slots][4 parts][128 rows]
    uint8_t *inu, *stash, *slots;
    uint32_t inu_stride;      // one [128 x Kmax] fp16 tile
};
__device__ __forceinline__ SmemW tw_smem(uint8_t* smem, const TwDims& dm) {
    SmemW m;
    m.bars = reinterpret_cast<BarsW*>(smem);
    m.cur = reinterpret_cast<TwStep*>(smem + kTwStepOff);
    m.part_buf = reinterpret_cast<float*>(smem + kTwPartOff);
    m.inu_stride = (uint32_t)(128 * dm.Kmax * 2);
    m.inu = smem + kTwHeader;                       // [IN hi, IN lo, U hi, U lo] x [128 x Kmax] fp16, canonical K-major
    m.stash = m.inu + (size_t)4 * m.inu_stride;     // [CPmax / 8][128 rows][8] fp32: density means
    m.slots = m.stash + (size_t)128 * dm.CPmax * 4;
    return m;
}
__device__ __forceinline__ void tw_load_step(const TwStep* steps, int step, TwStep* cur, int tid) {
    __syncthreads();
    const int nw = sizeof(TwStep) / 4;
    const int* s = reinterpret_cast<const int*>(steps + step);
    int* d = reinterpret_cast<int*>(cur);
    for (int i = tid; i < nw; i += kTwThreads) d[i] = __ldg(s + i);
    __syncthreads();
}
// item = (tile, jstep); every role walks exactly these loops so that the CTA-wide step-table reloads line up
#define TW_ITEM_BEGIN(SEQ)                                                                           \
    int cur_step = -1;                                                                             \
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {                     \
        const int jstep = dm.step_hi - (int)(item / tiles_per_step); /* heavy (late) first */      \
        const long long tile = item % tiles_per_step;                                              \
        const int nsteps = (SEQ) == 0 ? 1 : ((SEQ) == 1 ? jstep - dm.klo + 1 : dm.T);           \
        (void)tile;
#define TW_STEP_BEGIN(SEQ)                                                                           \
        for (int si = 0; si < nsteps; ++si) {                                                      \
            const int step = (SEQ) == 2 ? si : jstep - si;                                      \
            if (step != cur_step) { tw_load_step(steps, step, m.cur, (int)threadIdx.x); cur_step = step; } \
            const TwStep& ts = *m.cur;
#define TW_STEP_END }
#define TW_ITEM_END }

// image job of schedule position ji: the sampler runs [q density][blocks n-1 .. 0][r density] on the forward image
__device__ __forceinline__ int tw_img_job(const TwStep& ts, int ji, bool inverse) {
    if (!inverse) return ji;
    const int nb2 = 2 * (ts.n_hidden + 1), nb6 = 6 + ts.n_blocks * nb2;
    if (ji < 6) return nb6 + ji;
    if (ji < nb6) { const int bi = (ji - 6) / nb2; return 6 + (ts.n_blocks - 1 - bi) * nb2 + (ji - 6 - bi * nb2); }
    return ji - nb6;
}

__device__ __forceinline__ void tw_idle(const TwStep* __restrict__ steps, const TwDims& dm, long long tiles_per_step, long long total_items,
                                        uint8_t* smem) {
    const SmemW m = tw_smem(smem, dm);
    TW_ITEM_BEGIN(dm.seq)
    TW_STEP_BEGIN(dm.seq)
        (void)ts;
    TW_STEP_END
    TW_ITEM_END
}

// ===== weight producer (one warp): streams the K-slabs of every half-job of the step through the ring =====
template <int NPASS>
__device__ __forceinline__ void tw_producer(const TwStep* __restrict__ steps, const uint8_t* __restrict__ img, const TwDims& dm,
                                            long long tiles_per_step, long long total_items, uint8_t* smem) {
    const SmemW m = tw_smem(smem, dm);
    BarsW* bars = m.bars;
    const int lane = threadIdx.x & 31;
    constexpr int PARTS = NPASS == 3 ? 2 : 1;
    int slot = 0;
    uint32_t phase = 0;
    TW_ITEM_BEGIN(dm.seq)
    TW_STEP_BEGIN(dm.seq)
        for (int ji = 0; ji < ts.njobs; ++ji) {
            const int pj = tw_img_job(ts, ji, dm.seq == 2);
            const int N = ts.job_n[pj], n0 = ts.job_n0[pj], Kp = ts.job_kp[pj];
            const uint8_t* srcp = img + (size_t)ts.job_img[pj] * 16;
            for (int half = 0; half < 2; ++half) {
                const int Nh = half == 0 ? n0 : N - n0;
                if (Nh == 0) break;
                int Ks = (dm.slot_bytes / (Nh * 2 * PARTS)) & ~15;
                Ks = Ks < 16 ? 16 : Ks;
                for (int k0 = 0; k0 < Kp; k0 += Ks) {
                    const int ksl = Ks < Kp - k0 ? Ks : Kp - k0;
                    const uint32_t bytes = (uint32_t)(Nh * ksl * 2 * PARTS);
                    if (lane == 0) {
                        mbar_wait(&bars->empty[slot], phase ^ 1u);
#ifdef FC_TW_NOWEIGHTS      /* timing experiment only: no weight traffic, the MMAs read whatever is in the slot */
                        mbar_arrive(&bars->full[slot]);
#else
                        mbar_expect_tx(&bars->full[slot], bytes);
                        bulk_load(m.slots + (size_t)slot * dm.slot_bytes, srcp, bytes, &bars->full[slot]);
#endif
                    }
                    __syncwarp();
                    srcp += bytes;
                    if (++slot == dm.nslots) { slot = 0; phase ^= 1u; }
                }
            }
        }
    TW_STEP_END
    TW_ITEM_END
}

// MMAs of one K-slab of a half-job, straight line.  TS: A from the other tensor-memory region in the in-place layout
// (K-step ks: hi at 16 ks, lo at 16 ks + 8); otherwise A from the shared-memory IN / U tiles.
template <int NPASS, int NK, bool TS>
__device__ __forceinline__ void issue_slab(uint32_t acc, uint32_t ta, uint64_t ahi, uint64_t alo, uint64_t dhi, uint64_t dlo, uint32_t idesc,
                                           uint32_t acc0) {
#pragma unroll
    for (int ks = 0; ks < NK; ++ks) {
        if (TS) umma_ts(acc, ta + ks * 16, dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
        else umma_bf16(acc, ahi + (uint64_t)(ks * 16), dhi + (uint64_t)(ks * 16), idesc, ks > 0 ? 1u : acc0);
    }
    if (NPASS == 3) {
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            if (TS) umma_ts(acc, ta + ks * 16 + 8, dhi + (uint64_t)(ks * 16), idesc, 1u);
            else umma_bf16(acc, alo + (uint64_t)(ks * 16), dhi + (uint64_t)(ks * 16), idesc, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < NK; ++ks) {
            if (TS) umma_ts(acc, ta + ks * 16, dlo + (uint64_t)(ks * 16), idesc, 1u);
            else umma_bf16(acc, ahi + (uint64_t)(ks * 16), dlo + (uint64_t)(ks * 16), idesc, 1u);
        }
    }
}
template <int NPASS, bool TS>
__device__ __forceinline__ void issue_slab_n(int nk, uint32_t acc, uint32_t ta, uint64_t ahi, uint64_t alo, uint64_t dhi, uint64_t dlo,
                                             uint32_t idesc, uint32_t acc0) {
    switch (nk) {
        case 1: issue_slab<NPASS, 1, TS>(acc, ta, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 2: issue_slab<NPASS, 2, TS>(acc, ta, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 3: issue_slab<NPASS, 3, TS>(acc, ta, ahi, alo, dhi, dlo, idesc, acc0); break;
        case 4: issue_slab<NPASS, 4, TS>(acc, ta, ahi, alo, dhi, dlo, idesc, acc0); break;
        default:        // wider slabs (narrow layers): two straight-line groups at a time
            for (int k = 0; k < nk; k += 2) {
                if (k + 1 < nk) issue_slab<NPASS, 2, TS>(acc, ta + k * 16, ahi + (uint64_t)(k * 16), alo + (uint64_t)(k * 16), dhi + (uint64_t)(k * 16), dlo + (uint64_t)(k * 16), idesc, k ? 1u : acc0);
                else issue_slab<NPASS, 1, TS>(acc, ta + k * 16, ahi + (uint64_t)(k * 16), alo + (uint64_t)(k * 16), dhi + (uint64_t)(k * 16), dlo + (uint64_t)(k * 16), idesc, k ? 1u : acc0);
            }
            break;
    }
}

// ===== MMA issuer (one warp): job by job, half by half, slab by slab =====
template <int NPASS>
__device__ __forceinline__ void tw_issuer(const TwStep* __restrict__ steps, const TwDims& dm, long long tiles_per_step, long long total_items,
                                          uint8_t* smem, uint32_t tmem_base) {
    const SmemW m = tw_smem(smem, dm);
    BarsW* bars = m.bars;
    constexpr int PARTS = NPASS == 3 ? 2 : 1;
    int slot = 0;
    uint32_t phase = 0, opnd_cnt = 0;
    TW_ITEM_BEGIN(dm.seq)
    TW_STEP_BEGIN(dm.seq)
#pragma unroll 1
        for (int ji = 0; ji < ts.njobs; ++ji) {
            const int Kp = ts.job_kp[ji], N = ts.job_n[ji], n0 = ts.job_n0[ji], src = ts.job_src[ji], reg = ts.job_reg[ji];
            const uint32_t sboA = (uint32_t)(Kp >> 3) * 128u;
            const uint32_t sa = smem_u32(m.inu) + (uint32_t)(src == 1 ? 2 : 0) * m.inu_stride;
            const uint64_t ahi = umma_desc(sa, 128, sboA), alo = umma_desc(sa + m.inu_stride, 128, sboA);
            const uint32_t ta = tmem_base + (uint32_t)(1 - reg) * kTwRegion;        // A operand region (src == 2)
            uint32_t leader;
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
            // opnd[0]: the accumulator region is drained and the first `ksplit` K columns of the operand are converted (all of
            // it when ksplit == 0 / shared-memory operands); opnd[1]: the rest.  Starting the K-low slabs under the second
            // half of the producing layer's epilogue is safe: they read converted columns only and write the region whose
            // last readers (that layer's MMAs) precede them in the in-order tensor pipe.
            const int ksplit = ts.job_ksplit[ji];
            mbar_wait(&bars->opnd[0], opnd_cnt & 1u);
            bool all_ready = false;
            if (ksplit == 0) { mbar_wait(&bars->opnd[1], opnd_cnt & 1u); all_ready = true; }
            opnd_cnt++;
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int Nh = half == 0 ? n0 : N - n0;
                if (Nh == 0) break;
                const uint32_t acc = tmem_base + (uint32_t)reg * kTwRegion + (uint32_t)(half == 0 ? 0 : n0);
                const uint32_t idesc = umma_idesc_f16(128, Nh);
                int Ks = (dm.slot_bytes / (Nh * 2 * PARTS)) & ~15;
                Ks = Ks < 16 ? 16 : Ks;
                uint32_t accum = 0u;
#pragma unroll 1
                for (int k0 = 0; k0 < Kp; k0 += Ks) {
                    const int ksl = Ks < Kp - k0 ? Ks : Kp - k0, nk = ksl >> 4, kg0 = k0 >> 4;
                    if (!all_ready && k0 + ksl > ksplit) { mbar_wait(&bars->opnd[1], (opnd_cnt - 1u) & 1u); all_ready = true; tc_fence_after(); }
                    mbar_wait(&bars->full[slot], phase);
                    tc_fence_after();
                    const uint32_t b_hi = smem_u32(m.slots + (size_t)slot * dm.slot_bytes), sboB = (uint32_t)(ksl >> 3) * 128u;
                    const uint64_t dhi = umma_desc(b_hi, 128, sboB), dlo = umma_desc(b_hi + (uint32_t)(Nh * ksl * 2), 128, sboB);
                    if (leader) {
                        if (src == 2) issue_slab_n<NPASS, true>(nk, acc, ta + (uint32_t)(kg0 * 16), 0, 0, dhi, dlo, idesc, accum);
                        else issue_slab_n<NPASS, false>(nk, acc, 0, ahi + (uint64_t)(kg0 * 16), alo + (uint64_t)(kg0 * 16), dhi, dlo, idesc, accum);
                        accum = 1u;
                        umma_commit(&bars->empty[slot]);          // slab consumed -> slot back to the producer
                    }
                    __syncwarp();
                    if (++slot == dm.nslots) { slot = 0; phase ^= 1u; }
                }
                if (leader) umma_commit(&bars->done[half]);
                __syncwarp();
            }
        }
    TW_STEP_END
    TW_ITEM_END
}

// per-thread view of the tile
struct XW {
    int row, part, lane;
    uint32_t tb;              // tensor-memory base incl. this warp's lane quadrant
    BarsW* bars;
    uint32_t cnt0, cnt1;      // phases of done[0] / done[1] consumed so far
    float* part_buf;
    const float* tails;
};
__device__ __forceinline__ void wait_half(XW& x, int half) {
    if (half == 0) { mbar_wait(&x.bars->done[0], x.cnt0 & 1u); x.cnt0++; }
    else { mbar_wait(&x.bars->done[1], x.cnt1 & 1u); x.cnt1++; }
    tc_fence_after();
}
// this warp's tensor-memory accesses and shared-memory operand stores are complete -> the next job may be issued
// which: 0 -> opnd[0] only (first half of a layer converted), 1 -> opnd[1] only, 2 -> both (everything else)
__device__ __forceinline__ void arrive_w(const XW& x, int which = 2) {
    tmem_wait_ld();
    tmem_wait_st();
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (x.lane == 0) {
        if (which != 1) mbar_arrive(&x.bars->opnd[0]);
        if (which != 0) mbar_arrive(&x.bars->opnd[1]);
    }
}
__device__ __forceinline__ float row_sum_w(const XW& x, int slot, float v) {      // slot in {0, 1}
    float* p = x.part_buf + slot * 512;
    p[x.part * 128 + x.row] = v;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    return (p[x.row] + p[128 + x.row]) + (p[256 + x.row] + p[384 + x.row]);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// 16 accumulator values of this thread's row (+ bias) -> activation -> IN PLACE the K-step of the next layer's TS operand
// (8 packed hi columns, 8 packed lo columns), or (DOT) contracted with the surviving row of the final H -> 2 Linear.
template <int NPASS, int ACT, bool DOT>
__device__ __forceinline__ void epi_math16(float* u, uint32_t acc, int c0, const float* __restrict__ bias, const float* __restrict__ wlast,
                                           float& dot) {
    const float4 b0 = ldg4(bias + c0), b1 = ldg4(bias + c0 + 4), b2 = ldg4(bias + c0 + 8), b3 = ldg4(bias + c0 + 12);
    u[0] += b0.x; u[1] += b0.y; u[2] += b0.z; u[3] += b0.w; u[4] += b1.x; u[5] += b1.y; u[6] += b1.z; u[7] += b1.w;
    u[8] += b2.x; u[9] += b2.y; u[10] += b2.z; u[11] += b2.w; u[12] += b3.x; u[13] += b3.y; u[14] += b3.z; u[15] += b3.w;
    if (ACT == ACT_T) {
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
            if (NPASS == 3) tanh_acc2(u[q], u[q + 1]);
            else { asm("tanh.approx.f32 %0, %0;" : "+f"(u[q])); asm("tanh.approx.f32 %0, %0;" : "+f"(u[q + 1])); }
        }
    } else if (DOT) {
#pragma unroll
        for (int q = 0; q < 16; ++q) u[q] = fmaxf(u[q], 0.0f);
    }
    if (DOT) {
#pragma unroll
        for (int q = 0; q < 16; q += 4) {
            const float4 w = ldg4(wlast + c0 + q);
            dot = fmaf(u[q], w.x, dot); dot = fmaf(u[q + 1], w.y, dot); dot = fmaf(u[q + 2], w.z, dot); dot = fmaf(u[q + 3], w.w, dot);
        }
    } else {
        uint32_t hi[8], lo[8];
        split8<NPASS, ACT == ACT_R>(u, hi, lo);
        split8<NPASS, ACT == ACT_R>(u + 8, hi + 4, lo + 4);
        tmem_st8u(acc + (uint32_t)c0, hi);
        if (NPASS == 3) tmem_st8u(acc + (uint32_t)c0 + 8u, lo);
    }
}

// The 16-column groups c0, c0 + 64, ... < cend of this thread.  (Two groups in flight per iteration was tried: 32 live
// accumulator values + their hi / lo pairs do not fit the register budget of a 640-thread CTA and spill 600 B.)
template <int NPASS, int ACT, bool DOT>
__device__ __forceinline__ void epi_span(uint32_t acc, int c0, int cend, const float* __restrict__ bias, const float* __restrict__ wlast, float& dot) {
#pragma unroll 1
    for (; c0 < cend; c0 += 64) {
        float u[16];
        tmem_ld16(acc + (uint32_t)c0, u);
        tmem_wait_ld();
        epi_math16<NPASS, ACT, DOT>(u, acc, c0, bias, wlast, dot);
    }
}

// Epilogue of one net-layer: first half as soon as its MMAs are done, second half after the others (its MMAs ran under
// the first half's epilogue).  This thread owns the 16-column groups part, part + 4, ... of each half.
// A converting (non-DOT) layer hands its operand over itself: opnd[0] after the first half (two-half layers), opnd[1] /
// both at the end; a DOT layer leaves both arrivals to the caller.
template <int NPASS, int ACT, bool DOT>
__device__ __noinline__ float epi_wide(XW& x, uint32_t acc, const float* __restrict__ bias, const float* __restrict__ wlast, int W, int n0) {
    float dot = 0.0f;
    wait_half(x, 0);
    epi_span<NPASS, ACT, DOT>(acc, x.part * 16, n0, bias, wlast, dot);
    if (n0 < W) {
        if (!DOT) arrive_w(x, 0);
        wait_half(x, 1);
        epi_span<NPASS, ACT, DOT>(acc, n0 + x.part * 16, W, bias, wlast, dot);
        if (!DOT) arrive_w(x, 1);
    } else if (!DOT) {
        arrive_w(x, 2);
    }
    return dot;
}

// ===== compute warps: thread = (row, column part); MODE 0 per-step density, 1 chain, 2 sampler (CIF_step.inverse) =====
template <int NPASS, int MODE>
__device__ __forceinline__ void tw_compute(const TwStep* __restrict__ steps, const float* __restrict__ tails, const RowArgs& a, const SampleOut& o,
                                           const TwDims& dm, long long tiles_per_step, long long total_items, uint8_t* smem, uint32_t tmem_base) {
    constexpr int NPARTS = 4;
    constexpr bool INV = MODE == 2;
    const SmemW m = tw_smem(smem, dm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    XW x;
    x.lane = lane; x.part = warp >> 2; x.row = (warp & 3) * 32 + lane;
    x.tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    x.bars = m.bars; x.cnt0 = 0; x.cnt1 = 0; x.part_buf = m.part_buf;
    x.tails = tails;
    const uint32_t inu_stride = m.inu_stride;
    const uint32_t tX = x.tb;                                          // region 0: where every L3 / last layer lands
    // this thread's chunks of the density means: stash[(d0 / 8) * 128 + row] (8 floats), written and re-read by the same thread
    float* stash = reinterpret_cast<float*>(m.stash) + (size_t)x.row * 8;
    TW_ITEM_BEGIN(MODE)
        long long n = tile * 128 + x.row;
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
    TW_STEP_BEGIN(MODE)
        const int C = ts.C, c = ts.c, Kin = ts.Kin, Ku = ts.Ku, Wd = ts.Wd, CPd = ts.CPd, Hc = ts.Hc;
        if (INV) lacc = 0.0f;       // the sampler reports every step's own log-det term
        // shared addresses of this row inside the first-layer operand tiles [IN hi, IN lo, U hi, U lo]
        const uint32_t in_row = smem_u32(m.inu) + (uint32_t)((x.row >> 3) * (Kin >> 3) * 128 + (x.row & 7) * 16);
        const uint32_t u_row = smem_u32(m.inu) + 2u * inu_stride + (uint32_t)((x.row >> 3) * (Ku >> 3) * 128 + (x.row & 7) * 16);
        int ji = 0;                 // schedule position of the job whose epilogue runs next
        // ---- input operand [z (2), cond (C0), prefix (2*step), 0...] ----
        {
            const float* cp = src_at(a.cond, ag, pt, INV ? step : jstep) - 2;     // chain mode: cond[:, j] for every chain step
            // sampler: the prefix is the row's own earlier samples, read back from the output (written by the part-0
            // threads before the CTA-wide barrier of the step-table reload; coherent loads, not the read-only path)
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
                t3::sts128(in_row + (uint32_t)(k0 >> 3) * 128u, hi);
                if (NPASS == 3) t3::sts128(in_row + inu_stride + (uint32_t)(k0 >> 3) * 128u, lo);
            }
        }
        arrive_w(x);

#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            // ============ density nets: shift net (means -> stash), then log-scale net (-> accumulator) ============
            const float *bias_mu = nullptr, *bias_ls = nullptr;
#pragma unroll 1
            for (int net = 0; net < 2; ++net) {
#pragma unroll 1
                for (int l = 0; l < 2; ++l) {
                    const float* bias = x.tails + ts.job_tail[tw_img_job(ts, ji, INV)];
                    const uint32_t acc = x.tb + (uint32_t)ts.job_reg[ji] * kTwRegion;
                    const int n0 = ts.job_n0[ji];
                    ++ji;
                    epi_wide<NPASS, ACT_R, false>(x, acc, bias, nullptr, Wd, n0);      // waits for the halves and hands the operand over itself
                }
                const float* b3 = x.tails + ts.job_tail[tw_img_job(ts, ji, INV)];
                ++ji;
                if (net == 0) bias_mu = b3; else bias_ls = b3;
                wait_half(x, 0);                 // L3 is narrow (<= 96 columns): one half, lands in region 0
                if (net == 0) {                  // park the means (+ bias) while the log-scale net runs through both regions
                    for (int d0 = x.part * 8; d0 < CPd; d0 += 8 * NPARTS) {
                        float mu[8];
                        tmem_ld8(tX + (uint32_t)d0, mu);
                        const float4 ba = ldg4(b3 + d0), bb = ldg4(b3 + d0 + 4);
                        tmem_wait_ld();
                        float* sp = stash + (size_t)(d0 >> 3) * 1024;
                        *reinterpret_cast<float4*>(sp) = make_float4(mu[0] + ba.x, mu[1] + ba.y, mu[2] + ba.z, mu[3] + ba.w);
                        *reinterpret_cast<float4*>(sp + 4) = make_float4(mu[4] + bb.x, mu[5] + bb.y, mu[6] + bb.z, mu[7] + bb.w);
                    }
                    arrive_w(x);                 // log-scale net L1: operand IN is in place, region 0 is free again
                }
            }
            // ---- L3 outputs: means in the stash, log-stddevs in the accumulator ----
            if (which == 0) {
                // sample u = exp(ls) eps + mu (fastpredNF.py:727-733) -> [u, mx] operand; log density of the drawn u
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
                        {
                            const float* sp = stash + (size_t)(d0 >> 3) * 1024;
                            const float4 ma = *reinterpret_cast<const float4*>(sp), mb = *reinterpret_cast<const float4*>(sp + 4);
                            mu[0] = ma.x; mu[1] = ma.y; mu[2] = ma.z; mu[3] = ma.w; mu[4] = mb.x; mu[5] = mb.y; mu[6] = mb.z; mu[7] = mb.w;
                        }
                        tmem_ld8(tX + (uint32_t)d0, ls);
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
                        for (int i = 0; i < 8; ++i) {      // branch-free, see kernels_tc3.cuh (padded columns: mu = ls = 0)
                            const float mm = mu[i], l = ls[i] + __ldg(bias_ls + d0 + i);
                            const float en = d0 + i < C ? e[i] : 0.0f;
                            u[i] = clamp_h(fmaf(t3::fast_exp(l), en, mm));
                            p_ls += l;
                            p_e2 = fmaf(en, en, p_e2);
                        }
                    }
                    if (d0 == (C & ~7)) {        // mx = z * mask of the first coupling sits at columns C, C+1 (C is even)
                        const int cm = C & 7;
                        const float a0 = m0 == 0 ? clamp_h(z0) : 0.0f, a1 = m0 == 1 ? clamp_h(z1) : 0.0f;
#pragma unroll
                        for (int i = 0; i < 8; i += 2) { u[i] = i == cm ? a0 : u[i]; u[i + 1] = i == cm ? a1 : u[i + 1]; }
                    }
                    uint32_t hi[4], lo[4];
                    split8<3, false>(u, hi, lo);     // u is always kept as a hi+lo pair: the second density needs it back at full precision
                    t3::sts128(u_row + (uint32_t)(d0 >> 3) * 128u, hi);
                    t3::sts128(u_row + inu_stride + (uint32_t)(d0 >> 3) * 128u, lo);
                }
                const float s_ls = row_sum_w(x, 0, p_ls), s_e2 = row_sum_w(x, 1, p_e2);
                const float lp_u = (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_e2;
                lacc += INV ? lp_u : -lp_u;                                                // forward: -log r, inverse: +log q
                arrive_w(x);                                                               // first coupling job may start
            } else {
                // log density of u under the second density (fastpredNF.py:713-724), u reconstructed from its hi+lo operand
                float p_ls = 0.0f, p_q = 0.0f;
                for (int d0 = x.part * 8; d0 < CPd; d0 += 8 * NPARTS) {
                    float mu[8], ls[8];
                    uint32_t uh[4], ul[4];
                    {
                        const float* sp = stash + (size_t)(d0 >> 3) * 1024;
                        const float4 ma = *reinterpret_cast<const float4*>(sp), mb = *reinterpret_cast<const float4*>(sp + 4);
                        mu[0] = ma.x; mu[1] = ma.y; mu[2] = ma.z; mu[3] = ma.w; mu[4] = mb.x; mu[5] = mb.y; mu[6] = mb.z; mu[7] = mb.w;
                    }
                    tmem_ld8(tX + (uint32_t)d0, ls);
                    t3::lds128u(u_row + (uint32_t)(d0 >> 3) * 128u, uh);
                    t3::lds128u(u_row + inu_stride + (uint32_t)(d0 >> 3) * 128u, ul);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {      // branch-free: padded columns carry mu = ls = 0 and u is masked
                        const __half2 h2 = *reinterpret_cast<const __half2*>(&uh[i >> 1]), l2 = *reinterpret_cast<const __half2*>(&ul[i >> 1]);
                        const float2 hf = __half22float2(h2), lf = __half22float2(l2);
                        const float u = d0 + i < C ? ((i & 1) ? hf.y + lf.y : hf.x + lf.x) : 0.0f;     // columns C, C+1 of the operand hold mx, not u
                        const float mm = mu[i], l = ls[i] + __ldg(bias_ls + d0 + i);
                        const float df = (u - mm) * t3::fast_exp(-l);
                        p_q = fmaf(df, df, p_q);
                        p_ls += l;
                    }
                }
                const float s_ls = row_sum_w(x, 0, p_ls), s_q = row_sum_w(x, 1, p_q);
                const float lp_u = (-0.5f * (float)C * kLog2Pi - s_ls) - 0.5f * s_q;
                lacc += INV ? -lp_u : lp_u;                                                // forward: +log q, inverse: -log r
                break;                                                                     // end of the step: nothing to hand over
            }

            // ============ n_blocks x (masked affine coupling + BatchNorm) ============
#pragma unroll 1
            for (int bi = 0; bi < ts.n_blocks; ++bi) {
                const int b = INV ? ts.n_blocks - 1 - bi : bi;
                const int j = 1 - ts.masked_dim[b];
                float ps = 0.0f, pt_ = 0.0f, bs = 0.0f, bt = 0.0f;
#pragma unroll 1
                for (int net = 0; net < 2; ++net) {
#pragma unroll 1
                    for (int l = 0; l <= ts.n_hidden; ++l) {
                        const bool last = l == ts.n_hidden;
                        const float* tail = x.tails + ts.job_tail[tw_img_job(ts, ji, INV)];
                        const uint32_t acc = x.tb + (uint32_t)ts.job_reg[ji] * kTwRegion;
                        const int n0 = ts.job_n0[ji];
                        ++ji;
                        if (!last) {
                            if (net == 0) epi_wide<NPASS, ACT_T, false>(x, acc, tail, nullptr, Hc, n0);
                            else epi_wide<NPASS, ACT_R, false>(x, acc, tail, nullptr, Hc, n0);
                        } else {
                            float pd;
                            if (net == 0) pd = epi_wide<NPASS, ACT_T, true>(x, acc, tail, tail + Hc, Hc, n0);
                            else pd = epi_wide<NPASS, ACT_R, true>(x, acc, tail, tail + Hc, Hc, n0);
                            if (net == 0) { ps = pd; bs = __ldg(tail + 2 * Hc); arrive_w(x); }   // t-net L0: operand U unchanged, region 0 drained
                            else { pt_ = pd; bt = __ldg(tail + 2 * Hc); }
                        }
                    }
                }
                // both nets done: affine update of the unmasked coordinate + BatchNorm (TFCondARFlow.py:360-381, 303-315)
                const float s = row_sum_w(x, 0, ps) + bs, t = row_sum_w(x, 1, pt_) + bt;
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
                uint32_t dst_hi;              // shared address of the (hi) pair; the lo pair sits one tile further
                int owner;
                if (more) {
                    const int mn = ts.masked_dim[bn];
                    p0 = mn == 0 ? clamp_h(z0) : 0.0f; p1 = mn == 1 ? clamp_h(z1) : 0.0f;
                    dst_hi = u_row + (uint32_t)(C >> 3) * 128u + (uint32_t)(C & 7) * 2u;
                    owner = (C >> 3) % NPARTS;
                } else {
                    p0 = clamp_h(z0); p1 = clamp_h(z1);
                    dst_hi = in_row;
                    owner = 0;
                }
                if (x.part == owner) {
                    const __half2 h = __floats2half2_rn(p0, p1);
                    const float2 f = __half22float2(h);
                    const __half2 lo2 = __floats2half2_rn(p0 - f.x, p1 - f.y);
                    t3::sts32(dst_hi, *reinterpret_cast<const uint32_t*>(&h));
                    t3::sts32(dst_hi + inu_stride, *reinterpret_cast<const uint32_t*>(&lo2));
                }
                arrive_w(x);
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
    TW_STEP_END
        if (!INV && x.part == 0) {      // whole warps: store_result reduces K importance samples across lanes
            if (MODE == 0) {
                const float lp = normal_lp2(z0, z1, prev0, prev1, m.cur->std[0], m.cur->std[1]) + lacc;
                store_result(a, ag, rem, jstep, lp, valid);
            } else {      // base = Normal(base_pos, std[0]); ldj = mean over the chain steps of this result index
                const float lp = normal_lp2(z0, z1, prev0, prev1, __ldg(&steps[0].std[0]), __ldg(&steps[0].std[1])) + lacc / (float)(jstep - dm.jlo + 1);
                store_result(a, ag, rem, jstep - dm.jlo, lp, valid);
            }
        }
    TW_ITEM_END
}

}  // namespace tw

template <int NPASS, int MODE>
__global__ void __launch_bounds__(kTwThreads, 1)
k_cif_wide(const TwStep* __restrict__ steps, const uint8_t* __restrict__ img, const float* __restrict__ tails, const __grid_constant__ RowArgs a,
           const __grid_constant__ TwDims dm, long long tiles_per_step, long long total_items, const __grid_constant__ SampleOut o) {
    using namespace tc;
    using namespace tw;
    extern __shared__ __align__(128) uint8_t smem_tw[];
    uint8_t* smem = smem_tw;
    BarsW* bars = reinterpret_cast<BarsW*>(smem);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 240);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < dm.nslots; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->done[0], 1);
        mbar_init(&bars->done[1], 1);
        mbar_init(&bars->opnd[0], 16);
        mbar_init(&bars->opnd[1], 16);
        mbar_fence_init();
    }
    if (warp == 16) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tslot);

    // warp-specialised register split as in the two-tile kernel: 96 at launch, service warpgroup down to 56, compute up to 104
    if (warp < 16) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTwComputeRegs));
        tw_compute<NPASS, MODE>(steps, tails, a, o, dm, tiles_per_step, total_items, smem, tmem_base);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTwServiceRegs));
        if (warp == 16) tw_producer<NPASS>(steps, img, dm, tiles_per_step, total_items, smem);
        else if (warp == 17) tw_issuer<NPASS>(steps, dm, tiles_per_step, total_items, smem, tmem_base);
        else tw_idle(steps, dm, tiles_per_step, total_items, smem);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// mode: 0 per-step density (a8), 1 chain (a9: a.seq_klo / a.seq_jlo), 2 sampler (a10)
template <int NPASS>
inline int tcw_launch(TwDims dm, const TwStep* d_steps, const unsigned char* d_img, const float* d_tails, const RowArgs& a, const SampleOut& o,
                      int mode, int num_sms, int smem_optin, cudaStream_t stream, uint64_t* launches, std::string* msg) {
    if (!dm.supported || d_img == nullptr) {
        *msg = "the tensor-core path does not support this model shape (needs hidden_size % 16 == 0, hidden_size <= 256, an even "
               "cond_size and 2 * (cond_size + 2 * pred_len) <= 256); use FC_PREC_FP32";
        return 1;
    }
    const int T = dm.T;
    const long long tiles = (a.N + 127) / 128;
    dm.seq = mode;
    if (mode == 0) { dm.step_lo = 0; dm.step_hi = T - 1; }
    else if (mode == 1) { dm.step_lo = a.seq_jlo; dm.step_hi = T - 1; dm.klo = a.seq_klo; dm.jlo = a.seq_jlo; }
    else { dm.step_lo = 0; dm.step_hi = 0; }
    const long long items = tiles * (dm.step_hi - dm.step_lo + 1);
    const size_t smem = tcw_smem_bytes(dm);
    if (smem > (size_t)smem_optin) { *msg = "wide tensor-core kernel: shared-memory budget exceeded"; return 1; }
    const int grid = (int)(items < num_sms ? items : num_sms);
    cudaError_t e = cudaSuccess;
    auto go = [&](auto kern) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<grid, kTwThreads, smem, stream>>>(d_steps, d_img, d_tails, a, dm, tiles, items, o);
    };
    if (mode == 0) go(k_cif_wide<NPASS, 0>);
    else if (mode == 1) go(k_cif_wide<NPASS, 1>);
    else go(k_cif_wide<NPASS, 2>);
    if (e != cudaSuccess) { *msg = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
    (*launches)++;
    e = cudaGetLastError();
    if (e != cudaSuccess) { *msg = std::string("k_cif_wide launch: ") + cudaGetErrorString(e); return 1; }
    return 0;
}

}  // namespace fc
