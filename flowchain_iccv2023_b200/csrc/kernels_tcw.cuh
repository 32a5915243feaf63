// kernels_tcw.cuh -- f16x3 / f16 tensor-core path for WIDE models: hidden sizes up to 256 (BASELINE config 5) and
// conditioning lengths up to 64 (the reference's tuning range, src/main.py:222-229), i.e. every shape the two-tile
// kernel (kernels_tc3.cuh, layers <= 80 wide) refuses.
//
// Same arithmetic as the two-tile kernel (split-fp16 operands, tcgen05.mma kind::f16, fp32 accumulate in tensor
// memory, tc_f16x3.cuh), different resource plan, because a 256-wide layer needs 256 accumulator columns and a
// [128 x 256] hi+lo operand is 128 KB of shared memory:
//   * ONE 128-row tile per CTA, 16 epilogue warps (thread = (row = TMEM lane, one of 4 column parts)) + a service
//     warpgroup (TMA weight producer, MMA issuer);
//   * the two nets of a sub-network (s|t, shift|log-scale) run ONE AFTER THE OTHER through a single accumulator
//     [0, A) and a single hidden-operand buffer in shared memory (SS-mode MMAs; the epilogue of layer l overwrites the
//     operand layer l consumed, which is safe because its MMAs have completed); the means of a density net are parked
//     in a stash accumulator [A, A + CP) while the log-scale net runs; the narrow first-layer operands [z|w, y] and
//     [u, mx] live in tensor memory (TS-mode MMAs) behind the stash;
//   * a 256 x 256 weight matrix (hi + lo = 256 KB) does not fit next to the operands, so every net-layer's weights
//     are packed as K-SLABS ([N x ks] hi tile + lo tile, <= 16 KB, each in the canonical K-major layout) that stream
//     from L2 through a TMA ring; a slab is released by the tcgen05.commit of the MMAs that read it;
//   * biases / last-row vectors are read by the epilogue warps straight from L2/L1 (warp-uniform float4 loads).
// One job = one net-layer.  Protocol: the compute warps arrive on `opnd` once before every job (operand written,
// accumulator drained); the issuer waits for it, issues the job's MMAs slab by slab and commits `mma_done`; the
// compute warps wait for `mma_done` before every epilogue.
#pragma once
#include "kernels_tc3.cuh"

namespace fc {

constexpr int kTwMaxJobs = 12 + 2 * (kMaxHidden + 1) * kMaxBlocks;
constexpr int kTwThreads = 16 * 32 + 128;       // 16 compute warps + a service warpgroup (producer, issuer, 2 idle)
constexpr int kTwComputeRegs = 104, kTwServiceRegs = 56;
constexpr int kTwHeader = 8192;                 // barriers, step table, row-partial buffers
constexpr int kTwStepOff = 256, kTwPartOff = 3072;
constexpr int kTwSlotBytes = 16384;
constexpr int kTwMaxSlots = 8;

struct TwStep {
    int32_t C, c, Kin, Ku, Wd, CPd, Hc;
    int32_t n_blocks, n_hidden, njobs, eps_off, seq_eps_off;
    int32_t masked_dim[kMaxBlocks];
    float bn_scale[kMaxBlocks][2], bn_shift[kMaxBlocks][2], bn_ldj[kMaxBlocks];
    float bn_iscale[kMaxBlocks][2], bn_ishift[kMaxBlocks][2];
    float std[2];
    // jobs in forward order: [r density: net 0 L1 L2 L3, net 1 L1 L2 L3][block b: s-net L0..Ln, t-net L0..Ln] x n_blocks[q density]
    uint32_t job_img[kTwMaxJobs];      // byte offset / 16 of the job's first slab in the image
    uint32_t job_tail[kTwMaxJobs];     // float offset of the job's fp32 tail (bias[N] (+ last-row vector[N], its bias))
    uint16_t job_kp[kTwMaxJobs], job_n[kTwMaxJobs], job_ks[kTwMaxJobs];   // K (padded), N (padded), slab width
    uint8_t job_src[kTwMaxJobs];       // A operand: 0 = IN (tensor memory), 1 = U (tensor memory), 2 = hidden (shared memory)
    uint8_t job_dst[kTwMaxJobs];       // accumulator: 0 = [0, N), 1 = stash [A, A + N)
};
static_assert(sizeof(TwStep) <= kTwPartOff - kTwStepOff, "TwStep must fit the header");

struct TwDims {
    int32_t supported;
    int32_t T, C0;
    int32_t Wmax;                          // widest hidden operand: stride of the shared-memory operand tiles
    int32_t stash_col;                     // A = widest accumulator
    int32_t in_hi, in_lo, u_hi, u_lo;      // tensor-memory columns of the first-layer operands
    int32_t slot_bytes, nslots;
    int32_t step_lo, step_hi;              // per-step mode: steps covered; chain mode: result indices
    int32_t seq;                           // 0 per-step density, 1 chain, 2 sampler
    int32_t klo, jlo;
};

// slab width of a job: as many 16-wide K steps as fit a ring slot
inline int tw_slab_k(int N, int Kp, int parts, int slot_bytes) {
    int ks = (slot_bytes / (N * 2 * parts)) & ~15;
    if (ks < 16) ks = 16;
    return ks < Kp ? ks : Kp;
}

// Packs the wide image.  dm.supported = 0 for shapes this kernel does not cover.
inline void tcw_pack(const fc_model_desc& d, const std::vector<StepPlan>& plans, const std::vector<TcStepSrc>& src, int npass,
                     std::vector<unsigned char>& img, std::vector<float>& tails, std::vector<TwStep>& steps, TwDims& dm,
                     int smem_optin) {
    img.clear(); tails.clear(); steps.clear();
    memset(&dm, 0, sizeof dm);
    const int T = d.pred_len, H = d.hidden_size, parts = npass == 3 ? 2 : 1;
    const int Cmax = d.cond_size + (T - 1) * 2, cmax = Cmax + 2;
    const int Wd_max = pad16i(2 * cmax), Kmax = pad16i(cmax), CPmax = pad16i(Cmax);
    if (H % 16 != 0 || H > 256 || (d.cond_size & 1) || Wd_max > 256) return;
    const int A = Wd_max > H ? Wd_max : H;
    if (A + CPmax + 2 * Kmax > 512) return;                       // tensor-memory columns
    dm.T = T; dm.C0 = d.cond_size; dm.Wmax = A; dm.stash_col = A;
    dm.in_hi = A + CPmax; dm.in_lo = dm.in_hi + Kmax / 2; dm.u_hi = dm.in_hi + Kmax; dm.u_lo = dm.u_hi + Kmax / 2;
    dm.slot_bytes = kTwSlotBytes;
    const long long ring = (long long)smem_optin - kTwHeader - (long long)2 * 128 * A * 2;
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
        // one job: W [rows x K] (row-major, nn.Linear.weight) -> K-slabs of canonical K-major tiles; tail = bias (+ extra)
        auto job = [&](const float* W, const float* bias, int rows, int K, int N, int Kp, int src_, int dst_, const float* extra, int n_extra) {
            while (img.size() % 128) img.push_back(0);
            ts.job_img[nj] = (uint32_t)(img.size() / 16);
            const int Ks = tw_slab_k(N, Kp, parts, dm.slot_bytes);
            for (int k0 = 0; k0 < Kp; k0 += Ks) {
                const int ksl = Ks < Kp - k0 ? Ks : Kp - k0;
                const int kv = K - k0 < 0 ? 0 : (K - k0 < ksl ? K - k0 : ksl);      // valid (unpadded) columns of this slab
                sub.assign((size_t)rows * (kv > 0 ? kv : 1), 0.0f);
                for (int r = 0; r < rows; ++r)
                    for (int k = 0; k < kv; ++k) sub[(size_t)r * kv + k] = W[(size_t)r * K + k0 + k];
                for (int part = 0; part < parts; ++part) {
                    const size_t o = img.size();
                    img.resize(o + (size_t)N * ksl * 2);
                    pack_kmajor_f16(img.data() + o, sub.data(), rows, kv, N, ksl, part);
                }
            }
            while (tails.size() % 4) tails.push_back(0.0f);
            ts.job_tail[nj] = (uint32_t)tails.size();
            const size_t o = tails.size();
            tails.resize(o + (size_t)N + n_extra, 0.0f);
            memcpy(tails.data() + o, bias, (size_t)rows * 4);
            if (n_extra) memcpy(tails.data() + o + N, extra, (size_t)n_extra * 4);
            ts.job_kp[nj] = (uint16_t)Kp; ts.job_n[nj] = (uint16_t)N; ts.job_ks[nj] = (uint16_t)Ks;
            ts.job_src[nj] = (uint8_t)src_; ts.job_dst[nj] = (uint8_t)dst_;
            ++nj;
        };
        auto dens = [&](int which) {
            const int h2 = 2 * c;
            for (int net = 0; net < 2; ++net) {
                job(s.dens_W[which][net][0], s.dens_b[which][net][0], h2, c, ts.Wd, ts.Kin, 0, 0, nullptr, 0);
                job(s.dens_W[which][net][1], s.dens_b[which][net][1], h2, h2, ts.Wd, ts.Wd, 2, 0, nullptr, 0);
                job(s.dens_W[which][net][2], s.dens_b[which][net][2], C, h2, ts.CPd, ts.Wd, 2, net == 0 ? 1 : 0, nullptr, 0);
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
                    job(Wl, bl, H, Kl, H, l == 0 ? ts.Ku : H, l == 0 ? 1 : 2, 0, extra.empty() ? nullptr : extra.data(), (int)extra.size());
                }
        }
        dens(1);
        ts.njobs = nj;
    }
    while (img.size() % 128) img.push_back(0);
    while (tails.size() % 4) tails.push_back(0.0f);
    dm.supported = 1;
}

inline size_t tcw_smem_bytes(const TwDims& dm) { return (size_t)kTwHeader + (size_t)2 * 128 * dm.Wmax * 2 + (size_t)dm.nslots * dm.slot_bytes; }

namespace tw {
using namespace tc;
using namespace t2;

struct BarsW {
    uint64_t full[kTwMaxSlots], empty[kTwMaxSlots];
    uint64_t mma_done, opnd;
};
static_assert(sizeof(BarsW) <= 240, "barrier block");

struct SmemW {
    BarsW* bars;
    TwStep* cur;
    float* part_buf;          // [2 slots][4 parts][128 rows]
    uint8_t *hbuf, *slots;
    uint32_t h_stride;
};
__device__ __forceinline__ SmemW tw_smem(uint8_t* smem, const TwDims& dm) {
    SmemW m;
    m.bars = reinterpret_cast<BarsW*>(smem);
    m.cur = reinterpret_cast<TwStep*>(smem + kTwStepOff);
    m.part_buf = reinterpret_cast<float*>(smem + kTwPartOff);
    m.h_stride = (uint32_t)(128 * dm.Wmax * 2);
    m.hbuf = smem + kTwHeader;                     // [hi, lo] x [128 x Wmax] fp16
    m.slots = m.hbuf + (size_t)2 * m.h_stride;
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

// ===== weight producer (one warp): streams the K-slabs of every job of the step through the ring =====
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
            const int N = ts.job_n[pj], Kp = ts.job_kp[pj], Ks = ts.job_ks[pj];
            const uint8_t* srcp = img + (size_t)ts.job_img[pj] * 16;
            for (int k0 = 0; k0 < Kp; k0 += Ks) {
                const int ksl = Ks < Kp - k0 ? Ks : Kp - k0;
                const uint32_t bytes = (uint32_t)(N * ksl * 2 * PARTS);
                if (lane == 0) {
                    mbar_wait(&bars->empty[slot], phase ^ 1u);
                    mbar_expect_tx(&bars->full[slot], bytes);
                    bulk_load(m.slots + (size_t)slot * dm.slot_bytes, srcp, bytes, &bars->full[slot]);
                }
                __syncwarp();
                srcp += bytes;
                if (++slot == dm.nslots) { slot = 0; phase ^= 1u; }
            }
        }
    TW_STEP_END
    TW_ITEM_END
}

// ===== MMA issuer (one warp): job by job, slab by slab =====
template <int NPASS>
__device__ __forceinline__ void tw_issuer(const TwStep* __restrict__ steps, const TwDims& dm, long long tiles_per_step, long long total_items,
                                          uint8_t* smem, uint32_t tmem_base) {
    const SmemW m = tw_smem(smem, dm);
    BarsW* bars = m.bars;
    int slot = 0;
    uint32_t phase = 0, opnd_cnt = 0;
    const uint32_t sa_hi = smem_u32(m.hbuf), sa_lo = sa_hi + m.h_stride;
    TW_ITEM_BEGIN(dm.seq)
    TW_STEP_BEGIN(dm.seq)
#pragma unroll 1
        for (int ji = 0; ji < ts.njobs; ++ji) {
            const int Kp = ts.job_kp[ji], N = ts.job_n[ji], Ks = ts.job_ks[ji], src = ts.job_src[ji];
            const uint32_t acc = tmem_base + (ts.job_dst[ji] ? (uint32_t)dm.stash_col : 0u);
            const uint32_t idesc = umma_idesc_f16(128, N), sboA = (uint32_t)(Kp >> 3) * 128u;
            const uint32_t ta_hi = tmem_base + (uint32_t)(src == 0 ? dm.in_hi : dm.u_hi), ta_lo = tmem_base + (uint32_t)(src == 0 ? dm.in_lo : dm.u_lo);
            const uint64_t ahi = umma_desc(sa_hi, 128, sboA), alo = umma_desc(sa_lo, 128, sboA);
            uint32_t leader;
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
            // operand of this job written, accumulator drained by the previous epilogue
            mbar_wait(&bars->opnd, opnd_cnt & 1u);
            opnd_cnt++;
            tc_fence_after();
            uint32_t accum = 0u;
#pragma unroll 1
            for (int k0 = 0; k0 < Kp; k0 += Ks) {
                const int ksl = Ks < Kp - k0 ? Ks : Kp - k0, nk = ksl >> 4, kg0 = k0 >> 4;
                mbar_wait(&bars->full[slot], phase);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(m.slots + (size_t)slot * dm.slot_bytes), sboB = (uint32_t)(ksl >> 3) * 128u;
                const uint64_t dhi = umma_desc(b_hi, 128, sboB), dlo = umma_desc(b_hi + (uint32_t)(N * ksl * 2), 128, sboB);
                if (leader) {
                    // straight-line issue (compile-time operand offsets inside the slab), see t3::issue_job
                    const uint32_t tah = ta_hi + (uint32_t)(kg0 * 8), tal = ta_lo + (uint32_t)(kg0 * 8);
                    const uint64_t sah = ahi + (uint64_t)(kg0 * 16), sal = alo + (uint64_t)(kg0 * 16);
                    if (src < 2) {
                        switch (nk) {
                            case 1: t3::issue_job<NPASS, 1, true>(acc, tah, tal, 0, 0, dhi, dlo, idesc, accum); break;
                            case 2: t3::issue_job<NPASS, 2, true>(acc, tah, tal, 0, 0, dhi, dlo, idesc, accum); break;
                            case 3: t3::issue_job<NPASS, 3, true>(acc, tah, tal, 0, 0, dhi, dlo, idesc, accum); break;
                            case 4: t3::issue_job<NPASS, 4, true>(acc, tah, tal, 0, 0, dhi, dlo, idesc, accum); break;
                            default: t3::issue_job_loop<NPASS, true>(nk, acc, tah, tal, 0, 0, dhi, dlo, idesc, accum); break;
                        }
                    } else {
                        switch (nk) {
                            case 1: t3::issue_job<NPASS, 1, false>(acc, 0, 0, sah, sal, dhi, dlo, idesc, accum); break;
                            case 2: t3::issue_job<NPASS, 2, false>(acc, 0, 0, sah, sal, dhi, dlo, idesc, accum); break;
                            case 3: t3::issue_job<NPASS, 3, false>(acc, 0, 0, sah, sal, dhi, dlo, idesc, accum); break;
                            case 4: t3::issue_job<NPASS, 4, false>(acc, 0, 0, sah, sal, dhi, dlo, idesc, accum); break;
                            default: t3::issue_job_loop<NPASS, false>(nk, acc, 0, 0, sah, sal, dhi, dlo, idesc, accum); break;
                        }
                    }
                    accum = 1u;
                    umma_commit(&bars->empty[slot]);          // slab consumed -> slot back to the producer
                }
                __syncwarp();
                if (++slot == dm.nslots) { slot = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(&bars->mma_done);
            __syncwarp();
        }
    TW_STEP_END
    TW_ITEM_END
}

// per-thread view of the tile
struct XW {
    int row, part, lane;
    uint32_t tb;              // tensor-memory base incl. this warp's lane quadrant
    BarsW* bars;
    uint32_t mma_cnt;
    float* part_buf;
    uint32_t h_hi, h_lo;      // shared addresses of the hidden operand tiles
    const float* tails;
};
__device__ __forceinline__ void wait_mma_w(XW& x) {
    mbar_wait(&x.bars->mma_done, x.mma_cnt & 1u);
    x.mma_cnt++;
    tc_fence_after();
}
// this warp's tensor-memory accesses and shared-memory operand stores are complete -> the next job may be issued
__device__ __forceinline__ void arrive_w(const XW& x) {
    tmem_wait_ld();
    tmem_wait_st();
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (x.lane == 0) mbar_arrive(&x.bars->opnd);
}
__device__ __forceinline__ float row_sum_w(const XW& x, int slot, float v) {      // slot in {0, 1}
    float* p = x.part_buf + slot * 512;
    p[x.part * 128 + x.row] = v;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    return (p[x.row] + p[128 + x.row]) + (p[256 + x.row] + p[384 + x.row]);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 8 accumulator values (+ bias) -> activation -> hi / lo fp16 rows of the next layer's shared-memory operand tile,
// or (DOT) contracted with the surviving row of the final H -> 2 Linear
template <int NPASS, int ACT, bool DOT, bool WRITE>
__device__ __forceinline__ void epi_chunk_w(float* u, int c0, const float4 b0, const float4 b1, const float* __restrict__ wlast, uint32_t ohi, uint32_t olo,
                                            uint32_t koff, float& dot) {
    u[0] += b0.x; u[1] += b0.y; u[2] += b0.z; u[3] += b0.w; u[4] += b1.x; u[5] += b1.y; u[6] += b1.z; u[7] += b1.w;
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
        const float4 w0 = ldg4(wlast + c0), w1 = ldg4(wlast + c0 + 4);
        dot = fmaf(u[0], w0.x, dot); dot = fmaf(u[1], w0.y, dot); dot = fmaf(u[2], w0.z, dot); dot = fmaf(u[3], w0.w, dot);
        dot = fmaf(u[4], w1.x, dot); dot = fmaf(u[5], w1.y, dot); dot = fmaf(u[6], w1.z, dot); dot = fmaf(u[7], w1.w, dot);
    }
    if (WRITE) {
        uint32_t hi[4], lo[4];
        split8<NPASS, ACT == ACT_R>(u, hi, lo);
        const uint32_t off = koff + (uint32_t)(c0 >> 3) * 128u;      // canonical K-major tile: 8 K elements of a row = 16 bytes
        t3::sts128(ohi + off, hi);
        if (NPASS == 3) t3::sts128(olo + off, lo);
    }
}

// hidden-layer epilogue of one net: this thread owns the 8-column chunks part, part + 4, ... of its row
template <int NPASS, int ACT, bool DOT, bool WRITE>
__device__ __noinline__ float epi_wide(uint32_t acc, uint32_t ohi, uint32_t olo, const float* __restrict__ bias, const float* __restrict__ wlast,
                                       int part, int W, uint32_t koff) {
    float dot = 0.0f;
#pragma unroll 1
    for (int c0 = part * 8; c0 < W; c0 += 64) {
        const int c1 = c0 + 32;
        const bool two = c1 < W;
        float v0[8], v1[8];
        tmem_ld8(acc + (uint32_t)c0, v0);
        if (two) tmem_ld8(acc + (uint32_t)c1, v1);
        const float4 b00 = ldg4(bias + c0), b01 = ldg4(bias + c0 + 4);
        float4 b10 = b00, b11 = b01;
        if (two) { b10 = ldg4(bias + c1); b11 = ldg4(bias + c1 + 4); }
        tmem_wait_ld();
        epi_chunk_w<NPASS, ACT, DOT, WRITE>(v0, c0, b00, b01, wlast, ohi, olo, koff, dot);
        if (two) epi_chunk_w<NPASS, ACT, DOT, WRITE>(v1, c1, b10, b11, wlast, ohi, olo, koff, dot);
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
    x.bars = m.bars; x.mma_cnt = 0; x.part_buf = m.part_buf;
    x.h_hi = smem_u32(m.hbuf); x.h_lo = x.h_hi + m.h_stride;
    x.tails = tails;
    const uint32_t tIN_hi = x.tb + (uint32_t)dm.in_hi, tIN_lo = x.tb + (uint32_t)dm.in_lo;
    const uint32_t tU_hi = x.tb + (uint32_t)dm.u_hi, tU_lo = x.tb + (uint32_t)dm.u_lo;
    const uint32_t tACC = x.tb, tMU = x.tb + (uint32_t)dm.stash_col;
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
        const uint32_t koff_d = (uint32_t)((x.row >> 3) * (Wd >> 3) * 128 + (x.row & 7) * 16);
        const uint32_t koff_c = (uint32_t)((x.row >> 3) * (Hc >> 3) * 128 + (x.row & 7) * 16);
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
                tmem_st4(tIN_hi + (uint32_t)(k0 >> 1), hi);
                if (NPASS == 3) tmem_st4(tIN_lo + (uint32_t)(k0 >> 1), lo);
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
                    ++ji;
                    wait_mma_w(x);
                    epi_wide<NPASS, ACT_R, false, true>(tACC, x.h_hi, x.h_lo, bias, nullptr, x.part, Wd, koff_d);
                    arrive_w(x);
                }
                const float* b3 = x.tails + ts.job_tail[tw_img_job(ts, ji, INV)];
                ++ji;
                if (net == 0) bias_mu = b3; else bias_ls = b3;
                wait_mma_w(x);
                if (net == 0) arrive_w(x);      // log-scale net L1: operand IN is in place, the accumulator is free
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
                        tmem_ld8(tMU + (uint32_t)d0, mu);
                        tmem_ld8(tACC + (uint32_t)d0, ls);
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
                            const float mm = mu[i] + __ldg(bias_mu + d0 + i), l = ls[i] + __ldg(bias_ls + d0 + i);
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
                    tmem_st4(tU_hi + (uint32_t)(d0 >> 1), hi);
                    tmem_st4(tU_lo + (uint32_t)(d0 >> 1), lo);
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
                    tmem_ld8(tMU + (uint32_t)d0, mu);
                    tmem_ld8(tACC + (uint32_t)d0, ls);
                    tmem_ld4(tU_hi + (uint32_t)(d0 >> 1), uh);
                    tmem_ld4(tU_lo + (uint32_t)(d0 >> 1), ul);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {      // branch-free: padded columns carry mu = ls = 0 and u is masked
                        const __half2 h2 = *reinterpret_cast<const __half2*>(&uh[i >> 1]), l2 = *reinterpret_cast<const __half2*>(&ul[i >> 1]);
                        const float2 hf = __half22float2(h2), lf = __half22float2(l2);
                        const float u = d0 + i < C ? ((i & 1) ? hf.y + lf.y : hf.x + lf.x) : 0.0f;     // columns C, C+1 of the operand hold mx, not u
                        const float mm = mu[i] + __ldg(bias_mu + d0 + i), l = ls[i] + __ldg(bias_ls + d0 + i);
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
                        ++ji;
                        wait_mma_w(x);
                        if (!last) {
                            if (net == 0) epi_wide<NPASS, ACT_T, false, true>(tACC, x.h_hi, x.h_lo, tail, nullptr, x.part, Hc, koff_c);
                            else epi_wide<NPASS, ACT_R, false, true>(tACC, x.h_hi, x.h_lo, tail, nullptr, x.part, Hc, koff_c);
                            arrive_w(x);
                        } else {
                            float pd;
                            if (net == 0) pd = epi_wide<NPASS, ACT_T, true, false>(tACC, x.h_hi, x.h_lo, tail, tail + Hc, x.part, Hc, koff_c);
                            else pd = epi_wide<NPASS, ACT_R, true, false>(tACC, x.h_hi, x.h_lo, tail, tail + Hc, x.part, Hc, koff_c);
                            if (net == 0) { ps = pd; bs = __ldg(tail + 2 * Hc); arrive_w(x); }   // t-net L0: operand U unchanged, accumulator drained
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
                uint32_t dst_hi, dst_lo;
                int owner;
                if (more) {
                    const int mn = ts.masked_dim[bn];
                    p0 = mn == 0 ? clamp_h(z0) : 0.0f; p1 = mn == 1 ? clamp_h(z1) : 0.0f;
                    dst_hi = tU_hi + (uint32_t)(C >> 1); dst_lo = tU_lo + (uint32_t)(C >> 1);
                    owner = (C >> 3) % NPARTS;
                } else {
                    p0 = clamp_h(z0); p1 = clamp_h(z1);
                    dst_hi = tIN_hi; dst_lo = tIN_lo;
                    owner = 0;
                }
                if (x.part == owner) {
                    const __half2 h = __floats2half2_rn(p0, p1);
                    const float2 f = __half22float2(h);
                    const __half2 lo2 = __floats2half2_rn(p0 - f.x, p1 - f.y);
                    tmem_st1(dst_hi, *reinterpret_cast<const uint32_t*>(&h));
                    tmem_st1(dst_lo, *reinterpret_cast<const uint32_t*>(&lo2));
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
        mbar_init(&bars->mma_done, 1);
        mbar_init(&bars->opnd, 16);
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
               "cond_size, and widest layer + padded cond width + 2 x padded input width <= 512 tensor-memory columns); use FC_PREC_FP32";
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
