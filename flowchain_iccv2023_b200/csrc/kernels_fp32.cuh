// kernels_fp32.cuh -- fp32 CUDA-core path of the fused CIF-step kernels (parity path, 1e-4 nats).
//
// One CTA owns a tile of R rows (row = one query point of one agent) and runs a whole CIF step
// (reference CIF_step.forward/.inverse, fastpredNF.py:769-781) without touching HBM in between:
//   r/q density MLPs -> Gaussian sample / log-prob -> n_blocks x (masked affine coupling + BatchNorm).
// Activations live in shared memory feature-major ([feature][row]) so the register-tiled layer GEMM
// reads rows as float4 and weights (W^T, [K][NP], L2/L1-resident) as float4 broadcasts.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "plan.h"

namespace fc {

constexpr int kThreads = 256;
constexpr float kLog2Pi = 1.8378770664093453f;       // log(2 pi)
constexpr float kHalfLog2Pi = 0.9189385332046727f;   // log(sqrt(2 pi))

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

struct Dims {       // shared-memory carve-up (in floats), filled on the host
    int in_rows;    // 2 + C_max
    int u_rows;     // C_max + 2
    int h_rows;     // max(2*HP_max, 2*HH, 2*CP_max)
    int T, C0;
};

struct Prob {       // one dense layer: out[NP][R] = act(W^T[K][NP]^T * in[K][R] + b)
    const float* in;
    const float* W;
    const float* b;
    float* out;
    int K, NP, act;
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: production-mode noise (the reference draws torch.randn_like; parity
// runs inject a tape instead).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// Box-Muller on fast intrinsics: the noise only has to be N(0,1) to ~1e-6 (it is a Monte-Carlo draw, and both
// precision paths must see the SAME values, which they do because this one function serves all kernels).
// The angle is kept in [-pi, pi) where __sincosf is accurate to ~2^-21.
__device__ __noinline__ float4 philox_normal4(unsigned long long seed, long long row, unsigned key, unsigned blk) {
    const uint4 r = philox4x32_10(make_uint4((unsigned)row, (unsigned)((unsigned long long)row >> 32), key, blk),
                                  make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
    const float u0 = ((r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((r.z >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float a0 = ((float)(int)(r.y >> 8) - 8388607.5f) * (3.14159265358979f / 8388608.0f);
    const float a1 = ((float)(int)(r.w >> 8) - 8388607.5f) * (3.14159265358979f / 8388608.0f);
    const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
    float s0, c0, s1, c1;
    __sincosf(a0, &s0, &c0);
    __sincosf(a1, &s1, &c1);
    return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

// 8 normals for counter blocks (blk, blk + 1): the two Philox streams and Box-Muller pairs are independent, so the
// compiler interleaves them (the single-stream version is latency-bound: 10 serial rounds).  Same values as two calls
// of philox_normal4.
struct Normal8 { float4 a, b; };
__device__ __noinline__ Normal8 philox_normal8(unsigned long long seed, long long row, unsigned key, unsigned blk) {
    float out[8];
    uint4 c0 = make_uint4((unsigned)row, (unsigned)((unsigned long long)row >> 32), key, blk), c1 = c0;
    c1.w = blk + 1u;
    uint2 k0 = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const unsigned h00 = __umulhi(0xD2511F53u, c0.x), l00 = 0xD2511F53u * c0.x, h01 = __umulhi(0xCD9E8D57u, c0.z), l01 = 0xCD9E8D57u * c0.z;
        const unsigned h10 = __umulhi(0xD2511F53u, c1.x), l10 = 0xD2511F53u * c1.x, h11 = __umulhi(0xCD9E8D57u, c1.z), l11 = 0xCD9E8D57u * c1.z;
        c0 = make_uint4(h01 ^ c0.y ^ k0.x, l01, h00 ^ c0.w ^ k0.y, l00);
        c1 = make_uint4(h11 ^ c1.y ^ k0.x, l11, h10 ^ c1.w ^ k0.y, l10);
        k0.x += 0x9E3779B9u;
        k0.y += 0xBB67AE85u;
    }
    const unsigned rr[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float u = ((rr[2 * p] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float ang = ((float)(int)(rr[2 * p + 1] >> 8) - 8388607.5f) * (3.14159265358979f / 8388608.0f);
        const float rad = sqrtf(-2.0f * __logf(u));
        float s, c;
        __sincosf(ang, &s, &c);
        out[2 * p] = rad * c;
        out[2 * p + 1] = rad * s;
    }
    Normal8 r;
    r.a = make_float4(out[0], out[1], out[2], out[3]);
    r.b = make_float4(out[4], out[5], out[6], out[7]);
    return r;
}

// packed fp32 pairs (64-bit register pair) for FFMA2
__device__ __forceinline__ uint64_t pk2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

// ---------------------------------------------------------------------------------------------
// register-tiled layer: thread tile 4 rows x 8 outputs; two independent problems per call
// (s_net | t_net, or shift_net | log_scale_net) so all 256 threads have work.
// ---------------------------------------------------------------------------------------------
template <int R>
__device__ __forceinline__ void layer2(const Prob& p0, const Prob& p1) {
    constexpr int RG = R / 4;
    const int nt0 = RG * (p0.NP >> 3);
    const int nt = nt0 + RG * (p1.NP >> 3);
    for (int t = threadIdx.x; t < nt; t += kThreads) {
        const bool second = t >= nt0;
        const Prob& p = second ? p1 : p0;
        const int tt = second ? t - nt0 : t;
        const int rg = tt % RG, cg = tt / RG;
        const float* __restrict__ a = p.in + rg * 4;
        const float* __restrict__ w = p.W + cg * 8;
        // accumulators as packed fp32 pairs along the output dimension: Blackwell's FFMA2 does two fmas per instruction
        // (same rounding as fmaf, so results are bit-identical), which takes a third of the instructions out of this
        // issue-bound loop
        uint64_t acc2[4][4];
        {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.b + cg * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.b + cg * 8 + 4));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc2[i][0] = pk2(b0.x, b0.y); acc2[i][1] = pk2(b0.z, b0.w);
                acc2[i][2] = pk2(b1.x, b1.y); acc2[i][3] = pk2(b1.z, b1.w);
            }
        }
        const int K = p.K, NP = p.NP;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(a + k * R);
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + k * NP));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + k * NP + 4));
            const float ar[4] = {av.x, av.y, av.z, av.w};
            const uint64_t wp[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t aa = pk2(ar[i], ar[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc2[i][j] = fma2(aa, wp[j], acc2[i][j]);
            }
        }
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) upk2(acc2[i][j], acc[i][2 * j], acc[i][2 * j + 1]);
        const int act = p.act;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float v = acc[i][j];
                if (act == ACT_RELU) v = fmaxf(v, 0.0f);
                else if (act == ACT_TANH) v = tanhf(v);
                o[i] = v;
            }
            *reinterpret_cast<float4*>(p.out + (cg * 8 + j) * R + rg * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

template <int R>
struct Tile {
    float *IN, *U, *HA, *HB, *Z, *ACC, *PART, *PREV, *SEQ;
    long long* ROWN;         // clamped global row index per tile row
    long long *ROWA, *ROWP;  // agent / point index per tile row
    const float* wbuf;
    StepPlan* sp;            // current step plan (shared memory copy)
};

template <int R>
__device__ __forceinline__ Tile<R> carve(float* smem, const Dims& dm, const float* wbuf) {
    Tile<R> t;
    float* p = smem;
    t.sp = reinterpret_cast<StepPlan*>(p);
    p += (sizeof(StepPlan) + 15) / 16 * 4;
    t.ROWN = reinterpret_cast<long long*>(p); p += 2 * R;
    t.ROWA = reinterpret_cast<long long*>(p); p += 2 * R;
    t.ROWP = reinterpret_cast<long long*>(p); p += 2 * R;
    t.IN = p;   p += dm.in_rows * R;
    t.U = p;    p += dm.u_rows * R;
    t.HA = p;   p += dm.h_rows * R;
    t.HB = p;   p += dm.h_rows * R;
    t.Z = p;    p += 2 * R;
    t.ACC = p;  p += R;
    t.PREV = p; p += 2 * R;
    t.PART = p; p += 2 * (kThreads / R) * R;
    t.SEQ = p;  // 2*T*R floats, sampler only
    t.wbuf = wbuf;
    return t;
}

inline size_t smem_bytes_fp32(const Dims& dm, int R, bool sampler) {
    size_t fl = (sizeof(StepPlan) + 15) / 16 * 4 + 6 * R + (size_t)(dm.in_rows + dm.u_rows + 2 * dm.h_rows) * R +
                5 * R + 2 * kThreads + (sampler ? 2 * (size_t)dm.T * R : 0);
    return fl * sizeof(float);
}

__device__ __forceinline__ void load_plan(StepPlan* dst, const StepPlan* src) {
    const int n = sizeof(StepPlan) / 4;
    const int* s = reinterpret_cast<const int*>(src);
    int* d = reinterpret_cast<int*>(dst);
    for (int i = threadIdx.x; i < n; i += kThreads) d[i] = __ldg(s + i);
}

template <int R>
__device__ __forceinline__ void init_rows(const Tile<R>& t, const RowArgs& a, long long row0) {
    if (threadIdx.x < R) {
        long long n = row0 + threadIdx.x;
        if (n > a.N - 1) n = a.N - 1;
        t.ROWN[threadIdx.x] = n;
        t.ROWA[threadIdx.x] = n / a.PPA;
        t.ROWP[threadIdx.x] = (n % a.PPA) / a.K;
    }
}

__device__ __forceinline__ const float* src_at(const Src& s, long long ag, long long pt, int step) {
    return s.p + ag * s.sa + pt * s.sp + (long long)step * s.st;
}

// IN rows [2, 2+C0) <- cond[.., cstep, :];  IN rows [2+C0, 2+C0+2k) <- prefix[.., :k, :] (or SEQ in smem)
template <int R>
__device__ __forceinline__ void load_y(const Tile<R>& t, const RowArgs& a, const Dims& dm, int cstep, int k,
                                       bool prefix_from_seq) {
    const int C0 = dm.C0;
    for (int idx = threadIdx.x; idx < R * C0; idx += kThreads) {
        const int r = idx / C0, d = idx - r * C0;
        t.IN[(2 + d) * R + r] = __ldg(src_at(a.cond, t.ROWA[r], t.ROWP[r], cstep) + d);
    }
    const int k2 = 2 * k;
    if (prefix_from_seq) {
        for (int idx = threadIdx.x; idx < R * k2; idx += kThreads) {
            const int e = idx / R, r = idx - e * R;
            t.IN[(2 + C0 + e) * R + r] = t.SEQ[e * R + r];
        }
    } else {
        for (int idx = threadIdx.x; idx < R * k2; idx += kThreads) {
            const int r = idx / k2, e = idx - r * k2;
            t.IN[(2 + C0 + e) * R + r] = __ldg(src_at(a.prefix, t.ROWA[r], t.ROWP[r], e >> 1) + (e & 1));
        }
    }
}

// r/q density MLPs on IN -> means in HA rows [0,CP), log-stddevs in HA rows [CP, 2CP)
// (reference DiagonalGaussianConditionalDensity._means_and_log_stddevs, fastpredNF.py:709-710)
template <int R>
__device__ __forceinline__ void density_nets(const Tile<R>& t, int which) {
    const StepPlan& sp = *t.sp;
    const LayerRef(&L)[2][3] = sp.dens[which];
    {
        Prob a{t.IN, t.wbuf + L[0][0].w, t.wbuf + L[0][0].b, t.HA, L[0][0].K, L[0][0].NP, ACT_RELU};
        Prob b{t.IN, t.wbuf + L[1][0].w, t.wbuf + L[1][0].b, t.HA + sp.HP * R, L[1][0].K, L[1][0].NP, ACT_RELU};
        layer2<R>(a, b);
    }
    __syncthreads();
    {
        Prob a{t.HA, t.wbuf + L[0][1].w, t.wbuf + L[0][1].b, t.HB, L[0][1].K, L[0][1].NP, ACT_RELU};
        Prob b{t.HA + sp.HP * R, t.wbuf + L[1][1].w, t.wbuf + L[1][1].b, t.HB + sp.HP * R, L[1][1].K, L[1][1].NP, ACT_RELU};
        layer2<R>(a, b);
    }
    __syncthreads();
    {
        Prob a{t.HB, t.wbuf + L[0][2].w, t.wbuf + L[0][2].b, t.HA, L[0][2].K, L[0][2].NP, ACT_NONE};
        Prob b{t.HB + sp.HP * R, t.wbuf + L[1][2].w, t.wbuf + L[1][2].b, t.HA + sp.CP * R, L[1][2].K, L[1][2].NP, ACT_NONE};
        layer2<R>(a, b);
    }
    __syncthreads();
}

// diagonal_gaussian_sample / diagonal_gaussian_log_prob (fastpredNF.py:713-733) on the tile:
// SAMPLE: u = exp(ls)*eps + mu -> U rows [0,C); both: ACC += sign * log N(u; mu, exp(ls)).
template <int R, bool SAMPLE>
__device__ __forceinline__ void gauss_phase(const Tile<R>& t, const RowArgs& a, float sign, long long eps_col,
                                            unsigned key) {
    constexpr int NSL = kThreads / R;
    const StepPlan& sp = *t.sp;
    const int C = sp.C, CP = sp.CP;
    const int r = threadIdx.x % R, sl = threadIdx.x / R;
    float s_ls = 0.0f, s_q = 0.0f;
    for (int d0 = sl * 4; d0 < C; d0 += 4 * NSL) {
        float e4[4] = {0.f, 0.f, 0.f, 0.f};
        if (SAMPLE) {
            if (a.eps != nullptr) {
                const float* ep = a.eps + t.ROWN[r] * a.eps_rs + eps_col + d0;
#pragma unroll
                for (int dd = 0; dd < 4; ++dd)
                    if (d0 + dd < C) e4[dd] = __ldg(ep + dd);
            } else {
                const float4 g = philox_normal4(a.seed, t.ROWN[r], key, (unsigned)(d0 >> 2));
                e4[0] = g.x; e4[1] = g.y; e4[2] = g.z; e4[3] = g.w;
            }
        }
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
            const int d = d0 + dd;
            if (d < C) {
                const float mu = t.HA[d * R + r];
                const float ls = t.HA[(CP + d) * R + r];
                const float sd = expf(ls);
                float u;
                if (SAMPLE) {
                    u = __fadd_rn(__fmul_rn(sd, e4[dd]), mu);
                    t.U[d * R + r] = u;
                } else {
                    u = t.U[d * R + r];
                }
                const float df = u - mu;
                s_q += (df * df) / (sd * sd);
                s_ls += ls;
            }
        }
    }
    t.PART[sl * R + r] = s_ls;
    t.PART[(NSL + sl) * R + r] = s_q;
    __syncthreads();
    if (threadIdx.x < R) {
        float sls = 0.0f, sq = 0.0f;
#pragma unroll
        for (int s = 0; s < NSL; ++s) {
            sls += t.PART[s * R + r];
            sq += t.PART[(NSL + s) * R + r];
        }
        const float lp = (-0.5f * (float)C * kLog2Pi - sls) - 0.5f * sq;
        t.ACC[r] += sign * lp;
    }
    __syncthreads();
}

// s_net / t_net of coupling block b on U = [u (C), mx (2)]; returns after the last hidden layer with the
// activations in *hs (s half rows [0,HH), t half rows [HH,2HH)).
template <int R>
__device__ __forceinline__ const float* coupling_nets(const Tile<R>& t, int b) {
    const StepPlan& sp = *t.sp;
    const int HH = sp.HH;
    float* cur = t.HA;
    float* nxt = t.HB;
    {
        const LayerRef& ls = sp.cpl[b][0][0];
        const LayerRef& lt = sp.cpl[b][1][0];
        Prob a{t.U, t.wbuf + ls.w, t.wbuf + ls.b, cur, ls.K, ls.NP, ACT_TANH};
        Prob c{t.U, t.wbuf + lt.w, t.wbuf + lt.b, cur + HH * R, lt.K, lt.NP, ACT_RELU};
        layer2<R>(a, c);
    }
    __syncthreads();
    for (int l = 1; l <= sp.n_hidden; ++l) {
        const LayerRef& ls = sp.cpl[b][0][l];
        const LayerRef& lt = sp.cpl[b][1][l];
        Prob a{cur, t.wbuf + ls.w, t.wbuf + ls.b, nxt, ls.K, ls.NP, ACT_TANH};
        Prob c{cur + HH * R, t.wbuf + lt.w, t.wbuf + lt.b, nxt + HH * R, lt.K, lt.NP, ACT_RELU};
        layer2<R>(a, c);
        __syncthreads();
        float* tmp = cur; cur = nxt; nxt = tmp;
    }
    return cur;
}

// One full CIF step on the tile.  In: Z (point), IN rows >= 2 (conditioning y).  Out: Z, ACC += ldj + log q - log r.
template <int R, bool INVERSE>
__device__ __forceinline__ void cif_step(const Tile<R>& t, const RowArgs& a, long long eps_col, unsigned key) {
    const StepPlan& sp = *t.sp;
    const int r = threadIdx.x;
    if (r < R) { t.IN[r] = t.Z[r]; t.IN[R + r] = t.Z[R + r]; }
    __syncthreads();
    // forward samples u ~ r(u|z,y) (enters with -log r); inverse samples u ~ q(u|w,y) (+log q)
    density_nets<R>(t, INVERSE ? 1 : 0);
    gauss_phase<R, true>(t, a, INVERSE ? 1.0f : -1.0f, eps_col, key);

    const int C = sp.C, HH = sp.HH;
    for (int bi = 0; bi < sp.n_blocks; ++bi) {
        const int b = INVERSE ? sp.n_blocks - 1 - bi : bi;
        const int m = sp.masked_dim[b], j = 1 - m;
        if (r < R) {
            float z0 = t.Z[r], z1 = t.Z[R + r];
            if (INVERSE) {   // BatchNorm.inverse first (modules reversed, TFCondARFlow.py:264-269)
                z0 = fmaf(sp.bn_iscale[b][0], z0, sp.bn_ishift[b][0]);
                z1 = fmaf(sp.bn_iscale[b][1], z1, sp.bn_ishift[b][1]);
                t.Z[r] = z0; t.Z[R + r] = z1;
                t.ACC[r] -= sp.bn_ldj[b];
            }
            t.U[(C + m) * R + r] = m == 0 ? z0 : z1;   // mx = x * mask
            t.U[(C + j) * R + r] = 0.0f;
        }
        __syncthreads();
        const float* hs = coupling_nets<R>(t, b);
        if (r < R) {
            const float* ws = t.wbuf + sp.cpl_last_w[b][0];
            const float* wt = t.wbuf + sp.cpl_last_w[b][1];
            float s = sp.cpl_last_b[b][0], tt = sp.cpl_last_b[b][1];
            for (int k = 0; k < HH; ++k) {
                s = fmaf(hs[k * R + r], __ldg(ws + k), s);
                tt = fmaf(hs[(HH + k) * R + r], __ldg(wt + k), tt);
            }
            const float log_s = tanhf(s);
            float zz[2] = {t.Z[r], t.Z[R + r]};
            if (!INVERSE) {
                zz[j] = fmaf(zz[j], expf(log_s), tt);              // u = x*exp(log_s) + t
                zz[0] = fmaf(sp.bn_scale[b][0], zz[0], sp.bn_shift[b][0]);
                zz[1] = fmaf(sp.bn_scale[b][1], zz[1], sp.bn_shift[b][1]);
                t.ACC[r] += log_s + sp.bn_ldj[b];
            } else {
                zz[j] = (zz[j] - tt) * expf(-log_s);               // x = (u - t)*exp(-log_s)
                t.ACC[r] -= log_s;
            }
            t.Z[r] = zz[0]; t.Z[R + r] = zz[1];
        }
        __syncthreads();
    }
    if (r < R) { t.IN[r] = t.Z[r]; t.IN[R + r] = t.Z[R + r]; }
    __syncthreads();
    density_nets<R>(t, INVERSE ? 0 : 1);
    gauss_phase<R, false>(t, a, INVERSE ? -1.0f : 1.0f, 0, 0);
}

// sum_d Normal(loc_d, sd_d).log_prob(v_d)   (torch.distributions.Normal.log_prob)
__device__ __forceinline__ float normal_lp2(float v0, float v1, float l0, float l1, float s0, float s1) {
    const float d0 = v0 - l0, d1 = v1 - l1;
    const float a = -(d0 * d0) / (2.0f * (s0 * s0)) - logf(s0) - kHalfLog2Pi;
    const float b = -(d1 * d1) / (2.0f * (s1 * s1)) - logf(s1) - kHalfLog2Pi;
    return a + b;
}

// ---------------------------------------------------------------------------------------------
// a8: per-step ("separate") density.  grid = (row tiles, T); heavy (late) steps are scheduled first.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 2)
k_separate_fp32(const StepPlan* __restrict__ plans, const float* __restrict__ wbuf, RowArgs a, Dims dm) {
    extern __shared__ __align__(16) float smem[];
    Tile<R> t = carve<R>(smem, dm, wbuf);
    const int step = dm.T - 1 - blockIdx.y;
    const long long row0 = (long long)blockIdx.x * R;
    load_plan(t.sp, plans + step);
    init_rows<R>(t, a, row0);
    __syncthreads();
    const int r = threadIdx.x;
    if (r < R) {
        float q0, q1;
        load_query(a, t.ROWA[r], t.ROWP[r], step, q0, q1);
        t.Z[r] = q0; t.Z[R + r] = q1;
        const float* pp = step == 0 ? src_at(a.base, t.ROWA[r], t.ROWP[r], 0)
                                    : src_at(a.prefix, t.ROWA[r], t.ROWP[r], step - 1);
        t.PREV[r] = __ldg(pp); t.PREV[R + r] = __ldg(pp + 1);
        t.ACC[r] = 0.0f;
    }
    load_y<R>(t, a, dm, step, step, false);
    __syncthreads();
    cif_step<R, false>(t, a, t.sp->eps_off, (unsigned)step);
    if (r < R) {          // whole warps: store_result reduces K importance samples across lanes
        const float lp = normal_lp2(t.Z[r], t.Z[R + r], t.PREV[r], t.PREV[R + r], t.sp->std[0], t.sp->std[1]) + t.ACC[r];
        const long long n = row0 + r;
        store_result(a, t.ROWA[r], n % a.PPA, step, lp, n < a.N);
    }
}

// ---------------------------------------------------------------------------------------------
// a9: full-chain ("sequential") density.  grid = (row tiles, result indices); CTA (tile, j) applies
// net[j], net[j-1], ..., net[klo] with the point kept on-chip (fastpredNF.py:548-569).
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 2)
k_sequential_fp32(const StepPlan* __restrict__ plans, const float* __restrict__ wbuf, RowArgs a, Dims dm) {
    extern __shared__ __align__(16) float smem[];
    Tile<R> t = carve<R>(smem, dm, wbuf);
    const int j = dm.T - 1 - blockIdx.y;
    const long long row0 = (long long)blockIdx.x * R;
    init_rows<R>(t, a, row0);
    __syncthreads();
    const int r = threadIdx.x;
    if (r < R) {
        const float* xp = src_at(a.x, t.ROWA[r], t.ROWP[r], j);
        t.Z[r] = __ldg(xp); t.Z[R + r] = __ldg(xp + 1);
        const float* pp = src_at(a.base, t.ROWA[r], t.ROWP[r], 0);
        t.PREV[r] = __ldg(pp); t.PREV[R + r] = __ldg(pp + 1);
        t.ACC[r] = 0.0f;
    }
    for (int k = j; k >= a.seq_klo; --k) {
        __syncthreads();
        load_plan(t.sp, plans + k);
        load_y<R>(t, a, dm, j, k, false);
        __syncthreads();
        const long long col = (long long)t.sp->seq_eps_off + (long long)(j - k) * t.sp->C;
        cif_step<R, false>(t, a, col, (unsigned)(k * 64 + j + 4096));
    }
    __syncthreads();
    if (r < R) {
        const float s0 = __ldg(&plans[0].std[0]), s1 = __ldg(&plans[0].std[1]);
        const float lp = normal_lp2(t.Z[r], t.Z[R + r], t.PREV[r], t.PREV[R + r], s0, s1) +
                         t.ACC[r] / (float)(j - a.seq_jlo + 1);
        const long long n = row0 + r;
        store_result(a, t.ROWA[r], n % a.PPA, j - a.seq_jlo, lp, n < a.N);
    }
}

// ---------------------------------------------------------------------------------------------
// a10: autoregressive sampler with log-probs (fastpredNF.py:472-477, 586-607).  grid = row tiles.
// ---------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kThreads, 2)
k_sample_fp32(const StepPlan* __restrict__ plans, const float* __restrict__ wbuf, RowArgs a, Dims dm, SampleOut o) {
    extern __shared__ __align__(16) float smem[];
    Tile<R> t = carve<R>(smem, dm, wbuf);
    const long long row0 = (long long)blockIdx.x * R;
    init_rows<R>(t, a, row0);
    __syncthreads();
    const int r = threadIdx.x;
    const bool valid = r < R && row0 + r < a.N;
    float base_lp = 0.0f, cum = 0.0f;
    if (r < R) {
        const float* pp = src_at(a.base, t.ROWA[r], t.ROWP[r], 0);
        const float b0 = __ldg(pp), b1 = __ldg(pp + 1);
        const float s0 = __ldg(&plans[0].std[0]), s1 = __ldg(&plans[0].std[1]);
        float e0, e1;
        if (a.z0 != nullptr) {
            e0 = __ldg(a.z0 + t.ROWN[r] * 2); e1 = __ldg(a.z0 + t.ROWN[r] * 2 + 1);
        } else {
            const float4 g = philox_normal4(a.seed, t.ROWN[r], 0xFFFFu, 0u);
            e0 = g.x; e1 = g.y;
        }
        const float u0 = __fadd_rn(__fmul_rn(e0, s0), b0), u1 = __fadd_rn(__fmul_rn(e1, s1), b1);
        t.Z[r] = u0; t.Z[R + r] = u1;
        base_lp = normal_lp2(u0, u1, b0, b1, s0, s1);
        if (valid) o.base_lp[row0 + r] = base_lp;
    }
    const int T = dm.T;
    for (int k = 0; k < T; ++k) {
        __syncthreads();
        load_plan(t.sp, plans + k);
        if (r < R) t.ACC[r] = 0.0f;
        load_y<R>(t, a, dm, k, k, true);
        __syncthreads();
        cif_step<R, true>(t, a, t.sp->eps_off, (unsigned)k);
        if (r < R) {
            const float z0 = t.Z[r], z1 = t.Z[R + r];
            t.SEQ[(2 * k) * R + r] = z0;
            t.SEQ[(2 * k + 1) * R + r] = z1;
            const float l = t.ACC[r];
            cum += l;
            if (valid) {
                const long long n = row0 + r;
                o.seq[(n * T + k) * 2] = z0;
                o.seq[(n * T + k) * 2 + 1] = z1;
                o.ldjs[n * T + k] = l;
                o.log_prob[n * T + k] = base_lp + cum / (float)(k + 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K importance samples: out(a,g,t) = logsumexp_k tmp[(a,g,k), t] - log K   (extension)
// ---------------------------------------------------------------------------------------------
__global__ void k_logsumexp_k(const float* __restrict__ tmp, int K, int T, long long AG, long long G,
                              float* __restrict__ out, long long sa, long long sp, long long st) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over AG*T
    if (idx >= AG * T) return;
    const long long ag = idx / T;
    const int tt = (int)(idx - ag * T);
    const float* p = tmp + (ag * K) * T + tt;
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, p[(long long)k * T]);
    float s = 0.0f;
    for (int k = 0; k < K; ++k) s += expf(p[(long long)k * T] - m);
    const float v = (m == -INFINITY) ? -INFINITY : m + logf(s) - logf((float)K);
    out[(ag / G) * sa + (ag % G) * sp + (long long)tt * st] = v;
}

// ---------------------------------------------------------------------------------------------
// a11/a12: rapid update (fastpredNF.py:114, 142-164), one CTA per agent.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float s = 0.0f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) s += red[i];
    return s;
}

__global__ void k_base_normalizer(const float* __restrict__ base_lp, float* __restrict__ norm, long long S) {
    __shared__ float red[32];
    const float* p = base_lp + (long long)blockIdx.x * S;
    float acc = 0.0f;
    for (long long s = threadIdx.x; s < S; s += blockDim.x) acc += expf(p[s]);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) norm[blockIdx.x] = acc;
}

// Rapid update: ONE launch, one thread-block CLUSTER of kUpdCluster CTAs per agent (so that one agent -- the reference's
// only working case, B = 1 -- already spreads over 8 SMs and 1000 agents fill the chip 8000 CTAs deep):
//   phase 1: every CTA evaluates bp[s] = exp(base log-prob of its samples at the newest observed position) ONCE, keeps
//            it in shared memory and block-reduces its partial sum;
//   exchange: the partial sums are read from the peers' shared memory (DSMEM) in rank order -> deterministic total;
//   phase 2: lp'[s] = log(bp / total * normalizer) in place, then either (full) a flat, coalesced copy of
//            prob_st_0[:, :, t:] with the log-prob column replaced by lp'[s] + ldjs[s, t+j] -- 8 independent loads in
//            flight per thread, no partial-sum round trip through HBM, no second pass over the positions -- or (lazy)
//            just lp'(A, S): 40 KB per agent instead of 1.3 MB, for consumers that add ldjs themselves.
constexpr int kUpdCluster = 8, kUpdClusterBig = 16;
constexpr int kUpdThreads = 256, kUpdMaxThreads = 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t rank) {
    uint32_t la = (uint32_t)__cvta_generic_to_shared(local_ptr), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

// chunk = samples per CTA (multiple of 4); dynamic shared memory: chunk floats.  per_magic = ceil(2^40 / per).
template <bool LAZY>
__global__ void __launch_bounds__(kUpdMaxThreads)
k_update(const float* __restrict__ new_pos, int t, int T, const float* __restrict__ prob0, const float* __restrict__ ldjs,
         const float* __restrict__ normalizer, float* __restrict__ out, long long S, int chunk, float sd0, float sd1,
         unsigned long long per_magic) {
    extern __shared__ __align__(16) float upd_bp[];          // bp, then lp', of this CTA's samples
    __shared__ float red[32];
    __shared__ float my_partial;
    const long long ag = blockIdx.y;
    const int nthr = (int)blockDim.x;          // 256 when many agents fill the chip, 1024 when a few agents must not crawl
    const uint32_t rank = cluster_ctarank();
    const long long s0 = (long long)rank * chunk;
    const int ns = (int)(S - s0 < (long long)chunk ? (S - s0 > 0 ? S - s0 : 0) : chunk);
    const float np0 = __ldg(new_pos + ag * 2), np1 = __ldg(new_pos + ag * 2 + 1);
    float acc = 0.0f;
    // strided gather (8 of every 12 T bytes): 8 independent sample loads in flight per thread before any arithmetic
    for (int sl0 = threadIdx.x; sl0 < ns; sl0 += 8 * nthr) {
        float px[8], py[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int sl = sl0 + k * nthr;
            const float* q = prob0 + ((ag * S + s0 + (sl < ns ? sl : ns - 1)) * T + (t - 1)) * 3;
            px[k] = __ldg(q); py[k] = __ldg(q + 1);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int sl = sl0 + k * nthr;
            if (sl < ns) {
                const float v = expf(normal_lp2(px[k], py[k], np0, np1, sd0, sd1));
                upd_bp[sl] = v;
                acc += v;
            }
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) my_partial = acc;
    cluster_sync_all();
    float total = 0.0f;
    const uint32_t csize = cluster_nctarank();                // 8, or 16 (non-portable size) when only a few agents are in flight
    for (uint32_t r = 0; r < csize; ++r) total += ld_dsmem_f32(&my_partial, r);     // rank order: deterministic
    cluster_sync_all();                                       // no CTA may exit (or reuse my_partial) while peers still read it
    const float nz = __ldg(normalizer + ag);
    for (int sl = threadIdx.x; sl < ns; sl += nthr) {
        const float lp = logf(upd_bp[sl] / total * nz);       // same order of operations as fastpredNF.py:154-155
        if (LAZY) out[ag * S + s0 + sl] = lp;
        else upd_bp[sl] = lp;
    }
    if (LAZY) return;
    __syncthreads();
    const int Tt = T - t, per = Tt * 3;
    const float* src = prob0 + ((ag * S + s0) * T + t) * 3;   // sample sl's tail row block starts at src + sl * 3T
    const float* lj = ldjs + (ag * S + s0) * T + t;
    float* dst = out + (ag * S + s0) * per;
    const int nelem = ns * per, T3 = 3 * T;
    constexpr int U = 8;
    for (int e0 = threadIdx.x; e0 < nelem; e0 += U * nthr) {
        float v[U];
        int sls[U];
        bool isl[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int e = e0 + k * nthr;
            const int ec = e < nelem ? e : nelem - 1;
            const int sl = (int)(((unsigned long long)(unsigned)ec * per_magic) >> 40);     // ec / per
            const int r = ec - sl * per;
            const int j = (r * 43) >> 7, c = r - j * 3;                                    // r / 3 for r < 128
            sls[k] = sl;
            isl[k] = c == 2;
            const float* p = c == 2 ? lj + sl * T + j : src + sl * T3 + r;
            v[k] = __ldg(p);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int e = e0 + k * nthr;
            if (e < nelem) dst[e] = isl[k] ? upd_bp[sls[k]] + v[k] : v[k];
        }
    }
}

}  // namespace fc
