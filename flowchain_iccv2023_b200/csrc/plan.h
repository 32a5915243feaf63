// plan.h -- host/device shared descriptors of the packed FlowChain weights and of the row sources.
#pragma once
#include <stdint.h>

namespace fc {

constexpr int kMaxBlocks = 8;    // coupling+BN pairs per CIF step (reference default 3, stress 6)
constexpr int kMaxHidden = 4;    // hidden HxH layers per conditioner net (reference default 2)
constexpr int kMaxSteps = 32;    // T

// One dense layer of the fp32 path: W^T stored [K][NP] (NP = outputs padded to 8), bias [NP].
struct LayerRef {
    int32_t w;      // float offset into the packed fp32 buffer
    int32_t b;
    int32_t K;      // rows of W^T the kernel iterates (padded input width)
    int32_t NP;     // padded outputs
};

// Packed description of one CIF step (reference CIF_step, fastpredNF.py:756-781).
struct StepPlan {
    int32_t C;        // conditioning width C_i = C0 + 2 i (dim of the auxiliary variable u)
    int32_t c;        // C + 2 = input width of the r/q density nets
    int32_t HP;       // density hidden width 2c padded to 8
    int32_t CP;       // C padded to 8
    int32_t HH;       // coupling hidden width H padded to 8
    int32_t eps_off;  // column of this step in the separate noise tape
    int32_t seq_eps_off;  // offset of block k in the sequential noise tape
    int32_t n_blocks, n_hidden;
    // density nets: [which: 0 = r_u_given_z, 1 = q_u_given_w][net: 0 = shift, 1 = log_scale][layer 0..2]
    LayerRef dens[2][2][3];
    // coupling nets: [block][net: 0 = s (tanh), 1 = t (relu)][layer 0 .. n_hidden]; the last Linear (H -> 2)
    // is reduced to the single surviving output column (1 - mask) and stored as a vector.
    LayerRef cpl[kMaxBlocks][2][kMaxHidden + 1];
    int32_t cpl_last_w[kMaxBlocks][2];   // [HH] vector
    float cpl_last_b[kMaxBlocks][2];
    int32_t masked_dim[kMaxBlocks];      // index m with mask[m] == 1 (passes through); j = 1 - m is transformed
    float bn_scale[kMaxBlocks][2], bn_shift[kMaxBlocks][2];      // forward:  y = scale * x + shift
    float bn_iscale[kMaxBlocks][2], bn_ishift[kMaxBlocks][2];    // inverse:  x = iscale * y + ishift
    float bn_ldj[kMaxBlocks];                                    // sum_d (log_gamma - 0.5 log(var + eps))
    float std[2];                                                // clamp(var[step], 1e-4)
    // ---- bf16 tensor-core path (byte offsets into the packed bf16 image) ----
    int32_t tc_off;       // start of this step's image
    int32_t tc_bytes;     // size (multiple of 16)
};

// A strided view: element (agent, point, step, d) at p[agent*sa + point*sp + step*st + d].
struct Src {
    const float* p;
    long long sa, sp, st;
};

struct RowArgs {
    Src x;         // query / trajectory points (.., 2)
    Src prefix;    // conditioning-prefix trajectory (.., 2)
    Src cond;      // encoder conditioning (.., C0)
    Src base;      // base position (.., 2); st unused
    const float* eps;      // noise tape or nullptr
    long long eps_rs;      // row stride of eps
    const float* z0;       // sampler: (N, 2) standard normals or nullptr
    unsigned long long seed;
    long long N;           // rows
    long long PPA;         // rows per agent (G*K); agent = n / PPA, point = (n % PPA) / K
    int K;
    int prefix_is_query;   // 1: prefix/prev come from the row's own trajectory x (reference semantics)
    float* out;            // element (agent, n % PPA, t) at out[agent*out_sa + (n%PPA)*out_sp + t*out_st]
    long long out_sa, out_sp, out_st;
    int seq_klo;           // sequential: lowest chain step applied (0, or T - n_step)
    int seq_jlo;           // sequential: first result index written (0, or T - n_step - 1)
    int fuse_k;            // 1: K = 2^j <= 32 importance samples of a point sit in consecutive rows (= lanes): logsumexp_k - log K
                           // is reduced with warp shuffles in registers and ONE value per (point, step) is stored
    // position-space density maps (per-step density with a shared prefix only): the query of row
    // (agent, point) at step t is z = (x[point] - qshift[agent, t]) * qscale[agent] -- the caller's METRIC lattice mapped
    // into the standardised state space of that agent and step -- and every result of the agent gets out_shift[agent]
    // (the log-Jacobian of that change of variables) added and, with out_exp, is stored as exp(.) (a probability density)
    const float* qshift;   // (A, T, 2) or nullptr
    long long qshift_sa, qshift_st;
    const float* qscale;   // (A, 2) or nullptr
    const float* out_shift;   // (A) or nullptr
    int out_exp;
    int out_sys;           // 1: `out` is an NVSwitch multicast address of a symmetric buffer: every store is replicated to
                           // all peers, so the density maps are all-gathered by the stores themselves (multimem.st, system scope)
};

#ifdef __CUDACC__
// the one place a log-density leaves the chip
__device__ __forceinline__ void store_out(const RowArgs& a, long long idx, float v) {
    // multicast (multimem) addresses may only be accessed with multimem.* instructions (PTX ISA, "multimem.st")
    if (a.out_sys) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a.out + idx), "f"(v) : "memory");
    else a.out[idx] = v;
}
// query point of row (agent, point) at `step` (see RowArgs::qscale)
__device__ __forceinline__ void load_query(const RowArgs& a, long long ag, long long pt, int step, float& z0, float& z1) {
    const float* xp = a.x.p + ag * a.x.sa + pt * a.x.sp + (long long)step * a.x.st;
    z0 = __ldg(xp); z1 = __ldg(xp + 1);
    if (a.qscale != nullptr) {
        const float* qs = a.qshift + ag * a.qshift_sa + (long long)step * a.qshift_st;
        z0 = (z0 - __ldg(qs)) * __ldg(a.qscale + ag * 2);
        z1 = (z1 - __ldg(qs + 1)) * __ldg(a.qscale + ag * 2 + 1);
    }
}
__device__ __forceinline__ float finish_out(const RowArgs& a, long long ag, float v) {
    if (a.out_shift != nullptr) v += __ldg(a.out_shift + ag);
    return a.out_exp ? expf(v) : v;
}
// Result of row (agent ag, point-row rem = n % PPA) at output step `tidx`.  Must be reached by all 32 lanes of the warp
// (rows are consecutive lanes; `valid` masks rows past N).  K > 1 fused: CIF importance-sample bound
// log p ~= logsumexp_k lp_k - log K, rows ordered (a, g, k).
__device__ __forceinline__ void store_result(const RowArgs& a, long long ag, long long rem, long long tidx, float lp, bool valid) {
    if (a.fuse_k) {
        float m = valid ? lp : -INFINITY;
        for (int o = 1; o < a.K; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = (valid && m != -INFINITY) ? expf(lp - m) : 0.0f;
        for (int o = 1; o < a.K; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (valid && (rem % a.K) == 0) {
            const float v = (m == -INFINITY) ? -INFINITY : m + logf(s) - logf((float)a.K);
            store_out(a, ag * a.out_sa + (rem / a.K) * a.out_sp + tidx * a.out_st, finish_out(a, ag, v));
        }
    } else if (valid) {
        store_out(a, ag * a.out_sa + rem * a.out_sp + tidx * a.out_st, finish_out(a, ag, lp));
    }
}
#endif

struct SampleOut {
    float* seq;       // (N, T, 2)
    float* log_prob;  // (N, T)
    float* ldjs;      // (N, T)
    float* base_lp;   // (N)
};

}  // namespace fc
