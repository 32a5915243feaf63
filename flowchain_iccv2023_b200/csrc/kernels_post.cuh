// kernels_post.cuh -- the consumers immediately after the density hot path, HBM-bound:
//   k_to_trajectories : TP_metrics.unstandardize + output_to_trajectories (src/metrics/TP_metrics.py:68-112) fused:
//                       standardised state -> metric positions (x std, cumsum * dt + offset), log p -> p;
//   k_best_of_n       : the ADE/FDE best-of-N reduction of src/main.py:162-180 + TP_metrics.__call__ / evaluate_helper
//                       (TP_metrics.py:34-46, 163-191) for a whole batch of candidates with an on-device min.
// (The position-space density maps that replace TP_visualizer.prob_to_grid need no kernel of their own: the density
// kernels take an affine query transform and an exp / Jacobian epilogue, see RowArgs::qscale / out_shift in plan.h.)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace fc {

constexpr int kTrajRows = 128;      // (agent, sample) rows per CTA

// rows (A*S) of [Tk][ncol] floats; one thread scans one row in shared memory, global traffic is flat and coalesced.
//   state_mode 0 ('state_p', TP_metrics.py:71-77):  pos_t = st_t * std + offset
//   state_mode 1 ('state_v', TP_metrics.py:79-90):  pos_t = cumsum_t(st * std) * dt + offset
// ncol = 3: last column log p -> p = exp(log p) (TP_metrics.py:107-110); ncol = 2: positions only (("pred", 0)).
__global__ void __launch_bounds__(kTrajRows)
k_to_trajectories(const float* __restrict__ in, const float* __restrict__ offset, const float* __restrict__ dt, float std0, float std1,
                  int state_mode, float* __restrict__ out, long long rows, long long S, int Tk, int ncol) {
    extern __shared__ float traj_s[];                  // [kTrajRows][per + 1] (odd stride: conflict-free row scans)
    const int per = Tk * ncol, ld = per | 1;
    const long long r0 = (long long)blockIdx.x * kTrajRows;
    const int nr = (int)(rows - r0 < kTrajRows ? rows - r0 : kTrajRows);
    const float* src = in + r0 * per;
    for (int e = threadIdx.x; e < nr * per; e += kTrajRows) {
        const int r = e / per, c = e - r * per;
        traj_s[r * ld + c] = __ldg(src + e);
    }
    __syncthreads();
    if ((int)threadIdx.x < nr) {
        const long long ag = (r0 + threadIdx.x) / S;
        const float o0 = __ldg(offset + ag * 2), o1 = __ldg(offset + ag * 2 + 1);
        const float d = dt != nullptr ? __ldg(dt + ag) : 1.0f;
        float* row = traj_s + threadIdx.x * ld;
        float c0 = 0.0f, c1 = 0.0f;
        for (int t = 0; t < Tk; ++t) {
            const float v0 = __fmul_rn(row[t * ncol], std0), v1 = __fmul_rn(row[t * ncol + 1], std1);
            if (state_mode == 1) {
                c0 = __fadd_rn(c0, v0); c1 = __fadd_rn(c1, v1);                         // torch.cumsum in fp32
                row[t * ncol] = __fadd_rn(__fmul_rn(c0, d), o0);
                row[t * ncol + 1] = __fadd_rn(__fmul_rn(c1, d), o1);
            } else {
                row[t * ncol] = __fadd_rn(v0, o0);
                row[t * ncol + 1] = __fadd_rn(v1, o1);
            }
            if (ncol == 3) row[t * ncol + 2] = expf(row[t * ncol + 2]);
        }
    }
    __syncthreads();
    float* dst = out + r0 * per;
    for (int e = threadIdx.x; e < nr * per; e += kTrajRows) {
        const int r = e / per, c = e - r * per;
        dst[e] = traj_s[r * ld + c];
    }
}

// One warp per agent; lanes stride over the N candidates.  cand_st (A,N,T,2) standardised; gt (A,T,2) metric positions.
// NaN coordinates of a candidate are replaced by fallback_st[a] first (fastpredNF.py:131-136).  ADE over all T steps,
// FDE at the last; the two minima are taken independently (evaluate_helper stacks trials and takes min per agent).
__global__ void __launch_bounds__(128)
k_best_of_n(const float* __restrict__ cand, const float* __restrict__ gt, const float* __restrict__ last_obs, const float* __restrict__ dt,
            const float* __restrict__ fallback, float std0, float std1, int state_mode, float* __restrict__ ade_min,
            float* __restrict__ fde_min, int* __restrict__ ade_arg, int* __restrict__ fde_arg, long long A, long long N, int T) {
    const long long ag = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (ag >= A) return;
    const int lane = threadIdx.x & 31;
    const float o0 = __ldg(last_obs + ag * 2), o1 = __ldg(last_obs + ag * 2 + 1);
    const float d = dt != nullptr ? __ldg(dt + ag) : 1.0f;
    const float f0 = fallback != nullptr ? __ldg(fallback + ag * 2) : 0.0f, f1 = fallback != nullptr ? __ldg(fallback + ag * 2 + 1) : 0.0f;
    const float* g = gt + ag * T * 2;
    float best_a = INFINITY, best_f = INFINITY;
    int arg_a = 0x7fffffff, arg_f = 0x7fffffff;
    for (long long n = lane; n < N; n += 32) {
        const float* c = cand + (ag * N + n) * T * 2;
        float c0 = 0.0f, c1 = 0.0f, sum = 0.0f, fin = 0.0f;
        for (int t = 0; t < T; ++t) {
            float s0 = __ldg(c + 2 * t), s1 = __ldg(c + 2 * t + 1);
            if (fallback != nullptr) { s0 = isnan(s0) ? f0 : s0; s1 = isnan(s1) ? f1 : s1; }
            const float v0 = __fmul_rn(s0, std0), v1 = __fmul_rn(s1, std1);
            float p0, p1;
            if (state_mode == 1) {
                c0 = __fadd_rn(c0, v0); c1 = __fadd_rn(c1, v1);
                p0 = __fadd_rn(__fmul_rn(c0, d), o0); p1 = __fadd_rn(__fmul_rn(c1, d), o1);
            } else {
                p0 = __fadd_rn(v0, o0); p1 = __fadd_rn(v1, o1);
            }
            const float e0 = p0 - __ldg(g + 2 * t), e1 = p1 - __ldg(g + 2 * t + 1);
            fin = sqrtf(e0 * e0 + e1 * e1);
            sum += fin;
        }
        const float ade = sum / (float)T;
        // NaN errors never win (torch.min propagates NaN; a NaN candidate here can only come from a NaN ground truth)
        if (ade < best_a) { best_a = ade; arg_a = (int)n; }
        if (fin < best_f) { best_f = fin; arg_f = (int)n; }
    }
    for (int o = 16; o > 0; o >>= 1) {          // (value, first index) minimum across the warp
        const float oa = __shfl_xor_sync(0xffffffffu, best_a, o), of = __shfl_xor_sync(0xffffffffu, best_f, o);
        const int ia = __shfl_xor_sync(0xffffffffu, arg_a, o), jf = __shfl_xor_sync(0xffffffffu, arg_f, o);
        if (oa < best_a || (oa == best_a && ia < arg_a)) { best_a = oa; arg_a = ia; }
        if (of < best_f || (of == best_f && jf < arg_f)) { best_f = of; arg_f = jf; }
    }
    if (lane == 0) {
        ade_min[ag] = best_a; fde_min[ag] = best_f;
        if (ade_arg != nullptr) ade_arg[ag] = arg_a;
        if (fde_arg != nullptr) fde_arg[ag] = arg_f;
    }
}

}  // namespace fc
