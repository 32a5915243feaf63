/* flowchain_b200.h -- C-ABI of the B200-native FlowChain analytical-density hot path.
 *
 * The reference (meaten/FlowChain-ICCV2023) is pure Python/PyTorch and has NO FFI/plugin layer;
 * its boundary for this path is the flow object's methods and three methods of the model wrapper
 * (SURVEY.md §8b).  Each entry point below names the reference method it replaces; a maintainer
 * binds it with the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer to contiguous float32 unless stated otherwise; the
 *    caller owns all buffers; calls enqueue on `stream` (a cudaStream_t passed as void*), never
 *    synchronise, and are CUDA-graph capturable.  Two documented exceptions: fc_load_weights (one-off,
 *    synchronises `stream`), and the FIRST fc_grid_log_prob call with a K that is not a power of two <= 32 at a
 *    new (larger) size, which grows a handle-owned staging buffer (it returns an error instead when `stream`
 *    is being captured: run it once outside the capture);
 *  - calls run on the handle's device and restore the caller's current device before returning;
 *  - return value 0 = ok, non-zero = error; message via fc_last_error(); no C++ exceptions cross;
 *  - a handle is thread-compatible (one caller at a time); distinct handles are independent;
 *  - noise: `eps` (and `z0`) inject the CIF auxiliary draws (noise tape, layouts below); when
 *    NULL the kernels draw their own Philox4x32-10 normals from `seed` (production mode);
 *  - D (input_size) must be 2 and the coupling masks one-hot, as in every shipped config.
 *
 * Noise-tape layouts (rows N):
 *    separate / sampler : eps (N, E), E = sum_i C_i, step i at columns [off_i, off_i + C_i), C_i = C0 + 2 i
 *    sequential         : eps (N, sum_k (T-k) C_k); block k = the (N, T-k, C_k) draw of chain step k,
 *                         result index j at position j-k inside the block
 */
#ifndef FLOWCHAIN_B200_H
#define FLOWCHAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fc_handle fc_handle;

/* MODEL.FLOW.* knobs (reference: src/default_params.py:30-35; ctor fastpredNF.py:58-70, 815-826). */
typedef struct fc_model_desc {
    int32_t pred_len;     /* T  : chained CIF steps (DATA.PREDICT_LENGTH)        */
    int32_t input_size;   /* D  : flow dimension, must be 2                      */
    int32_t cond_size;    /* C0 : MODEL.FLOW.CONDITIONING_LENGTH                 */
    int32_t hidden_size;  /* H  : MODEL.FLOW.HIDDEN_SIZE                         */
    int32_t n_hidden;     /*      MODEL.FLOW.N_HIDDEN                            */
    int32_t n_blocks;     /*      MODEL.FLOW.N_BLOCKS (coupling+BN pairs / step) */
} fc_model_desc;

/* arithmetic path of the conditioner / auxiliary-density MLPs */
enum { FC_PREC_FP32 = 0,    /* CUDA-core fp32, parity 1e-4 (+rel) nats                                         */
       FC_PREC_BF16 = 1,    /* RETIRED (single-pass bf16 operands, ~1e-1 nats, failed the 5e-3 bar): every call returns an
                               error naming FC_PREC_F16 / FC_PREC_F16X3; the value is kept so the numbering is stable       */
       FC_PREC_F16X3 = 2,   /* tcgen05, split-fp16 operands (a_hi+a_lo)(W_hi+W_lo), 3 MMA passes, fp32 accumulate:
                               fp32-grade, meets the 5e-3 nats tensor-core parity bar                             */
       FC_PREC_F16 = 3 };   /* tcgen05, fp16 operands / fp32 accumulate, single pass: ~3e-2 nats                  */
/* Coverage: FC_PREC_FP32 every entry point and model shape.  FC_PREC_F16X3 / FC_PREC_F16: fc_log_prob,
 * fc_log_prob_sequential, fc_grid_log_prob[_multicast] (per-step and chain) and fc_sample_with_log_prob for
 * hidden_size % 16 == 0 and an even cond_size, on one of two kernels chosen at fc_load_weights:
 *   - every layer <= 80 wide (hidden_size <= 80 and 2*(cond_size + 2*pred_len) <= 80, e.g. the default ETH model):
 *     two resident 128-row tiles per CTA (csrc/kernels_tc3.cuh);
 *   - wider, up to hidden_size 256 / cond_size 64 (BASELINE config 5, the reference's tuning range
 *     src/main.py:222-229), as long as  2*(cond_size + 2*pred_len) <= 256  (two 256-column tensor-memory regions):
 *     one tile per CTA, in-place accumulator -> operand conversion, K-slab weight streaming (csrc/kernels_tcw.cuh);
 * other shapes return an error that names FC_PREC_FP32. */

/* prefix modes of the dense-grid evaluation */
enum { FC_PREFIX_SELF = 0,     /* x[:, i] = g for every step (the literal reference call)   */
       FC_PREFIX_SHARED = 1 }; /* prefix = per-agent conditioning trajectory cond_traj      */

/* Replaces: fastpredNF_TP.__init__ building self.flow (fastpredNF.py:58-70).  device = CUDA ordinal. */
int fc_create(fc_handle** out, const fc_model_desc* desc, int device);
void fc_destroy(fc_handle* h);
/* Last error text of `h` (or of the failed fc_create when h == NULL). */
const char* fc_last_error(const fc_handle* h);

/* Number of floats of the canonical weight blob for `desc` (layout: flowchain_iccv2023_b200/packing.py). */
size_t fc_weight_blob_floats(const fc_model_desc* desc);
/* Replaces: reading flow.state_dict() (fastpredNF.py:233-243 load path).  `host_blob` is a HOST pointer to
 * the canonical fp32 blob; it is re-laid-out for the kernels and uploaded.  `version` is remembered so the
 * host can re-pack after a training update().  Synchronises `stream` (one-off). */
int fc_load_weights(fc_handle* h, const float* host_blob, size_t n_floats, uint64_t version, void* stream);
uint64_t fc_weights_version(const fc_handle* h);

/* Replaces: fastpredNF.log_prob -> log_prob_separate -> forward_separate (fastpredNF.py:450-465, 571-584).
 * base_pos (N,2), x (N,T,2), cond (N,T,C0), eps (N,E)|NULL -> out (N,T). */
int fc_log_prob(fc_handle* h, const float* base_pos, const float* x, const float* cond,
                const float* eps, uint64_t seed, float* out, int64_t N, int precision, void* stream);

/* Replaces: fastpredNF.log_prob_sequential -> forward_sequential(seq, cond, n_step) (fastpredNF.py:454-458,
 * 548-569).  n_step < 0 means None.  out (N, n_out), n_out = T (None) or n_step+1. */
int fc_log_prob_sequential(fc_handle* h, const float* base_pos, const float* x, const float* cond,
                           const float* eps, uint64_t seed, int n_step, float* out, int64_t N,
                           int precision, void* stream);

/* Working version of the reference's dead predict_forward (fastpredNF.py:166-184):
 * log p for every agent x grid point x step without materialising the expanded rows.
 * base_pos (A,2), cond (A,T,C0), cond_traj (A,T,2) (FC_PREFIX_SHARED only, else NULL), grid (G,2),
 * eps (A*G*K, E or sequential width)|NULL.  sequential != 0 evaluates the full chain (a9) instead of
 * the per-step density (a8).  K >= 1 importance samples of the CIF auxiliary variable (K = 1 is the
 * reference; K > 1 returns logsumexp_k - log K, rows ordered (a, g, k)).
 * out: element (a, g, t) at out[a*out_agent_stride + g*out_point_stride + t*out_step_stride]. */
int fc_grid_log_prob(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                     const float* grid, const float* eps, uint64_t seed, int prefix_mode, int sequential,
                     int K, float* out, int64_t out_agent_stride, int64_t out_point_stride,
                     int64_t out_step_stride, int64_t A, int64_t G, int precision, void* stream);

/* Multi-GPU variant of fc_grid_log_prob (K = 1): `out_multicast` is the NVSwitch MULTICAST address of a symmetric
 * buffer (cuMulticast* / torch symmetric memory), offset to this rank's block of the global (A_total, T, G) maps.
 * The kernel's 4-byte result stores (multimem.st, system scope) are replicated to every peer by the switch, i.e. the
 * all-gather that the reference-side wrapper would run after the density evaluation (one ncclAllGather)
 * is performed by the stores themselves, tile by tile, with no second kernel.  The caller runs one cross-rank barrier
 * before reading the maps.  Replaces: flow.log_prob over rows = A*G (fastpredNF.py:450-465) + the all-gather. */
int fc_grid_log_prob_multicast(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                               const float* grid, const float* eps, uint64_t seed, int prefix_mode, int sequential,
                               float* out_multicast, int64_t out_agent_stride, int64_t out_point_stride,
                               int64_t out_step_stride, int64_t A, int64_t G, int precision, void* stream);

/* Replaces: fastpredNF.sample_with_log_prob -> fastpredNF_separate_cond.inverse (fastpredNF.py:472-477,
 * 586-607).  base_pos (B,2), cond (B,T,C0) are per agent and expanded over the S samples on the fly
 * (the reference expands them, fastpredNF.py:106-108).  z0 (B*S,2)|NULL is the standard-normal base
 * draw (u0 = base_pos + std0*z0), eps (B*S,E)|NULL.
 * Outputs: seq (B,S,T,2), log_prob (B,S,T), ldjs (B,S,T), base_lp (B,S).
 * precision: FC_PREC_FP32 (CUDA cores) or FC_PREC_F16X3 / FC_PREC_F16 (the tensor-core kernel run in the inverse direction). */
int fc_sample_with_log_prob(fc_handle* h, const float* base_pos, const float* cond, const float* z0,
                            const float* eps, uint64_t seed, float* seq, float* log_prob, float* ldjs,
                            float* base_lp, int64_t B, int64_t S, int precision, void* stream);

/* normalizer[a] = sum_s exp(base_lp[a,s])  (fastpredNF.py:114 base_prob_normalizer_tmp). */
int fc_base_normalizer(fc_handle* h, const float* base_lp, float* normalizer, int64_t A, int64_t S, void* stream);

/* Replaces: fastpredNF_TP.predict_from_new_obs (fastpredNF.py:142-164), vmapped over agents (the reference
 * only works at B = 1).  new_pos (A,2) = gt_st[:, t-1]; prob_st_0 (A,S,T,3); ldjs (A,S,T); normalizer (A)
 * -> prob_st_t (A,S,T-t,3).  0 < time_step < T. */
int fc_update(fc_handle* h, const float* new_pos, int time_step, const float* prob_st_0, const float* ldjs,
              const float* normalizer, float* prob_st_t, int64_t A, int64_t S, void* stream);

/* Lazy variant of fc_update for consumers that add `ldjs` themselves (the only form whose HBM
 * traffic allows < 100 us per 1000-agent batch): writes ONLY base_lp_t (A,S) = log(bp / sum_s bp * normalizer), the
 * per-sample term that fastpredNF.py:150-156 adds to ldjs_tmp[:, :, t:]; prob_st_t[a,s,j,2] == base_lp_t[a,s] + ldjs[a,s,t+j]. */
int fc_update_lazy(fc_handle* h, const float* new_pos, int time_step, const float* prob_st_0, const float* normalizer,
                   float* base_lp_t, int64_t A, int64_t S, void* stream);

/* ---- consumers immediately after the hot path -------------------------------------------- */

/* Replaces: TP_visualizer.prob_to_grid + griddata_on_cluster (src/visualization/TP_visualizer.py:118-144, 220-251) and the
 * interp2d ground-truth look-up (:165-179) by DIRECT evaluation of the analytic per-step density on the caller's METRIC
 * lattice, conditioned on a prefix trajectory (teacher forcing with the ground truth, or the ML prediction):
 *   out[a, g, t] = p_t(grid_xy[g] | cond_traj[a, :t])   (as_prob != 0)   or its log   (as_prob == 0)
 * The change of variables is fused into the density kernel: the query of agent a at step t is
 *   z = (grid_xy[g] - origin[a, t]) * inv_scale[a]      (standardised state of that agent / step)
 * and log_jac[a] (= -log(std_x std_y dt^2) for 'state_v') is added before the optional exp; see INTEGRATION.md for how
 * origin / inv_scale / log_jac follow from TP_metrics.unstandardize / output_to_trajectories (TP_metrics.py:68-112).
 * base_pos (A,2), cond (A,T,C0), cond_traj (A,T,2) standardised, grid_xy (G,2), origin (A,T,2), inv_scale (A,2),
 * log_jac (A)|NULL, eps (A*G, E)|NULL.  out strides as in fc_grid_log_prob. */
int fc_grid_density_maps(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                         const float* grid_xy, const float* origin, const float* inv_scale, const float* log_jac,
                         const float* eps, uint64_t seed, int as_prob, float* out, int64_t out_agent_stride,
                         int64_t out_point_stride, int64_t out_step_stride, int64_t A, int64_t G, int precision, void* stream);

/* Replaces: TP_metrics.unstandardize + output_to_trajectories (src/metrics/TP_metrics.py:68-112) for one tensor:
 * state_st (A,S,Tk,ncol) standardised state (+ log p when ncol == 3) -> out (same shape): metric positions (+ p = exp(log p)).
 *   state_mode 0 ('state_p'): pos = st * std + offset[a];  1 ('state_v'): pos = cumsum_t(st * std) * dt[a] + offset[a].
 * offset (A,2) = obs[:, -1, :2] for key k = 0, the metric ground truth gt[:, k-1, :2] for an update key k > 0; dt (A)|NULL
 * (= 1); std0 / std1 = the dataset's standardisation scale.  ("pred_st", 0) (B,T,2) is the S = 1, ncol = 2 case. */
int fc_to_trajectories(fc_handle* h, const float* state_st, const float* offset, const float* dt, float std0, float std1,
                       int state_mode, float* out, int64_t A, int64_t S, int Tk, int ncol, void* stream);

/* Replaces: the best-of-N ADE/FDE loop of src/main.py:162-180 (N_TRIAL predict calls + deepcopies) and
 * TP_metrics.__call__ / evaluate_helper (src/metrics/TP_metrics.py:34-46, 163-191) for one batch: cand_st (A,N,T,2) are the
 * N standardised candidates per agent (one sampler launch with S = N), gt (A,T,2) the METRIC ground truth, last_obs (A,2),
 * dt (A)|NULL, fallback_st (A,2)|NULL replaces NaN coordinates (fastpredNF.py:131-136).
 * -> ade_min (A) = min_n mean_t |pred_n - gt|, fde_min (A) = min_n |pred_n[T-1] - gt[T-1]| (independent minima, as the
 * reference's stack + min), ade_arg / fde_arg (A) int32 | NULL = the winning candidates. */
int fc_best_of_n(fc_handle* h, const float* cand_st, const float* gt, const float* last_obs, const float* dt,
                 const float* fallback_st, float std0, float std1, int state_mode, float* ade_min, float* fde_min,
                 int32_t* ade_arg, int32_t* fde_arg, int64_t A, int64_t N, int T, void* stream);

/* Diagnostic: D[128 x N] = A[128 x K] * B[N x K]^T on the tcgen05 tensor cores (bf16 operands, fp32
 * accumulate) through the same descriptor / TMEM / mbarrier helpers the fused kernel uses.  HOST pointers,
 * row-major fp32; N multiple of 16 in [16,256], K in [1,256].  mode 0: bf16, A and B tiles in shared memory;
 * mode 1: fp16, A operand resident in tensor memory (written with tcgen05.st).  Error text via fc_last_error(NULL). */
int fc_tc_gemm_selftest(const float* A, const float* B, float* D, int N, int K, int device, int mode);

/* Number of kernels this library has launched on behalf of `h` since creation (bench.py's gpu_launches). */
uint64_t fc_launch_count(const fc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* FLOWCHAIN_B200_H */
