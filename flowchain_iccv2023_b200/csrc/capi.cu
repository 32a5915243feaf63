// capi.cu -- C-ABI (include/flowchain_b200.h) over the fused CIF-chain kernels: handle, weight
// re-layout ("packing") and launch logic.  No torch types; plain pointers and sizes only.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/flowchain_b200.h"
#include "kernels_fp32.cuh"
#include "kernels_post.cuh"
#include "kernels_tc3.cuh"
#include "kernels_tcw.cuh"
#include "plan.h"
#include "tc_f16x3.cuh"
#include "tc_host.cuh"

using namespace fc;

struct fc_handle {
    fc_model_desc desc{};
    int device = 0;
    std::string err;
    uint64_t version = 0;
    uint64_t launches = 0;
    bool loaded = false;
    std::vector<StepPlan> plans;      // host copy
    StepPlan* d_plans = nullptr;
    float* d_wbuf = nullptr;          // packed fp32 weights
    size_t wbuf_floats = 0;
    // accurate tensor-core path: [0] = 3-pass split fp16 (f16x3), [1] = single-pass fp16
    unsigned char* d_t2img[2] = {nullptr, nullptr};
    fc::T2Step* d_t2steps[2] = {nullptr, nullptr};
    fc::T2Dims t2dims[2]{};
    std::vector<fc::T2Step> t2steps_host[2];
    // wide tensor-core path (kernels_tcw.cuh), same index convention; used when the two-tile kernel refuses the shape
    unsigned char* d_twimg[2] = {nullptr, nullptr};
    float* d_twtails[2] = {nullptr, nullptr};
    fc::TwStep* d_twsteps[2] = {nullptr, nullptr};
    fc::TwDims twdims[2]{};
    float* d_scratch = nullptr;       // K>1 importance-sample staging
    size_t scratch_floats = 0;
    Dims dims{};
    int E = 0, Eseq = 0;
    int smem_optin = 0;
    int num_sms = 0;
};

static thread_local std::string g_err;

static int fail(fc_handle* h, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_err = buf;
    return 1;
}

#define FC_CUDA(h, call)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail(h, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

static inline int pad8(int v) { return (v + 7) / 8 * 8; }

// Entry points run on the handle's device and give the caller's current device back on return.
struct DevGuard {
    int prev = -1, dev;
    explicit DevGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

static int check_desc(const fc_model_desc* d, fc_handle* h) {
    if (!d) return fail(h, "null model desc");
    if (d->input_size != 2) return fail(h, "input_size must be 2 (got %d)", d->input_size);
    if (d->pred_len < 1 || d->pred_len > kMaxSteps) return fail(h, "pred_len %d out of range [1,%d]", d->pred_len, kMaxSteps);
    if (d->n_blocks < 1 || d->n_blocks > kMaxBlocks) return fail(h, "n_blocks %d out of range [1,%d]", d->n_blocks, kMaxBlocks);
    if (d->n_hidden < 0 || d->n_hidden > kMaxHidden) return fail(h, "n_hidden %d out of range [0,%d]", d->n_hidden, kMaxHidden);
    if (d->cond_size < 1 || d->hidden_size < 1) return fail(h, "cond_size / hidden_size must be positive");
    return 0;
}

extern "C" size_t fc_weight_blob_floats(const fc_model_desc* d) {
    if (!d) return 0;
    const size_t T = d->pred_len, D = d->input_size, H = d->hidden_size;
    size_t n = T * D;
    for (size_t i = 0; i < T; ++i) {
        const size_t C = d->cond_size + i * D, c = C + D;
        const size_t net = (c * H + H) + (size_t)d->n_hidden * (H * H + H) + (H * D + D);
        n += (size_t)d->n_blocks * (D + 2 * net + 4 * D);
        const size_t dn = (c * 2 * c + 2 * c) + (2 * c * 2 * c + 2 * c) + (2 * c * C + C);
        n += 4 * dn;
    }
    return n;
}

extern "C" int fc_create(fc_handle** out, const fc_model_desc* desc, int device) {
    if (!out) return fail(nullptr, "null out pointer");
    *out = nullptr;
    if (check_desc(desc, nullptr)) return 1;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, "device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10) return fail(nullptr, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    fc_handle* h = new fc_handle();
    h->desc = *desc;
    h->device = device;
    h->smem_optin = (int)prop.sharedMemPerBlockOptin;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return 0;
}

extern "C" void fc_destroy(fc_handle* h) {
    if (!h) return;
    DevGuard guard(h->device);
    cudaFree(h->d_plans);
    cudaFree(h->d_wbuf);
    for (int i = 0; i < 2; ++i) {
        cudaFree(h->d_t2img[i]); cudaFree(h->d_t2steps[i]);
        cudaFree(h->d_twimg[i]); cudaFree(h->d_twtails[i]); cudaFree(h->d_twsteps[i]);
    }
    cudaFree(h->d_scratch);
    delete h;
}

extern "C" const char* fc_last_error(const fc_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }
extern "C" uint64_t fc_weights_version(const fc_handle* h) { return h ? h->version : 0; }
extern "C" uint64_t fc_launch_count(const fc_handle* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------------------
// weight re-layout.  Canonical blob order (flowchain_iccv2023_b200/packing.py):
//   var[T*D]
//   for step i: for block b: mask[D]; s_net layers (W[out][in], bias[out]) x (n_hidden+2); t_net likewise;
//                            log_gamma[D], beta[D], running_mean[D], running_var[D]
//               for dens in (r_u_given_z, q_u_given_w): for net in (shift, log_scale): 3 layers (W, bias)
// ---------------------------------------------------------------------------------------------
namespace {
struct Reader {
    const float* p;
    size_t n, pos = 0;
    const float* take(size_t k) {
        const float* r = p + pos;
        pos += k;
        return r;
    }
};

// append W^T padded: dst[k][n] = W[n][k]; K_pad rows, NP cols
int32_t put_layer(std::vector<float>& buf, const float* W, const float* b, int n_out, int n_in, int K_pad, int NP,
                  LayerRef* ref) {
    const size_t w_off = buf.size();
    buf.resize(w_off + (size_t)K_pad * NP + NP, 0.0f);
    for (int k = 0; k < n_in; ++k)
        for (int n = 0; n < n_out; ++n) buf[w_off + (size_t)k * NP + n] = W[(size_t)n * n_in + k];
    const size_t b_off = w_off + (size_t)K_pad * NP;
    for (int n = 0; n < n_out; ++n) buf[b_off + n] = b[n];
    ref->w = (int32_t)w_off;
    ref->b = (int32_t)b_off;
    ref->K = K_pad;
    ref->NP = NP;
    return (int32_t)w_off;
}
}  // namespace

extern "C" int fc_load_weights(fc_handle* h, const float* blob, size_t n_floats, uint64_t version, void* stream_) {
    if (!h) return fail(nullptr, "null handle");
    if (!blob) return fail(h, "null weight blob");
    const fc_model_desc& d = h->desc;
    if (n_floats != fc_weight_blob_floats(&d))
        return fail(h, "weight blob has %zu floats, expected %zu for this model", n_floats, fc_weight_blob_floats(&d));
    cudaStream_t stream = (cudaStream_t)stream_;
    DevGuard guard(h->device);
    const int T = d.pred_len, D = 2, H = d.hidden_size, HH = pad8(H);
    Reader rd{blob, n_floats};
    const float* var = rd.take((size_t)T * D);
    std::vector<float> buf;
    buf.reserve(n_floats * 2);
    std::vector<StepPlan> plans(T);
    std::vector<fc::TcStepSrc> tcsrc(T);
    int eoff = 0, seoff = 0, cmax = 0, hmax = 0;
    for (int i = 0; i < T; ++i) {
        StepPlan& sp = plans[i];
        memset(&sp, 0, sizeof sp);
        const int C = d.cond_size + i * D, c = C + D, HP = pad8(2 * c), CP = pad8(C);
        sp.C = C; sp.c = c; sp.HP = HP; sp.CP = CP; sp.HH = HH;
        sp.eps_off = eoff; sp.seq_eps_off = seoff;
        eoff += C; seoff += (T - i) * C;
        sp.n_blocks = d.n_blocks; sp.n_hidden = d.n_hidden;
        sp.std[0] = fmaxf(var[i * D], 1e-4f);
        sp.std[1] = fmaxf(var[i * D + 1], 1e-4f);
        cmax = c > cmax ? c : cmax;
        const int hr = 2 * HP > 2 * HH ? 2 * HP : 2 * HH;
        hmax = hr > hmax ? hr : hmax;
        fc::TcStepSrc& ts = tcsrc[i];
        for (int b = 0; b < d.n_blocks; ++b) {
            const float* mask = rd.take(D);
            const bool m0 = mask[0] == 1.0f && mask[1] == 0.0f, m1 = mask[0] == 0.0f && mask[1] == 1.0f;
            if (!m0 && !m1) return fail(h, "step %d block %d: coupling mask must be one-hot (got [%g,%g])", i, b, mask[0], mask[1]);
            sp.masked_dim[b] = m0 ? 0 : 1;
            const int j = 1 - sp.masked_dim[b];
            for (int net = 0; net < 2; ++net) {
                int n_in = c;
                for (int l = 0; l <= d.n_hidden; ++l) {
                    const float* W = rd.take((size_t)H * n_in);
                    const float* bias = rd.take(H);
                    put_layer(buf, W, bias, H, n_in, l == 0 ? n_in : HH, HH, &sp.cpl[b][net][l]);
                    ts.cpl_W[b][net][l] = W; ts.cpl_b[b][net][l] = bias;
                    n_in = H;
                }
                const float* W = rd.take((size_t)D * H);
                const float* bias = rd.take(D);
                const size_t off = buf.size();
                buf.resize(off + HH, 0.0f);
                for (int k = 0; k < H; ++k) buf[off + k] = W[(size_t)j * H + k];
                sp.cpl_last_w[b][net] = (int32_t)off;
                sp.cpl_last_b[b][net] = bias[j];
                ts.cpl_W[b][net][d.n_hidden + 1] = W; ts.cpl_b[b][net][d.n_hidden + 1] = bias;
            }
            const float* lg = rd.take(D);
            const float* beta = rd.take(D);
            const float* mean = rd.take(D);
            const float* rv = rd.take(D);
            double ldj = 0.0;
            for (int q = 0; q < D; ++q) {
                const double sd = std::sqrt((double)rv[q] + 1e-5);
                const double sc = std::exp((double)lg[q]) / sd;
                sp.bn_scale[b][q] = (float)sc;
                sp.bn_shift[b][q] = (float)((double)beta[q] - sc * (double)mean[q]);
                const double isc = std::exp(-(double)lg[q]) * sd;
                sp.bn_iscale[b][q] = (float)isc;
                sp.bn_ishift[b][q] = (float)((double)mean[q] - (double)beta[q] * isc);
                ldj += (double)lg[q] - 0.5 * std::log((double)rv[q] + 1e-5);
            }
            sp.bn_ldj[b] = (float)ldj;
        }
        for (int which = 0; which < 2; ++which)
            for (int net = 0; net < 2; ++net) {
                const int dims[4] = {c, 2 * c, 2 * c, C};
                for (int l = 0; l < 3; ++l) {
                    const float* W = rd.take((size_t)dims[l + 1] * dims[l]);
                    const float* bias = rd.take(dims[l + 1]);
                    put_layer(buf, W, bias, dims[l + 1], dims[l], l == 0 ? c : HP, l == 2 ? CP : HP, &sp.dens[which][net][l]);
                    ts.dens_W[which][net][l] = W; ts.dens_b[which][net][l] = bias;
                }
            }
        while (buf.size() % 4) buf.push_back(0.0f);
    }
    if (rd.pos != n_floats) return fail(h, "internal: blob parse consumed %zu of %zu floats", rd.pos, n_floats);
    h->E = eoff; h->Eseq = seoff;
    h->dims.in_rows = cmax; h->dims.u_rows = cmax; h->dims.h_rows = hmax; h->dims.T = T; h->dims.C0 = d.cond_size;

    FC_CUDA(h, cudaFree(h->d_plans)); h->d_plans = nullptr;
    FC_CUDA(h, cudaFree(h->d_wbuf)); h->d_wbuf = nullptr;
    FC_CUDA(h, cudaMalloc(&h->d_plans, sizeof(StepPlan) * T));
    FC_CUDA(h, cudaMalloc(&h->d_wbuf, buf.size() * sizeof(float)));
    FC_CUDA(h, cudaMemcpyAsync(h->d_plans, plans.data(), sizeof(StepPlan) * T, cudaMemcpyHostToDevice, stream));
    FC_CUDA(h, cudaMemcpyAsync(h->d_wbuf, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    for (int i = 0; i < 2; ++i) {
        std::vector<unsigned char> img2;
        std::vector<fc::T2Step> st2;
        fc::tc2_pack(d, plans, tcsrc, i == 0 ? 3 : 1, img2, st2, h->t2dims[i], h->smem_optin);
        if (h->t2dims[i].supported && !fc::tc3_fits(st2, h->smem_optin)) h->t2dims[i].supported = 0;    // -> wide kernel
        h->t2steps_host[i] = st2;
        FC_CUDA(h, cudaFree(h->d_t2img[i])); h->d_t2img[i] = nullptr;
        FC_CUDA(h, cudaFree(h->d_t2steps[i])); h->d_t2steps[i] = nullptr;
        if (h->t2dims[i].supported) {
            FC_CUDA(h, cudaMalloc(&h->d_t2img[i], img2.size()));
            // pageable sources: cudaMemcpyAsync returns once they are staged, so the vectors may go out of scope
            FC_CUDA(h, cudaMemcpyAsync(h->d_t2img[i], img2.data(), img2.size(), cudaMemcpyHostToDevice, stream));
            FC_CUDA(h, cudaMalloc(&h->d_t2steps[i], sizeof(fc::T2Step) * T));
            FC_CUDA(h, cudaMemcpyAsync(h->d_t2steps[i], st2.data(), sizeof(fc::T2Step) * T, cudaMemcpyHostToDevice, stream));
        }
    }
    for (int i = 0; i < 2; ++i) {       // wide image only for the shapes the two-tile kernel does not take
        FC_CUDA(h, cudaFree(h->d_twimg[i])); h->d_twimg[i] = nullptr;
        FC_CUDA(h, cudaFree(h->d_twtails[i])); h->d_twtails[i] = nullptr;
        FC_CUDA(h, cudaFree(h->d_twsteps[i])); h->d_twsteps[i] = nullptr;
        memset(&h->twdims[i], 0, sizeof h->twdims[i]);
        if (h->t2dims[i].supported) continue;
        std::vector<unsigned char> imgw;
        std::vector<float> tailsw;
        std::vector<fc::TwStep> stw;
        fc::tcw_pack(d, plans, tcsrc, i == 0 ? 3 : 1, imgw, tailsw, stw, h->twdims[i], h->smem_optin);
        if (!h->twdims[i].supported) continue;
        FC_CUDA(h, cudaMalloc(&h->d_twimg[i], imgw.size()));
        FC_CUDA(h, cudaMemcpyAsync(h->d_twimg[i], imgw.data(), imgw.size(), cudaMemcpyHostToDevice, stream));
        FC_CUDA(h, cudaMalloc(&h->d_twtails[i], tailsw.size() * sizeof(float)));
        FC_CUDA(h, cudaMemcpyAsync(h->d_twtails[i], tailsw.data(), tailsw.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        FC_CUDA(h, cudaMalloc(&h->d_twsteps[i], sizeof(fc::TwStep) * T));
        FC_CUDA(h, cudaMemcpyAsync(h->d_twsteps[i], stw.data(), sizeof(fc::TwStep) * T, cudaMemcpyHostToDevice, stream));
    }
    FC_CUDA(h, cudaStreamSynchronize(stream));
    h->plans = plans;
    h->wbuf_floats = buf.size();
    h->version = version;
    h->loaded = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
namespace {
constexpr int kR = 64;

int ready(fc_handle* h) {
    if (!h) return fail(nullptr, "null handle");
    if (!h->loaded) return fail(h, "weights not loaded (call fc_load_weights first)");
    return 0;
}
// every compute entry point: validate the handle, then run on its device (restored on return)
#define FC_ENTER(h)           \
    if (ready(h)) return 1;   \
    DevGuard guard_((h)->device)

template <typename Kern>
int prep_kernel(fc_handle* h, Kern kern, size_t smem) {
    if ((int)smem > h->smem_optin)
        return fail(h, "model too wide for the fp32 path: needs %zu B shared memory per CTA, device allows %d", smem, h->smem_optin);
    FC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return 0;
}

enum Mode { MODE_SEPARATE, MODE_SEQUENTIAL };

int launch_density(fc_handle* h, RowArgs a, Mode mode, int precision, cudaStream_t stream) {
    if (a.N <= 0) return 0;
    const int T = h->desc.pred_len;
    const long long tiles = (a.N + kR - 1) / kR;
    if (tiles > 0x7fffffffLL) return fail(h, "too many rows (%lld)", a.N);
    if (precision == FC_PREC_BF16)
        return fail(h, "FC_PREC_BF16 (single-pass bf16 operands) was retired: ~1e-1 nats, it fails the 5e-3 tensor-core bar; use "
                       "FC_PREC_F16 (same speed, ~3e-2 nats) or FC_PREC_F16X3 (5e-3 nats)");
    if (precision == FC_PREC_F16X3 || precision == FC_PREC_F16) {
        std::string msg;
        const int i = precision == FC_PREC_F16X3 ? 0 : 1;
        int rc;
        if (h->t2dims[i].supported) {       // narrow models (every layer <= 80 wide): two resident tiles per CTA
            if (mode == MODE_SEQUENTIAL)
                rc = i == 0 ? fc::tc3_launch_sequential<3>(h->t2dims[0], h->t2steps_host[0], h->d_t2steps[0], h->d_t2img[0], a, h->num_sms, h->smem_optin, stream, &h->launches, &msg)
                            : fc::tc3_launch_sequential<1>(h->t2dims[1], h->t2steps_host[1], h->d_t2steps[1], h->d_t2img[1], a, h->num_sms, h->smem_optin, stream, &h->launches, &msg);
            else
                rc = i == 0 ? fc::tc3_launch_separate<3>(h->t2dims[0], h->t2steps_host[0], h->d_t2steps[0], h->d_t2img[0], a, h->num_sms, h->smem_optin, stream, &h->launches, &msg)
                            : fc::tc3_launch_separate<1>(h->t2dims[1], h->t2steps_host[1], h->d_t2steps[1], h->d_t2img[1], a, h->num_sms, h->smem_optin, stream, &h->launches, &msg);
        } else {                            // wide models (hidden up to 256, cond up to 64): one tile per CTA, K-slab weight streaming
            const int m = mode == MODE_SEQUENTIAL ? 1 : 0;
            rc = i == 0 ? fc::tcw_launch<3>(h->twdims[0], h->d_twsteps[0], h->d_twimg[0], h->d_twtails[0], a, SampleOut{}, m, h->num_sms, h->smem_optin, stream, &h->launches, &msg)
                        : fc::tcw_launch<1>(h->twdims[1], h->d_twsteps[1], h->d_twimg[1], h->d_twtails[1], a, SampleOut{}, m, h->num_sms, h->smem_optin, stream, &h->launches, &msg);
        }
        if (rc) return fail(h, "%s", msg.c_str());
        return 0;
    }
    if (precision != FC_PREC_FP32) return fail(h, "unknown precision %d", precision);
    // 64-row tiles normally; models too wide for that (e.g. HIDDEN_SIZE 256) fall back to 32-row tiles
    const bool wide = (int)smem_bytes_fp32(h->dims, 64, false) > h->smem_optin;
    const int R = wide ? 32 : 64;
    const size_t smem = smem_bytes_fp32(h->dims, R, false);
    const long long rtiles = (a.N + R - 1) / R;
    if (rtiles > 0x7fffffffLL) return fail(h, "too many rows (%lld)", a.N);
    if (mode == MODE_SEPARATE) {
        dim3 grid((unsigned)rtiles, (unsigned)T);
        if (wide) {
            if (prep_kernel(h, k_separate_fp32<32>, smem)) return 1;
            k_separate_fp32<32><<<grid, kThreads, smem, stream>>>(h->d_plans, h->d_wbuf, a, h->dims);
        } else {
            if (prep_kernel(h, k_separate_fp32<64>, smem)) return 1;
            k_separate_fp32<64><<<grid, kThreads, smem, stream>>>(h->d_plans, h->d_wbuf, a, h->dims);
        }
    } else {
        dim3 grid((unsigned)rtiles, (unsigned)(T - a.seq_jlo));
        if (wide) {
            if (prep_kernel(h, k_sequential_fp32<32>, smem)) return 1;
            k_sequential_fp32<32><<<grid, kThreads, smem, stream>>>(h->d_plans, h->d_wbuf, a, h->dims);
        } else {
            if (prep_kernel(h, k_sequential_fp32<64>, smem)) return 1;
            k_sequential_fp32<64><<<grid, kThreads, smem, stream>>>(h->d_plans, h->d_wbuf, a, h->dims);
        }
    }
    h->launches++;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

int seq_bounds(fc_handle* h, int n_step, int* klo, int* jlo) {
    const int T = h->desc.pred_len;
    if (n_step < 0) { *klo = 0; *jlo = 0; return 0; }
    if (n_step < 1 || n_step > T - 1) return fail(h, "n_step must be None (<0) or in [1, %d], got %d", T - 1, n_step);
    *klo = T - n_step;
    *jlo = T - n_step - 1;
    return 0;
}
}  // namespace

extern "C" int fc_log_prob(fc_handle* h, const float* base_pos, const float* x, const float* cond, const float* eps,
                           uint64_t seed, float* out, int64_t N, int precision, void* stream) {
    FC_ENTER(h);
    if (N < 0) return fail(h, "negative row count");
    if (N > 0 && (!base_pos || !x || !cond || !out)) return fail(h, "null tensor pointer");
    const int T = h->desc.pred_len, C0 = h->desc.cond_size;
    RowArgs a{};
    a.x = Src{x, 2LL * T, 0, 2};
    a.prefix = a.x;
    a.cond = Src{cond, (long long)T * C0, 0, C0};
    a.base = Src{base_pos, 2, 0, 0};
    a.eps = eps; a.eps_rs = h->E; a.seed = seed;
    a.N = N; a.PPA = 1; a.K = 1;
    a.out = out; a.out_sa = T; a.out_sp = 0; a.out_st = 1;
    return launch_density(h, a, MODE_SEPARATE, precision, (cudaStream_t)stream);
}

extern "C" int fc_log_prob_sequential(fc_handle* h, const float* base_pos, const float* x, const float* cond,
                                      const float* eps, uint64_t seed, int n_step, float* out, int64_t N,
                                      int precision, void* stream) {
    FC_ENTER(h);
    if (N < 0) return fail(h, "negative row count");
    if (N > 0 && (!base_pos || !x || !cond || !out)) return fail(h, "null tensor pointer");
    const int T = h->desc.pred_len, C0 = h->desc.cond_size;
    RowArgs a{};
    if (seq_bounds(h, n_step, &a.seq_klo, &a.seq_jlo)) return 1;
    a.x = Src{x, 2LL * T, 0, 2};
    a.prefix = a.x;
    a.cond = Src{cond, (long long)T * C0, 0, C0};
    a.base = Src{base_pos, 2, 0, 0};
    a.eps = eps; a.eps_rs = h->Eseq; a.seed = seed;
    a.N = N; a.PPA = 1; a.K = 1;
    a.out = out; a.out_sa = T - a.seq_jlo; a.out_sp = 0; a.out_st = 1;
    return launch_density(h, a, MODE_SEQUENTIAL, precision, (cudaStream_t)stream);
}

namespace {
struct QueryMap {          // affine query transform + output epilogue of the position-space maps (RowArgs::qscale ...)
    const float* origin = nullptr;      // (A, T, 2)
    const float* inv_scale = nullptr;   // (A, 2)
    const float* log_jac = nullptr;     // (A)
    int as_prob = 0;
};
}  // namespace

static int grid_log_prob_impl(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                              const float* grid, const float* eps, uint64_t seed, int prefix_mode, int sequential,
                              int K, float* out, int64_t out_agent_stride, int64_t out_point_stride,
                              int64_t out_step_stride, int64_t A, int64_t G, int precision, void* stream_, int multicast,
                              const QueryMap* qm = nullptr) {
    FC_ENTER(h);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (A < 0 || G < 0) return fail(h, "negative size");
    if (K < 1 || K > 4096) return fail(h, "K must be in [1, 4096], got %d", K);
    if (A == 0 || G == 0) return 0;
    if (!base_pos || !cond || !grid || !out) return fail(h, "null tensor pointer");
    if (prefix_mode != FC_PREFIX_SELF && prefix_mode != FC_PREFIX_SHARED) return fail(h, "unknown prefix_mode %d", prefix_mode);
    if (prefix_mode == FC_PREFIX_SHARED && !cond_traj) return fail(h, "FC_PREFIX_SHARED needs cond_traj");
    const int T = h->desc.pred_len, C0 = h->desc.cond_size;
    RowArgs a{};
    a.x = Src{grid, 0, 2, 0};
    a.prefix = prefix_mode == FC_PREFIX_SHARED ? Src{cond_traj, 2LL * T, 0, 2} : a.x;
    a.cond = Src{cond, (long long)T * C0, 0, C0};
    a.base = Src{base_pos, 2, 0, 0};
    a.eps = eps; a.eps_rs = sequential ? h->Eseq : h->E; a.seed = seed;
    a.N = A * G * K; a.PPA = G * K; a.K = K;
    a.seq_klo = 0; a.seq_jlo = 0;
    a.out_sys = multicast ? 1 : 0;
    if (qm != nullptr) {
        a.qshift = qm->origin; a.qshift_sa = 2LL * T; a.qshift_st = 2;
        a.qscale = qm->inv_scale; a.out_shift = qm->log_jac; a.out_exp = qm->as_prob;
    }
    // K = 2^j <= 32: the K samples of a point are consecutive rows = consecutive lanes, and logsumexp_k - log K is reduced
    // with warp shuffles inside the density kernel (one store per point and step); other K go through a scratch buffer
    const bool fuse_k = K > 1 && K <= 32 && (K & (K - 1)) == 0;
    if (K == 1 || fuse_k) {
        a.fuse_k = fuse_k ? 1 : 0;
        a.out = out; a.out_sa = out_agent_stride; a.out_sp = out_point_stride; a.out_st = out_step_stride;
        return launch_density(h, a, sequential ? MODE_SEQUENTIAL : MODE_SEPARATE, precision, stream);
    }
    const size_t need = (size_t)a.N * T;
    if (need > h->scratch_floats) {          // grows outside graph capture only (the one call that allocates: documented in the header)
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        FC_CUDA(h, cudaStreamIsCapturing(stream, &cs));
        if (cs != cudaStreamCaptureStatusNone)
            return fail(h, "K = %d needs a %zu-float staging buffer that is not allocated yet: run the call once outside graph capture", K, need);
        FC_CUDA(h, cudaStreamSynchronize(stream));
        FC_CUDA(h, cudaFree(h->d_scratch));
        h->d_scratch = nullptr; h->scratch_floats = 0;
        FC_CUDA(h, cudaMalloc(&h->d_scratch, need * sizeof(float)));
        h->scratch_floats = need;
    }
    a.out = h->d_scratch; a.out_sa = (long long)a.PPA * T; a.out_sp = T; a.out_st = 1;
    if (launch_density(h, a, sequential ? MODE_SEQUENTIAL : MODE_SEPARATE, precision, stream)) return 1;
    const long long AG = A * G, total = AG * T;
    k_logsumexp_k<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(h->d_scratch, K, T, AG, G, out, out_agent_stride,
                                                                       out_point_stride, out_step_stride);
    h->launches++;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int fc_grid_log_prob(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                                const float* grid, const float* eps, uint64_t seed, int prefix_mode, int sequential,
                                int K, float* out, int64_t out_agent_stride, int64_t out_point_stride,
                                int64_t out_step_stride, int64_t A, int64_t G, int precision, void* stream) {
    return grid_log_prob_impl(h, base_pos, cond, cond_traj, grid, eps, seed, prefix_mode, sequential, K, out, out_agent_stride,
                              out_point_stride, out_step_stride, A, G, precision, stream, 0);
}

extern "C" int fc_grid_log_prob_multicast(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                                          const float* grid, const float* eps, uint64_t seed, int prefix_mode, int sequential,
                                          float* out_multicast, int64_t out_agent_stride, int64_t out_point_stride,
                                          int64_t out_step_stride, int64_t A, int64_t G, int precision, void* stream) {
    return grid_log_prob_impl(h, base_pos, cond, cond_traj, grid, eps, seed, prefix_mode, sequential, 1, out_multicast,
                              out_agent_stride, out_point_stride, out_step_stride, A, G, precision, stream, 1);
}

extern "C" int fc_sample_with_log_prob(fc_handle* h, const float* base_pos, const float* cond, const float* z0,
                                       const float* eps, uint64_t seed, float* seq, float* log_prob, float* ldjs,
                                       float* base_lp, int64_t B, int64_t S, int precision, void* stream) {
    FC_ENTER(h);
    if (B < 0 || S < 0) return fail(h, "negative size");
    if (B == 0 || S == 0) return 0;
    if (!base_pos || !cond || !seq || !log_prob || !ldjs || !base_lp) return fail(h, "null tensor pointer");
    if (precision != FC_PREC_FP32 && precision != FC_PREC_F16X3 && precision != FC_PREC_F16)
        return fail(h, "the sampler runs with FC_PREC_FP32, FC_PREC_F16X3 or FC_PREC_F16");
    const int T = h->desc.pred_len, C0 = h->desc.cond_size;
    RowArgs a{};
    a.cond = Src{cond, (long long)T * C0, 0, C0};
    a.base = Src{base_pos, 2, 0, 0};
    a.x = a.base; a.prefix = a.base;
    a.eps = eps; a.eps_rs = h->E; a.z0 = z0; a.seed = seed;
    a.N = B * S; a.PPA = S; a.K = 1;
    SampleOut o{seq, log_prob, ldjs, base_lp};
    if (precision != FC_PREC_FP32) {     // tensor-core sampler: the two-tile kernel run in the inverse direction
        std::string msg;
        const int i = precision == FC_PREC_F16X3 ? 0 : 1;
        int rc;
        if (h->t2dims[i].supported)
            rc = i == 0 ? fc::tc3_launch_sample<3>(h->t2dims[0], h->t2steps_host[0], h->d_t2steps[0], h->d_t2img[0], a, o, h->num_sms, h->smem_optin, (cudaStream_t)stream, &h->launches, &msg)
                        : fc::tc3_launch_sample<1>(h->t2dims[1], h->t2steps_host[1], h->d_t2steps[1], h->d_t2img[1], a, o, h->num_sms, h->smem_optin, (cudaStream_t)stream, &h->launches, &msg);
        else
            rc = i == 0 ? fc::tcw_launch<3>(h->twdims[0], h->d_twsteps[0], h->d_twimg[0], h->d_twtails[0], a, o, 2, h->num_sms, h->smem_optin, (cudaStream_t)stream, &h->launches, &msg)
                        : fc::tcw_launch<1>(h->twdims[1], h->d_twsteps[1], h->d_twimg[1], h->d_twtails[1], a, o, 2, h->num_sms, h->smem_optin, (cudaStream_t)stream, &h->launches, &msg);
        if (rc) return fail(h, "%s", msg.c_str());
        return 0;
    }
    const bool wide = (int)smem_bytes_fp32(h->dims, 64, true) > h->smem_optin;
    const int R = wide ? 32 : 64;
    const long long tiles = (a.N + R - 1) / R;
    if (tiles > 0x7fffffffLL) return fail(h, "too many rows");
    const size_t smem = smem_bytes_fp32(h->dims, R, true);
    if (wide) {
        if (prep_kernel(h, k_sample_fp32<32>, smem)) return 1;
        k_sample_fp32<32><<<(unsigned)tiles, kThreads, smem, (cudaStream_t)stream>>>(h->d_plans, h->d_wbuf, a, h->dims, o);
    } else {
        if (prep_kernel(h, k_sample_fp32<64>, smem)) return 1;
        k_sample_fp32<64><<<(unsigned)tiles, kThreads, smem, (cudaStream_t)stream>>>(h->d_plans, h->d_wbuf, a, h->dims, o);
    }
    h->launches++;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int fc_base_normalizer(fc_handle* h, const float* base_lp, float* normalizer, int64_t A, int64_t S, void* stream) {
    FC_ENTER(h);
    if (A <= 0 || S <= 0) return A < 0 || S < 0 ? fail(h, "negative size") : 0;
    if (!base_lp || !normalizer) return fail(h, "null tensor pointer");
    k_base_normalizer<<<(unsigned)A, 512, 0, (cudaStream_t)stream>>>(base_lp, normalizer, S);
    h->launches++;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

static int update_impl(fc_handle* h, const float* new_pos, int t, const float* prob_st_0, const float* ldjs,
                       const float* normalizer, float* out, int64_t A, int64_t S, void* stream, bool lazy) {
    FC_ENTER(h);
    const int T = h->desc.pred_len;
    if (!(0 < t && t < T)) return fail(h, "time_step for update need to be 0~%d, got %d", T - 1, t);
    if (A < 0 || S < 0) return fail(h, "negative size");
    if (A == 0 || S == 0) return 0;
    if (!new_pos || !prob_st_0 || !normalizer || !out || (!lazy && !ldjs)) return fail(h, "null tensor pointer");
    if (A > 65535) return fail(h, "too many agents for one update launch (%lld)", (long long)A);
    // one cluster of 8 CTAs per agent (16 -- a non-portable cluster size -- while the agents alone cannot fill the chip:
    // the reference's B = 1 call then spreads over 16 SMs); a CTA keeps exp(base_lp') of its samples in shared memory
    const int per = (T - t) * 3;
    const unsigned long long magic = ((1ull << 40) + per - 1) / per;
    auto kern = lazy ? k_update<true> : k_update<false>;
    // whether the 16-CTA cluster can be scheduled is a property of the device: ask once, fall back to the portable 8
    int csize = A * kUpdCluster * 2 <= (int64_t)h->num_sms ? kUpdClusterBig : kUpdCluster;
    for (;;) {
        const long long chunk = ((S + csize - 1) / csize + 3) / 4 * 4;
        const size_t smem = (size_t)chunk * sizeof(float);
        if (smem + 1024 > (size_t)h->smem_optin) return fail(h, "too many samples per agent for the update kernel (%lld)", (long long)S);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cudaLaunchConfig_t cfg{};
        cfg.stream = (cudaStream_t)stream;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (smem > 48 * 1024) FC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (csize > 8) FC_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3(csize, (unsigned)A, 1);
        // few agents (the reference's B = 1 case): the cluster is all there is, so give each CTA 1024 threads worth of loads in flight
        cfg.blockDim = dim3(A * kUpdCluster < 2 * (int64_t)h->num_sms ? kUpdMaxThreads : kUpdThreads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        if (csize > 8) {
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) {
                (void)cudaGetLastError();
                csize = kUpdCluster;
                continue;
            }
        }
        FC_CUDA(h, cudaLaunchKernelEx(&cfg, kern, new_pos, t, T, prob_st_0, ldjs, normalizer, out, (long long)S, (int)chunk,
                                      h->plans[t].std[0], h->plans[t].std[1], magic));
        break;
    }
    h->launches += 1;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int fc_update(fc_handle* h, const float* new_pos, int t, const float* prob_st_0, const float* ldjs,
                         const float* normalizer, float* prob_st_t, int64_t A, int64_t S, void* stream) {
    return update_impl(h, new_pos, t, prob_st_0, ldjs, normalizer, prob_st_t, A, S, stream, false);
}

extern "C" int fc_update_lazy(fc_handle* h, const float* new_pos, int t, const float* prob_st_0, const float* normalizer,
                              float* base_lp_t, int64_t A, int64_t S, void* stream) {
    return update_impl(h, new_pos, t, prob_st_0, nullptr, normalizer, base_lp_t, A, S, stream, true);
}

extern "C" int fc_grid_density_maps(fc_handle* h, const float* base_pos, const float* cond, const float* cond_traj,
                                    const float* grid_xy, const float* origin, const float* inv_scale, const float* log_jac,
                                    const float* eps, uint64_t seed, int as_prob, float* out, int64_t out_agent_stride,
                                    int64_t out_point_stride, int64_t out_step_stride, int64_t A, int64_t G, int precision,
                                    void* stream) {
    if (!h) return fail(nullptr, "null handle");
    if (!cond_traj || !origin || !inv_scale) return fail(h, "fc_grid_density_maps needs cond_traj, origin and inv_scale");
    QueryMap qm;
    qm.origin = origin; qm.inv_scale = inv_scale; qm.log_jac = log_jac; qm.as_prob = as_prob ? 1 : 0;
    return grid_log_prob_impl(h, base_pos, cond, cond_traj, grid_xy, eps, seed, FC_PREFIX_SHARED, 0, 1, out, out_agent_stride,
                              out_point_stride, out_step_stride, A, G, precision, stream, 0, &qm);
}

extern "C" int fc_to_trajectories(fc_handle* h, const float* state_st, const float* offset, const float* dt, float std0, float std1,
                                  int state_mode, float* out, int64_t A, int64_t S, int Tk, int ncol, void* stream) {
    FC_ENTER(h);
    if (A < 0 || S < 0 || Tk < 0) return fail(h, "negative size");
    if (ncol != 2 && ncol != 3) return fail(h, "ncol must be 2 (positions) or 3 (positions + log p), got %d", ncol);
    if (state_mode != 0 && state_mode != 1) return fail(h, "state_mode must be 0 (state_p) or 1 (state_v); 'state_a' is asserted out in the reference too");
    if (A == 0 || S == 0 || Tk == 0) return 0;
    if (!state_st || !offset || !out) return fail(h, "null tensor pointer");
    const long long rows = (long long)A * S;
    const int per = Tk * ncol;
    const size_t smem = (size_t)kTrajRows * (per | 1) * sizeof(float);
    if (smem > 48 * 1024) FC_CUDA(h, cudaFuncSetAttribute(k_to_trajectories, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if ((int)smem > h->smem_optin) return fail(h, "rows too long for fc_to_trajectories (%d floats)", per);
    const long long blocks = (rows + kTrajRows - 1) / kTrajRows;
    if (blocks > 0x7fffffffLL) return fail(h, "too many rows");
    k_to_trajectories<<<(unsigned)blocks, kTrajRows, smem, (cudaStream_t)stream>>>(state_st, offset, dt, std0, std1, state_mode, out, rows, S, Tk, ncol);
    h->launches++;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int fc_best_of_n(fc_handle* h, const float* cand_st, const float* gt, const float* last_obs, const float* dt,
                            const float* fallback_st, float std0, float std1, int state_mode, float* ade_min, float* fde_min,
                            int32_t* ade_arg, int32_t* fde_arg, int64_t A, int64_t N, int T, void* stream) {
    FC_ENTER(h);
    if (A < 0 || N < 0 || T < 1) return fail(h, "bad size");
    if (state_mode != 0 && state_mode != 1) return fail(h, "state_mode must be 0 (state_p) or 1 (state_v)");
    if (A == 0) return 0;
    if (N == 0) return fail(h, "no candidates");
    if (!cand_st || !gt || !last_obs || !ade_min || !fde_min) return fail(h, "null tensor pointer");
    const long long blocks = (A + 3) / 4;
    k_best_of_n<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(cand_st, gt, last_obs, dt, fallback_st, std0, std1, state_mode, ade_min,
                                                                   fde_min, ade_arg, fde_arg, A, N, T);
    h->launches++;
    FC_CUDA(h, cudaGetLastError());
    return 0;
}

extern "C" int fc_tc_gemm_selftest(const float* A, const float* B, float* D, int N, int K, int device, int mode) {
    std::string msg;
    int rc = fc::tc_selftest(A, B, D, N, K, device, &msg, mode);
    if (rc) g_err = msg;
    return rc;
}

#if FC_T3_TRACE
// development aid (-DFC_T3_TRACE=1 builds only, not part of the shipped ABI): event trace of CTA 0
extern "C" int fc_debug_t3trace(unsigned long long* out, int reset) {   // 20 x 1024 entries, (tag << 48) | clock
    if (out && cudaMemcpyFromSymbol(out, fc::g_t3trace, sizeof(unsigned long long) * 20 * 1024) != cudaSuccess) return 1;
    if (reset) {
        static unsigned long long z[20 * 1024];
        if (cudaMemcpyToSymbol(fc::g_t3trace, z, sizeof z) != cudaSuccess) return 1;
    }
    return 0;
}
#endif
