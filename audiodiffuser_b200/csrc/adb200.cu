// adb200 — C-ABI entry points (include/adb200.h) over the sm_100a kernels in this directory.
// Single translation unit: the kernel headers are included here so that device globals (the
// spin-guard word) exist exactly once.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/adb200.h"
#include "ptx.cuh"
#include "edm_kernels.cuh"
#include "wavenet_f32.cuh"
#include "wavenet_tc.cuh"
#include "wavenet_tc2.cuh"
#include "wavenet_tc3.cuh"
#include "cl_ops.cuh"
#include "cl_conv_tc.cuh"
#include "cl_conv_gn_tc.cuh"
#include "train_kernels.cuh"
#include "train_tc.cuh"

using namespace adb;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(expr)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess)                                                                        \
            return fail(ADB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define REQUIRE(cond, ...)                                 \
    do {                                                   \
        if (!(cond)) return fail(ADB_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// Kernel launches issued by the op-level entry points (adb_edm_*, adb_cl_*, training step); the DiffWave handle keeps its
// own per-class counters (adb_wavenet_timers). bench.py reports their sum as gpu_launches.
static long long g_launches = 0;
#define KL(n) (g_launches += (n))
extern "C" long long adb_launch_count(int reset) {
    const long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" const char* adb_last_error(void) { return g_err.c_str(); }
extern "C" int adb_version(void) { return 100; }

extern "C" int adb_device_check(int device) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(ADB_ERR_UNSUPPORTED, "device %d is sm_%d%d; adb200 contains sm_100a code only", device, prop.major,
                    prop.minor);
    return ADB_OK;
}

// debug: in-kernel cycle counters of the residual-block kernel (see wavenet_tc.cuh); not part of the product ABI
extern "C" int adb_debug_tc_cycles(unsigned long long* out16, int reset) {
    CK(cudaDeviceSynchronize());
    if (out16) CK(cudaMemcpyFromSymbol(out16, g_tc_cycles, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {};
        CK(cudaMemcpyToSymbol(g_tc_cycles, z, sizeof z));
    }
    return ADB_OK;
}

// host view of the multi-layer wavefront launch (wavenet_tc3.cuh, ML = true), for the CPU test of its dependency structure: walks every
// item of a chunk of `bc` samples in launch order and checks that each tile it must wait for belongs to an EARLIER item (a pair only
// ever waits for items dealt before its own: no wait cycle). Returns the number of violations; min_distance = the smallest index
// distance between an item and one of its dependencies, pipelined_ok = whether the launch would use the software-pipelined job order.
extern "C" int adb_debug_ml_order(int bc, int L, int layers, int cycle, int S, int pairs, long long* min_distance, int* n_items,
                                  int* pipelined_ok) {
    const int tiles_per_b = (L + TC_TILE_T - 1) / TC_TILE_T, num_tiles = bc * tiles_per_b;
    if (S > bc) S = bc;
    const int full_sp = bc / S, rem = bc % S;
    const int items_per_sp = layers * ((S * tiles_per_b + 1) / 2);
    const int items = full_sp * items_per_sp + layers * ((rem * tiles_per_b + 1) / 2);
    const int groups_min = rem ? (rem * tiles_per_b + 1) / 2 : (S * tiles_per_b + 1) / 2;
    long long best = 1LL << 60;
    int bad = 0;
    for (int n = 0; n < items; ++n) {
        int layer, tile0, t_end;
        ml_item_decode(n, items_per_sp, S, tiles_per_b, num_tiles, layer, tile0, t_end);
        if (layer == 0) continue;
        const int sp = n / items_per_sp, t_begin = sp * S * tiles_per_b, groups = (t_end - t_begin + 1) >> 1;
        for (int rank = 0; rank < 2; ++rank) {
            const int tile = tile0 + rank;
            if (tile >= t_end) continue;
            int lo, hi;
            ml_dep_range(layer, tile, cycle, tiles_per_b, lo, hi);
            for (int j = lo; j <= hi; ++j) {
                if (j < t_begin || j >= t_end) { ++bad; continue; }            // a dependency outside the sub-pass would never be signalled
                const int dep = sp * items_per_sp + (layer - 1) * groups + ((j - t_begin) >> 1);
                if (dep >= n) ++bad;
                else if (n - dep < best) best = n - dep;
            }
        }
    }
    if (min_distance) *min_distance = best;
    if (n_items) *n_items = items;
    if (pipelined_ok) *pipelined_ok = groups_min > pairs + 17 ? 1 : 0;
    return bad;
}

// host view of the z-stash kernel's job order (wavenet_tc3.cuh: zs_job_at), for the CPU tests of the schedule: writes up to `cap`
// (type, group) pairs and returns how many jobs the sequence has
extern "C" int adb_debug_zs_job_order(int n_groups, int write_h, int pipelined, int* types, int* groups, int cap) {
    const int slots = pipelined ? zs_num_slots<true>(n_groups, write_h) : zs_num_slots<false>(n_groups, write_h);
    int count = 0;
    for (int s = 0; s < slots; ++s) {
        int type = 0, gi = 0;
        const bool ok = pipelined ? zs_job_at<true>(s, n_groups, write_h, type, gi) : zs_job_at<false>(s, n_groups, write_h, type, gi);
        if (!ok) continue;
        if (count < cap && types && groups) { types[count] = type; groups[count] = gi; }
        ++count;
    }
    return count;
}

extern "C" int adb_check_async(void) {
    CK(cudaDeviceSynchronize());
    SpinGuardState st;
    CK(cudaMemcpyFromSymbol(&st, g_spin_guard, sizeof st));
    if (st.abort_flag) {
        SpinGuardState zero = {0, 0, 0, 0};
        cudaMemcpyToSymbol(g_spin_guard, &zero, sizeof zero);
        return fail(ADB_ERR_PIPELINE, "in-kernel barrier wait timed out: site=%u block=%u aux=%u", st.site, st.block,
                    st.aux);
    }
    return ADB_OK;
}

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------
// EDM elementwise entry points
// ------------------------------------------------------------------------------------------------
static EdmArgs edm_args(const float* x, const float* in1, const float* in2, float* o0, float* o1, int64_t n_per,
                        int64_t total) {
    EdmArgs p;
    memset(&p, 0, sizeof p);
    p.x = x; p.in1 = in1; p.in2 = in2; p.out0 = o0; p.out1 = o1;
    p.n_per = n_per; p.total = total;
    return p;
}
static inline float sd2_of(float sigma_data) {
    const double d = static_cast<double>(sigma_data);
    return static_cast<float>(d * d);
}

extern "C" int adb_edm_precond_in(const float* x, const float* sigmas, int sigma_stride, float sigma_data, float* net_in,
                                  float* c_noise, int B, int64_t n_per, void* stream) {
    REQUIRE(x && sigmas && net_in && B > 0 && n_per > 0, "adb_edm_precond_in: bad arguments");
    REQUIRE(sigma_stride == 0 || sigma_stride == 1, "sigma_stride must be 0 or 1");
    EdmArgs p = edm_args(x, nullptr, nullptr, net_in, nullptr, n_per, n_per * B);
    p.sigmas = sigmas; p.sigma_stride = sigma_stride; p.sigma_data = sigma_data; p.sd2 = sd2_of(sigma_data);
    KL(1); CK(edm_launch<OP_SCALE_IN>(p, nullptr, S(stream)));
    if (c_noise) {
        KL(1);
        edm_cnoise_kernel<<<(B + 255) / 256, 256, 0, S(stream)>>>(sigmas, sigma_stride, c_noise, B);
        CK(cudaGetLastError());
    }
    return ADB_OK;
}

// c_in(sigma_b) and c_noise(sigma_b) alone (diffusion.py:235, :240): for callers that fold the input scale into their own first
// kernel (the U-Net's WAVenc re-layout, adb_cl_wavenc_prep) instead of running adb_edm_precond_in over the state
__global__ void precond_prepare_kernel(const float* __restrict__ sigmas, int stride, float sigma_data, float sd2,
                                       float* __restrict__ c_noise, float* __restrict__ c_in, int B);
extern "C" int adb_edm_precond_coef(const float* sigmas, int sigma_stride, float sigma_data, float* c_in, float* c_noise, int B,
                                    void* stream) {
    REQUIRE(sigmas && c_in && c_noise && B > 0, "adb_edm_precond_coef: bad arguments");
    REQUIRE(sigma_stride == 0 || sigma_stride == 1, "sigma_stride must be 0 or 1");
    KL(1); precond_prepare_kernel<<<(B + 255) / 256, 256, 0, S(stream)>>>(sigmas, sigma_stride, sigma_data, sd2_of(sigma_data), c_noise, c_in, B);
    CK(cudaGetLastError());
    return ADB_OK;
}

extern "C" int adb_edm_precond_out(const float* x, const float* f, const float* f_null, float cond_scale,
                                   const float* sigmas, int sigma_stride, float sigma_data, float* out, int B,
                                   int64_t n_per, void* stream) {
    REQUIRE(x && f && sigmas && out && B > 0 && n_per > 0, "adb_edm_precond_out: bad arguments");
    REQUIRE(sigma_stride == 0 || sigma_stride == 1, "sigma_stride must be 0 or 1");
    EdmArgs p = edm_args(x, f, f_null, out, nullptr, n_per, n_per * B);
    p.sigmas = sigmas; p.sigma_stride = sigma_stride; p.sigma_data = sigma_data; p.sd2 = sd2_of(sigma_data);
    p.cond_scale = cond_scale;
    KL(1);
    if (f_null) CK(edm_launch<OP_COMBINE_CFG>(p, nullptr, S(stream)));
    else        CK(edm_launch<OP_COMBINE>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_scale(const float* x, float a, float* out, int64_t n, void* stream) {
    REQUIRE(x && out && n > 0, "adb_edm_scale: bad arguments");
    EdmArgs p = edm_args(x, nullptr, nullptr, out, nullptr, n, n);
    p.a = a;
    KL(1); CK(edm_launch<OP_SCALE>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_axpy(const float* x, const float* e, float a, float* out, int64_t n, void* stream) {
    REQUIRE(x && e && out && n > 0, "adb_edm_axpy: bad arguments");
    EdmArgs p = edm_args(x, e, nullptr, out, nullptr, n, n);
    p.a = a;
    KL(1); CK(edm_launch<OP_AXPY>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_euler(const float* x, const float* den, float sigma, float h, float* d, float* x_next, int64_t n,
                             void* stream) {
    REQUIRE(x && den && d && x_next && n > 0, "adb_edm_euler: bad arguments");
    EdmArgs p = edm_args(x, den, nullptr, d, x_next, n, n);
    p.s0 = sigma; p.h = h;
    KL(1); CK(edm_launch<OP_EULER>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_rk2(const float* x, const float* d, const float* x1, const float* den1, float sigma1, float h,
                           float w0, float w1, float* out, int64_t n, void* stream) {
    REQUIRE(x && d && x1 && den1 && out && n > 0, "adb_edm_rk2: bad arguments");
    EdmArgs p = edm_args(x, den1, x1, out, nullptr, n, n);
    p.s1 = sigma1; p.h = h; p.w0 = w0; p.w1 = w1;
    if (w0 == 0.5f && w1 == 0.5f) {
        p.hh = 0.5f * h;                       // 0.5 * (sigma_next - sigma_hat), sampler_edm.py:367
        KL(1); CK(edm_launch<OP_HEUN>(p, d, S(stream)));
    } else {
        KL(1); CK(edm_launch<OP_RK2>(p, d, S(stream)));
    }
    return ADB_OK;
}

extern "C" int adb_edm_lincomb(const float* x, const float* d, const float* d_old, float a, float e, float c0, float c1,
                               float* out, int64_t n, void* stream) {
    REQUIRE(x && d && out && n > 0, "adb_edm_lincomb: bad arguments");
    EdmArgs p = edm_args(x, d, d_old, out, nullptr, n, n);
    p.a = a; p.s0 = e; p.w0 = c0; p.w1 = c1;
    KL(1);
    if (d_old) CK(edm_launch<OP_LINCOMB3>(p, nullptr, S(stream)));
    else       CK(edm_launch<OP_LINCOMB2>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_clamp(const float* x, float* out, int64_t n, void* stream) {
    REQUIRE(x && out && n > 0, "adb_edm_clamp: bad arguments");
    EdmArgs p = edm_args(x, nullptr, nullptr, out, nullptr, n, n);
    KL(1); CK(edm_launch<OP_CLAMP>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_ema_lerp(float* ema, const float* params, float weight, int64_t n, void* stream) {
    REQUIRE(ema && params && n > 0, "adb_ema_lerp: bad arguments");
    EdmArgs p = edm_args(ema, params, nullptr, ema, nullptr, n, n);
    p.a = weight;
    KL(1); CK(edm_launch<OP_LERP>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_lincomb_n(const float* x, float a, const float* const* terms, const float* coefs, int k, int clamp,
                                 float* out, int64_t n, void* stream) {
    REQUIRE(x && out && n > 0, "adb_edm_lincomb_n: bad arguments");
    REQUIRE(k >= 0 && k <= 4, "adb_edm_lincomb_n: 0..4 terms supported, got %d", k);
    REQUIRE(k == 0 || (terms && coefs), "adb_edm_lincomb_n: terms / coefs missing");
    LincombArgs p{};
    p.x = x; p.a = a; p.out = out; p.n = n; p.clamp = clamp;
    for (int i = 0; i < k; ++i) {
        REQUIRE(terms[i], "adb_edm_lincomb_n: term %d is NULL", i);
        p.m[i] = terms[i]; p.c[i] = coefs[i];
    }
    KL(1);
    switch (k) {
        case 0: CK(lincomb_n_launch<0>(p, S(stream))); break;
        case 1: CK(lincomb_n_launch<1>(p, S(stream))); break;
        case 2: CK(lincomb_n_launch<2>(p, S(stream))); break;
        case 3: CK(lincomb_n_launch<3>(p, S(stream))); break;
        default: CK(lincomb_n_launch<4>(p, S(stream))); break;
    }
    return ADB_OK;
}

extern "C" int adb_pcm16_encode(const float* x, int16_t* pcm, int64_t n, void* stream) {
    REQUIRE(x && pcm && n > 0, "adb_pcm16_encode: bad arguments");
    const bool vec = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(pcm) % 16 == 0);
    const int64_t items = vec ? (n + 7) / 8 : n;
    const int64_t blocks = std::min<int64_t>((items + 255) / 256, 148 * 16);
    KL(1);
    if (vec) pcm16_kernel<true><<<(unsigned)blocks, 256, 0, S(stream)>>>(x, pcm, n);
    else     pcm16_kernel<false><<<(unsigned)blocks, 256, 0, S(stream)>>>(x, pcm, n);
    CK(cudaGetLastError());
    return ADB_OK;
}

extern "C" int adb_edm_churn_rng(const float* x, float* out, float a, float s_noise, uint64_t seed, int step, int64_t sample0,
                                 int B, int64_t n_per, void* stream) {
    REQUIRE(x && out && B > 0 && n_per > 0, "adb_edm_churn_rng: bad arguments");
    REQUIRE(step >= 0 && sample0 >= 0, "adb_edm_churn_rng: step and sample0 must be non-negative");
    REQUIRE((n_per + 3) / 4 <= 0xFFFFFFFFLL, "adb_edm_churn_rng: sample too long for the 32-bit group counter");
    KL(1); CK(edm_churn_rng_launch(x, out, a, s_noise, seed, step, sample0, B, n_per, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_heun_mid(const float* x, const float* f1, float sigma, float sigma_data, float h, float* d, float* x1,
                                int64_t n, void* stream) {
    REQUIRE(x && f1 && d && x1 && n > 0, "adb_edm_heun_mid: bad arguments");
    const PrecondCoef c = precond_coef(sigma, sigma_data, sd2_of(sigma_data));
    EdmArgs p = edm_args(x, f1, nullptr, d, x1, n, n);
    p.s0 = sigma; p.h = h; p.c_skip0 = c.c_skip; p.c_out0 = c.c_out;
    KL(1); CK(edm_launch<OP_MID>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_heun_post(const float* x, const float* d, const float* f2, float sigma1, float sigma_data, float h,
                                 float* x_next, int64_t n, void* stream) {
    REQUIRE(x && d && f2 && x_next && n > 0, "adb_edm_heun_post: bad arguments");
    const PrecondCoef c = precond_coef(sigma1, sigma_data, sd2_of(sigma_data));
    EdmArgs p = edm_args(x, f2, nullptr, x_next, nullptr, n, n);
    p.s1 = sigma1; p.h = h; p.hh = 0.5f * h; p.c_skip1 = c.c_skip; p.c_out1 = c.c_out;
    KL(1); CK(edm_launch<OP_POST>(p, d, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_euler_raw(const float* x, const float* f, float sigma, float sigma_data, float h, float* x_next, int64_t n,
                                 void* stream) {
    REQUIRE(x && f && x_next && n > 0, "adb_edm_euler_raw: bad arguments");
    const PrecondCoef c = precond_coef(sigma, sigma_data, sd2_of(sigma_data));
    EdmArgs p = edm_args(x, f, nullptr, x_next, nullptr, n, n);
    p.s0 = sigma; p.h = h; p.c_skip0 = c.c_skip; p.c_out0 = c.c_out;
    KL(1); CK(edm_launch<OP_EULER_RAW>(p, nullptr, S(stream)));
    return ADB_OK;
}

extern "C" int adb_edm_noise_in(const float* x, const float* noise, const float* sigmas, float sigma_data, float* x_noisy,
                                float* net_in, float* c_noise, int B, int64_t n_per, void* stream) {
    REQUIRE(x && noise && sigmas && x_noisy && net_in && B > 0 && n_per > 0, "adb_edm_noise_in: bad arguments");
    EdmArgs p = edm_args(x, noise, nullptr, x_noisy, net_in, n_per, n_per * B);
    p.sigmas = sigmas; p.sigma_stride = 1; p.sigma_data = sigma_data; p.sd2 = sd2_of(sigma_data);
    KL(1); CK(edm_launch<OP_NOISE_IN>(p, nullptr, S(stream)));
    if (c_noise) {
        KL(1);
        edm_cnoise_kernel<<<(B + 255) / 256, 256, 0, S(stream)>>>(sigmas, 1, c_noise, B);
        CK(cudaGetLastError());
    }
    return ADB_OK;
}

extern "C" int adb_edm_dsm_loss_masked(const float* x, const float* x_noisy, const float* f, const float* sigmas,
                                      float sigma_data, const unsigned char* mask, float* loss, int B, int64_t n_per, void* stream) {
    REQUIRE(x && x_noisy && f && sigmas && loss && B > 0 && n_per > 0, "adb_edm_dsm_loss: bad arguments");
    CK(cudaMemsetAsync(loss, 0, sizeof(float) * B, S(stream)));
    int chunks = static_cast<int>((n_per + 8191) / 8192);
    if (chunks < 1) chunks = 1;
    if (chunks > 64) chunks = 64;
    KL(1);
    edm_dsm_loss_kernel<<<B * chunks, 256, 0, S(stream)>>>(x, x_noisy, f, sigmas, sigma_data, sd2_of(sigma_data), mask, loss,
                                                           n_per, chunks);
    CK(cudaGetLastError());
    return ADB_OK;
}

extern "C" int adb_edm_dsm_loss(const float* x, const float* x_noisy, const float* f, const float* sigmas,
                                float sigma_data, float* loss, int B, int64_t n_per, void* stream) {
    return adb_edm_dsm_loss_masked(x, x_noisy, f, sigmas, sigma_data, nullptr, loss, B, n_per, stream);
}

// ------------------------------------------------------------------------------------------------
// DiffWave handle
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

struct LayerW {
    // views into the flat parameter copy
    const float *v1, *g1, *b1, *wp, *bp, *v2, *g2, *b2;
    // folded fp32 (owned)
    float *w1f, *w2f;
    // transposed / tap-reversed copies for the data-gradient convolutions (training step)
    float *w2T, *w1d;
    // the same three matrices packed for cl_conv_tc_kernel (bf16 training step): forward recompute, dz, dx
    __nv_bfloat16 *w1p, *w2Tp, *w1dp;     // w1p: columns permuted to [128 gate | 128 filter] per 256-wide tile (CL_MODE_GATE_FWD)
    float* b1p;                           // conv bias in the same column order
};

struct TimerPair {
    cudaEvent_t a, b;
    int cls;
};

struct HMap {
    const void* ptr;
    int B, L, kind;
    CUtensorMap map;
};

struct adb_wavenet {
    int C = 0, layers = 0, cycle = 0;
    int64_t n_params = 0;
    float* params = nullptr;          // device copy of the flat vector
    std::vector<LayerW> L;
    const float *v_in, *g_in, *b_in, *fc1w, *fc1b, *fc2w, *fc2b, *v_sp, *g_sp, *b_sp, *w_out, *b_out;
    float* w_in_f = nullptr;          // folded [C]
    float* wsp_f = nullptr;           // folded [C][C] (ci, co)
    float* wspT = nullptr;            // [co][ci] (training step)
    float* d_scale = nullptr;         // g / ||v|| per weight-normed conv: [0] input, [1 + 2l] dilated, [2 + 2l] output, [last] skip
    WnJob* d_jobs = nullptr;
    float* wpT = nullptr;             // [layers][512][C] transposed diffusion projections (refold)
    __nv_bfloat16* wsp_p = nullptr;   // skip projection [C -> C] in cl_conv_tc blocks (training tail on the tensor cores)
    __nv_bfloat16* wspT_p = nullptr;  // its transpose (dskip = ds2 Wsp)
    const float** fold_tab = nullptr; // 6 x (layers*3) operand pointers of the two batched fold GEMMs (refold)
    std::vector<int64_t> counts, dst_off;   // flat-vector pieces (state_dict order) and their offsets in `params`
    const float** d_wp = nullptr;     // device arrays of per-layer pointers
    const float** d_bp = nullptr;
    // tensor-core path (C == 256 only)
    __nv_bfloat16* wtc = nullptr;     // [layers][32][256][64]
    __nv_bfloat16* ws16 = nullptr;    // [layers][4][256][64] bf16 skip half of W2 (training: bf16 z stash, bf16 skip GEMM)
    CUtensorMap tm_ws16;              // its map with 128-row boxes (CTA-pair halves)
    __nv_bfloat16* wsp_tc = nullptr;  // [4][256][64]
    float* mtab = nullptr;            // [512][layers*3*512]  (W1[tap] Wp)^T
    float* cvec = nullptr;            // [layers*3*512]       W1[tap] bp (+ b1 for the centre tap)
    CUtensorMap tm_w, tm_w2, tm_w4, tm_wsp;     // weight maps with box rows 256 / 128 / 64 (cluster 1 / 2 / 4)
    CUtensorMap tm_wsp2;                        // skip-projection weights with 128-row boxes (CTA-pair halves, fused skip GEMM + tail)
    int fuse_tail = 1;                          // z-stash path: tail fused into the skip GEMM's epilogue when one GEMM covers all layers
                                                // (ADB_FUSE_TAIL=0: separate skip GEMM + tail kernels)
    int cluster = 2;                            // CTAs per cluster for the single-CTA residual-block kernel (ADB_TC_CLUSTER)
    // Residual-block kernel of the bf16 sampling path, fixed when the handle is created (ADB_BLOCK_KERNEL):
    //   3 (default) z-stash kernel + skip GEMM (wavenet_tc3.cuh)   2 pair kernel with the fp16 skip stash (wavenet_tc2.cuh)
    //   1 pair kernel, fp32 skip read-modify-write in every block   0 single-CTA kernel (wavenet_tc.cuh)
    // The training forward and the per-block debug entry point always use the pair kernel (they need every block's skip sum).
    int block_kernel = 3;
    int zs_pipe = 1;                            // z-stash kernel: software-pipelined job order (ADB_ZS_PIPE=0: plain per-group order)
    int zs_hi_roles = 0;                        // z-stash kernel: producer / MMA issuer on the highest warp ids (ADB_ZS_HI_ROLES)
    int zs_ml_S = 0;                            // > 0: ONE wavefront launch for all blocks of a chunk, sub-passes of this many samples
                                                // kept L2-resident across the blocks (ADB_ZS_ML; 0 = one launch per block)
    int pair = 1;                               // derived: block_kernel != 0
    int no_stash = 0;                           // derived: block_kernel == 1
    int chunk = 256;                            // samples per pass of the bf16 stack (ADB_CHUNK): bounds the workspace
    int64_t stash_budget = 80LL << 30;          // bytes the z stash may take (ADB_STASH_GB); fewer layers per skip GEMM if it does not fit
    int dbg = 0;                                // ADB_DEBUG builds: ADB_DEBUG_FLAGS at create (2 = in-kernel cycle accounting)
    RefoldLayer* d_refold = nullptr;            // per-block pointer table of refold_layers_kernel
    float* skip_bias = nullptr;                 // [256] sum over layers of the skip half of b2
    const float** d_b2 = nullptr;               // per-layer b2 pointers (device)
    std::vector<HMap> hmaps;
    bool tc_ready = false;
    // timing
    bool timing = false;
    std::vector<TimerPair> timers;
    size_t timers_used = 0;
    double ms_acc[ADB_TIMER_COUNT] = {0, 0, 0, 0};
    int64_t launches[ADB_TIMER_COUNT] = {0, 0, 0, 0};
    std::vector<void*> owned;
};

struct ScopedTimer {
    adb_wavenet* n;
    cudaStream_t s;
    int idx = -1;
    ScopedTimer(adb_wavenet* net, int cls, cudaStream_t st, int nlaunch = 1) : n(net), s(st) {
        n->launches[cls] += nlaunch;
        if (!n->timing) return;
        if (n->timers_used == n->timers.size()) {
            TimerPair t;
            cudaEventCreate(&t.a);
            cudaEventCreate(&t.b);
            n->timers.push_back(t);
        }
        idx = static_cast<int>(n->timers_used++);
        n->timers[idx].cls = cls;
        cudaEventRecord(n->timers[idx].a, s);
    }
    ~ScopedTimer() {
        if (idx >= 0) cudaEventRecord(n->timers[idx].b, s);
    }
};

extern "C" int adb_wavenet_set_timing(adb_wavenet* net, int enabled) {
    REQUIRE(net, "null handle");
    net->timing = enabled != 0;
    net->timers_used = 0;
    for (int i = 0; i < ADB_TIMER_COUNT; ++i) { net->ms_acc[i] = 0; net->launches[i] = 0; }
    return ADB_OK;
}

extern "C" int adb_wavenet_timers(adb_wavenet* net, double* ms_out, int64_t* launches_out) {
    REQUIRE(net, "null handle");
    CK(cudaDeviceSynchronize());
    for (size_t i = 0; i < net->timers_used; ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, net->timers[i].a, net->timers[i].b));
        net->ms_acc[net->timers[i].cls] += ms;
    }
    net->timers_used = 0;
    for (int i = 0; i < ADB_TIMER_COUNT; ++i) {
        if (ms_out) ms_out[i] = net->ms_acc[i];
        if (launches_out) launches_out[i] = net->launches[i];
    }
    return ADB_OK;
}

extern "C" int64_t adb_wavenet_param_count(int C, int layers) {
    int64_t n = 0;
    n += C + 1 + C;                                   // input_projection: bias, g, v[C][1][1]
    n += 512 * 128 + 512 + 512 * 512 + 512;           // fc_t1, fc_t2
    n += static_cast<int64_t>(layers) * (2 * C + 1 + 2LL * C * C * 3 + C * 512LL + C + 2 * C + 1 + 2LL * C * C);
    n += C + 1 + static_cast<int64_t>(C) * C;         // skip_projection
    n += C + 1;                                       // output_projection (ZeroConv1d): weight [1][C][1], bias
    return n;
}

template <typename T>
static cudaError_t dmalloc(adb_wavenet* n, T** p, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
    if (e == cudaSuccess) n->owned.push_back(*p);
    return e;
}

static int make_weight_map(CUtensorMap* map, const void* base, uint64_t rows, uint32_t box_rows = 256) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(ADB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {64, rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ADB_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", static_cast<int>(r));
    return ADB_OK;
}

// 3-D maps over channels-last activations [B][L][256].
//   kind 0: bf16 load map, box [64 ch][128 t][1] — out-of-range time coordinates (negative or >= L) are
//           zero-filled, which is the convolution's zero padding;
//   kind 1: bf16 store map, box [64 ch][32 t][1]; kind 2: fp32 store / reduce map, box [32 ch][32 t][1].
static int make_act_map(CUtensorMap* map, const void* base, int B, int L, int kind) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(ADB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    if (kind == 4) {        // bf16 store map for the pair kernel's residual output: box [32 ch][32 t][1], 64-byte swizzle
        cuuint64_t dims4[3] = {256, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
        cuuint64_t strides4[2] = {512, static_cast<cuuint64_t>(L) * 512};
        cuuint32_t box4[3] = {32, 32, 1};
        cuuint32_t estr4[3] = {1, 1, 1};
        CUresult r4 = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims4, strides4, box4, estr4,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r4 != CUDA_SUCCESS) return fail(ADB_ERR_CUDA, "cuTensorMapEncodeTiled(act kind 4) failed: %d", static_cast<int>(r4));
        return ADB_OK;
    }
    const bool f32 = (kind == 2);
    const cuuint64_t esz = f32 ? 4 : 2;
    cuuint64_t dims[3] = {256, static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {256 * esz, static_cast<cuuint64_t>(L) * 256 * esz};
    cuuint32_t box[3] = {f32 ? 32u : 64u, kind == 0 ? 128u : 32u, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ADB_ERR_CUDA, "cuTensorMapEncodeTiled(act kind %d) failed: %d", kind, static_cast<int>(r));
    return ADB_OK;
}

static int get_act_map(adb_wavenet* n, const void* ptr, int B, int L, int kind, CUtensorMap* out) {
    for (auto& m : n->hmaps)
        if (m.ptr == ptr && m.B == B && m.L == L && m.kind == kind) { *out = m.map; return ADB_OK; }
    if (n->hmaps.size() > 512) n->hmaps.clear();       // training keeps one activation buffer per block: ~150 maps
    HMap m;
    m.ptr = ptr; m.B = B; m.L = L; m.kind = kind;
    int rc = make_act_map(&m.map, ptr, B, L, kind);
    if (rc) return rc;
    n->hmaps.push_back(m);
    *out = m.map;
    return ADB_OK;
}

extern "C" void adb_wavenet_destroy(adb_wavenet* net) {
    if (!net) return;
    for (void* p : net->owned) cudaFree(p);
    for (auto& t : net->timers) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    delete net;
}

// Everything derived from the parameters: weight-norm scales, folded fp32 weights (+ transposed copies for the
// backward), tensor-core packing and the step-embedding fold tables. Kernels only, legacy default stream.
static int refold(adb_wavenet* n) {
    const int C = n->C, layers = n->layers, njobs = 2 * layers + 2;
    wn_scale_kernel<<<njobs, 256>>>(n->d_jobs, n->d_scale);
    scale_copy_kernel<<<1, 256>>>(n->v_in, n->d_scale + 0, n->w_in_f, C);
    pack_conv_f32_kernel<<<grid_for(static_cast<long long>(C) * C), 256>>>(n->v_sp, n->d_scale + njobs - 1, n->wsp_f, C, C, 1);
    transpose_f32_kernel<<<grid_for(static_cast<long long>(C) * C), 256>>>(n->wsp_f, n->wspT, C, C);
    // every block's folded / transposed / packed weights in ONE launch (13 launches per block before)
    refold_layers_kernel<<<dim3(148 * 2, layers), 256>>>(n->d_refold, C);
    CK(cudaGetLastError());
    n->tc_ready = false;
    if (C == TC_C) {
        const long long ldm = static_cast<long long>(layers) * 1536;
        {
            // Step-embedding fold tables for all layers and taps in two launches (operands through pointer tables):
            //   mtab[k][l,tap,co] = sum_ci Wp_l[ci][k] * W1_l[tap][ci][co]
            //   cvec[l,tap,co]    = sum_ci bp_l[ci]    * W1_l[tap][ci][co] (+ b1_l[co] for the centre tap)
            const int nb = layers * 3;
            ConvF32Args a;
            memset(&a, 0, sizeof a);
            a.nb = nb; a.Cin = C; a.Cout = 2 * C; a.taps = 1; a.dil = 1; a.ldw = 2 * C; a.in_scale = 1.f;
            a.w_tab = n->fold_tab + 1 * nb;
            a.in_tab = n->fold_tab + 0 * nb; a.out_tab = const_cast<float* const*>(n->fold_tab + 2 * nb); a.bias_tab = nullptr;
            a.L = 512; a.ldo = ldm;
            CK(conv_cl_f32(a, 0));
            a.in_tab = n->fold_tab + 3 * nb; a.out_tab = const_cast<float* const*>(n->fold_tab + 4 * nb); a.bias_tab = n->fold_tab + 5 * nb;
            a.L = 1; a.ldo = 2 * C;
            CK(conv_cl_f32(a, 0));
        }
        pack_tc_tail_kernel<<<64, 256>>>(n->wsp_f, n->wsp_tc);
        skip_bias_sum_kernel<<<1, 256>>>(n->d_b2, layers, n->skip_bias);
        cl_pack_conv_tc_kernel<<<grid_for(static_cast<long long>(C) * C), 256>>>(n->wsp_f, n->wsp_p, C, C, 1);
        cl_pack_conv_tc_kernel<<<grid_for(static_cast<long long>(C) * C), 256>>>(n->wspT, n->wspT_p, C, C, 1);
        CK(cudaGetLastError());
        n->tc_ready = true;
    }
    return ADB_OK;
}

extern "C" int adb_wavenet_create(adb_wavenet** out, int C, int layers, int cycle, const float* params, int64_t n_params,
                                  int on_device) {
    REQUIRE(out && params, "adb_wavenet_create: null argument");
    REQUIRE(C >= 64 && C % 64 == 0, "residual_channels must be a positive multiple of 64 (got %d)", C);
    REQUIRE(layers >= 1 && cycle >= 1, "bad layer / cycle count");
    REQUIRE(n_params == adb_wavenet_param_count(C, layers), "parameter vector has %lld values, expected %lld",
            static_cast<long long>(n_params), static_cast<long long>(adb_wavenet_param_count(C, layers)));
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int rc = adb_device_check(dev);
    if (rc) return rc;

    adb_wavenet* n = new adb_wavenet();
    n->C = C; n->layers = layers; n->cycle = cycle; n->n_params = n_params;
#define CKN(expr)                                                                                   \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            adb_wavenet_destroy(n);                                                                 \
            return fail(ADB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                           \
    } while (0)
    // ---- carve the flat vector in state_dict order (wavenet.py:153-168) into a copy whose tensors
    //      each start on a 256-byte boundary (the kernels use 128-bit loads on them) ----
    std::vector<int64_t> counts;
    counts.insert(counts.end(), {C, 1, C, 512 * 128, 512, 512 * 512, 512});
    for (int l = 0; l < layers; ++l)
        counts.insert(counts.end(), {2 * C, 1, 2LL * C * C * 3, C * 512LL, C, 2 * C, 1, 2LL * C * C});
    counts.insert(counts.end(), {C, 1, static_cast<int64_t>(C) * C, C, 1});
    std::vector<int64_t> dst_off(counts.size());
    int64_t padded = 0, src_total = 0;
    for (size_t i = 0; i < counts.size(); ++i) {
        dst_off[i] = padded;
        padded += (counts[i] + 63) / 64 * 64;
        src_total += counts[i];
    }
    if (src_total != n_params) { adb_wavenet_destroy(n); return fail(ADB_ERR_INVALID, "internal: parameter carve mismatch"); }
    n->counts = counts;
    n->dst_off = dst_off;
    CKN(dmalloc(n, &n->params, static_cast<size_t>(padded)));
    CKN(cudaMemset(n->params, 0, sizeof(float) * padded));
    {
        int64_t src_off = 0;
        for (size_t i = 0; i < counts.size(); ++i) {
            CKN(cudaMemcpy(n->params + dst_off[i], params + src_off, sizeof(float) * counts[i],
                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
            src_off += counts[i];
        }
    }
    size_t piece = 0;
    auto take = [&](int64_t cnt) {
        const float* r = n->params + dst_off[piece];
        (void)cnt;
        ++piece;
        return r;
    };
    n->b_in = take(C); n->g_in = take(1); n->v_in = take(C);
    n->fc1w = take(512 * 128); n->fc1b = take(512); n->fc2w = take(512 * 512); n->fc2b = take(512);
    n->L.resize(layers);
    for (int l = 0; l < layers; ++l) {
        LayerW& w = n->L[l];
        w.b1 = take(2 * C); w.g1 = take(1); w.v1 = take(2LL * C * C * 3);
        w.wp = take(C * 512LL); w.bp = take(C);
        w.b2 = take(2 * C); w.g2 = take(1); w.v2 = take(2LL * C * C);
    }
    n->b_sp = take(C); n->g_sp = take(1); n->v_sp = take(static_cast<int64_t>(C) * C);
    n->w_out = take(C); n->b_out = take(1);

    // ---- buffers derived from the parameters (filled by refold) ----
    const int njobs = 2 * layers + 2;
    std::vector<WnJob> jobs(njobs);
    jobs[0] = {n->v_in, n->g_in, C};
    for (int l = 0; l < layers; ++l) {
        jobs[1 + 2 * l] = {n->L[l].v1, n->L[l].g1, 2LL * C * C * 3};
        jobs[2 + 2 * l] = {n->L[l].v2, n->L[l].g2, 2LL * C * C};
    }
    jobs[njobs - 1] = {n->v_sp, n->g_sp, static_cast<long long>(C) * C};
    CKN(dmalloc(n, &n->d_jobs, njobs));
    CKN(dmalloc(n, &n->d_scale, njobs));
    CKN(cudaMemcpy(n->d_jobs, jobs.data(), sizeof(WnJob) * njobs, cudaMemcpyHostToDevice));
    CKN(dmalloc(n, &n->w_in_f, C));
    CKN(dmalloc(n, &n->wsp_f, static_cast<size_t>(C) * C));
    CKN(dmalloc(n, &n->wspT, static_cast<size_t>(C) * C));
    CKN(dmalloc(n, &n->wpT, static_cast<size_t>(layers) * 512 * C));
    std::vector<const float*> h_wp(layers), h_bp(layers);
    for (int l = 0; l < layers; ++l) {
        LayerW& w = n->L[l];
        CKN(dmalloc(n, &w.w1f, 3ULL * C * 2 * C));
        CKN(dmalloc(n, &w.w2f, 2ULL * C * C));
        CKN(dmalloc(n, &w.w2T, 2ULL * C * C));
        CKN(dmalloc(n, &w.w1d, 3ULL * 2 * C * C));
        w.w1p = w.w2Tp = w.w1dp = nullptr;
        w.b1p = nullptr;
        if (C == TC_C) {
            CKN(dmalloc(n, &w.b1p, 2ULL * C));
            CKN(dmalloc(n, &w.w1p, 3ULL * C * 2 * C));
            CKN(dmalloc(n, &w.w2Tp, 2ULL * C * C));
            CKN(dmalloc(n, &w.w1dp, 3ULL * 2 * C * C));
        }
        h_wp[l] = w.wp; h_bp[l] = w.bp;
    }
    CKN(dmalloc(n, &n->d_wp, layers));
    CKN(dmalloc(n, &n->d_bp, layers));
    CKN(cudaMemcpy(n->d_wp, h_wp.data(), sizeof(float*) * layers, cudaMemcpyHostToDevice));
    CKN(cudaMemcpy(n->d_bp, h_bp.data(), sizeof(float*) * layers, cudaMemcpyHostToDevice));
    {
        std::vector<const float*> h_b2(layers);
        for (int l = 0; l < layers; ++l) h_b2[l] = n->L[l].b2;
        CKN(dmalloc(n, &n->d_b2, layers));
        CKN(cudaMemcpy(n->d_b2, h_b2.data(), sizeof(float*) * layers, cudaMemcpyHostToDevice));
        CKN(dmalloc(n, &n->skip_bias, 256));
    }
    if (C == TC_C) {
        CKN(dmalloc(n, &n->wtc, static_cast<size_t>(layers) * 32 * 256 * 64));
        CKN(dmalloc(n, &n->ws16, static_cast<size_t>(layers) * 4 * 256 * 64));
        CKN(dmalloc(n, &n->wsp_tc, 4ULL * 256 * 64));
        const long long ldm = static_cast<long long>(layers) * 1536;
        CKN(dmalloc(n, &n->wsp_p, static_cast<size_t>(C) * C));
        CKN(dmalloc(n, &n->wspT_p, static_cast<size_t>(C) * C));
        CKN(dmalloc(n, &n->mtab, 512ULL * ldm));
        CKN(dmalloc(n, &n->cvec, static_cast<size_t>(ldm)));
        {
            const int nb = layers * 3;
            std::vector<const float*> tab(6 * static_cast<size_t>(nb), nullptr);
            for (int l = 0; l < layers; ++l)
                for (int tap = 0; tap < 3; ++tap) {
                    const int e = l * 3 + tap;
                    tab[0 * nb + e] = n->wpT + static_cast<size_t>(l) * 512 * C;                       // mtab: input
                    tab[1 * nb + e] = n->L[l].w1f + static_cast<size_t>(tap) * C * 2 * C;              // both: weights
                    tab[2 * nb + e] = n->mtab + static_cast<size_t>(e) * 512;                          // mtab: output
                    tab[3 * nb + e] = n->L[l].bp;                                                      // cvec: input
                    tab[4 * nb + e] = n->cvec + static_cast<size_t>(e) * 512;                          // cvec: output
                    tab[5 * nb + e] = (tap == 1) ? n->L[l].b1 : nullptr;                               // cvec: bias
                }
            CKN(dmalloc(n, &n->fold_tab, tab.size()));
            CKN(cudaMemcpy(n->fold_tab, tab.data(), sizeof(float*) * tab.size(), cudaMemcpyHostToDevice));
        }
        int rc2 = make_weight_map(&n->tm_w, n->wtc, static_cast<uint64_t>(layers) * 32 * 256);
        if (!rc2) rc2 = make_weight_map(&n->tm_w2, n->wtc, static_cast<uint64_t>(layers) * 32 * 256, 128);
        if (!rc2) rc2 = make_weight_map(&n->tm_w4, n->wtc, static_cast<uint64_t>(layers) * 32 * 256, 64);
        if (!rc2) rc2 = make_weight_map(&n->tm_ws16, n->ws16, static_cast<uint64_t>(layers) * 4 * 256, 128);
        if (!rc2) rc2 = make_weight_map(&n->tm_wsp, n->wsp_tc, 4 * 256);
        if (!rc2) rc2 = make_weight_map(&n->tm_wsp2, n->wsp_tc, 4 * 256, 128);
        {
            const char* e = getenv("ADB_TC_CLUSTER");
            if (e) n->cluster = atoi(e);
            if (n->cluster != 1 && n->cluster != 2 && n->cluster != 4) n->cluster = 2;
            const char* be = getenv("ADB_BLOCK_KERNEL");
            if (be) n->block_kernel = atoi(be);
            if (n->block_kernel < 0 || n->block_kernel > 3) n->block_kernel = 3;
            n->pair = n->block_kernel != 0;
            n->no_stash = n->block_kernel == 1;
            const char* pe = getenv("ADB_ZS_PIPE");
            if (pe) n->zs_pipe = atoi(pe) != 0;
            const char* he = getenv("ADB_ZS_HI_ROLES");
            if (he) n->zs_hi_roles = atoi(he) != 0;
            const char* me = getenv("ADB_ZS_ML");
            if (me && atoi(me) >= 0) n->zs_ml_S = atoi(me);
            // the wavefront needs a block's tiles to complete within ~1.5 tile times of their first load: the software-pipelined job
            // order keeps two groups in flight per pair and would make the pairs wait on each other
            if (n->zs_ml_S > 0 && !pe) n->zs_pipe = 0;
            const char* fe = getenv("ADB_FUSE_TAIL");
            if (fe) n->fuse_tail = atoi(fe) != 0;
            const char* ce = getenv("ADB_CHUNK");
            if (ce && atoi(ce) > 0) n->chunk = atoi(ce);
            const char* se = getenv("ADB_STASH_GB");
            if (se && atof(se) > 0) n->stash_budget = static_cast<int64_t>(atof(se) * (1LL << 30));
#ifdef ADB_DEBUG
            const char* de = getenv("ADB_DEBUG_FLAGS");
            if (de) n->dbg = atoi(de);
#endif
        }
        if (rc2) { adb_wavenet_destroy(n); return rc2; }
        CKN(cudaFuncSetAttribute(wavenet_block_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_BLOCK_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_block_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_TAIL_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_block_zs_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_block_zs_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_block_zs_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_block_zs_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_skip_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SKIP_GEMM_SMEM_BYTES));
        CKN(cudaFuncSetAttribute(wavenet_skip_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SKIP_TAIL_SMEM_BYTES));
    }
    {
        // per-block pointer table of refold_layers_kernel
        std::vector<RefoldLayer> tab(layers);
        for (int l = 0; l < layers; ++l) {
            const LayerW& w = n->L[l];
            RefoldLayer& r = tab[l];
            r.v1 = w.v1; r.v2 = w.v2; r.wp = w.wp; r.b1 = w.b1;
            r.s1 = n->d_scale + 1 + 2 * l; r.s2 = n->d_scale + 2 + 2 * l;
            r.w1f = w.w1f; r.w2f = w.w2f; r.w2T = w.w2T; r.w1d = w.w1d;
            r.wpT = n->wpT + static_cast<size_t>(l) * 512 * C;
            r.b1p = w.b1p;
            r.wtc = n->wtc ? n->wtc + static_cast<size_t>(l) * 32 * 256 * 64 : nullptr;
            r.ws16 = n->ws16 ? n->ws16 + static_cast<size_t>(l) * 4 * 256 * 64 : nullptr;
            r.w1p = w.w1p; r.w2Tp = w.w2Tp; r.w1dp = w.w1dp;
        }
        CKN(dmalloc(n, &n->d_refold, layers));
        CKN(cudaMemcpy(n->d_refold, tab.data(), sizeof(RefoldLayer) * layers, cudaMemcpyHostToDevice));
    }
    rc = refold(n);
    if (rc) { adb_wavenet_destroy(n); return rc; }
    CKN(cudaDeviceSynchronize());
#undef CKN
    *out = n;
    return ADB_OK;
}

// Replace the parameters of an existing handle (same configuration) and rebuild everything derived from them. Used
// after an optimizer step: no allocation, kernels only (legacy default stream).
extern "C" int adb_wavenet_load_params(adb_wavenet* n, const float* params, int64_t n_params, int on_device) {
    REQUIRE(n && params, "adb_wavenet_load_params: null argument");
    REQUIRE(n_params == n->n_params, "parameter vector has %lld values, expected %lld", static_cast<long long>(n_params),
            static_cast<long long>(n->n_params));
    int64_t src_off = 0;
    for (size_t i = 0; i < n->counts.size(); ++i) {
        if (on_device) CK(cudaMemcpyAsync(n->params + n->dst_off[i], params + src_off, sizeof(float) * n->counts[i], cudaMemcpyDeviceToDevice, 0));
        else CK(cudaMemcpy(n->params + n->dst_off[i], params + src_off, sizeof(float) * n->counts[i], cudaMemcpyHostToDevice));
        src_off += n->counts[i];
    }
    return refold(n);
}

// ------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------
struct Workspace {
    // shared
    float *emb, *c_noise, *c_in;        // [B][512], [B], [B]
    float *xhat, *dslope, *fbuf, *xnext;  // sampler state, [B][L] each
    // fp32 path
    float *proj, *hA, *hB, *y, *z, *o, *skip, *s2;
    // bf16 path: the residual stack runs over the batch in passes of Bc samples, so these are sized by Bc
    float* E;                            // [Bc][layers*1536]
    __nv_bfloat16 *hbA, *hbB;
    // z-stash kernel: [G][Bc][L][C] fp16 bits, the gated activations of G consecutive blocks (skip GEMM operand);
    // pair kernel: [Bc][L][C] fp16 bits, an even block's skip term consumed by the next block
    __nv_bfloat16* stash;
    unsigned int* ml_flags;              // multi-layer block launch: [layers][Bc * tiles] completion counters
    int Bc, G;
    int64_t total;
};

// layers per skip GEMM: all of them if the stash fits the budget, otherwise the largest count that does
static int zs_group_layers(const adb_wavenet* n, int Bc, int L) {
    const int64_t per_layer = static_cast<int64_t>(Bc) * L * n->C * 2;
    int64_t g = n->stash_budget / per_layer;
    if (g < 1) g = 1;
    if (g > n->layers) g = n->layers;
    return static_cast<int>(g);
}

// `whole_batch`: the training step and the per-block debug entry point run the stack over all B samples at once with the
// pair kernel (they need every block's fp32 skip sum / saved inputs); sampling runs passes of at most n->chunk samples.
static Workspace carve(const adb_wavenet* n, int B, int L, int precision, void* base, bool whole_batch = false) {
    Workspace w;
    memset(&w, 0, sizeof w);
    uint8_t* p = reinterpret_cast<uint8_t*>(base);
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        uint8_t* r = p ? p + off : nullptr;
        off += (bytes + 1023) / 1024 * 1024;
        return r;
    };
    const int64_t BL = static_cast<int64_t>(B) * L, C = n->C;
    w.emb = reinterpret_cast<float*>(take(B * 512LL * 4));
    w.c_noise = reinterpret_cast<float*>(take(B * 4LL));
    w.c_in = reinterpret_cast<float*>(take(B * 4LL));
    w.xhat = reinterpret_cast<float*>(take(BL * 4));
    w.dslope = reinterpret_cast<float*>(take(BL * 4));
    w.fbuf = reinterpret_cast<float*>(take(BL * 4));
    w.xnext = reinterpret_cast<float*>(take(BL * 4));
    if (precision == ADB_PRECISION_FP32) {
        w.Bc = B; w.G = 1;
        w.skip = reinterpret_cast<float*>(take(BL * C * 4));
        w.proj = reinterpret_cast<float*>(take(static_cast<int64_t>(n->layers) * B * C * 4));
        w.hA = reinterpret_cast<float*>(take(BL * C * 4));
        w.hB = reinterpret_cast<float*>(take(BL * C * 4));
        w.y = reinterpret_cast<float*>(take(BL * 2 * C * 4));
        w.z = reinterpret_cast<float*>(take(BL * C * 4));
        w.o = reinterpret_cast<float*>(take(BL * 2 * C * 4));
        w.s2 = reinterpret_cast<float*>(take(BL * C * 4));
    } else {
        w.Bc = (whole_batch || B < n->chunk) ? B : n->chunk;
        const bool zs = !whole_batch && n->block_kernel == 3;
        w.G = zs ? zs_group_layers(n, w.Bc, L) : 1;
        const int64_t BcL = static_cast<int64_t>(w.Bc) * L;
        w.skip = reinterpret_cast<float*>(take(BcL * C * 4));
        w.E = reinterpret_cast<float*>(take(static_cast<int64_t>(w.Bc) * n->layers * 1536 * 4));
        w.hbA = reinterpret_cast<__nv_bfloat16*>(take(BcL * C * 2));
        w.hbB = reinterpret_cast<__nv_bfloat16*>(take(BcL * C * 2));
        w.stash = reinterpret_cast<__nv_bfloat16*>(take(BcL * C * 2 * w.G));
        if (zs && n->zs_ml_S > 0)
            w.ml_flags = reinterpret_cast<unsigned int*>(take(static_cast<int64_t>(n->layers) * w.Bc * ((L + TC_TILE_T - 1) / TC_TILE_T) * 4));
    }
    w.total = off;
    return w;
}

extern "C" int64_t adb_wavenet_workspace_bytes(const adb_wavenet* net, int B, int L, int precision) {
    if (!net || B <= 0 || L <= 0) return -1;
    return carve(net, B, L, precision, nullptr).total;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void cvt_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = __bfloat162float(in[i]);
}

__global__ void fill2_kernel(float* a, float va, float* b, float vb, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = va; b[i] = vb; }
}

// c_noise[b], c_in[b] from device sigmas (diffusion.py:235, :240)
__global__ void precond_prepare_kernel(const float* __restrict__ sigmas, int stride, float sigma_data, float sd2,
                                       float* __restrict__ c_noise, float* __restrict__ c_in, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) {
        const PrecondCoef c = precond_coef(sigmas[b * stride], sigma_data, sd2);
        c_noise[b] = c.c_noise;
        c_in[b] = c.c_in;
    }
}

static int check_forward_args(adb_wavenet* n, int B, int L, int precision, void* ws, int64_t ws_bytes) {
    REQUIRE(n, "null handle");
    REQUIRE(B > 0 && L > 0, "bad shape B=%d L=%d", B, L);
    REQUIRE(precision == ADB_PRECISION_FP32 || precision == ADB_PRECISION_BF16, "unknown precision %d", precision);
    if (precision == ADB_PRECISION_BF16 && !n->tc_ready)
        return fail(ADB_ERR_UNSUPPORTED, "the bf16 tensor-core path is built for residual_channels == 256 (got %d)", n->C);
    const int64_t need = adb_wavenet_workspace_bytes(n, B, L, precision);
    REQUIRE(ws && ws_bytes >= need, "workspace too small: %lld < %lld bytes", static_cast<long long>(ws_bytes),
            static_cast<long long>(need));
    REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024-byte aligned");
    return ADB_OK;
}

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// bf16 sampling path with the z-stash kernel (wavenet_tc3.cuh): the batch is processed in passes of w.Bc samples; every
// block leaves its gated activations in the stash and one skip GEMM per w.G blocks contracts them into the fp32 skip sum.
// `emb` ([B][512], embed_mlp_kernel) is already in w.emb.
// Training forward (h_save / y_save non-null, one pass over the whole batch): block l reads h_save slot l and writes slot l + 1, and
// leaves its pre-gate activations in y_save slot l; the stash then holds every block's z (w.G == layers).
// `uniform`: every sample has the same noise level (the sampler's scalar sigma): the step embedding and the E constants are computed
// for ONE sample (w.emb row 0) and shared (the per-sample E GEMM was 0.8 % of an evaluation).
static int forward_bf16_zs(adb_wavenet* n, const float* x, const float* in_scale, int in_scale_stride, float* out, int B, int L,
                           const Workspace& w, cudaStream_t st, __nv_bfloat16* h_save = nullptr, __nv_bfloat16* y_save = nullptr,
                           bool uniform = false) {
    REQUIRE(!h_save || (w.Bc == B && w.G == n->layers), "internal: the training forward needs a whole-batch workspace with a full stash");
    const int C = n->C, layers = n->layers;
    const int num_sms = sm_count();
    const int tiles_per_b = (L + TC_TILE_T - 1) / TC_TILE_T;
    // the training step needs the fp32 skip sum itself (tail backward), so it keeps the separate kernels
    const bool fused_tail = n->fuse_tail && w.G == layers && !h_save;
    for (int b0 = 0; b0 < B; b0 += w.Bc) {
        const int bc = (B - b0 < w.Bc) ? B - b0 : w.Bc;
        const long long BL = static_cast<long long>(bc) * L;
        {
            ScopedTimer t(n, ADB_TIMER_AUX, st, 2);
            ConvF32Args a;
            memset(&a, 0, sizeof a);
            a.in = uniform ? w.emb : w.emb + static_cast<long long>(b0) * 512; a.w = n->mtab; a.bias = n->cvec; a.out = w.E;
            a.nb = 1; a.L = uniform ? 1 : bc; a.Cin = 512; a.Cout = layers * 1536; a.taps = 1; a.dil = 1;
            a.ldw = static_cast<long long>(layers) * 1536; a.ldo = a.ldw; a.in_scale = 1.f;
            CK(conv_cl_f32(a, st));
            in_proj_kernel<true><<<grid_for(BL * (C / 8)), 256, 0, st>>>(
                x + static_cast<long long>(b0) * L, in_scale ? in_scale + static_cast<long long>(b0) * in_scale_stride : nullptr,
                in_scale_stride, n->w_in_f, n->b_in, h_save ? h_save : w.hbA, bc, L, C);
            CK(cudaGetLastError());
        }
        const int num_tiles = tiles_per_b * bc;
        int pairs = (num_tiles + 1) / 2;
        if (pairs > num_sms / 2) pairs = num_sms / 2;
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(TC_THREADS); lc.stream = st;
        cudaLaunchAttribute la[1];
        la[0].id = cudaLaunchAttributeClusterDimension;
        la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
        lc.attrs = la; lc.numAttrs = 1;
        CUtensorMap m_zst, m_zld, m_skip;
        int rc = get_act_map(n, w.stash, w.G * bc, L, 1, &m_zst);
        if (!rc) rc = get_act_map(n, w.stash, w.G * bc, L, 0, &m_zld);
        if (!rc) rc = get_act_map(n, w.skip, bc, L, 2, &m_skip);
        if (rc) return rc;
        __nv_bfloat16 *hin = h_save ? h_save : w.hbA, *hout = (h_save && layers > 1) ? h_save + BL * C : w.hbB;
        const bool ml = n->zs_ml_S > 0 && !h_save && w.G == layers && w.ml_flags != nullptr && layers > 1;
        if (ml) {
            // all blocks of this chunk in ONE wavefront launch (wavenet_tc3.cuh, ML = true)
            ScopedTimer t(n, ADB_TIMER_CONV, st, layers);
            CUtensorMap m_hA, m_hB, m_hAo, m_hBo;
            rc = get_act_map(n, w.hbA, bc, L, 0, &m_hA);
            if (!rc) rc = get_act_map(n, w.hbB, bc, L, 0, &m_hB);
            if (!rc) rc = get_act_map(n, w.hbA, bc, L, 4, &m_hAo);
            if (!rc) rc = get_act_map(n, w.hbB, bc, L, 4, &m_hBo);
            if (rc) return rc;
            CK(cudaMemsetAsync(w.ml_flags, 0, sizeof(unsigned int) * layers * num_tiles, st));
            BlockZsParams bp;
            memset(&bp, 0, sizeof bp);
            bp.E = w.E; bp.b2 = nullptr; bp.h_in = w.hbA; bp.h_in2 = w.hbB; bp.h_out_dbg = nullptr; bp.y_out = nullptr;
            bp.B = bc; bp.L = L; bp.layer = 0; bp.layers = layers; bp.dil = 1;
            bp.tiles_per_b = tiles_per_b; bp.num_tiles = num_tiles;
            bp.write_h = 1;                       // uniform job structure; the last block's h' stores are skipped in the kernel
            bp.zrow0 = 0; bp.hi_roles = n->zs_hi_roles; bp.dbg = n->dbg;
            bp.b2_tab = n->d_b2; bp.ml_flags = w.ml_flags; bp.cycle = n->cycle; bp.e_uniform = uniform ? 1 : 0;
            bp.ml_S = n->zs_ml_S < bc ? n->zs_ml_S : bc;
            const int full_sp = bc / bp.ml_S, rem = bc % bp.ml_S;
            bp.ml_items_per_sp = layers * ((bp.ml_S * tiles_per_b + 1) / 2);
            bp.ml_items = full_sp * bp.ml_items_per_sp + layers * ((rem * tiles_per_b + 1) / 2);
            lc.dynamicSmemBytes = TC3_SMEM_BYTES;
            // even blocks read ping (hbA) and write pong (hbB), odd blocks the other way round
            // The pipelined order issues G1a of a pair's NEXT item before G2r of the current one: the next item's dependency wait then
            // sits in front of the current item's last job. That is only safe while the next item (pairs items later) cannot depend,
            // directly or transitively, on the current one, i.e. while a layer of a sub-pass has more groups than pairs + the widest
            // neighbourhood (17 tiles); otherwise the plain order, whose waits only ever precede an item's own jobs.
            // (the ragged last sub-pass counts: it is the one with the fewest groups per layer)
            const int groups_min = rem ? (rem * tiles_per_b + 1) / 2 : (bp.ml_S * tiles_per_b + 1) / 2;
            const bool use_pipe = n->zs_pipe && groups_min > pairs + 17;
            if (use_pipe) CK(cudaLaunchKernelEx(&lc, wavenet_block_zs_kernel<true, true>, m_hA, n->tm_w2, m_hBo, m_zst, m_hB, m_hAo, bp));
            else          CK(cudaLaunchKernelEx(&lc, wavenet_block_zs_kernel<false, true>, m_hA, n->tm_w2, m_hBo, m_zst, m_hB, m_hAo, bp));
        }
        for (int l = 0; l < layers; ++l) {
            const int slot = l % w.G;
            if (!ml) {
                ScopedTimer t(n, ADB_TIMER_CONV, st);
                CUtensorMap m_h, m_hout;
                rc = get_act_map(n, hin, bc, L, 0, &m_h);
                if (!rc) rc = get_act_map(n, hout, bc, L, 4, &m_hout);
                if (rc) return rc;
                BlockZsParams bp;
                memset(&bp, 0, sizeof bp);
                bp.E = w.E; bp.b2 = n->L[l].b2; bp.h_in = hin; bp.h_out_dbg = hout;
                bp.y_out = y_save ? y_save + static_cast<long long>(l) * BL * 2 * C : nullptr;
                // training: the stash slot of this block receives z as bf16 (what the backward's weight-gradient GEMM takes; the fp16
                // stash of the sampling path would need a re-typing pass per block) and the skip GEMM below runs in bf16
                bp.zb_out = y_save ? w.stash + static_cast<long long>(slot) * BL * C : nullptr;
                bp.B = bc; bp.L = L; bp.layer = l; bp.layers = layers; bp.dil = 1 << (l % n->cycle);
                bp.tiles_per_b = tiles_per_b; bp.num_tiles = num_tiles;
                bp.write_h = (l + 1 < layers) ? 1 : 0;
                bp.zrow0 = slot * bc;
                bp.hi_roles = n->zs_hi_roles;
                bp.e_uniform = uniform ? 1 : 0;
                bp.dbg = n->dbg;
                lc.dynamicSmemBytes = TC3_SMEM_BYTES;
                if (n->zs_pipe) CK(cudaLaunchKernelEx(&lc, wavenet_block_zs_kernel<true, false>, m_h, n->tm_w2, m_hout, m_zst, m_h, m_hout, bp));
                else            CK(cudaLaunchKernelEx(&lc, wavenet_block_zs_kernel<false, false>, m_h, n->tm_w2, m_hout, m_zst, m_h, m_hout, bp));
            }
            if (h_save) { hin = hout; hout = (l + 2 < layers) ? hin + BL * C : w.hbB; }
            else { __nv_bfloat16* tmp = hin; hin = hout; hout = tmp; }
            if (fused_tail) {
                if (l + 1 == layers) {
                    // one GEMM over all layers with the tail in its epilogue: the fp32 skip sum never reaches HBM
                    ScopedTimer t(n, ADB_TIMER_SKIP, st);
                    SkipTailParams sp;
                    sp.bias = n->skip_bias; sp.b_sp = n->b_sp; sp.w_out = n->w_out; sp.b_out = n->b_out;
                    sp.out = out + static_cast<long long>(b0) * L;
                    sp.scale = static_cast<float>(sqrt(1.0 / layers));
                    sp.B = bc; sp.L = L; sp.tiles_per_b = tiles_per_b; sp.num_tiles = num_tiles; sp.G = layers;
                    lc.dynamicSmemBytes = SKIP_TAIL_SMEM_BYTES;
                    CK(cudaLaunchKernelEx(&lc, wavenet_skip_tail_kernel, m_zld, n->tm_w2, n->tm_wsp2, sp));
                }
            } else if (slot == w.G - 1 || l + 1 == layers) {
                ScopedTimer t(n, ADB_TIMER_SKIP, st);
                SkipGemmParams sp;
                sp.G = slot + 1; sp.layer0 = l - slot;
                sp.accumulate = sp.layer0 > 0 ? 1 : 0;
                sp.bias = sp.accumulate ? nullptr : n->skip_bias;
                sp.B = bc; sp.L = L; sp.tiles_per_b = tiles_per_b; sp.num_tiles = num_tiles;
                sp.w_bf16 = y_save ? 1 : 0;
                lc.dynamicSmemBytes = SKIP_GEMM_SMEM_BYTES;
                CK(cudaLaunchKernelEx(&lc, wavenet_skip_gemm_kernel, m_zld, y_save ? n->tm_ws16 : n->tm_w2, m_skip, sp));
            }
        }
        if (!fused_tail) {
            ScopedTimer t(n, ADB_TIMER_TAIL, st);
            TailTcParams tp;
            tp.skip = w.skip; tp.b_sp = n->b_sp; tp.w_out = n->w_out; tp.b_out = n->b_out;
            tp.out = out + static_cast<long long>(b0) * L;
            tp.scale = static_cast<float>(sqrt(1.0 / layers));
            tp.B = bc; tp.L = L; tp.tiles_per_b = tiles_per_b; tp.num_tiles = num_tiles;
            wavenet_tail_tc_kernel<<<num_tiles < num_sms ? num_tiles : num_sms, 256, TC_TAIL_SMEM_BYTES, st>>>(n->tm_wsp, tp);
            CK(cudaGetLastError());
        }
    }
    return ADB_OK;
}

static int forward_impl(adb_wavenet* n, const float* x, const float* c_noise, const float* in_scale, int in_scale_stride,
                        float* out, int B, int L, int precision, const Workspace& w, float* dump_h, float* dump_skip,
                        int dump_layers, cudaStream_t st, __nv_bfloat16* h_save = nullptr, __nv_bfloat16* y_save = nullptr,
                        bool uniform_sigma = false) {
    // h_save (bf16 path, training): [layers][B][L][C]; block l reads slot l and writes slot l + 1 instead of ping-ponging.
    // y_save (z-stash path, training): [layers][B][L][2C], every block's pre-gate activations, so the backward recomputes nothing.
    const int C = n->C, layers = n->layers;
    const long long BL = static_cast<long long>(B) * L;
    {
        ScopedTimer t(n, ADB_TIMER_AUX, st);
        // one embedding for the whole batch when the z-stash path runs with a scalar sigma (see forward_bf16_zs)
        const bool zs_uniform = uniform_sigma && precision == ADB_PRECISION_BF16 && dump_layers == 0 && n->block_kernel == 3 && !h_save;
        embed_mlp_kernel<<<zs_uniform ? 1 : B, 512, 0, st>>>(c_noise, n->fc1w, n->fc1b, n->fc2w, n->fc2b, w.emb);
        CK(cudaGetLastError());
    }
    if (precision == ADB_PRECISION_FP32) {
        {
            ScopedTimer t(n, ADB_TIMER_AUX, st, 2);
            embed_proj_kernel<<<dim3(B, layers), 256, 0, st>>>(w.emb, n->d_wp, n->d_bp, w.proj, B, C);
            in_proj_kernel<false><<<grid_for(BL * (C / 8)), 256, 0, st>>>(x, in_scale, in_scale_stride, n->w_in_f, n->b_in,
                                                                          w.hA, B, L, C);
            CK(cudaGetLastError());
        }
        float *hin = w.hA, *hout = w.hB;
        for (int l = 0; l < layers; ++l) {
            ScopedTimer t(n, ADB_TIMER_CONV, st, 4);
            const LayerW& lw = n->L[l];
            ConvF32Args a;
            memset(&a, 0, sizeof a);
            a.in = hin; a.padd = w.proj + static_cast<long long>(l) * B * C; a.w = lw.w1f; a.bias = lw.b1; a.out = w.y;
            a.nb = B; a.L = L; a.Cin = C; a.Cout = 2 * C; a.taps = 3; a.dil = 1 << (l % n->cycle);
            a.ldw = 2 * C; a.ldo = 2 * C; a.in_scale = 1.f;
            CK(conv_cl_f32(a, st));
            gate_f32_kernel<<<grid_for(BL * C), 256, 0, st>>>(w.y, w.z, BL, C);
            memset(&a, 0, sizeof a);
            a.in = w.z; a.w = lw.w2f; a.bias = lw.b2; a.out = w.o;
            a.nb = B; a.L = L; a.Cin = C; a.Cout = 2 * C; a.taps = 1; a.dil = 1; a.ldw = 2 * C; a.ldo = 2 * C; a.in_scale = 1.f;
            CK(conv_cl_f32(a, st));
            res_skip_f32_kernel<<<grid_for(BL * C), 256, 0, st>>>(hin, w.o, hout, w.skip, BL, C, l == 0 ? 1 : 0);
            CK(cudaGetLastError());
            if (l < dump_layers) {
                CK(cudaMemcpyAsync(dump_h + l * BL * C, hout, sizeof(float) * BL * C, cudaMemcpyDeviceToDevice, st));
                CK(cudaMemcpyAsync(dump_skip + l * BL * C, w.skip, sizeof(float) * BL * C, cudaMemcpyDeviceToDevice, st));
            }
            float* tmp = hin; hin = hout; hout = tmp;
        }
        {
            ScopedTimer t(n, ADB_TIMER_TAIL, st, 2);
            ConvF32Args a;
            memset(&a, 0, sizeof a);
            a.in = w.skip; a.w = n->wsp_f; a.bias = n->b_sp; a.out = w.s2;
            a.nb = B; a.L = L; a.Cin = C; a.Cout = C; a.taps = 1; a.dil = 1; a.ldw = C; a.ldo = C;
            a.in_scale = static_cast<float>(sqrt(1.0 / layers));      // wavenet.py:151
            a.relu = 1;                                               // wavenet.py:178
            CK(conv_cl_f32(a, st));
            out_proj_f32_kernel<<<static_cast<unsigned>((BL * 32 + 255) / 256), 256, 0, st>>>(w.s2, n->w_out, n->b_out, out, BL, C);
            CK(cudaGetLastError());
        }
        return ADB_OK;
    }

    // ---------------- bf16 tensor-core path ----------------
    if (dump_layers == 0 && n->block_kernel == 3 && (!h_save || y_save))
        return forward_bf16_zs(n, x, in_scale, in_scale_stride, out, B, L, w, st, h_save, y_save, uniform_sigma && !h_save);
    REQUIRE(w.Bc == B, "internal: the pair-kernel path needs a whole-batch workspace (B=%d, pass=%d)", B, w.Bc);
    {
        ScopedTimer t(n, ADB_TIMER_AUX, st, 2);
        ConvF32Args a;
        memset(&a, 0, sizeof a);
        a.in = w.emb; a.w = n->mtab; a.bias = n->cvec; a.out = w.E;
        a.nb = 1; a.L = B; a.Cin = 512; a.Cout = layers * 1536; a.taps = 1; a.dil = 1;
        a.ldw = static_cast<long long>(layers) * 1536; a.ldo = a.ldw; a.in_scale = 1.f;
        CK(conv_cl_f32(a, st));
        in_proj_kernel<true><<<grid_for(BL * (C / 8)), 256, 0, st>>>(x, in_scale, in_scale_stride, n->w_in_f, n->b_in,
                                                                     h_save ? h_save : w.hbA,
                                                                     B, L, C);
        CK(cudaGetLastError());
    }
    const int num_sms = sm_count();
    const int tiles_per_b = (L + TC_TILE_T - 1) / TC_TILE_T;
    const int num_tiles = tiles_per_b * B;
    const int grid = num_tiles < num_sms ? num_tiles : num_sms;
    // single-CTA kernel: persistent clusters of `cl` CTAs (weights multicast inside a cluster)
    const int cl = n->cluster;
    int max_clusters = num_sms / cl;
    if (cl > 1 && !n->pair) {
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(num_sms / cl * cl); qc.blockDim = dim3(TC_THREADS); qc.dynamicSmemBytes = TC_BLOCK_SMEM_BYTES;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = cl; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        qc.attrs = qa; qc.numAttrs = 1;
        int mc = 0;
        if (cudaOccupancyMaxActiveClusters(&mc, wavenet_block_tc_kernel, &qc) == cudaSuccess && mc > 0) max_clusters = mc;
    }
    const int groups = (num_tiles + cl - 1) / cl;
    const int grid_block = (groups < max_clusters ? groups : max_clusters) * cl;
    __nv_bfloat16 *hin = h_save ? h_save : w.hbA, *hout = (h_save && layers > 1) ? h_save + BL * C : w.hbB;
    // pair kernel: even layers stash their skip term as fp16, odd layers add it and touch the fp32 sum once for both
    // (off when per-layer skip sums are dumped for the debug entry point, or with ADB_BLOCK_KERNEL=1)
    const bool use_stash = n->pair && dump_layers == 0 && !n->no_stash;
    bool skip_written = false;
    for (int l = 0; l < layers; ++l) {
        ScopedTimer t(n, ADB_TIMER_CONV, st);
        CUtensorMap m_h, m_hout, m_skip;
        int rc = get_act_map(n, hin, B, L, 0, &m_h);
        if (!rc) rc = get_act_map(n, hout, B, L, 1, &m_hout);
        if (!rc) rc = get_act_map(n, w.skip, B, L, 2, &m_skip);
        if (rc) return rc;
        BlockTcParams bp;
        bp.E = w.E; bp.b2 = n->L[l].b2; bp.h_in = hin; bp.h_out = hout; bp.skip = w.skip;
        bp.B = B; bp.L = L; bp.layer = l; bp.layers = layers; bp.dil = 1 << (l % n->cycle);
        bp.tiles_per_b = tiles_per_b; bp.num_tiles = num_tiles;
        bp.first_layer = (l == 0); bp.write_h = (l + 1 < layers) || (l < dump_layers);
        bp.add_stash = (use_stash && (l & 1)) ? 1 : 0;
        if (use_stash && !(l & 1) && l + 1 < layers) bp.skip_mode = 2;
        else { bp.skip_mode = skip_written ? 0 : 1; skip_written = true; }
        bp.dbg = n->dbg;
        bp.cluster = cl;
        if (n->pair) {
            int pairs = (num_tiles + 1) / 2;
            if (pairs > num_sms / 2) pairs = num_sms / 2;
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(TC_THREADS); lc.dynamicSmemBytes = TC2_SMEM_BYTES; lc.stream = st;
            cudaLaunchAttribute la[1];
            la[0].id = cudaLaunchAttributeClusterDimension;
            la[0].val.clusterDim.x = 2; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
            lc.attrs = la; lc.numAttrs = 1;
            bp.cluster = 2;
            CUtensorMap m_hout4, m_stash_ld, m_stash_st;
            rc = get_act_map(n, hout, B, L, 4, &m_hout4);
            if (!rc) rc = get_act_map(n, w.stash, B, L, 0, &m_stash_ld);
            if (!rc) rc = get_act_map(n, w.stash, B, L, 4, &m_stash_st);
            if (rc) return rc;
            CK(cudaLaunchKernelEx(&lc, wavenet_block_pair_kernel, m_h, n->tm_w2, m_skip, m_hout4, m_stash_ld, m_stash_st, bp));
        } else {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(grid_block); lc.blockDim = dim3(TC_THREADS); lc.dynamicSmemBytes = TC_BLOCK_SMEM_BYTES; lc.stream = st;
            cudaLaunchAttribute la[1];
            la[0].id = cudaLaunchAttributeClusterDimension;
            la[0].val.clusterDim.x = cl; la[0].val.clusterDim.y = 1; la[0].val.clusterDim.z = 1;
            lc.attrs = la; lc.numAttrs = 1;
            const CUtensorMap& m_w = cl == 1 ? n->tm_w : (cl == 2 ? n->tm_w2 : n->tm_w4);
            CK(cudaLaunchKernelEx(&lc, wavenet_block_tc_kernel, m_h, m_w, m_hout, m_skip, bp));
        }
        if (l < dump_layers) {
            cvt_bf16_f32_kernel<<<grid_for(BL * C), 256, 0, st>>>(hout, dump_h + l * BL * C, BL * C);
            CK(cudaMemcpyAsync(dump_skip + l * BL * C, w.skip, sizeof(float) * BL * C, cudaMemcpyDeviceToDevice, st));
        }
        if (h_save) { hin = hout; hout = (l + 2 < layers) ? hin + BL * C : w.hbB; }
        else { __nv_bfloat16* tmp = hin; hin = hout; hout = tmp; }
    }
    {
        ScopedTimer t(n, ADB_TIMER_TAIL, st);
        TailTcParams tp;
        tp.skip = w.skip; tp.b_sp = n->b_sp; tp.w_out = n->w_out; tp.b_out = n->b_out; tp.out = out;
        tp.scale = static_cast<float>(sqrt(1.0 / layers));
        tp.B = B; tp.L = L; tp.tiles_per_b = tiles_per_b; tp.num_tiles = num_tiles;
        wavenet_tail_tc_kernel<<<grid, 256, TC_TAIL_SMEM_BYTES, st>>>(n->tm_wsp, tp);
        CK(cudaGetLastError());
    }
    return ADB_OK;
}

extern "C" int adb_wavenet_forward_debug(adb_wavenet* n, const float* x, const float* c_noise, const float* in_scale,
                                         int in_scale_stride, float* out, int B, int L, int precision, void* ws,
                                         int64_t ws_bytes, float* dump_h, float* dump_skip, int dump_layers, void* stream) {
    int rc = check_forward_args(n, B, L, precision, ws, ws_bytes);
    if (rc) return rc;
    REQUIRE(x && c_noise && out, "null tensor argument");
    REQUIRE(dump_layers >= 0 && dump_layers <= n->layers, "bad dump_layers");
    REQUIRE(dump_layers == 0 || (dump_h && dump_skip), "dump buffers missing");
    const Workspace w = carve(n, B, L, precision, ws, dump_layers > 0);
    REQUIRE(ws_bytes >= w.total, "workspace too small for the per-block dump: %lld < %lld bytes",
            static_cast<long long>(ws_bytes), static_cast<long long>(w.total));
    return forward_impl(n, x, c_noise, in_scale, in_scale_stride, out, B, L, precision, w, dump_h, dump_skip, dump_layers,
                        S(stream));
}

extern "C" int adb_wavenet_forward(adb_wavenet* n, const float* x, const float* c_noise, const float* in_scale,
                                   int in_scale_stride, float* out, int B, int L, int precision, void* ws, int64_t ws_bytes,
                                   void* stream) {
    return adb_wavenet_forward_debug(n, x, c_noise, in_scale, in_scale_stride, out, B, L, precision, ws, ws_bytes, nullptr,
                                     nullptr, 0, stream);
}

extern "C" int adb_wavenet_denoise(adb_wavenet* n, const float* x, const float* sigmas, int sigma_stride, float sigma_data,
                                   float* out, int B, int L, int precision, void* ws, int64_t ws_bytes, void* stream) {
    int rc = check_forward_args(n, B, L, precision, ws, ws_bytes);
    if (rc) return rc;
    REQUIRE(x && sigmas && out, "null tensor argument");
    REQUIRE(sigma_stride == 0 || sigma_stride == 1, "sigma_stride must be 0 or 1");
    const Workspace w = carve(n, B, L, precision, ws);
    cudaStream_t st = S(stream);
    precond_prepare_kernel<<<(B + 255) / 256, 256, 0, st>>>(sigmas, sigma_stride, sigma_data, sd2_of(sigma_data), w.c_noise,
                                                            w.c_in, B);
    CK(cudaGetLastError());
    rc = forward_impl(n, x, w.c_noise, w.c_in, 1, w.fbuf, B, L, precision, w, nullptr, nullptr, 0, st, nullptr, nullptr, sigma_stride == 0);
    if (rc) return rc;
    return adb_edm_precond_out(x, w.fbuf, nullptr, 1.0f, sigmas, sigma_stride, sigma_data, out, B, L, stream);
}

// ------------------------------------------------------------------------------------------------
// device-resident EDM sampling trajectory
// ------------------------------------------------------------------------------------------------
static int net_eval(adb_wavenet* n, const float* x, float sigma, float sigma_data, float* f_out, int B, int L, int precision,
                    const Workspace& w, cudaStream_t st) {
    const PrecondCoef c = precond_coef(sigma, sigma_data, sd2_of(sigma_data));
    fill2_kernel<<<(B + 255) / 256, 256, 0, st>>>(w.c_noise, c.c_noise, w.c_in, c.c_in, B);
    CK(cudaGetLastError());
    n->launches[ADB_TIMER_AUX]++;
    return forward_impl(n, x, w.c_noise, w.c_in, 1, f_out, B, L, precision, w, nullptr, nullptr, 0, st, nullptr, nullptr, true);
}

extern "C" int adb_wavenet_sample_edm_seeded(adb_wavenet* n, const float* noise, const float* sigmas_host, int n_sigmas,
                                             int num_steps, float sigma_data, float s_tmin, float s_tmax, float s_churn,
                                             float s_noise, int use_heun, float alpha, const float* eps, uint64_t churn_seed,
                                             int64_t sample0, float* x_out, int B, int L, int precision, void* ws,
                                             int64_t ws_bytes, int* nfe_out, void* stream) {
    int rc = check_forward_args(n, B, L, precision, ws, ws_bytes);
    if (rc) return rc;
    REQUIRE(noise && sigmas_host && x_out, "null tensor argument");
    REQUIRE(num_steps >= 1, "num_steps must be >= 1");
    const bool alpha_mode = alpha > 0.f;
    REQUIRE(n_sigmas >= num_steps, "schedule has %d sigmas, sampler needs %d", n_sigmas, num_steps);
    const Workspace w = carve(n, B, L, precision, ws);
    cudaStream_t st = S(stream);
    const long long N = static_cast<long long>(B) * L;
    const float sd2 = sd2_of(sigma_data);
    int nfe = 0;

    // x = sigmas[0] * noise   (sampler_edm.py:380 / :291)
    float* x = w.xhat;
    {
        ScopedTimer t(n, ADB_TIMER_STEP, st);
        rc = adb_edm_scale(noise, sigmas_host[0], x, N, stream);
        if (rc) return rc;
    }

    if (!alpha_mode) {
        // gamma_i = min(s_churn / N, sqrt(2) - 1) where s_tmin <= sigma_i <= s_tmax   (sampler_edm.py:383-387)
        const float gamma_on = static_cast<float>(fmin(static_cast<double>(s_churn) / num_steps, sqrt(2.0) - 1.0));
        REQUIRE(sample0 >= 0, "sample0 must be non-negative");
        for (int i = 0; i < num_steps; ++i) {
            const float sigma = sigmas_host[i];
            const float sigma_next = (i + 1 < n_sigmas) ? sigmas_host[i + 1] : 0.0f;   // t_N = 0 appended (:377)
            const float gamma = (sigma >= s_tmin && sigma <= s_tmax) ? gamma_on : 0.0f;
            float sigma_hat = sigma;
            if (gamma > 0.f) {
                const float gs = gamma * sigma;
                sigma_hat = sigma + gs;                                                    // :343
                const float a = sqrtf(sigma_hat * sigma_hat - sigma * sigma);              // :347
                ScopedTimer t(n, ADB_TIMER_STEP, st);
                // x_hat = x + a * (s_noise * eps_i), each product rounded like the reference (:346-347); eps_i is the caller's
                // tensor (parity tests) or drawn in the kernel from (churn_seed, sample0 + b, i)
                if (eps) {
                    EdmArgs p = edm_args(x, eps + static_cast<long long>(i) * N, nullptr, x, nullptr, N, N);
                    p.a = a; p.w0 = s_noise;
                    KL(1); CK(edm_launch<OP_CHURN>(p, nullptr, st));
                } else {
                    KL(1); CK(edm_churn_rng_launch(x, x, a, s_noise, churn_seed, i, sample0, B, L, st));
                }
            }
            rc = net_eval(n, x, sigma_hat, sigma_data, w.fbuf, B, L, precision, w, st);
            if (rc) return rc;
            ++nfe;
            const PrecondCoef c0 = precond_coef(sigma_hat, sigma_data, sd2);
            const float h = sigma_next - sigma_hat;
            EdmArgs p = edm_args(x, w.fbuf, nullptr, nullptr, nullptr, N, N);
            p.s0 = sigma_hat; p.h = h; p.c_skip0 = c0.c_skip; p.c_out0 = c0.c_out;
            if (sigma_next != 0.f && use_heun) {
                {
                    ScopedTimer t(n, ADB_TIMER_STEP, st);
                    p.out0 = w.dslope; p.out1 = w.xnext;
                    KL(1); CK(edm_launch<OP_MID>(p, nullptr, st));
                }
                rc = net_eval(n, w.xnext, sigma_next, sigma_data, w.fbuf, B, L, precision, w, st);
                if (rc) return rc;
                ++nfe;
                const PrecondCoef c1 = precond_coef(sigma_next, sigma_data, sd2);
                ScopedTimer t(n, ADB_TIMER_STEP, st);
                p.out0 = x; p.out1 = nullptr;
                p.s1 = sigma_next; p.hh = 0.5f * h; p.c_skip1 = c1.c_skip; p.c_out1 = c1.c_out;
                KL(1); CK(edm_launch<OP_POST>(p, w.dslope, st));
            } else {
                ScopedTimer t(n, ADB_TIMER_STEP, st);
                p.out0 = x;
                KL(1); CK(edm_launch<OP_EULER_RAW>(p, nullptr, st));
            }
        }
    } else {
        // EDMAlphaSampler (sampler_edm.py:251-300): num_steps - 1 iterations, no sigma = 0 appended
        for (int i = 0; i + 1 < num_steps; ++i) {
            const float sigma = sigmas_host[i], sigma_next = sigmas_host[i + 1];
            const float h = sigma_next - sigma;
            rc = net_eval(n, x, sigma, sigma_data, w.fbuf, B, L, precision, w, st);
            if (rc) return rc;
            ++nfe;
            const PrecondCoef c0 = precond_coef(sigma, sigma_data, sd2);
            const float ah = alpha * h;
            const float sigma_p = sigma + ah;                                  // :268
            EdmArgs p = edm_args(x, w.fbuf, nullptr, nullptr, nullptr, N, N);
            p.s0 = sigma; p.c_skip0 = c0.c_skip; p.c_out0 = c0.c_out;
            if (sigma_p != 0.f && use_heun) {
                {
                    ScopedTimer t(n, ADB_TIMER_STEP, st);
                    p.h = ah;                                                  // x_p = x + alpha h d   (:270)
                    p.out0 = w.dslope; p.out1 = w.xnext;
                    KL(1); CK(edm_launch<OP_MID>(p, nullptr, st));
                }
                rc = net_eval(n, w.xnext, sigma_p, sigma_data, w.fbuf, B, L, precision, w, st);
                if (rc) return rc;
                ++nfe;
                // x_next = x + h ((1 - 1/2a) d + (1/2a) d_p)   (:271-278)
                const PrecondCoef c1 = precond_coef(sigma_p, sigma_data, sd2);
                ScopedTimer t(n, ADB_TIMER_STEP, st);
                p.out0 = x; p.out1 = nullptr;
                p.s1 = sigma_p; p.a = h; p.c_skip1 = c1.c_skip; p.c_out1 = c1.c_out;
                p.w0 = static_cast<float>(1.0 - 0.5 / alpha); p.w1 = static_cast<float>(0.5 / alpha);
                KL(1); CK(edm_launch<OP_POST_RK2>(p, w.dslope, st));
            } else {
                ScopedTimer t(n, ADB_TIMER_STEP, st);
                p.h = h;
                p.out0 = x;
                KL(1); CK(edm_launch<OP_EULER_RAW>(p, nullptr, st));
            }
        }
    }
    if (x_out != x) CK(cudaMemcpyAsync(x_out, x, sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
    if (nfe_out) *nfe_out = nfe;
    return ADB_OK;
}

extern "C" int adb_wavenet_sample_edm(adb_wavenet* n, const float* noise, const float* sigmas_host, int n_sigmas,
                                      int num_steps, float sigma_data, float s_tmin, float s_tmax, float s_churn,
                                      float s_noise, int use_heun, float alpha, const float* eps, float* x_out, int B, int L,
                                      int precision, void* ws, int64_t ws_bytes, int* nfe_out, void* stream) {
    return adb_wavenet_sample_edm_seeded(n, noise, sigmas_host, n_sigmas, num_steps, sigma_data, s_tmin, s_tmax, s_churn, s_noise,
                                         use_heun, alpha, eps, 0ULL, 0, x_out, B, L, precision, ws, ws_bytes, nfe_out, stream);
}

// d loss[b] / d F of adb_edm_dsm_loss(_masked), times the per-sample upstream gradient: what autograd needs to carry the DSM
// loss back into ANY differentiable backbone (generic training path; the fused DiffWave step has its own backward).
extern "C" int adb_edm_dsm_loss_grad(const float* x, const float* x_noisy, const float* f, const float* sigmas, float sigma_data,
                                     const unsigned char* mask, const float* upstream, float* d_f, int B, int64_t n_per, void* stream) {
    REQUIRE(x && x_noisy && f && sigmas && upstream && d_f && B > 0 && n_per > 0, "adb_edm_dsm_loss_grad: bad arguments");
    KL(1);
    dsm_loss_bwd_kernel<<<grid_for(static_cast<long long>(B) * n_per), 256, 0, S(stream)>>>(x, x_noisy, f, sigmas, sigma_data,
                                                                                           sd2_of(sigma_data), 1.0f, d_f, B, n_per, upstream, mask);
    CK(cudaGetLastError());
    return ADB_OK;
}

#include "cl_api.inc"
#include "train.inc"
