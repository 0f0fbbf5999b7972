// Fused elementwise kernels for EDM preconditioning and the Heun / Euler sampler update.
//
// Reference arithmetic being replaced (one fused pass instead of ~20 ATen kernels per step):
//   EluDiffusion.get_scale_weights      src/models/components/diffusion.py:232-241
//   Diffusion.denoise_fn (eq. 7 + clip) src/models/components/diffusion.py:46-63, utils.py:20-22
//   EDMSampler.step                     src/models/components/sampler_edm.py:343-367
//   EDMAlphaSampler.step                src/models/components/sampler_edm.py:259-280
//   Diffusion.forward (DSM loss)        src/models/components/diffusion.py:76-95
//
// The state is fp32 [B, n] (n = C*L contiguous per sample). Every kernel is a grid-stride loop over
// 128-bit vectors (4 floats) with a scalar tail path when n % 4 != 0 or a pointer is unaligned.
// Rounding follows the reference's operation order (separate mul / add roundings, true division) so
// results agree with the torch-CPU oracle to the last bit wherever libm is not involved.
#pragma once
#include "ptx.cuh"

namespace adb {

struct PrecondCoef {
    float c_skip, c_out, c_in, c_noise;
};

// Same operation order as diffusion.py:236-240 evaluated on fp32 tensors.
// `sd2` is (float)(double(sigma_data)^2): the reference squares the python double first and only then
// rounds it to fp32 when it meets the fp32 tensor.
__host__ __device__ __forceinline__ PrecondCoef precond_coef(float sigma, float sigma_data, float sd2) {
    PrecondCoef c;
    const float s2 = sigma * sigma;
#ifdef __CUDA_ARCH__
    const float sum = __fadd_rn(s2, sd2);
    const float rs = __fdiv_rn(1.0f, __fsqrt_rn(sum));          // (..) ** -0.5 == 1/sqrt in ATen
    c.c_skip = __fdiv_rn(sd2, sum);
    c.c_out = __fmul_rn(__fmul_rn(sigma, sigma_data), rs);
    c.c_in = rs;
    c.c_noise = __fmul_rn(logf(sigma), 0.25f);
#else
    const float sum = s2 + sd2;
    const float rs = 1.0f / sqrtf(sum);
    c.c_skip = sd2 / sum;
    c.c_out = (sigma * sigma_data) * rs;
    c.c_in = rs;
    c.c_noise = logf(sigma) * 0.25f;
#endif
    return c;
}

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

// D = clamp(c_skip * x + c_out * F)   (diffusion.py:60,63)
__device__ __forceinline__ float denoised(float x, float f, float c_skip, float c_out) {
    return clamp1(__fadd_rn(__fmul_rn(c_skip, x), __fmul_rn(c_out, f)));
}

enum EdmOp : int {
    OP_SCALE_IN = 0,     // out0 = c_in(sigma_b) * x
    OP_COMBINE = 1,      // out0 = clamp(c_skip x + c_out F)
    OP_COMBINE_CFG = 2,  // F = Fn + (F - Fn) * cond_scale first (diffusion.py:52-54)
    OP_CHURN = 3,        // out0 = x + a * (w0 * eps), w0 = s_noise               (sampler_edm.py:346-347)
    OP_EULER = 4,        // d = (x - D)/s0 ; out0 = d ; out1 = x + h d           (sampler_edm.py:354-357)
    OP_HEUN = 5,         // d2 = (x1 - D)/s1 ; out0 = x + hh (d + d2)            (sampler_edm.py:366-367)
    OP_RK2 = 6,          // d2 = (x1 - D)/s1 ; out0 = x + h (w0 d + w1 d2)       (sampler_edm.py:277-278)
    OP_AXPY = 7,         // out0 = x + a * in1                                    (sampler_edm.py:280, 273)
    OP_MID = 8,          // fused: D1 from raw F; d; x' -> out0 = d, out1 = x'
    OP_POST = 9,         // fused: x' = x + h d; D2 from raw F; x_next -> out0
    OP_EULER_RAW = 10,   // fused: D1 from raw F; out0 = x + h d (final Euler step / use_heun=False)
    OP_NOISE_IN = 11,    // training: out0 = x + sigma_b * noise ; out1 = c_in(sigma_b) * out0   (diffusion.py:79,50/57)
    OP_SCALE = 12,       // out0 = a * x                                          (sampler_edm.py:380)
    OP_POST_RK2 = 13,    // fused general RK2: x' = x + h d; D2; out0 = x + a (w0 d + w1 d2)   (sampler_edm.py:270-278)
    OP_LINCOMB2 = 14,    // out0 = a x - s0 in1                                   (DPM-Solver++(2M) 1st-order step, sampler_edm.py:1098)
    OP_LINCOMB3 = 15,    // out0 = a x - s0 (w0 in1 - w1 in2)                     (2nd-order multistep, sampler_edm.py:1107-1108)
    OP_CLAMP = 16,       // out0 = clamp(x, -1, 1)                                (sampler_edm.py:1131)
    OP_LERP = 17,        // out0 = torch.lerp(x, in1, a)                          (EMA update, src/models/phema.py:107, :151)
};

struct EdmArgs {
    const float* x;       // primary state input
    const float* in1;     // F / D / eps / d  (op dependent)
    const float* in2;     // x1 / Fnull / d   (op dependent)
    float* out0;
    float* out1;
    const float* sigmas;  // device sigma(s) for the per-sample ops (SCALE_IN / COMBINE* / NOISE_IN)
    int sigma_stride;     // 0: one sigma for the whole batch, 1: sigma per sample
    float sigma_data;
    float sd2;            // (float)(double(sigma_data)^2)
    float cond_scale;
    // host scalars for the sampler ops
    float a, s0, s1, h, hh, w0, w1, c_skip0, c_out0, c_skip1, c_out1;
    long long n_per;      // elements per sample
    long long total;      // B * n_per
};

// One generic kernel: loads up to 4 inputs, applies the op, stores up to 2 outputs.
// in3 is only used by HEUN / RK2 / POST (the slope d).
template <int OP, int VEC>
__global__ void __launch_bounds__(256) edm_kernel(EdmArgs p, const float* __restrict__ in3) {
    const long long nvec = p.total / VEC;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const long long base = v * VEC;
        float cs = 0.f, co = 0.f, ci = 0.f, sg = 0.f;
        if constexpr (OP == OP_SCALE_IN || OP == OP_COMBINE || OP == OP_COMBINE_CFG || OP == OP_NOISE_IN) {
            const long long b = p.sigma_stride ? (base / p.n_per) : 0;
            sg = __ldg(p.sigmas + b * p.sigma_stride);
            const PrecondCoef c = precond_coef(sg, p.sigma_data, p.sd2);
            cs = c.c_skip; co = c.c_out; ci = c.c_in;
        }
        float x[VEC], i1[VEC], i2[VEC], i3[VEC], o0[VEC], o1[VEC];
        constexpr bool need1 = (OP != OP_SCALE_IN && OP != OP_SCALE && OP != OP_CLAMP);
        constexpr bool need2 = (OP == OP_COMBINE_CFG || OP == OP_HEUN || OP == OP_RK2 || OP == OP_LINCOMB3);
        constexpr bool need3 = (OP == OP_HEUN || OP == OP_RK2 || OP == OP_POST || OP == OP_POST_RK2);
        constexpr bool two_out = (OP == OP_EULER || OP == OP_MID || OP == OP_NOISE_IN);
        if constexpr (VEC == 4) {
            const float4 t = *reinterpret_cast<const float4*>(p.x + base);
            x[0] = t.x; x[1] = t.y; x[2] = t.z; x[3] = t.w;
            if constexpr (need1) { const float4 u = *reinterpret_cast<const float4*>(p.in1 + base); i1[0] = u.x; i1[1] = u.y; i1[2] = u.z; i1[3] = u.w; }
            if constexpr (need2) { const float4 u = *reinterpret_cast<const float4*>(p.in2 + base); i2[0] = u.x; i2[1] = u.y; i2[2] = u.z; i2[3] = u.w; }
            if constexpr (need3) { const float4 u = *reinterpret_cast<const float4*>(in3 + base); i3[0] = u.x; i3[1] = u.y; i3[2] = u.z; i3[3] = u.w; }
        } else {
            x[0] = p.x[base];
            if constexpr (need1) i1[0] = p.in1[base];
            if constexpr (need2) i2[0] = p.in2[base];
            if constexpr (need3) i3[0] = in3[base];
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            if constexpr (OP == OP_HEUN) {
                // x = x_hat, i1 = D(x1), i2 = x1, i3 = d
                const float d2 = __fdiv_rn(__fsub_rn(i2[k], i1[k]), p.s1);
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.hh, __fadd_rn(i3[k], d2)));
            } else if constexpr (OP == OP_RK2) {
                const float d2 = __fdiv_rn(__fsub_rn(i2[k], i1[k]), p.s1);
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.h, __fadd_rn(__fmul_rn(p.w0, i3[k]), __fmul_rn(p.w1, d2))));
            } else if constexpr (OP == OP_MID) {
                const float D = denoised(x[k], i1[k], p.c_skip0, p.c_out0);
                const float d = __fdiv_rn(__fsub_rn(x[k], D), p.s0);
                o0[k] = d;
                o1[k] = __fadd_rn(x[k], __fmul_rn(p.h, d));
            } else if constexpr (OP == OP_POST) {
                // x = x_hat, i1 = raw F at x', i3 = d ; x' recomputed exactly as OP_MID stored it
                const float x1 = __fadd_rn(x[k], __fmul_rn(p.h, i3[k]));
                const float D = denoised(x1, i1[k], p.c_skip1, p.c_out1);
                const float d2 = __fdiv_rn(__fsub_rn(x1, D), p.s1);
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.hh, __fadd_rn(i3[k], d2)));
            } else if constexpr (OP == OP_POST_RK2) {
                const float x1 = __fadd_rn(x[k], __fmul_rn(p.h, i3[k]));
                const float D = denoised(x1, i1[k], p.c_skip1, p.c_out1);
                const float d2 = __fdiv_rn(__fsub_rn(x1, D), p.s1);
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.a, __fadd_rn(__fmul_rn(p.w0, i3[k]), __fmul_rn(p.w1, d2))));
            } else if constexpr (OP == OP_EULER_RAW) {
                const float D = denoised(x[k], i1[k], p.c_skip0, p.c_out0);
                const float d = __fdiv_rn(__fsub_rn(x[k], D), p.s0);
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.h, d));
            } else if constexpr (OP == OP_NOISE_IN) {
                const float xn = __fadd_rn(x[k], __fmul_rn(sg, i1[k]));
                o0[k] = xn;
                o1[k] = __fmul_rn(ci, xn);
            } else if constexpr (OP == OP_SCALE_IN) {
                o0[k] = __fmul_rn(ci, x[k]);
            } else if constexpr (OP == OP_COMBINE) {
                o0[k] = denoised(x[k], i1[k], cs, co);
            } else if constexpr (OP == OP_COMBINE_CFG) {
                // i1 = F(cond), i2 = F(null): null + (cond - null) * scale   (diffusion.py:54)
                const float f = __fadd_rn(i2[k], __fmul_rn(__fsub_rn(i1[k], i2[k]), p.cond_scale));
                o0[k] = denoised(x[k], f, cs, co);
            } else if constexpr (OP == OP_CHURN) {
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.a, __fmul_rn(p.w0, i1[k])));
            } else if constexpr (OP == OP_AXPY) {
                o0[k] = __fadd_rn(x[k], __fmul_rn(p.a, i1[k]));
            } else if constexpr (OP == OP_EULER) {
                const float d = __fdiv_rn(__fsub_rn(x[k], i1[k]), p.s0);
                o0[k] = d;
                o1[k] = __fadd_rn(x[k], __fmul_rn(p.h, d));
            } else if constexpr (OP == OP_SCALE) {
                o0[k] = __fmul_rn(p.a, x[k]);
            } else if constexpr (OP == OP_LINCOMB2) {
                o0[k] = __fsub_rn(__fmul_rn(p.a, x[k]), __fmul_rn(p.s0, i1[k]));
            } else if constexpr (OP == OP_LINCOMB3) {
                const float dd = __fsub_rn(__fmul_rn(p.w0, i1[k]), __fmul_rn(p.w1, i2[k]));
                o0[k] = __fsub_rn(__fmul_rn(p.a, x[k]), __fmul_rn(p.s0, dd));
            } else if constexpr (OP == OP_CLAMP) {
                o0[k] = clamp1(x[k]);
            } else if constexpr (OP == OP_LERP) {
                // torch.lerp: weight < 0.5 ? a + w (b - a) : b - (b - a) (1 - w)
                const float diff = __fsub_rn(i1[k], x[k]);
                o0[k] = p.a < 0.5f ? __fmaf_rn(p.a, diff, x[k]) : __fsub_rn(i1[k], __fmul_rn(diff, __fsub_rn(1.0f, p.a)));
            }
        }
        if constexpr (VEC == 4) {
            *reinterpret_cast<float4*>(p.out0 + base) = make_float4(o0[0], o0[1], o0[2], o0[3]);
            if constexpr (two_out) *reinterpret_cast<float4*>(p.out1 + base) = make_float4(o1[0], o1[1], o1[2], o1[3]);
        } else {
            p.out0[base] = o0[0];
            if constexpr (two_out) p.out1[base] = o1[0];
        }
    }
}

// c_noise[b] = 0.25 * ln(sigma_b)  (diffusion.py:235) — B values, one tiny launch.
__global__ void edm_cnoise_kernel(const float* __restrict__ sigmas, int sigma_stride, float* __restrict__ c_noise,
                                  int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) c_noise[b] = __fmul_rn(logf(sigmas[b * sigma_stride]), 0.25f);
}

// DSM loss (diffusion.py:92-95): loss_b = lambda(sigma_b) * sum_i (D_i - x_i)^2 / n, with
// D = clamp(c_skip x_noisy + c_out F). One block per (sample, chunk); per-sample atomics finish it.
// `mask` (optional, one byte per element): the reference's x_mask — masked-out elements count with weight 0.01
// (diffusion.py:80-83, :92).
__global__ void __launch_bounds__(256) edm_dsm_loss_kernel(const float* __restrict__ x, const float* __restrict__ x_noisy,
                                                           const float* __restrict__ F, const float* __restrict__ sigmas,
                                                           float sigma_data, float sd2, const unsigned char* __restrict__ mask,
                                                           float* __restrict__ loss, long long n_per, int chunks) {
    const int b = blockIdx.x / chunks;
    const int ch = blockIdx.x % chunks;
    const float sg = sigmas[b];
    const PrecondCoef c = precond_coef(sg, sigma_data, sd2);
    const long long per = (n_per + chunks - 1) / chunks;
    const long long lo = ch * per;
    const long long hi = (lo + per < n_per) ? lo + per : n_per;
    float acc = 0.f;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const long long g = b * n_per + i;
        const float D = denoised(x_noisy[g], F[g], c.c_skip, c.c_out);
        const float e = D - x[g];
        const float wgt = (mask == nullptr || mask[g]) ? 1.0f : 0.01f;
        acc = fmaf(e * e, wgt, acc);
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        acc = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
        if (threadIdx.x == 0) {
            // lambda(sigma) = (s^2 + sd^2) * (s*sd)^-2   (diffusion.py:245)
            const float sd = sigma_data;
            const float w = (sg * sg + sd2) / ((sg * sd) * (sg * sd));
            atomicAdd(loss + b, acc * w / static_cast<float>(n_per));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Churn noise drawn in the kernel (EDMSampler.step: `epsilon = randn_like(x)`, sampler_edm.py:346-347).
// Counter-based Philox4x32-10 keyed by the trajectory's 64-bit seed; the counter is
// (group of 4 elements inside the sample, global sample index, sampler step), so a sample's noise depends neither on the
// batch it is drawn in nor on the rank / world size that owns it, and nothing of size [steps][B][n] ever exists in memory.
// Normal variates by Box-Muller on 24-bit uniforms: four per Philox block.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
        const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        c1 = static_cast<uint32_t>(p1); c3 = static_cast<uint32_t>(p0); c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// four N(0,1) values of (seed, sample, step, group)
__device__ __forceinline__ void churn_normals(unsigned long long seed, long long sample, int step, uint32_t group, float (&z)[4]) {
    uint32_t r[4];
    philox4x32_10(group, static_cast<uint32_t>(sample), static_cast<uint32_t>(step), static_cast<uint32_t>(static_cast<unsigned long long>(sample) >> 32),
                  static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = (static_cast<float>(r[2 * h] >> 8) + 0.5f) * 5.9604644775390625e-08f;      // (0, 1), 2^-24 grid
        const float u2 = static_cast<float>(r[2 * h + 1] >> 8) * 5.9604644775390625e-08f;           // [0, 1)
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z[2 * h] = rad * cs;
        z[2 * h + 1] = rad * sn;
    }
}

// out = x + a * (s_noise * eps), eps ~ N(0,1) from churn_normals; one thread per group of 4 consecutive elements of a sample
template <bool VEC>
__global__ void __launch_bounds__(256) edm_churn_rng_kernel(const float* __restrict__ x, float* __restrict__ out, float a, float s_noise,
                                                            unsigned long long seed, int step, long long sample0, long long n_per,
                                                            long long groups_per_sample, long long total_groups) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long gidx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; gidx < total_groups; gidx += stride) {
        const long long b = gidx / groups_per_sample;
        const long long g = gidx - b * groups_per_sample;
        float z[4];
        churn_normals(seed, sample0 + b, step, static_cast<uint32_t>(g), z);
        const long long base = b * n_per + 4 * g;
        if (VEC) {
            const float4 v = *reinterpret_cast<const float4*>(x + base);
            float4 o;
            o.x = __fadd_rn(v.x, __fmul_rn(a, __fmul_rn(s_noise, z[0])));
            o.y = __fadd_rn(v.y, __fmul_rn(a, __fmul_rn(s_noise, z[1])));
            o.z = __fadd_rn(v.z, __fmul_rn(a, __fmul_rn(s_noise, z[2])));
            o.w = __fadd_rn(v.w, __fmul_rn(a, __fmul_rn(s_noise, z[3])));
            *reinterpret_cast<float4*>(out + base) = o;
        } else {
            const long long lim = n_per - 4 * g < 4 ? n_per - 4 * g : 4;
            for (long long k = 0; k < lim; ++k) out[base + k] = __fadd_rn(x[base + k], __fmul_rn(a, __fmul_rn(s_noise, z[k])));
        }
    }
}

inline cudaError_t edm_churn_rng_launch(const float* x, float* out, float a, float s_noise, unsigned long long seed, int step,
                                        long long sample0, int B, long long n_per, cudaStream_t stream) {
    const long long gps = (n_per + 3) / 4, total = gps * B;
    if (total <= 0) return cudaSuccess;
    const bool vec = (n_per % 4 == 0) && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (vec) edm_churn_rng_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, out, a, s_noise, seed, step, sample0, n_per, gps, total);
    else     edm_churn_rng_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(x, out, a, s_noise, seed, step, sample0, n_per, gps, total);
    return cudaGetLastError();
}

template <int OP>
inline cudaError_t edm_launch(const EdmArgs& p, const float* in3, cudaStream_t stream) {
    if (p.total <= 0) return cudaSuccess;
    auto aligned = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec = (p.n_per % 4 == 0) && aligned(p.x) && aligned(p.in1) && aligned(p.in2) && aligned(p.out0) &&
                     aligned(p.out1) && aligned(in3);
    const long long work = vec ? p.total / 4 : p.total;
    long long blocks = (work + 255) / 256;
    const long long cap = 148LL * 16;      // 16 resident 256-thread CTAs per SM's worth of grid-stride
    if (blocks > cap) blocks = cap;
    if (vec) edm_kernel<OP, 4><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, in3);
    else     edm_kernel<OP, 1><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, in3);
    return cudaGetLastError();
}

// out = a x + sum_{k<K} c[k] m[k]  (optionally clamped to [-1, 1]): the update of every DPM-Solver / UniPC step is such a
// combination of the state and up to four network outputs with host-computed scalars (sampler_edm.py:562-704, :870-987).
// (K + 2) * 4 bytes of HBM traffic per element, one launch per update.
struct LincombArgs {
    const float* x;
    const float* m[4];
    float a, c[4];
    float* out;
    long long n;
    int clamp;
};

template <int K, bool VEC>
__global__ void __launch_bounds__(256) lincomb_n_kernel(LincombArgs p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const long long nv = p.n / 4;
        for (; i < nv; i += stride) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x) + i);
            float4 mv[K > 0 ? K : 1];
#pragma unroll
            for (int k = 0; k < K; ++k) mv[k] = __ldg(reinterpret_cast<const float4*>(p.m[k]) + i);
            float4 r = make_float4(p.a * xv.x, p.a * xv.y, p.a * xv.z, p.a * xv.w);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                r.x = fmaf(p.c[k], mv[k].x, r.x); r.y = fmaf(p.c[k], mv[k].y, r.y);
                r.z = fmaf(p.c[k], mv[k].z, r.z); r.w = fmaf(p.c[k], mv[k].w, r.w);
            }
            if (p.clamp) {
                r.x = fminf(fmaxf(r.x, -1.f), 1.f); r.y = fminf(fmaxf(r.y, -1.f), 1.f);
                r.z = fminf(fmaxf(r.z, -1.f), 1.f); r.w = fminf(fmaxf(r.w, -1.f), 1.f);
            }
            reinterpret_cast<float4*>(p.out)[i] = r;
        }
    } else {
        for (; i < p.n; i += stride) {
            float r = p.a * p.x[i];
#pragma unroll
            for (int k = 0; k < K; ++k) r = fmaf(p.c[k], p.m[k][i], r);
            if (p.clamp) r = fminf(fmaxf(r, -1.f), 1.f);
            p.out[i] = r;
        }
    }
}

template <int K>
inline cudaError_t lincomb_n_launch(const LincombArgs& p, cudaStream_t stream) {
    auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    bool vec = (p.n % 4 == 0) && aligned(p.x) && aligned(p.out);
    for (int k = 0; k < K; ++k) vec = vec && aligned(p.m[k]);
    const long long work = vec ? p.n / 4 : p.n;
    const long long blocks = std::min<long long>((work + 255) / 256, 148LL * 16);
    if (vec) lincomb_n_kernel<K, true><<<(unsigned)blocks, 256, 0, stream>>>(p);
    else     lincomb_n_kernel<K, false><<<(unsigned)blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// float -> 16-bit PCM (round half to even of x * 2^15, saturated): the device half of writing 16-bit WAV files.
__device__ __forceinline__ short pcm16_of(float v) {
    const int q = __float2int_rn(v * 32768.0f);          // NaN -> 0, +-inf saturate to int32
    return (short)max(-32768, min(32767, q));
}

template <bool VEC>
__global__ void __launch_bounds__(256) pcm16_kernel(const float* __restrict__ x, short* __restrict__ out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const long long nv = n / 8;
        for (; i < nv; i += stride) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
            const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
            union { short s[8]; uint4 u; } r;
            r.s[0] = pcm16_of(a.x); r.s[1] = pcm16_of(a.y); r.s[2] = pcm16_of(a.z); r.s[3] = pcm16_of(a.w);
            r.s[4] = pcm16_of(b.x); r.s[5] = pcm16_of(b.y); r.s[6] = pcm16_of(b.z); r.s[7] = pcm16_of(b.w);
            reinterpret_cast<uint4*>(out)[i] = r.u;
        }
        if (i == nv) for (long long j = nv * 8; j < n; ++j) out[j] = pcm16_of(x[j]);   // ragged tail: one thread
    } else {
        for (; i < n; i += stride) out[i] = pcm16_of(x[i]);
    }
}

}  // namespace adb
