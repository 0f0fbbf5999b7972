// Channels-last building blocks of the 1-D U-Net backbone (fp32 CUDA-core path + shared small kernels).
//
// Reference: src/models/backbones/unet1d.py — ConvBlock1d :163-207 (GroupNorm -> scale/shift -> SiLU -> conv),
// ResnetBlock1d :257-316, Downsample1d / Upsample1d :214-255, WAVenc1d / WAVdec1d :572-622, LayerNorm /
// LayerNorm1d :16-45, FeedForward1d :49-61, TransformerBlock1d :67-122, time embedding :128-148;
// src/models/backbones/attention_utils.py:78-184 (self-attention branch).
//
// Activations are channels-last [B][L][C] (the reference is [B][C][L]); T = float (fp32 path) or
// __nv_bfloat16 (tensor-core path, statistics and softmax still in fp32). Every convolution of the network is
// expressed as ONE generic "GEMM-convolution"
//     Y[b][t][n] = act( bias[n] + sum_{j < taps} sum_{ci} X[b][t + off0 + j*dil][ci] * W[j][ci][n] ) (+ res[b][t][n])
// with rows of X outside [0, L_in) reading as zero (the conv's zero padding):
//   * k=3 "same" conv            taps 3, off0 -1, dil 1
//   * 1x1 conv / Linear          taps 1
//   * strided conv k=2f+1,s=f,p=f  on the free view [L/f][f*Cin]: taps 3, off0 -1, unused fine taps have zero weights
//   * ConvTranspose k=2f,s=f     taps 2, off0 0, dil -1, N = f*Cout, rows = L_in + 1, and the store maps
//                                (t, n) -> output row t*f + n/Cout - p, channel n % Cout (each output sample is the sum
//                                of exactly two (input row, tap) products, so no overlap-add pass is needed)
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace adb {

enum ClAct : int { CL_ACT_NONE = 0, CL_ACT_RELU = 1, CL_ACT_SILU = 2, CL_ACT_GELU = 3 };

__device__ __forceinline__ float cl_act(float x, int act) {
    switch (act) {
        case CL_ACT_RELU: return fmaxf(x, 0.f);
        case CL_ACT_SILU: return x / (1.0f + expf(-x));
        case CL_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));   // nn.GELU() default (erf)
        default: return x;
    }
}

// bf16 paths: MUFU-based exp / reciprocal (error ~1e-6 relative, far below the bf16 rounding of the result)
__device__ __forceinline__ float cl_act_fast(float x, int act) {
    switch (act) {
        case CL_ACT_RELU: return fmaxf(x, 0.f);
        case CL_ACT_SILU: return __fdividef(x, 1.0f + __expf(-x));
        case CL_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
        default: return x;
    }
}
template <typename T> __device__ __forceinline__ float cl_act_t(float x, int act);
template <> __device__ __forceinline__ float cl_act_t<float>(float x, int act) { return cl_act(x, act); }
template <> __device__ __forceinline__ float cl_act_t<__nv_bfloat16>(float x, int act) { return cl_act_fast(x, act); }

template <typename T> __device__ __forceinline__ float cl_ld(const T* p);
template <> __device__ __forceinline__ float cl_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float cl_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void cl_st(T* p, float v);
template <> __device__ __forceinline__ void cl_st<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void cl_st<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

struct ClConvArgs {
    const void* in;      // [B][L_in][Cin]
    const void* w;       // fp32 path: [taps][Cin][N] fp32 ; tensor-core path: packed blocks (cl_conv_tc.cuh)
    const float* bias;   // [N] or nullptr
    const void* res;     // [B][rows][N] or nullptr (normal store only)
    void* out;           // normal: [B][rows][N] ; transposed: [B][L_out][N / ups]
    int B, L_in, rows, Cin, N, taps, off0, dil, act;
    int ups;             // 0: normal store ; f > 0: transposed store with N = f * Cout
    int shift, L_out;    // transposed store: output row = t * ups + n / Cout - shift, kept if in [0, L_out)
    int cf_cout;         // > 0 (tensor-core path only): WAVdec store — the first ups * cf_cout columns of row t hold the output
                         // samples t * ups + phase - shift of cf_cout channels; out is fp32 channels-FIRST [B][cf_cout][L_out]
    // tensor-core path only (training step fusions); zero-initialised = plain behaviour
    int mode;            // CL_MODE_PLAIN / CL_MODE_GATE_FWD / CL_MODE_GATE_BWD
    long long ldo;       // row pitch of out (0: N)
    long long res_ld;    // row pitch of res (0: N)
    float out_scale;     // plain mode: out = (act(acc + bias) + res) * out_scale (0 is read as 1)
    const void* aux_in;  // GATE_BWD: pre-gate activations y [B][rows][2N]
    void* aux_out;       // GATE_FWD: pre-gate activations y [B][rows][N] (out receives z [B][rows][N/2])
};
// GATE_FWD: the N = 2C columns arrive permuted so that every 256-wide tile holds [128 gate | 128 filter] of the same 128
//   channels (weights / bias packed that way): writes y in the ORIGINAL layout (gate c at column c, filter c at C + c) to
//   aux_out and z = sigmoid(gate) tanh(filter) to out   (wavenet.py:110-112).
// GATE_BWD: acc = dz [N = C]; reads y from aux_in and writes dy = [dz tanh(f) s (1 - s) | dz s (1 - tanh(f)^2)] to out [2C].
enum ClMode : int { CL_MODE_PLAIN = 0, CL_MODE_GATE_FWD = 3, CL_MODE_GATE_BWD = 4 };

// fp32 tiled GEMM: 64 (rows) x 64 (n) x 16 (k), 256 threads, 4x4 outputs per thread. Cin % 16 == 0, N % 64 == 0.
__global__ void __launch_bounds__(256) cl_conv_f32_kernel(ClConvArgs p) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN];
    const int b = blockIdx.z, t0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    float acc[4][4] = {};
    const float* inb = static_cast<const float*>(p.in) + static_cast<long long>(b) * p.L_in * p.Cin;
    const float* w = static_cast<const float*>(p.w);
    const int ksteps = p.taps * p.Cin / BK;
    const int a_row = tid / 4, a_k = (tid % 4) * 4;
    const int b_k = tid / 16, b_c = (tid % 16) * 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        const int kglob = ks * BK, tap = kglob / p.Cin, ci0 = kglob % p.Cin;
        {
            const int t = t0 + a_row;
            const int s = t + p.off0 + tap * p.dil;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < p.rows && s >= 0 && s < p.L_in)
                v = *reinterpret_cast<const float4*>(inb + static_cast<long long>(s) * p.Cin + ci0 + a_k);
            As[a_k + 0][a_row] = v.x; As[a_k + 1][a_row] = v.y; As[a_k + 2][a_row] = v.z; As[a_k + 3][a_row] = v.w;
        }
        *reinterpret_cast<float4*>(&Bs[b_k][b_c]) =
            *reinterpret_cast<const float4*>(w + static_cast<long long>(kglob + b_k) * p.N + n0 + b_c);
        __syncthreads();
        float part[4][4] = {};      // per-slice partial sums keep the rounding error ~ K/16 + 16 terms
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
        __syncthreads();
    }
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    const int n = n0 + tx * 4;
    if (p.bias) { const float4 bb = *reinterpret_cast<const float4*>(p.bias + n); bias[0] = bb.x; bias[1] = bb.y; bias[2] = bb.z; bias[3] = bb.w; }
    float* out = static_cast<float*>(p.out);
    const float* res = static_cast<const float*>(p.res);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t >= p.rows) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = cl_act(acc[i][j] + bias[j], p.act);
        if (p.ups == 0) {
            const long long o = (static_cast<long long>(b) * p.rows + t) * p.N + n;
            if (res) { const float4 r = *reinterpret_cast<const float4*>(res + o); v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w; }
            *reinterpret_cast<float4*>(out + o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const int cout = p.N / p.ups, phase = n / cout, c = n % cout;
            const int to = t * p.ups + phase - p.shift;
            if (to >= 0 && to < p.L_out)
                *reinterpret_cast<float4*>(out + (static_cast<long long>(b) * p.L_out + to) * cout + c) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Small dense layers on [B][K] rows (time embedding MLP, per-block conditioning projections):
//   out[b][n] = bias[n] + sum_k W[n][k] * f(in[b][k]),  f = SiLU if silu_in (unet1d.py:271-276), W in torch layout
// One warp per output element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cl_linear_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out, int B, int K,
                                                        int N, int silu_in, int act) {
    const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= static_cast<long long>(B) * N) return;
    const int b = static_cast<int>(warp / N), n = static_cast<int>(warp % N);
    const float* x = in + static_cast<long long>(b) * K;
    const float* wr = w + static_cast<long long>(n) * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        float v = x[k];
        if (silu_in) v = v / (1.0f + expf(-v));
        acc = fmaf(wr[k], v, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[warp] = cl_act(acc + (bias ? bias[n] : 0.f), act);
}

// LabelEmbedder lookup (conditioner.py:94-106): out[b] = drop[b] ? null_row : table[labels[b]]. One block per sample.
__global__ void cl_label_embed_kernel(const float* __restrict__ table, const float* __restrict__ null_row,
                                      const long long* __restrict__ labels, const int* __restrict__ drop, float* __restrict__ out,
                                      int C, int num_classes) {
    const int b = blockIdx.x;
    const long long lab = labels[b];
    const bool use_null = (drop && drop[b]) || lab < 0 || lab >= num_classes;
    const float* src = use_null ? null_row : table + lab * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) out[static_cast<long long>(b) * C + c] = src[c];
}

// [t, sin(2 pi t w_j), cos(2 pi t w_j)]   (LearnedPositionalEmbedding, unet1d.py:128-142). out: [B][2*half + 1]
__global__ void cl_time_features_kernel(const float* __restrict__ t, const float* __restrict__ w, float* __restrict__ out,
                                        int B, int half) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int width = 2 * half + 1;
    if (i >= B * width) return;
    const int b = i / width, j = i % width;
    const float x = t[b];
    if (j == 0) { out[i] = x; return; }
    const int jj = (j - 1) % half;
    const float f = __fmul_rn(__fmul_rn(__fmul_rn(x, w[jj]), 2.0f), 3.14159265358979323846f);   // x * w * 2 * pi, left to right
    out[i] = (j - 1 < half) ? sinf(f) : cosf(f);
}

// fp32 -> bf16 with an optional SiLU (the operand of the conditioning GEMM, unet1d.py:279-283), and bf16 -> fp32
__global__ void cl_cast_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n, int silu) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        float v = in[i];
        if (silu) v = v / (1.0f + expf(-v));
        out[i] = __float2bfloat16_rn(v);
    }
}
__global__ void cl_cast_to_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = __bfloat162float(in[i]);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm (nn.GroupNorm(num_groups, C), unet1d.py:179) over (C/G channels x L) per sample.
//   pass 1: fp64 sum / sum of squares per (b, g) -> sums[b][g_off + g][2] of a [B][g_total][2] array (zeroed by the caller;
//           g_off / g_total place the groups of one half of a channel concatenation next to the other half's)
//   pass 2: y = (x - mean) * rstd * gamma + beta ; optional y = y * (scale + 1) + shift (unet1d.py:160-161) ; act
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cl_gn_stats_kernel(const T* __restrict__ in, double* __restrict__ sums, int L, int C,
                                                          int G, int chunks, int g_total, int g_off) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cpg = C / G;
    const int rows_per = (L + chunks - 1) / chunks;
    const int r0 = chunk * rows_per, r1 = min(L, r0 + rows_per);
    if (r0 >= r1) return;
    const T* base = in + static_cast<long long>(b) * L * C;
    __shared__ double red[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = 0; g < G; ++g) {
        double s = 0.0, ss = 0.0;
        const long long n = static_cast<long long>(r1 - r0) * cpg;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const int r = r0 + static_cast<int>(i / cpg), j = static_cast<int>(i % cpg);
            const double v = static_cast<double>(cl_ld<T>(base + static_cast<long long>(r) * C + g * cpg + j));
            s += v;
            ss += v * v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
        }
        if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, c = 0.0;
            for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) { a += red[0][w]; c += red[1][w]; }
            atomicAdd(&sums[(static_cast<long long>(b) * g_total + g_off + g) * 2], a);
            atomicAdd(&sums[(static_cast<long long>(b) * g_total + g_off + g) * 2 + 1], c);
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cl_gn_apply_kernel(const T* __restrict__ in, const double* __restrict__ sums,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ ss, long long ss_ld, T* __restrict__ out,
                                                          int B, int L, int C, int G, float eps, int act) {
    const long long total = static_cast<long long>(B) * L * C;
    const int cpg = C / G;
    const double cnt = static_cast<double>(cpg) * L;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % C);
        const int b = static_cast<int>(i / (static_cast<long long>(L) * C));
        const double* sp = sums + (static_cast<long long>(b) * G + c / cpg) * 2;
        const double mean = sp[0] / cnt;
        const double var = fmax(sp[1] / cnt - mean * mean, 0.0);
        const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        float y = (cl_ld<T>(in + i) - static_cast<float>(mean)) * rstd * gamma[c] + beta[c];
        if (ss) {
            const float* sb = ss + static_cast<long long>(b) * ss_ld;
            y = y * (sb[c] + 1.0f) + sb[C + c];
        }
        cl_st<T>(out + i, cl_act(y, act));
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the channel dim of each [C] row (nn.LayerNorm(C) unet1d.py:79 with gain+bias; LayerNorm1d
// unet1d.py:32-45 with gain only — in channels-last both are the same row-wise op). One warp per row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cl_layernorm_kernel(const T* __restrict__ in, const float* __restrict__ g,
                                                           const float* __restrict__ bvec, T* __restrict__ out, long long rows,
                                                           int C, float eps) {
    const long long row = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const T* x = in + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += cl_ld<T>(x + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = cl_ld<T>(x + c) - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / C + eps);
    for (int c = lane; c < C; c += 32) {
        float y = (cl_ld<T>(x + c) - mean) * rstd * g[c];
        if (bvec) y += bvec[c];
        cl_st<T>(out + row * C + c, y);
    }
}

// ------------------------------------------------------------------------------------------------
// Self-attention core (attention_utils.py:163-184): per (batch, head) softmax(q k^T d^-1/2) v with the softmax in
// fp32. q: [B][L][C], kv: [B][L][2C] (k = first C columns, v = last C), out: [B][L][C]; head h owns columns
// [h*d, (h+1)*d). K and V of one (b, h) live in shared memory; one warp per query row.
// ------------------------------------------------------------------------------------------------
constexpr int CL_ATT_KPAD = 4;      // K rows padded to d + 4 floats: 16-byte aligned rows whose float4 loads are bank-conflict free
template <typename T>
__global__ void __launch_bounds__(256) cl_attention_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ out,
                                                           int L, int C, int heads) {
    extern __shared__ __align__(16) float smem_att[];
    const int d = C / heads;                    // d % 4 == 0 (checked by the caller)
    const int ldk = d + CL_ATT_KPAD;
    const int Lp = (L + 3) & ~3;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* Ks = smem_att;                                  // [L][d + 4]
    float* Vs = Ks + static_cast<size_t>(L) * ldk;         // [L][d]
    float* Ps = Vs + static_cast<size_t>(L) * d;           // [warps][Lp]
    float* Qs = Ps + static_cast<size_t>(nwarps) * Lp;     // [warps][d]
    const T* kvb = kv + static_cast<long long>(b) * L * 2 * C;
    for (int i = threadIdx.x; i < L * d; i += blockDim.x) {
        const int r = i / d, c = i % d;
        Ks[r * ldk + c] = cl_ld<T>(kvb + static_cast<long long>(r) * 2 * C + h * d + c);
        Vs[r * d + c] = cl_ld<T>(kvb + static_cast<long long>(r) * 2 * C + C + h * d + c);
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(d));
    float* P = Ps + static_cast<size_t>(warp) * Lp;
    float* Q = Qs + static_cast<size_t>(warp) * d;
    for (int qi = blockIdx.y * nwarps + warp; qi < L; qi += gridDim.y * nwarps) {
        const T* qr = q + (static_cast<long long>(b) * L + qi) * C + h * d;
        for (int c = lane; c < d; c += 32) Q[c] = cl_ld<T>(qr + c);        // the query row once, then shared-memory broadcasts
        __syncwarp();
        float mx = -INFINITY;
        for (int j0 = 0; j0 < L; j0 += 128) {                              // four keys per lane share every q broadcast
            float s[4] = {0.f, 0.f, 0.f, 0.f};
            const float4* kr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                kr[u] = reinterpret_cast<const float4*>(Ks + static_cast<size_t>(min(j0 + lane + 32 * u, L - 1)) * ldk);
            const float4* q4 = reinterpret_cast<const float4*>(Q);
            for (int c = 0; c < d / 4; ++c) {                              // same summation order as a scalar loop over c
                const float4 qv = q4[c];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 kk = kr[u][c];
                    s[u] = fmaf(qv.x, kk.x, s[u]); s[u] = fmaf(qv.y, kk.y, s[u]);
                    s[u] = fmaf(qv.z, kk.z, s[u]); s[u] = fmaf(qv.w, kk.w, s[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + lane + 32 * u;
                if (j < L) { P[j] = s[u] * scale; mx = fmaxf(mx, s[u] * scale); }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < Lp; j += 32) {
            const float e = j < L ? expf(P[j] - mx) : 0.f;                 // the pad of P is zero so the loop below can run by fours
            P[j] = e;
            sum += e;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        const float inv = 1.0f / sum;
        for (int c0 = 0; c0 < d; c0 += 64) {                               // two adjacent value columns per lane, four keys per P load
            const int ca = min(c0 + 2 * lane, d - 2);
            const float* va = Vs + ca;
            const float4* p4 = reinterpret_cast<const float4*>(P);
            float a0 = 0.f, a1 = 0.f;
            for (int j = 0; j < Lp; j += 4) {
                const float4 p = p4[j >> 2];
                const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 v = *reinterpret_cast<const float2*>(va + static_cast<size_t>(min(j + u, L - 1)) * d);
                    a0 = fmaf(pv[u], v.x, a0);
                    a1 = fmaf(pv[u], v.y, a1);
                }
            }
            if (c0 + 2 * lane < d) {
                T* orow = out + (static_cast<long long>(b) * L + qi) * C + h * d + ca;
                cl_st<T>(orow, a0 * inv);
                cl_st<T>(orow + 1, a1 * inv);
            }
        }
        __syncwarp();
    }
}

// Self-attention core on the tensor cores (bf16 activations, head dimension 64, L % 16 == 0, L <= 128 — the U-Net's
// attention levels): one CTA per (batch, head), K as [L][64 + 8] and V transposed as [64][L + 8] in shared memory (both
// paddings make the mma fragment loads bank-conflict free), one warp per 16 query rows. S = Q K^T and O = P V are
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate); the whole score row block (16 x L) lives in registers, softmax in fp32,
// P rounded to bf16 for the second product exactly like `attn.to(sim.dtype)` (attention_utils.py:178-179).
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t att_pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

constexpr int ATT_D = 64, ATT_LMAX = 128, ATT_KPAD = 8, ATT_VPAD = 8;
__global__ void __launch_bounds__(256) cl_attention_mma_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                                               __nv_bfloat16* __restrict__ out, int L, int C, int heads) {
    extern __shared__ __align__(16) uint8_t smem_att_mma[];
    constexpr int LDK = ATT_D + ATT_KPAD;                      // 72 bf16 = 144 bytes per key row
    const int ldv = L + ATT_VPAD;                              // bf16 per row of V^T
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_att_mma);                  // [L][LDK]
    __nv_bfloat16* Vt = Ks + static_cast<size_t>(L) * LDK;                               // [64][ldv]
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const __nv_bfloat16* kvb = kv + static_cast<long long>(b) * L * 2 * C + h * ATT_D;
    for (int i = threadIdx.x; i < L * (ATT_D / 8); i += blockDim.x) {          // 16-byte pieces: 8 per key row
        const int r = i >> 3, c8 = (i & 7) * 8;
        const uint4 kk = *reinterpret_cast<const uint4*>(kvb + static_cast<long long>(r) * 2 * C + c8);
        *reinterpret_cast<uint4*>(Ks + r * LDK + c8) = kk;
        const uint4 vv = *reinterpret_cast<const uint4*>(kvb + static_cast<long long>(r) * 2 * C + C + c8);
        const __nv_bfloat16* ve = reinterpret_cast<const __nv_bfloat16*>(&vv);
#pragma unroll
        for (int e = 0; e < 8; ++e) Vt[(c8 + e) * ldv + r] = ve[e];
    }
    __syncthreads();
    const int g = lane >> 2, tig = lane & 3;
    const int nblk = L >> 3;                                   // key blocks of 8
    const float scale = 0.125f;                                // 64^-1/2
    for (int q0 = warp * 16; q0 < L; q0 += nwarps * 16) {
        const __nv_bfloat16* q_lo = q + (static_cast<long long>(b) * L + q0 + g) * C + h * ATT_D;
        const __nv_bfloat16* q_hi = q_lo + 8LL * C;
        uint32_t qa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            qa[ks][0] = *reinterpret_cast<const uint32_t*>(q_lo + ks * 16 + 2 * tig);
            qa[ks][1] = *reinterpret_cast<const uint32_t*>(q_hi + ks * 16 + 2 * tig);
            qa[ks][2] = *reinterpret_cast<const uint32_t*>(q_lo + ks * 16 + 8 + 2 * tig);
            qa[ks][3] = *reinterpret_cast<const uint32_t*>(q_hi + ks * 16 + 8 + 2 * tig);
        }
        float sc[ATT_LMAX / 8][4];
#pragma unroll
        for (int nb = 0; nb < ATT_LMAX / 8; ++nb) {
            sc[nb][0] = sc[nb][1] = sc[nb][2] = sc[nb][3] = 0.f;
            if (nb < nblk) {
                const __nv_bfloat16* krow = Ks + (nb * 8 + g) * LDK + 2 * tig;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    mma_bf16_16816(sc[nb], qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3],
                                   *reinterpret_cast<const uint32_t*>(krow + ks * 16), *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8));
            }
        }
        // softmax over the keys of rows q0 + g (values [.][0..1]) and q0 + g + 8 (values [.][2..3])
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < ATT_LMAX / 8; ++nb)
            if (nb < nblk) {
                sc[nb][0] *= scale; sc[nb][1] *= scale; sc[nb][2] *= scale; sc[nb][3] *= scale;
                m0 = fmaxf(m0, fmaxf(sc[nb][0], sc[nb][1]));
                m1 = fmaxf(m1, fmaxf(sc[nb][2], sc[nb][3]));
            }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < ATT_LMAX / 8; ++nb)
            if (nb < nblk) {
                sc[nb][0] = expf(sc[nb][0] - m0); sc[nb][1] = expf(sc[nb][1] - m0);
                sc[nb][2] = expf(sc[nb][2] - m1); sc[nb][3] = expf(sc[nb][3] - m1);
                l0 += sc[nb][0] + sc[nb][1];
                l1 += sc[nb][2] + sc[nb][3];
            }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        // O = P V with the normalised probabilities rounded to bf16 (the accumulator layout of S is the A layout of P)
        float oc[ATT_D / 8][4];
#pragma unroll
        for (int nb = 0; nb < ATT_D / 8; ++nb) oc[nb][0] = oc[nb][1] = oc[nb][2] = oc[nb][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < ATT_LMAX / 16; ++kk) {
            if (kk < (L >> 4)) {
                const uint32_t a0 = att_pack_bf16(sc[2 * kk][0] * inv0, sc[2 * kk][1] * inv0);
                const uint32_t a1 = att_pack_bf16(sc[2 * kk][2] * inv1, sc[2 * kk][3] * inv1);
                const uint32_t a2 = att_pack_bf16(sc[2 * kk + 1][0] * inv0, sc[2 * kk + 1][1] * inv0);
                const uint32_t a3 = att_pack_bf16(sc[2 * kk + 1][2] * inv1, sc[2 * kk + 1][3] * inv1);
#pragma unroll
                for (int nb = 0; nb < ATT_D / 8; ++nb) {
                    const __nv_bfloat16* vrow = Vt + (nb * 8 + g) * ldv + kk * 16 + 2 * tig;
                    mma_bf16_16816(oc[nb], a0, a1, a2, a3, *reinterpret_cast<const uint32_t*>(vrow),
                                   *reinterpret_cast<const uint32_t*>(vrow + 8));
                }
            }
        }
        __nv_bfloat16* o_lo = out + (static_cast<long long>(b) * L + q0 + g) * C + h * ATT_D;
        __nv_bfloat16* o_hi = o_lo + 8LL * C;
#pragma unroll
        for (int nb = 0; nb < ATT_D / 8; ++nb) {
            *reinterpret_cast<uint32_t*>(o_lo + nb * 8 + 2 * tig) = att_pack_bf16(oc[nb][0], oc[nb][1]);
            *reinterpret_cast<uint32_t*>(o_hi + nb * 8 + 2 * tig) = att_pack_bf16(oc[nb][2], oc[nb][3]);
        }
    }
}

// out[r][0:Ca] = a[r][:], out[r][Ca:Ca+Cb] = b[r][:] * scale_b   (UpsampleBlock1d.add_skip, unet1d.py:536-537)
template <typename T>
__global__ void __launch_bounds__(256) cl_concat_kernel(const T* __restrict__ a, const T* __restrict__ bsrc, float scale_b,
                                                        T* __restrict__ out, long long rows, int Ca, int Cb) {
    const int Ct = Ca + Cb;
    const long long total = rows * Ct;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / Ct;
        const int c = static_cast<int>(i % Ct);
        if (c < Ca) out[i] = a[r * Ca + c];
        else cl_st<T>(out + i, cl_ld<T>(bsrc + r * Cb + (c - Ca)) * scale_b);
    }
}

// ------------------------------------------------------------------------------------------------
// WAVenc1d (unet1d.py:572-594): x [B][Cin][L] channels-FIRST fp32 -> h [B][L/S][F] channels-last,
//   h[b][t][f] = sum_{c,k} w[f][c][k] * x[b][c][t*S + k - pad],  pad = W/2 - S/2, no bias. Cin*W is tiny (64).
// WAVdec1d (unet1d.py:596-622): h [B][Lc][F] channels-last -> y [B][Cout][Lc*S] channels-first fp32,
//   y[b][c][t] = sum_f sum_{i,k: i*S + k - pad = t} h[b][i][f] * w[f][c][k]   (two (i, k) pairs per t when W = 2S)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) cl_wavenc_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ out,
                                                        int B, int Cin, int L, int F, int W, int S, int pad) {
    const int Lc = (L + 2 * pad - W) / S + 1;
    const long long total = static_cast<long long>(B) * Lc * F;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(i % F);
        const int t = static_cast<int>((i / F) % Lc);
        const int b = static_cast<int>(i / (static_cast<long long>(F) * Lc));
        float acc = 0.f;
        for (int c = 0; c < Cin; ++c) {
            const float* xr = x + (static_cast<long long>(b) * Cin + c) * L;
            const float* wr = w + (static_cast<long long>(f) * Cin + c) * W;
            for (int k = 0; k < W; ++k) {
                const int s = t * S + k - pad;
                if (s >= 0 && s < L) acc = fmaf(wr[k], xr[s], acc);
            }
        }
        cl_st<T>(out + i, acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cl_wavdec_kernel(const T* __restrict__ h, const float* __restrict__ w, float* __restrict__ y,
                                                        int B, int Lc, int F, int Cout, int W, int S, int pad) {
    const int L = (Lc - 1) * S - 2 * pad + W;
    const long long total = static_cast<long long>(B) * Cout * L;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % L);
        const int c = static_cast<int>((i / L) % Cout);
        const int b = static_cast<int>(i / (static_cast<long long>(L) * Cout));
        float acc = 0.f;
        // k = t + pad - i*S in [0, W)
        const int u = t + pad;
        for (int ii = u / S; ii >= 0 && u - ii * S < W; --ii) {
            if (ii >= Lc) continue;
            const int k = u - ii * S;
            const T* hr = h + (static_cast<long long>(b) * Lc + ii) * F;
            for (int f = 0; f < F; ++f) acc = fmaf(cl_ld<T>(hr + f), w[(static_cast<long long>(f) * Cout + c) * W + k], acc);
        }
        y[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Fast paths (memory-bound, 128-bit vectorised) used when the shapes allow; the simple kernels above stay as
// the general fallback.
// ------------------------------------------------------------------------------------------------
template <typename T> struct ClVec;
template <> struct ClVec<float> {
    static constexpr int N = 4;
    __device__ static void load(const float* p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    __device__ static void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct ClVec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u); }
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
};

// GroupNorm statistics, one pass: thread owns one 16-byte channel vector (fixed group) and strides over the rows of
// its chunk (4 independent loads in flight); fp32 partial sums over <= 16 rows are promoted to fp64; the block
// combines per group with warp shuffles (fp64) and issues 2 global atomics per group.
// Needs (C / G) % VEC == 0, 256 % (C / VEC) == 0, (256 / G) threads per group. grid (chunks, B).
// Optional last stage of cl_gn_stats_vec_kernel (the fused GroupNorm-apply convolution, cl_conv_gn_tc.cuh): the block that
// finishes a sample LAST turns the group sums into the per-channel affine coefficients the convolution's transform warps
// apply (the arithmetic of cl_gn_apply_vec_kernel; `scale` = a constant factor on this input, folded into statistics and
// slope), stored HALVED because the consumer evaluates SiLU(y) = h tanh(h) + h with h = y / 2. It then re-zeroes its sums and
// its ticket, so neither needs a memset between uses.
struct GnCoefArgs {
    float* coef;            // [2][B][Cin_total]: slopes, then offsets; nullptr = statistics only
    int* tickets;           // [B], zero before the first use
    const float* gamma;     // [Cin_total]
    const float* beta;      // [Cin_total]
    const float* ss;        // [B][ss_ld]: scale at [c], shift at [Cin_total + c] (c over the concatenated channels), or nullptr
    long long ss_ld;
    float eps, scale;
    int c_off, Cin_total, B;
};

template <typename T>
__global__ void __launch_bounds__(256) cl_gn_stats_vec_kernel(const T* __restrict__ in, double* __restrict__ sums, int L, int C,
                                                              int G, int chunks, int g_total, int g_off, GnCoefArgs fin) {
    constexpr int VE = ClVec<T>::N;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cpg = C / G, vpr = C / VE;
    const int rows_per = (L + chunks - 1) / chunks;
    const int r0 = chunk * rows_per, r1 = min(L, r0 + rows_per);
    __shared__ double sd[2][256];
    pdl_wait();                    // programmatic dependent launch (a no-op otherwise): the input is the previous kernel's output
    pdl_launch_dependents();
    const T* base = in + static_cast<long long>(b) * L * C;
    const int rstep = 256 / vpr, rofs = threadIdx.x / vpr, v = threadIdx.x % vpr;
    double ds = 0.0, dss = 0.0;
    {
        const T* col = base + v * VE;
        int r = r0 + rofs;
        // eight independent 16-byte loads in flight per thread (the pass is latency-bound with four: 4.2 TB/s on a 268 MB tensor)
        for (; r + 7 * rstep < r1; r += 8 * rstep) {
            float x[8][VE];
#pragma unroll
            for (int u = 0; u < 8; ++u) ClVec<T>::load(col + static_cast<long long>(r + u * rstep) * C, x[u]);
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
            for (int u = 0; u < 8; u += 2)
#pragma unroll
                for (int i = 0; i < VE; ++i) {
                    s0 += x[u][i]; q0 = fmaf(x[u][i], x[u][i], q0);
                    s1 += x[u + 1][i]; q1 = fmaf(x[u + 1][i], x[u + 1][i], q1);
                }
            ds += static_cast<double>(s0 + s1);
            dss += static_cast<double>(q0 + q1);
        }
        for (; r + 3 * rstep < r1; r += 4 * rstep) {
            float x0[VE], x1[VE], x2[VE], x3[VE];
            ClVec<T>::load(col + static_cast<long long>(r) * C, x0);
            ClVec<T>::load(col + static_cast<long long>(r + rstep) * C, x1);
            ClVec<T>::load(col + static_cast<long long>(r + 2 * rstep) * C, x2);
            ClVec<T>::load(col + static_cast<long long>(r + 3 * rstep) * C, x3);
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
            for (int i = 0; i < VE; ++i) {
                s0 += x0[i] + x1[i]; q0 = fmaf(x0[i], x0[i], fmaf(x1[i], x1[i], q0));
                s1 += x2[i] + x3[i]; q1 = fmaf(x2[i], x2[i], fmaf(x3[i], x3[i], q1));
            }
            ds += static_cast<double>(s0 + s1);
            dss += static_cast<double>(q0 + q1);
        }
        for (; r < r1; r += rstep) {
            float x0[VE];
            ClVec<T>::load(col + static_cast<long long>(r) * C, x0);
            float s0 = 0.f, q0 = 0.f;
#pragma unroll
            for (int i = 0; i < VE; ++i) { s0 += x0[i]; q0 = fmaf(x0[i], x0[i], q0); }
            ds += static_cast<double>(s0);
            dss += static_cast<double>(q0);
        }
    }
    sd[0][threadIdx.x] = ds;
    sd[1][threadIdx.x] = dss;
    __syncthreads();
    const int vg = cpg / VE;                       // vectors per group within a row
    const int per = rstep * vg;                    // threads that accumulated into one group
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < G; g += 8) {
        double a = 0.0, c = 0.0;
        for (int j = lane; j < per; j += 32) {
            const int idx = (j / vg) * vpr + g * vg + (j % vg);
            a += sd[0][idx];
            c += sd[1][idx];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (lane == 0) {
            atomicAdd(&sums[(static_cast<long long>(b) * g_total + g_off + g) * 2], a);
            atomicAdd(&sums[(static_cast<long long>(b) * g_total + g_off + g) * 2 + 1], c);
        }
    }
    if (fin.coef == nullptr) return;
    __shared__ int s_last;
    __shared__ float s_mean[64], s_rstd[64];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&fin.tickets[b], 1) == static_cast<int>(gridDim.x) - 1;
    // the affine parameters of this thread's first two channels do not depend on the sums: request them before the ticket
    // result is known (the tail of the last block is latency, one L2 round trip shorter this way)
    float pg[2], pb[2], psc[2], psh[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = threadIdx.x + 256 * k, cg = fin.c_off + c;
        const bool ok = c < C;
        pg[k] = ok ? fin.gamma[cg] : 0.f;
        pb[k] = ok ? fin.beta[cg] : 0.f;
        psc[k] = (ok && fin.ss) ? fin.ss[static_cast<long long>(b) * fin.ss_ld + cg] + 1.0f : 1.0f;
        psh[k] = (ok && fin.ss) ? fin.ss[static_cast<long long>(b) * fin.ss_ld + fin.Cin_total + cg] : 0.f;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < G) {
        double* sp = sums + (static_cast<long long>(b) * g_total + g_off + threadIdx.x) * 2;
        const double inv = 1.0 / (static_cast<double>(cpg) * L), sc = static_cast<double>(fin.scale);
        const double mean = sc * (__ldcg(sp) * inv);
        const double var = fmax(sc * sc * (__ldcg(sp + 1) * inv) - mean * mean, 0.0);
        s_mean[threadIdx.x] = static_cast<float>(mean);
        s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(fin.eps)));
        sp[0] = 0.0; sp[1] = 0.0;
    }
    __syncthreads();
    float* ca = fin.coef + static_cast<long long>(b) * fin.Cin_total + fin.c_off;
    float* cb = ca + static_cast<long long>(fin.B) * fin.Cin_total;
    for (int c = threadIdx.x, k = 0; c < C; c += 256, ++k) {
        const int cg = fin.c_off + c;
        float g_, b_, sc = 1.0f, sh = 0.f;
        if (k < 2) { g_ = pg[k & 1]; b_ = pb[k & 1]; sc = psc[k & 1]; sh = psh[k & 1]; }
        else {
            g_ = fin.gamma[cg]; b_ = fin.beta[cg];
            if (fin.ss) { sc = fin.ss[static_cast<long long>(b) * fin.ss_ld + cg] + 1.0f; sh = fin.ss[static_cast<long long>(b) * fin.ss_ld + fin.Cin_total + cg]; }
        }
        float a = s_rstd[c / cpg] * g_;
        float bb = b_ - s_mean[c / cpg] * a;
        if (fin.ss) { a *= sc; bb = fmaf(bb, sc, sh); }
        ca[c] = 0.5f * a * fin.scale;          // the slope acts on the raw (unscaled) input
        cb[c] = 0.5f * bb;
    }
    if (threadIdx.x == 0) fin.tickets[b] = 0;
}

// GroupNorm apply: per-group mean / rstd are finished in fp64 by G threads, per-channel affine coefficients
// (normalisation, gain / bias and the optional (scale + 1, shift) folded) are built once per block in shared memory,
// then one vectorised fused pass. grid (chunks, B); C <= 2048, G <= 64.
template <typename T, int ACT>
__global__ void __launch_bounds__(256) cl_gn_apply_vec_kernel(const T* __restrict__ in, const double* __restrict__ sums,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ ss, long long ss_ld, T* __restrict__ out,
                                                              int L, int C, int G, float eps, int chunks) {
    constexpr int act = ACT;                      // compile-time: no per-element switch
    constexpr int VE = ClVec<T>::N;
    __shared__ float ca[2048], cb[2048];
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cpg = C / G;
    if (threadIdx.x < G) {
        const double cnt = static_cast<double>(cpg) * L;
        const double* sp = sums + (static_cast<long long>(b) * G + threadIdx.x) * 2;
        const double mean = sp[0] / cnt;
        const double var = fmax(sp[1] / cnt - mean * mean, 0.0);
        s_mean[threadIdx.x] = static_cast<float>(mean);
        s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float a = s_rstd[g] * gamma[c];
        float bb = beta[c] - s_mean[g] * a;
        if (ss) {
            const float sc = ss[static_cast<long long>(b) * ss_ld + c] + 1.0f, sh = ss[static_cast<long long>(b) * ss_ld + C + c];
            a *= sc;
            bb = fmaf(bb, sc, sh);
        }
        ca[c] = a; cb[c] = bb;
    }
    __syncthreads();
    const int vpr = C / VE;
    const int rows_per = (L + chunks - 1) / chunks;
    const int r0 = chunk * rows_per, r1 = min(L, r0 + rows_per);
    const int nvec = (r1 - r0) * vpr;               // < 2^31: rows_per * C / VE
    const T* src = in + (static_cast<long long>(b) * L + r0) * C;
    T* dst = out + (static_cast<long long>(b) * L + r0) * C;
    // 256 % vpr == 0 or vpr % 256 == 0 (checked by the caller): a thread's channel offset only depends on tid
    constexpr int UN = (VE == 4) ? 4 : 2;
    int i = threadIdx.x;
    for (; i + 256 * (UN - 1) < nvec; i += 256 * UN) {
        float x[UN][VE];
#pragma unroll
        for (int u = 0; u < UN; ++u) ClVec<T>::load(src + static_cast<long long>(i + 256 * u) * VE, x[u]);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int c0 = ((i + 256 * u) % vpr) * VE;
#pragma unroll
            for (int k = 0; k < VE; ++k) x[u][k] = cl_act_t<T>(fmaf(x[u][k], ca[c0 + k], cb[c0 + k]), act);
            ClVec<T>::store(dst + static_cast<long long>(i + 256 * u) * VE, x[u]);
        }
    }
    for (; i < nvec; i += 256) {
        const int c0 = (i % vpr) * VE;
        float x[VE];
        ClVec<T>::load(src + static_cast<long long>(i) * VE, x);
#pragma unroll
        for (int k = 0; k < VE; ++k) x[k] = cl_act_t<T>(fmaf(x[k], ca[c0 + k], cb[c0 + k]), act);
        ClVec<T>::store(dst + static_cast<long long>(i) * VE, x);
    }
}

// WAVenc1d with the filter bank (transposed to [Cin*W][F]) and the input window staged in shared memory.
// grid (ceil(Lc / 128), B); dynamic smem = (Cin*W*F + Cin*(128*S + W)) floats.
template <typename T>
__global__ void __launch_bounds__(256) cl_wavenc_smem_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ out,
                                                             int Cin, int L, int Lc, int F, int W, int S, int pad) {
    extern __shared__ float sm_enc[];
    constexpr int TT = 128;
    float* ws = sm_enc;                       // [Cin*W][F]
    const int span = TT * S + W;
    float* xs = ws + Cin * W * F;             // [Cin][span]
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    for (int i = threadIdx.x; i < Cin * W * F; i += blockDim.x) {
        const int f = i % F, ck = i / F;      // w is [F][Cin][W]
        ws[i] = w[static_cast<long long>(f) * Cin * W + ck];
    }
    for (int i = threadIdx.x; i < Cin * span; i += blockDim.x) {
        const int c = i / span, j = i % span;
        const int s = t0 * S - pad + j;
        xs[i] = (s >= 0 && s < L) ? x[(static_cast<long long>(b) * Cin + c) * L + s] : 0.f;
    }
    __syncthreads();
    const int rows = min(TT, Lc - t0);
    for (int i = threadIdx.x; i < rows * F; i += blockDim.x) {
        const int r = i / F, f = i % F;
        float acc = 0.f;
        for (int c = 0; c < Cin; ++c) {
            const float* xr = xs + c * span + r * S;
            const float* wr = ws + c * W * F + f;
            for (int k = 0; k < W; ++k) acc = fmaf(wr[k * F], xr[k], acc);
        }
        cl_st<T>(out + (static_cast<long long>(b) * Lc + t0 + r) * F + f, acc);
    }
}

// WAVdec1d for W == 2S (two (row, tap) pairs per output sample): filter bank [F][Cout][W] and the needed input rows
// staged in shared memory. grid (ceil(L / TO), B), TO = 32 * S output samples per block.
template <typename T>
__global__ void __launch_bounds__(256) cl_wavdec_smem_kernel(const T* __restrict__ h, const float* __restrict__ w, float* __restrict__ y,
                                                             int Lc, int L, int F, int Cout, int W, int S, int pad) {
    extern __shared__ float sm_dec[];
    const int TO = 32 * S;
    const int nrows = 34;                     // rows i0-1 .. i0+32 cover every output of the tile
    float* ws = sm_dec;                       // [F][Cout][W]
    float* hs = ws + F * Cout * W;            // [nrows][F]
    const int b = blockIdx.y, T0 = blockIdx.x * TO;
    const int ibase = (T0 + pad) / S - 1;
    for (int i = threadIdx.x; i < F * Cout * W; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < nrows * F; i += blockDim.x) {
        const int r = ibase + i / F;
        hs[i] = (r >= 0 && r < Lc) ? cl_ld<T>(h + (static_cast<long long>(b) * Lc + r) * F + i % F) : 0.f;
    }
    __syncthreads();
    const int n = min(TO, L - T0);
    for (int i = threadIdx.x; i < n * Cout; i += blockDim.x) {
        const int c = i / n, tt = i % n;
        const int u = T0 + tt + pad;
        const int i0 = u / S - ibase, k0 = u % S;          // rows i0 (tap k0) and i0 - 1 (tap k0 + S)
        const float* h0 = hs + i0 * F;
        const float* h1 = h0 - F;
        const float* w0 = ws + c * W + k0;
        float acc = 0.f;
        for (int f = 0; f < F; ++f) {
            acc = fmaf(h0[f], w0[f * Cout * W], acc);
            acc = fmaf(h1[f], w0[f * Cout * W + S], acc);
        }
        y[(static_cast<long long>(b) * Cout + c) * L + T0 + tt] = acc;
    }
}

// 16-byte vectorised concat (Ca, Cb multiples of the vector width; rows * (Ca + Cb) / VE < 2^31)
template <typename T>
__global__ void __launch_bounds__(256) cl_concat_vec_kernel(const T* __restrict__ a, const T* __restrict__ bsrc, float scale_b,
                                                            T* __restrict__ out, unsigned rows, int Ca, int Cb) {
    constexpr int VE = ClVec<T>::N;
    const unsigned va = Ca / VE, vt = (Ca + Cb) / VE;
    const unsigned total = rows * vt;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned r = i / vt, v = i - r * vt;
        float x[VE];
        if (v < va) {
            ClVec<T>::load(a + (static_cast<size_t>(r) * va + v) * VE, x);
        } else {
            ClVec<T>::load(bsrc + (static_cast<size_t>(r) * (vt - va) + (v - va)) * VE, x);
#pragma unroll
            for (int k = 0; k < VE; ++k) x[k] *= scale_b;
        }
        ClVec<T>::store(out + static_cast<size_t>(i) * VE, x);
    }
}

// Batched small dense layer with the weights read ONCE per 32 samples: out[b][n] = act(bias[n] + sum_k W[n][k] f(in[b][k])).
// lane = sample: the (optionally SiLU'd) inputs of up to 32 samples sit in shared memory as [32][K + 4] (16-byte loads,
// conflict-free); one warp owns CL_LIN_COLS output columns at a time and reads their weight rows with warp-uniform 16-byte
// loads (one broadcast transaction each), so the inner loop is 1 LDS.128 + 8 LDG.128 per 32 FMAs and needs no shuffles.
// grid (ceil(N / (8 * CL_LIN_COLS)), ceil(B / 32)), smem = 32 * (K + 4) floats; K % 4 == 0 and 16-byte aligned W rows
// (otherwise the caller uses cl_linear_kernel).
constexpr int CL_LIN_COLS = 8;
__global__ void __launch_bounds__(256) cl_linear_batched_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ out, int B, int K,
                                                                int N, int silu_in, int act) {
    extern __shared__ __align__(16) float xs[];     // [32][K + 4]
    const int ldx = K + 4;
    const int b0 = blockIdx.y * 32;
    const int nb = min(32, B - b0);
    for (int i = threadIdx.x; i < K * 32; i += blockDim.x) {
        const int bb = i / K, k = i - bb * K;       // consecutive threads read consecutive k of one sample (coalesced)
        float v = 0.f;
        if (bb < nb) {
            v = in[static_cast<long long>(b0 + bb) * K + k];
            if (silu_in) v = v / (1.0f + expf(-v));
        }
        xs[bb * ldx + k] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = (blockIdx.x * 8 + warp) * CL_LIN_COLS;
    if (n >= N) return;
    const float4* wr[CL_LIN_COLS];
#pragma unroll
    for (int c = 0; c < CL_LIN_COLS; ++c) wr[c] = reinterpret_cast<const float4*>(w + static_cast<long long>(min(n + c, N - 1)) * K);
    const float4* xr = reinterpret_cast<const float4*>(xs + lane * ldx);
    float acc[CL_LIN_COLS];
#pragma unroll
    for (int c = 0; c < CL_LIN_COLS; ++c) acc[c] = 0.f;
    for (int k4 = 0; k4 < K / 4; ++k4) {            // ascending k per column: same summation order as a scalar loop
        const float4 xv = xr[k4];
#pragma unroll
        for (int c = 0; c < CL_LIN_COLS; ++c) {
            const float4 wv = __ldg(wr[c] + k4);
            acc[c] = fmaf(wv.x, xv.x, acc[c]); acc[c] = fmaf(wv.y, xv.y, acc[c]);
            acc[c] = fmaf(wv.z, xv.z, acc[c]); acc[c] = fmaf(wv.w, xv.w, acc[c]);
        }
    }
    if (lane < nb) {
#pragma unroll
        for (int c = 0; c < CL_LIN_COLS; ++c)
            if (n + c < N) out[static_cast<long long>(b0 + lane) * N + n + c] = cl_act(acc[c] + (bias ? bias[n + c] : 0.f), act);
    }
}

// WAVenc1d on the tensor cores, step 1: x [B][Cin][L] channels-first fp32 -> bf16 channels-last rows [B][L/W + 1][W*Cin],
// shifted right by `pad` samples (zeros in front and behind) so that buffer row m = samples [mW - pad, (m+1)W - pad): with
// W = 2S an even output frame 2m reads exactly row m and an odd frame the second half of row m plus the first half of
// row m+1 — a 2-tap GEMM-convolution over these rows (weights laid out by the host, audiodiffuser_b200/backbones/unet1d.py).
__global__ void __launch_bounds__(256) cl_wavenc_prep_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int Cin,
                                                             int L, long long Lp, int pad, const float* __restrict__ scale) {
    // `scale` (per sample, or nullptr): the EDM input scale c_in(sigma) folded into the re-layout (diffusion.py:46-48 multiplies the
    // network input by it; a separate pass over a 268 MB state otherwise)
    if (Cin == 2 && (pad & 3) == 0 && (L & 3) == 0 && (Lp & 3) == 0) {
        // stereo: four consecutive samples per thread: two 16-byte loads, one 16-byte store (pad, L multiples of 4: all in or all out)
        const long long total4 = static_cast<long long>(B) * (Lp >> 2);
        for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const long long b = i / (Lp >> 2), j = (i - b * (Lp >> 2)) << 2;
            const long long s0 = j - pad;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (s0 >= 0 && s0 < L) {
                const float sc = scale ? __ldg(scale + b) : 1.0f;
                const float4 a = __ldg(reinterpret_cast<const float4*>(x + b * 2 * static_cast<long long>(L) + s0));
                const float4 c = __ldg(reinterpret_cast<const float4*>(x + (b * 2 + 1) * static_cast<long long>(L) + s0));
                o = make_uint4(pack_bf16x2(a.x * sc, c.x * sc), pack_bf16x2(a.y * sc, c.y * sc), pack_bf16x2(a.z * sc, c.z * sc),
                               pack_bf16x2(a.w * sc, c.w * sc));
            }
            *reinterpret_cast<uint4*>(out + (b * Lp + j) * 2) = o;
        }
        return;
    }
    const long long total = static_cast<long long>(B) * Lp;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / Lp, j = i - b * Lp;
        const long long s = j - pad;
        const bool in = s >= 0 && s < L;
        const float sc = scale ? __ldg(scale + b) : 1.0f;
        const float* xb = x + b * Cin * static_cast<long long>(L) + s;
        __nv_bfloat16* o = out + i * Cin;
        if (Cin == 2) {
            const float v0 = in ? __ldg(xb) * sc : 0.f, v1 = in ? __ldg(xb + L) * sc : 0.f;
            *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(v0, v1);
        } else {
            for (int c = 0; c < Cin; ++c) o[c] = __float2bfloat16(in ? __ldg(xb + static_cast<long long>(c) * L) * sc : 0.f);
        }
    }
}

// WAVenc1d, register-tiled: thread = 4 consecutive output rows x 4 consecutive filters; per (channel, tap) one 16-byte
// weight load and four input loads feed 16 FMAs. Same staging as cl_wavenc_smem_kernel. F % 4 == 0.
template <typename T>
__global__ void __launch_bounds__(256) cl_wavenc_tiled_kernel(const float* __restrict__ x, const float* __restrict__ w, T* __restrict__ out,
                                                              int Cin, int L, int Lc, int F, int W, int S, int pad) {
    extern __shared__ float sm_enc[];
    constexpr int TT = 128;
    float* ws = sm_enc;                       // [Cin*W][F]
    const int span = TT * S + W;
    float* xs = ws + Cin * W * F;             // [Cin][span]
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    for (int i = threadIdx.x; i < Cin * W * F; i += blockDim.x) {
        const int f = i % F, ck = i / F;
        ws[i] = w[static_cast<long long>(f) * Cin * W + ck];
    }
    for (int i = threadIdx.x; i < Cin * span; i += blockDim.x) {
        const int c = i / span, j = i % span;
        const int s = t0 * S - pad + j;
        xs[i] = (s >= 0 && s < L) ? x[(static_cast<long long>(b) * Cin + c) * L + s] : 0.f;
    }
    __syncthreads();
    const int f4n = F / 4;
    for (int i = threadIdx.x; i < (TT / 4) * f4n; i += blockDim.x) {
        const int r0 = (i / f4n) * 4, f0 = (i % f4n) * 4;
        float acc[4][4] = {};
        for (int c = 0; c < Cin; ++c) {
            const float* xr = xs + c * span + r0 * S;
            const float* wr = ws + c * W * F + f0;
            for (int k = 0; k < W; ++k) {
                const float4 wv = *reinterpret_cast<const float4*>(wr + k * F);
                const float xv[4] = {xr[k], xr[S + k], xr[2 * S + k], xr[3 * S + k]};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u][0] = fmaf(wv.x, xv[u], acc[u][0]); acc[u][1] = fmaf(wv.y, xv[u], acc[u][1]);
                    acc[u][2] = fmaf(wv.z, xv[u], acc[u][2]); acc[u][3] = fmaf(wv.w, xv[u], acc[u][3]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (t0 + r0 + u >= Lc) break;
            T* o = out + (static_cast<long long>(b) * Lc + t0 + r0 + u) * F + f0;
#pragma unroll
            for (int q = 0; q < 4; ++q) cl_st<T>(o + q, acc[u][q]);
        }
    }
}

}  // namespace adb
