// Backward-pass kernels of the DiffWave training step (fp32 CUDA-core path).
//
// Reference: the autograd graph of Diffusion.forward (src/models/components/diffusion.py:65-97) through
// WaveNetNoise (src/models/backbones/wavenet.py:94-180) as built by PyTorch; the forward arithmetic is the one of
// wavenet_f32.cuh. Every data-gradient GEMM reuses conv_cl_f32 (a transposed / tap-reversed convolution is again a
// channels-last convolution); this file adds the weight-gradient GEMM (reduction over batch x time), the element-wise
// derivative kernels, the bias / embedding reductions, the weight-norm chain rule and the fused AdamW update
// (torch.optim.AdamW semantics, configs/model/diffunet_complex.yaml:7-12).
#pragma once
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "cl_ops.cuh"

namespace adb {

// All element-wise kernels below work on 16-byte vectors (ClVec<T>: 4 floats or 8 bf16) of one row, with 32-bit
// index math (rows * C / VE < 2^31 is checked by the caller).

// out[b][t][c] = h[b][t][c] + p[b][c]   (wavenet.py:108-109, materialised for the weight gradient)
template <typename T>
__global__ void __launch_bounds__(256) add_bcast_kernel(const T* __restrict__ h, const float* __restrict__ p,
                                                        T* __restrict__ out, int B, int L, int C) {
    constexpr int VE = ClVec<T>::N;
    constexpr int U = 4;                                   // four independent 16-byte loads in flight per thread
    const int vpr = C / VE;
    const unsigned total = static_cast<unsigned>(B) * L * vpr;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
        float x[U][VE];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned i = i0 + u * stride;
            if (i < total) ClVec<T>::load(h + static_cast<size_t>(i) * VE, x[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned i = i0 + u * stride;
            if (i >= total) break;
            const unsigned r = i / vpr, v = i - r * vpr;
            const float* pp = p + static_cast<size_t>(r / L) * C + v * VE;
#pragma unroll
            for (int k = 0; k < VE; ++k) x[u][k] += pp[k];
            ClVec<T>::store(out + static_cast<size_t>(i) * VE, x[u]);
        }
    }
}

// z[r][c] = sigmoid(y[r][c]) * tanh(y[r][C + c])   (wavenet.py:111-112), any activation dtype
template <typename T>
__global__ void __launch_bounds__(256) gate_t_kernel(const T* __restrict__ y, T* __restrict__ z, long long rows, int C) {
    constexpr int VE = ClVec<T>::N;
    const int vpr = C / VE;
    const unsigned total = static_cast<unsigned>(rows) * vpr;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned r = i / vpr, v = i - r * vpr;
        const T* yr = y + static_cast<size_t>(r) * 2 * C + v * VE;
        float g[VE], f[VE];
        ClVec<T>::load(yr, g);
        ClVec<T>::load(yr + C, f);
#pragma unroll
        for (int k = 0; k < VE; ++k) g[k] = __fdividef(1.0f, 1.0f + __expf(-g[k])) * tanh_fast(f[k]);
        ClVec<T>::store(z + static_cast<size_t>(i) * VE, g);
    }
}

// z = sigmoid(g) tanh(f)  =>  dg = dz tanh(f) s (1 - s),  df = dz s (1 - tanh(f)^2)   (wavenet.py:111-112)
template <typename T>
__global__ void __launch_bounds__(256) gate_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dz,
                                                       T* __restrict__ dy, long long rows, int C) {
    constexpr int VE = ClVec<T>::N;
    constexpr bool FAST = (VE == 8);                 // bf16 path: MUFU exp / tanh (error far below the bf16 rounding)
    const int vpr = C / VE;
    const unsigned total = static_cast<unsigned>(rows) * vpr;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned r = i / vpr, v = i - r * vpr;
        const T* yr = y + static_cast<size_t>(r) * 2 * C + v * VE;
        float g[VE], f[VE], d[VE];
        ClVec<T>::load(yr, g);
        ClVec<T>::load(yr + C, f);
        ClVec<T>::load(dz + static_cast<size_t>(i) * VE, d);
#pragma unroll
        for (int k = 0; k < VE; ++k) {
            const float s = FAST ? __fdividef(1.0f, 1.0f + __expf(-g[k])) : 1.0f / (1.0f + expf(-g[k]));
            const float th = FAST ? tanh_fast(f[k]) : tanhf(f[k]);
            g[k] = d[k] * th * s * (1.0f - s);
            f[k] = d[k] * s * (1.0f - th * th);
        }
        T* dr = dy + static_cast<size_t>(r) * 2 * C + v * VE;
        ClVec<T>::store(dr, g);
        ClVec<T>::store(dr + C, f);
    }
}

// do[r][0:C] = dh_out[r] / sqrt(2) (0 if dh_out == nullptr: the last block's residual output is unused) ; do[r][C:2C] = dskip[r]
// (wavenet.py:114-115, :149)
template <typename T>
__global__ void __launch_bounds__(256) build_do_kernel(const T* __restrict__ dh_out, const T* __restrict__ dskip,
                                                       T* __restrict__ dout, long long rows, int C) {
    constexpr int VE = ClVec<T>::N;
    const int vpr = C / VE;
    const unsigned total = static_cast<unsigned>(rows) * vpr;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const unsigned r = i / vpr, v = i - r * vpr;
        float a[VE], s[VE];
        if (dh_out) {
            ClVec<T>::load(dh_out + static_cast<size_t>(i) * VE, a);
#pragma unroll
            for (int k = 0; k < VE; ++k) a[k] *= 0.70710678118654752f;
        } else {
#pragma unroll
            for (int k = 0; k < VE; ++k) a[k] = 0.f;
        }
        ClVec<T>::load(dskip + static_cast<size_t>(i) * VE, s);
        T* dr = dout + static_cast<size_t>(r) * 2 * C + v * VE;
        ClVec<T>::store(dr, a);
        ClVec<T>::store(dr + C, s);
    }
}

// out = a * x + b * y (y may be nullptr); n % VE == 0
template <typename T>
__global__ void __launch_bounds__(256) axpby_kernel(const T* __restrict__ x, float a, const T* __restrict__ y, float b,
                                                    T* __restrict__ out, long long n) {
    constexpr int VE = ClVec<T>::N;
    const unsigned total = static_cast<unsigned>(n / VE);
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        float xv[VE], yv[VE];
        ClVec<T>::load(x + static_cast<size_t>(i) * VE, xv);
        if (y) {
            ClVec<T>::load(y + static_cast<size_t>(i) * VE, yv);
#pragma unroll
            for (int k = 0; k < VE; ++k) xv[k] = fmaf(a, xv[k], b * yv[k]);
        } else {
#pragma unroll
            for (int k = 0; k < VE; ++k) xv[k] *= a;
        }
        ClVec<T>::store(out + static_cast<size_t>(i) * VE, xv);
    }
}


// ------------------------------------------------------------------------------------------------
// Everything a residual block derives from its parameters, for ALL blocks in one launch (grid.y = block). Every derived
// element is one scaled gather from the weight-normed tensors v1 [2C][C][3], v2 [2C][C] (scale = g / ||v||), the embedding
// projection wp [C][512] or the conv bias b1 [2C]; an optimizer step used to issue 13 small launches per block for this:
//   w1f [3][C][2C] fp32   folded dilated-conv weights (ci-major)            w2f [C][2C] fp32   folded 1x1 weights
//   w2T [2C][C]           = w2f transposed (dz = DO W2^T)                   w1d [3][2C][C]     tap-reversed transpose (dx)
//   wpT [512][C]          transposed embedding projection (fold tables)
//   wtc [32][256][64]     TMA blocks of the forward kernels (G1 bf16, G2 fp16; wavenet_tc.cuh pack order)
//   w1p / w2Tp / w1dp     the same three matrices as cl_conv_tc blocks [kb][N][64] bf16 (w1p with the [128 gate | 128 filter]
//                         column order of CL_MODE_GATE_FWD), b1p the conv bias in that column order
// ------------------------------------------------------------------------------------------------
struct RefoldLayer {
    const float *v1, *v2, *wp, *b1, *s1, *s2;
    float *w1f, *w2f, *w2T, *w1d, *wpT, *b1p;
    __nv_bfloat16 *wtc, *w1p, *w2Tp, *w1dp;
    __nv_bfloat16* ws16;    // [4][256][64] bf16 copy of the skip half of W2 (blocks 28..31 of wtc hold it in fp16): training's bf16 skip GEMM
};

__device__ __forceinline__ int gate_perm_src(int np, int C) {      // column np of the permuted layout <- original column
    const int j = np >> 8, r = np & 255;
    return r < 128 ? 128 * j + r : C + 128 * j + (r - 128);
}

__global__ void __launch_bounds__(256) refold_layers_kernel(const RefoldLayer* __restrict__ tab, int C) {
    const RefoldLayer L = tab[blockIdx.y];
    const float s1 = L.s1[0], s2 = L.s2[0];
    const long long N2 = 2LL * C, n_w1 = 3LL * C * N2, n_w2 = static_cast<long long>(C) * N2, n_wp = 512LL * C;
    const bool tc = L.wtc != nullptr;                      // C == 256
    const long long seg[10] = {n_w1, n_w2, n_w2, n_w1, n_wp, tc ? 32LL * 256 * 64 : 0, tc ? n_w1 : 0, tc ? n_w2 : 0, tc ? n_w1 : 0,
                               tc ? N2 : 0};
    long long total = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) total += seg[k];
    for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < total;
         g += static_cast<long long>(gridDim.x) * blockDim.x) {
        long long i = g;
        int k = 0;
        while (i >= seg[k]) { i -= seg[k]; ++k; }
        switch (k) {
            case 0: {   // w1f[tap][ci][co]
                const int co = static_cast<int>(i % N2), ci = static_cast<int>((i / N2) % C), tap = static_cast<int>(i / (N2 * C));
                L.w1f[i] = L.v1[(static_cast<long long>(co) * C + ci) * 3 + tap] * s1;
                break;
            }
            case 1: {   // w2f[ci][co]
                const int co = static_cast<int>(i % N2), ci = static_cast<int>(i / N2);
                L.w2f[i] = L.v2[static_cast<long long>(co) * C + ci] * s2;
                break;
            }
            case 2: L.w2T[i] = L.v2[i] * s2; break;          // w2T[co][ci] = w2f[ci][co]
            case 3: {   // w1d[tap][co][ci] = w1f[2 - tap][ci][co]
                const int ci = static_cast<int>(i % C), co = static_cast<int>((i / C) % N2), tap = static_cast<int>(i / (N2 * C));
                L.w1d[i] = L.v1[(static_cast<long long>(co) * C + ci) * 3 + (2 - tap)] * s1;
                break;
            }
            case 4: {   // wpT[k][c] = wp[c][k]
                const int c = static_cast<int>(i % C), kk = static_cast<int>(i / C);
                L.wpT[i] = L.wp[static_cast<long long>(c) * 512 + kk];
                break;
            }
            case 5: {   // forward TMA blocks (pack_tc_layer_kernel order)
                const int kk = static_cast<int>(i & 63), n = static_cast<int>((i >> 6) & 255), blk = static_cast<int>(i >> 14);
                if (blk < 24) {
                    const int j = blk / 12, kb = blk % 12, tap = kb >> 2, cib = kb & 3;
                    const int co = n < 128 ? 128 * j + n : 256 + 128 * j + (n - 128);
                    L.wtc[i] = __float2bfloat16_rn(L.v1[(static_cast<long long>(co) * C + cib * 64 + kk) * 3 + tap] * s1);
                } else {
                    const int j = (blk - 24) >> 2, kb = (blk - 24) & 3;
                    const float wv = L.v2[static_cast<long long>(256 * j + n) * C + kb * 64 + kk] * s2;
                    const __half hv = __float2half_rn(wv);
                    L.wtc[i] = *reinterpret_cast<const __nv_bfloat16*>(&hv);      // GEMM2 runs in fp16: store the fp16 bit pattern
                    if (j == 1 && L.ws16) L.ws16[i - 28LL * 256 * 64] = __float2bfloat16_rn(wv);
                }
                break;
            }
            case 6: {   // w1p: cl_conv_tc blocks of the column-permuted w1f (Cin = C, N = 2C, 3 taps)
                const int kk = static_cast<int>(i & 63), n = static_cast<int>((i >> 6) % N2), kb = static_cast<int>(i / (64 * N2));
                const int kbt = C / 64, tap = kb / kbt, ci = (kb % kbt) * 64 + kk;
                L.w1p[i] = __float2bfloat16_rn(L.v1[(static_cast<long long>(gate_perm_src(n, C)) * C + ci) * 3 + tap] * s1);
                break;
            }
            case 7: {   // w2Tp: blocks of w2T (Cin = 2C, N = C, 1 tap)
                const int kk = static_cast<int>(i & 63), n = static_cast<int>((i >> 6) % C), kb = static_cast<int>(i / (64LL * C));
                L.w2Tp[i] = __float2bfloat16_rn(L.v2[static_cast<long long>(kb * 64 + kk) * C + n] * s2);
                break;
            }
            case 8: {   // w1dp: blocks of w1d (Cin = 2C, N = C, 3 taps)
                const int kk = static_cast<int>(i & 63), n = static_cast<int>((i >> 6) % C), kb = static_cast<int>(i / (64LL * C));
                const int kbt = static_cast<int>(N2 / 64), tap = kb / kbt, co = (kb % kbt) * 64 + kk;
                L.w1dp[i] = __float2bfloat16_rn(L.v1[(static_cast<long long>(co) * C + n) * 3 + (2 - tap)] * s1);
                break;
            }
            default: L.b1p[i] = L.b1[gate_perm_src(static_cast<int>(i), C)]; break;
        }
    }
}

// dst (bf16) = src (fp32)

__global__ void __launch_bounds__(256) cvt_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n,
                                                           float scale = 1.0f) {
    const long long n4 = n / 4;                                 // n % 4 == 0 and 16-byte aligned buffers at every call site
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(in)[i];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x * scale, v.y * scale), hi = __floats2bfloat162_rn(v.z * scale, v.w * scale);
        uint2 u;
        u.x = *reinterpret_cast<unsigned*>(&lo); u.y = *reinterpret_cast<unsigned*>(&hi);
        reinterpret_cast<uint2*>(out)[i] = u;
    }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient: out[i][j] += sum_{b, t : 0 <= t + shift < L} A[b][t + shift][i] * G[b][t][j]
// (A: [nb][L][Ca] layer input, G: [nb][L][Cb] output gradient). Tile 64 x 64 outputs, each block reduces a slab of
// `rows_per_block` time steps of one sample and adds its partial tile with fp32 atomics. out row pitch = ldo.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wgrad_f32_kernel(const float* __restrict__ A, const float* __restrict__ G,
                                                        float* __restrict__ out, int L, int Ca, int Cb, int shift, long long ldo,
                                                        int rows_per_block, float a_scale) {
    constexpr int BK = 16;
    __shared__ float As[BK][64 + 4];
    __shared__ float Gs[BK][64];
    const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
    const int slabs = (L + rows_per_block - 1) / rows_per_block;
    const int b = blockIdx.z / slabs, slab = blockIdx.z % slabs;
    const int t_lo = slab * rows_per_block, t_hi = min(L, t_lo + rows_per_block);
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const float* Ab = A + static_cast<long long>(b) * L * Ca;
    const float* Gb = G + static_cast<long long>(b) * L * Cb;
    float acc[4][4] = {};
    const int l_row = tid / 16, l_c = (tid % 16) * 4;           // 16 rows x 64 channels per load pass
    for (int t = t_lo; t < t_hi; t += BK) {
        {
            const int tt = t + l_row, s = tt + shift;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vg = va;
            if (tt < t_hi) {
                vg = *reinterpret_cast<const float4*>(Gb + static_cast<long long>(tt) * Cb + j0 + l_c);
                if (s >= 0 && s < L) va = *reinterpret_cast<const float4*>(Ab + static_cast<long long>(s) * Ca + i0 + l_c);
            }
            *reinterpret_cast<float4*>(&As[l_row][l_c]) = va;
            *reinterpret_cast<float4*>(&Gs[l_row][l_c]) = vg;
        }
        __syncthreads();
        float part[4][4] = {};
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 g = *reinterpret_cast<const float4*>(&Gs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], gv[j], part[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j];
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            atomicAdd(out + static_cast<long long>(i0 + ty * 4 + i) * ldo + j0 + tx * 4 + j, acc[i][j] * a_scale);
}

// Column sums: out[(per_sample ? b : 0)][c] += sum_t in[b][t][c]. grid (chunks, nb), 256 threads; a thread owns one
// 16-byte channel vector and strides over the rows of its chunk. C <= 1024, 256 % (C / VE) == 0.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, float* __restrict__ out, int L, int C,
                                                     int chunks, int per_sample, int ld) {
    constexpr int VE = ClVec<T>::N;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int rows_per = (L + chunks - 1) / chunks;
    const int r0 = chunk * rows_per, r1 = min(L, r0 + rows_per);
    __shared__ float red[1024];
    for (int c = threadIdx.x; c < C; c += blockDim.x) red[c] = 0.f;
    __syncthreads();
    const int vpr = C / VE;
    const int rstep = 256 / vpr, rofs = threadIdx.x / vpr, v = threadIdx.x % vpr;
    float s[VE];
#pragma unroll
    for (int k = 0; k < VE; ++k) s[k] = 0.f;
    const T* col = in + static_cast<size_t>(b) * L * ld + v * VE;     // ld = row pitch (>= C)
    int r = r0 + rofs;
    for (; r + 3 * rstep < r1; r += 4 * rstep) {           // four independent 16-byte loads in flight per thread
        float x[4][VE];
#pragma unroll
        for (int u = 0; u < 4; ++u) ClVec<T>::load(col + static_cast<size_t>(r + u * rstep) * ld, x[u]);
#pragma unroll
        for (int k = 0; k < VE; ++k) s[k] += (x[0][k] + x[1][k]) + (x[2][k] + x[3][k]);
    }
    for (; r < r1; r += rstep) {
        float x[VE];
        ClVec<T>::load(col + static_cast<size_t>(r) * ld, x);
#pragma unroll
        for (int k = 0; k < VE; ++k) s[k] += x[k];
    }
#pragma unroll
    for (int k = 0; k < VE; ++k) atomicAdd(&red[v * VE + k], s[k]);
    __syncthreads();
    float* o = out + (per_sample ? static_cast<size_t>(b) * C : 0);
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(o + c, red[c]);
}

// Step-embedding term of the dilated conv's weight gradient (wavenet.py:108-110): the conv sees x = h + p_b inside [0, L) and
// zeros outside, so  dW1[tap][i][j] = sum_b sum_t h_b[t + s_tap][i] dy_b[t][j]  +  sum_b p_b[i] * (sum over the rows t whose
// tap stays inside the sample of dy_b[t][j]).  The first term is the tensor-core GEMM on h itself; this kernel adds the second
// from the three per-sample column sums S (all rows / first d rows / last d rows — a by-product of the weight-gradient GEMM, see
// WgradTcParams::colsum_mode) and also writes db1[j] = sum_b S_all[b][j].
//   tap 0 (t - d >= 0): S_all - S_lo ;  tap 1: S_all ;  tap 2 (t + d < L): S_all - S_hi.   gw1: [3][C][N] fp32, N = 2C.
// grid (N / 256, C / 8, 3 taps), 256 threads: a thread owns one column j and eight rows i; p of the block's rows sits in shared memory.
// grid.z = 3 * layers runs every block's correction in ONE launch (layer = z / 3): p, S, gw1 and db1 advance by the given strides.
__global__ void __launch_bounds__(256) wgrad_pcorr_kernel(const float* __restrict__ p /*[B][C]*/, const float* __restrict__ S /*[3][B][N]*/,
                                                          float* __restrict__ gw1, float* __restrict__ db1, int B, int C, int N,
                                                          long long gw_stride, long long db_stride) {
    __shared__ float ps[64 * 8];                              // [b][8 rows], B <= 64 per pass
    const int layer = blockIdx.z / 3;
    p += static_cast<long long>(layer) * B * C; S += static_cast<long long>(layer) * 3 * B * N;
    gw1 += layer * gw_stride; db1 += layer * db_stride;
    const int j = blockIdx.x * 256 + threadIdx.x, i0 = blockIdx.y * 8, tap = blockIdx.z % 3;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float tot = 0.f;
    for (int b0 = 0; b0 < B; b0 += 64) {
        const int nb = min(64, B - b0);
        __syncthreads();
        for (int e = threadIdx.x; e < nb * 8; e += 256) ps[e] = p[static_cast<long long>(b0 + (e >> 3)) * C + i0 + (e & 7)];
        __syncthreads();
        for (int b = 0; b < nb; ++b) {
            const float sa = S[(static_cast<long long>(0) * B + b0 + b) * N + j];
            const float sv = tap == 1 ? sa : sa - S[(static_cast<long long>(tap == 0 ? 1 : 2) * B + b0 + b) * N + j];
            tot += sa;
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(ps[b * 8 + r], sv, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) gw1[(static_cast<long long>(tap) * C + i0 + r) * N + j] += acc[r];
    if (tap == 1 && blockIdx.y == 0) db1[j] = tot;
}

// ------------------------------------------------------------------------------------------------
// Tail backward (wavenet.py:177-179): F = b_out + sum_c w_out[c] s2[c], s2 = relu(...)
//   ds2[r][c] = dF[r] w_out[c] [s2 > 0] ; dw_out[c] += sum_r dF[r] s2[r][c] ; db_out += sum_r dF[r]
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) tail_bwd_kernel(const float* __restrict__ dF, const T* __restrict__ s2,
                                                       const float* __restrict__ w_out, T* __restrict__ ds2,
                                                       float* __restrict__ dw_out, float* __restrict__ db_out, long long rows,
                                                       int C) {
    __shared__ float acc_w[1024];
    __shared__ float acc_b;
    for (int c = threadIdx.x; c < C; c += blockDim.x) acc_w[c] = 0.f;
    if (threadIdx.x == 0) acc_b = 0.f;
    __syncthreads();
    const int c = threadIdx.x % C;                  // C <= 256 here: fixed channel per thread
    const int rstep = blockDim.x / C, rofs = threadIdx.x / C;
    float lw = 0.f, lb = 0.f;
    const float wc = w_out[c];
    for (long long r = static_cast<long long>(blockIdx.x) * rstep + rofs; r < rows; r += static_cast<long long>(gridDim.x) * rstep) {
        const float g = dF[r], s = cl_ld<T>(s2 + r * C + c);
        cl_st<T>(ds2 + r * C + c, s > 0.f ? g * wc : 0.f);
        lw = fmaf(g, s, lw);
        if (c == 0) lb += g;
    }
    atomicAdd(&acc_w[c], lw);
    if (c == 0) atomicAdd(&acc_b, lb);
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) atomicAdd(dw_out + k, acc_w[k]);
    if (threadIdx.x == 0) atomicAdd(db_out, acc_b);
}

// Input projection backward (wavenet.py:172-174): h0 = relu(w_in x~ + b_in):
//   dw_in[c] += sum_r [h0 > 0] dh0[r][c] x~[r] ; db_in[c] += sum_r [h0 > 0] dh0[r][c] ; x~[r] = scale_b x[r]
template <typename T>
__global__ void __launch_bounds__(256) inproj_bwd_kernel(const T* __restrict__ dh0, const T* __restrict__ h0,
                                                         const float* __restrict__ x, const float* __restrict__ scale,
                                                         float* __restrict__ dw_in, float* __restrict__ db_in, int L,
                                                         long long rows, int C, int ld_dh, float dh_scale) {
    __shared__ float aw[1024], ab[1024];
    for (int c = threadIdx.x; c < C; c += blockDim.x) { aw[c] = 0.f; ab[c] = 0.f; }
    __syncthreads();
    const int c = threadIdx.x % C;
    const int rstep = blockDim.x / C, rofs = threadIdx.x / C;
    float lw = 0.f, lb = 0.f;
    for (long long r = static_cast<long long>(blockIdx.x) * rstep + rofs; r < rows; r += static_cast<long long>(gridDim.x) * rstep) {
        const float m = cl_ld<T>(h0 + r * C + c) > 0.f ? cl_ld<T>(dh0 + r * ld_dh + c) * dh_scale : 0.f;
        const float xv = __fmul_rn(scale[r / L], x[r]);
        lw = fmaf(m, xv, lw);
        lb += m;
    }
    atomicAdd(&aw[c], lw);
    atomicAdd(&ab[c], lb);
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) { atomicAdd(dw_in + k, aw[k]); atomicAdd(db_in + k, ab[k]); }
}

// Loss backward (diffusion.py:60-63, :92-95): loss_b = lambda_b / n * sum_i (D_i - x_i)^2, D = clamp(c_skip xn + c_out F)
//   dF_i = upstream * (2 lambda_b / n) (D_i - x_i) c_out [-1 <= c_skip xn + c_out F <= 1]
__global__ void __launch_bounds__(256) dsm_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ x_noisy,
                                                           const float* __restrict__ F, const float* __restrict__ sigmas,
                                                           float sigma_data, float sd2, float upstream, float* __restrict__ dF,
                                                           int B, long long n_per, const float* __restrict__ upstream_b = nullptr,
                                                           const unsigned char* __restrict__ mask = nullptr) {
    const long long total = static_cast<long long>(B) * n_per;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(i / n_per);
        const float sg = sigmas[b];
        const PrecondCoef c = precond_coef(sg, sigma_data, sd2);
        const float pre = __fadd_rn(__fmul_rn(c.c_skip, x_noisy[i]), __fmul_rn(c.c_out, F[i]));
        const float D = clamp1(pre);
        const float lam = (sg * sg + sd2) / ((sg * sigma_data) * (sg * sigma_data));
        const bool pass = pre >= -1.0f && pre <= 1.0f;           // torch.clamp passes the gradient on the closed interval
        const float up = upstream_b ? upstream_b[b] : upstream;                  // per-sample upstream gradient (generic autograd path)
        const float wgt = (mask == nullptr || mask[i]) ? 1.0f : 0.01f;          // x_mask weights (diffusion.py:80-83)
        dF[i] = pass ? up * wgt * 2.0f * lam / static_cast<float>(n_per) * (D - x[i]) * c.c_out : 0.f;
    }
}

// Per-layer step-embedding projection backward (wavenet.py:108): p = Wp emb + bp, given dp [B][C]:
//   dWp[c][k] += sum_b dp[b][c] emb[b][k] ; dbp[c] += sum_b dp[b][c] ; demb[b][k] += sum_c dp[b][c] Wp[c][k]
// grid (layers), 512 threads (k), small B.
__global__ void __launch_bounds__(512) embproj_bwd_kernel(const float* __restrict__ dp /*[layers][B][C]*/, const float* __restrict__ emb,
                                                          const float* const* __restrict__ wp, float* __restrict__ grad_layers,
                                                          long long layer_stride, long long wp_off, long long bp_off,
                                                          float* __restrict__ demb, int B, int C) {
    const int layer = blockIdx.x, k = threadIdx.x;
    const float* dpl = dp + static_cast<long long>(layer) * B * C;
    const float* w = wp[layer];
    float* dw = grad_layers + layer * layer_stride + wp_off;      // this block owns the layer's dWp / dbp: plain stores
    float* db = grad_layers + layer * layer_stride + bp_off;
    for (int c = 0; c < C; ++c) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc = fmaf(dpl[b * C + c], emb[b * 512 + k], acc);
        dw[static_cast<long long>(c) * 512 + k] = acc;
    }
    for (int c = k; c < C; c += 512) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += dpl[b * C + c];
        db[c] = acc;
    }
    for (int b = 0; b < B; ++b) {
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(dpl[b * C + c], w[static_cast<long long>(c) * 512 + k], acc);
        atomicAdd(demb + b * 512 + k, acc);
    }
}

// bf16 backward bookkeeping: with U_l = (gradient wrt the output of block l) / sqrt(2) and S[l + 1] = per-sample column sums
// of U_l (S[0]: of U_{-1} = grad wrt the first block's input / sqrt(2)):
//   db2_l[0:C] = sum_b S[l+1][b] ; db2_l[C:2C] = column sums of dskip ; dp_l[b] = colsum_b(dx_l) = sqrt(2) S[l][b] - S[l+1][b]
// grid (layers), C threads.
__global__ void train_finalize_kernel(const float* __restrict__ S /*[layers+1][B][C]*/, const float* __restrict__ dbskip,
                                      float* __restrict__ grad_layers, long long layer_stride, long long b2_off,
                                      float* __restrict__ dp /*[layers][B][C]*/, int B, int C) {
    const int l = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float tot = 0.f;
        for (int b = 0; b < B; ++b) {
            const float s1 = S[(static_cast<long long>(l + 1) * B + b) * C + c];
            tot += s1;
            dp[(static_cast<long long>(l) * B + b) * C + c] = 1.41421356237309505f * S[(static_cast<long long>(l) * B + b) * C + c] - s1;
        }
        float* b2 = grad_layers + l * layer_stride + b2_off;
        b2[c] = tot;
        b2[C + c] = dbskip[c];
    }
}

// dF[r] *= upstream[r / L]   (per-sample weights of the upstream gradient)
__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ v, const float* __restrict__ w, long long n, int L) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        v[i] *= w[i / L];
}

// Step-embedding MLP backward (wavenet.py:88-92, :139-141): recomputes the forward of one sample, then
//   d_a2 = demb swish'(a2) ; dfc2 += d_a2 e1^T ; d_e1 = fc2^T d_a2 ; d_a1 = d_e1 swish'(a1) ; dfc1 += d_a1 e0^T
// one block of 512 threads per sample; weight gradients accumulate with atomics over samples.
__device__ __forceinline__ float swish_grad(float a) {
    const float s = 1.0f / (1.0f + expf(-a));
    return s * (1.0f + a * (1.0f - s));
}
__global__ void __launch_bounds__(512) embed_mlp_bwd_kernel(const float* __restrict__ c_noise, const float* __restrict__ w1,
                                                            const float* __restrict__ b1, const float* __restrict__ w2,
                                                            const float* __restrict__ b2, const float* __restrict__ demb,
                                                            float* __restrict__ dw1, float* __restrict__ db1,
                                                            float* __restrict__ dw2, float* __restrict__ db2) {
    __shared__ float e0[128], e1[512], da2[512], de1[512];
    const int b = blockIdx.x, j = threadIdx.x;
    const float t = c_noise[b];
    if (j < 128) {
        const int jj = j % 64;
        const float arg = t * expf(-static_cast<float>(jj) * 4.0f / 63.0f);
        e0[j] = (j < 64) ? sinf(arg) : cosf(arg);
    }
    __syncthreads();
    float a1 = b1[j];
    for (int k = 0; k < 128; ++k) a1 = fmaf(w1[j * 128 + k], e0[k], a1);
    e1[j] = a1 / (1.0f + expf(-a1));
    __syncthreads();
    float a2 = b2[j];
    for (int k = 0; k < 512; ++k) a2 = fmaf(w2[j * 512 + k], e1[k], a2);
    const float d2 = demb[b * 512 + j] * swish_grad(a2);
    da2[j] = d2;
    atomicAdd(db2 + j, d2);
    __syncthreads();
    for (int k = 0; k < 512; ++k) atomicAdd(dw2 + j * 512 + k, d2 * e1[k]);
    float acc = 0.f;
    for (int jj = 0; jj < 512; ++jj) acc = fmaf(w2[jj * 512 + j], da2[jj], acc);      // d_e1[j] = sum_jj fc2[jj][j] d_a2[jj]
    const float d1 = acc * swish_grad(a1);
    de1[j] = d1;
    atomicAdd(db1 + j, d1);
    for (int k = 0; k < 128; ++k) atomicAdd(dw1 + j * 128 + k, d1 * e0[k]);
}

// ------------------------------------------------------------------------------------------------
// Weight-norm chain rule (wavenet.py:44-51, scalar g): w = v g / ||v||, given dW in the packed layout
// [taps][Cin][Cout] (what wgrad produces) and v in torch layout [Cout][Cin][taps]:
//   dot = sum dW v ; dg = dot / ||v|| ; dv = (g / ||v||) (dW - v dot / ||v||^2)
// pass 1 (many blocks): dot via atomics ; pass 2: element-wise, writing dv in torch layout.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) wn_bwd_dot_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                         float* __restrict__ dot, int Cout, int Cin, int taps) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    float acc = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const int ci = static_cast<int>((i / Cout) % Cin);
        const int tap = static_cast<int>(i / (static_cast<long long>(Cout) * Cin));
        acc = fmaf(dw[i], v[(static_cast<long long>(co) * Cin + ci) * taps + tap], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(dot, acc);
}
__global__ void __launch_bounds__(256) wn_bwd_apply_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                           const float* __restrict__ g, const float* __restrict__ scale,
                                                           const float* __restrict__ dot, float* __restrict__ dv,
                                                           float* __restrict__ dg, int Cout, int Cin, int taps) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    const float s = scale[0];                       // g / ||v||
    const float nrm = g[0] / s;                     // ||v||
    const float d = dot[0];
    const float k = d / (nrm * nrm);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const int ci = static_cast<int>((i / Cout) % Cin);
        const int tap = static_cast<int>(i / (static_cast<long long>(Cout) * Cin));
        const long long j = (static_cast<long long>(co) * Cin + ci) * taps + tap;
        dv[j] = s * (dw[i] - v[j] * k);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) dg[0] = d / nrm;
}

// The same two passes for BOTH weight-normed convolutions of EVERY residual block in one launch each (grid.y = 2 * layers:
// job y = block y / 2, conv y % 2: 0 = dilated conv (3 taps), 1 = output projection): 148 tiny launches per step otherwise.
// Parameters and gradients of the blocks lie at constant strides in the device parameter copy / the flat gradient.
struct WnLayersArgs {
    const float* gw;        // [layers][(3 + 1) * C * 2C]: per block dW1 (3 taps) then dW2, packed [taps][Cin][Cout]
    long long gw_stride;
    const float *v1, *g1, *v2, *g2;     // block 0 in the device parameter copy (pieces padded to 64 floats); + layer * param_stride
    float *dv1, *dg1, *dv2, *dg2;       // block 0 in the flat gradient (unpadded state_dict order); + layer * grad_stride
    long long param_stride, grad_stride;
    const float* scale;     // d_scale: [1 + 2 l] dilated conv, [2 + 2 l] output projection
    float* dots;            // [2 * layers], zero before the dot pass
    int C;
};
__global__ void __launch_bounds__(256) wn_bwd_dot_layers_kernel(WnLayersArgs a) {
    const int layer = blockIdx.y >> 1, which = blockIdx.y & 1;
    const int Cout = 2 * a.C, Cin = a.C, taps = which ? 1 : 3;
    const float* dw = a.gw + layer * a.gw_stride + (which ? 3LL * Cin * Cout : 0);
    const float* v = (which ? a.v2 : a.v1) + layer * a.param_stride;
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    float acc = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const int ci = static_cast<int>((i / Cout) % Cin);
        const int tap = static_cast<int>(i / (static_cast<long long>(Cout) * Cin));
        acc = fmaf(dw[i], v[(static_cast<long long>(co) * Cin + ci) * taps + tap], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(a.dots + blockIdx.y, acc);
}
__global__ void __launch_bounds__(256) wn_bwd_apply_layers_kernel(WnLayersArgs a) {
    const int layer = blockIdx.y >> 1, which = blockIdx.y & 1;
    const int Cout = 2 * a.C, Cin = a.C, taps = which ? 1 : 3;
    const float* dw = a.gw + layer * a.gw_stride + (which ? 3LL * Cin * Cout : 0);
    const float* v = (which ? a.v2 : a.v1) + layer * a.param_stride;
    const float* g = (which ? a.g2 : a.g1) + layer * a.param_stride;
    float* dv = (which ? a.dv2 : a.dv1) + layer * a.grad_stride;
    float* dg = (which ? a.dg2 : a.dg1) + layer * a.grad_stride;
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    const float s = a.scale[1 + which + 2 * layer];       // g / ||v||
    const float nrm = g[0] / s;                            // ||v||
    const float d = a.dots[blockIdx.y];
    const float k = d / (nrm * nrm);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const int ci = static_cast<int>((i / Cout) % Cin);
        const int tap = static_cast<int>(i / (static_cast<long long>(Cout) * Cin));
        const long long j = (static_cast<long long>(co) * Cin + ci) * taps + tap;
        dv[j] = s * (dw[i] - v[j] * k);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) dg[0] = d / nrm;
}

// torch.optim.AdamW step on flat vectors (decoupled weight decay; bias-corrected moments):
//   p *= 1 - lr wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                    float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gi = g[i] * grad_scale;
        float pi = p[i] * (1.0f - lr * wd);
        const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= (lr / bc1) * (mi / denom);
        p[i] = pi;
    }
}

}  // namespace adb
