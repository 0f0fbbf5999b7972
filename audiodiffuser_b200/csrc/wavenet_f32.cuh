// fp32 (CUDA-core) kernels for the DiffWave residual stack: the exact-arithmetic path used for
// fp32 parity (<= 1e-5 relative to the reference, BASELINE.json north_star) and the small dense
// pieces both precisions share (step-embedding MLP, per-layer embedding projections, weight
// folding / packing, input and output projections).
//
// Reference: src/models/backbones/wavenet.py — WeightNorm :44-51, Conv :68-82,
// diffusion_embedding :88-92, ResidualBlock :107-115, ResidualGroup :136-151, WaveNetNoise :170-180.
//
// Activations are channels-last fp32 [B][L][C] (the reference is [B][C][L]); the conversion
// happens for free at the two ends of the network where C == 1.
#pragma once
#include "ptx.cuh"

namespace adb {

// ------------------------------------------------------------------------------------------------
// Weight folding:  w = v * g / ||v||_F with a scalar g   (wavenet.py:44-51, g is 0-dim: :30)
// ------------------------------------------------------------------------------------------------
struct WnJob {
    const float* v;   // [n]
    const float* g;   // [1]
    long long n;
};

// scale[j] = g_j / ||v_j||_F ; one block per weight-normed conv.
__global__ void __launch_bounds__(256) wn_scale_kernel(const WnJob* __restrict__ jobs, float* __restrict__ scale) {
    const WnJob job = jobs[blockIdx.x];
    double acc = 0.0;      // fp64 partial sums: the reference's torch.norm is fp32 pairwise; fp64 is closer to exact
    for (long long i = threadIdx.x; i < job.n; i += blockDim.x) {
        const double t = job.v[i];
        acc += t * t;
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) scale[blockIdx.x] = job.g[0] / static_cast<float>(sqrt(red[0]));
}

// out[tap][ci][co] = v[co][ci][tap] * scale     (torch Conv1d weight layout is [Cout][Cin][k])
__global__ void pack_conv_f32_kernel(const float* __restrict__ v, const float* __restrict__ scale, float* __restrict__ out,
                                     int Cout, int Cin, int taps) {
    const long long total = static_cast<long long>(Cout) * Cin * taps;
    const float s = scale ? scale[0] : 1.0f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int co = static_cast<int>(i % Cout);
        const int ci = static_cast<int>((i / Cout) % Cin);
        const int tap = static_cast<int>(i / (static_cast<long long>(Cout) * Cin));
        out[i] = v[(static_cast<long long>(co) * Cin + ci) * taps + tap] * s;
    }
}

// out[c][r] = in[r][c]  (used once at load time for the embedding projections)
__global__ void transpose_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    const long long total = static_cast<long long>(rows) * cols;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
        out[static_cast<long long>(c) * rows + r] = in[i];
    }
}

__global__ void scale_copy_kernel(const float* __restrict__ in, const float* __restrict__ scale, float* __restrict__ out,
                                  long long n) {
    const float s = scale ? scale[0] : 1.0f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = in[i] * s;
}

// ------------------------------------------------------------------------------------------------
// Generic channels-last dilated convolution as a tiled fp32 GEMM (taps in {1,3}).
//   out[b][t][co] = bias[co] + sum_{tap,ci} w[tap][ci][co] * X[b][t + (tap - taps/2) * dil][ci]
//   X[..][s][ci] = in_scale * in[b][s][ci] + padd[b][ci]   if 0 <= s < L, else 0   (zero "same" padding
//   applied AFTER the embedding add, wavenet.py:109-110)
// Tile 64 (t) x 64 (co) x 16 (k), 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
struct ConvF32Args {
    const float* in;     // [nb][L][Cin]
    const float* padd;   // [nb][Cin] or nullptr
    const float* w;      // [taps][Cin][ldw] (co contiguous)
    const float* bias;   // [Cout] or nullptr
    float* out;          // [nb][L][ldo]
    int nb, L, Cin, Cout, taps, dil;
    long long ldw;       // row pitch of w (>= Cout)
    long long ldo;       // row pitch of out (>= Cout)
    float in_scale;
    int relu;
    // optional per-batch-entry operand tables (device arrays of nb pointers): entry b then uses in_tab[b] / w_tab[b] /
    // bias_tab[b] (may hold NULL) / out_tab[b] instead of the strided operands above — one launch for many small GEMMs
    // with unrelated operands (the per-layer, per-tap step-embedding fold tables).
    const float* const* in_tab;
    const float* const* w_tab;
    const float* const* bias_tab;
    float* const* out_tab;
};

__global__ void __launch_bounds__(256) conv_cl_f32_kernel(ConvF32Args p) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;      // thread tile: rows ty*4.., cols tx*4..
    float acc[4][4] = {};
    const float* inb = p.in_tab ? p.in_tab[b] : p.in + static_cast<long long>(b) * p.L * p.Cin;
    const float* wb = p.in_tab ? p.w_tab[b] : p.w;
    const float* biasb = p.in_tab ? (p.bias_tab ? p.bias_tab[b] : nullptr) : p.bias;
    const float* paddb = p.padd ? p.padd + static_cast<long long>(b) * p.Cin : nullptr;
    const int ksteps = p.taps * p.Cin / BK;
    // A-load mapping: 64 rows x 16 k = 1024 floats, 4 per thread: row = tid/4, k4 = (tid%4)*4
    const int a_row = tid / 4, a_k = (tid % 4) * 4;
    // B-load mapping: 16 k x 64 co = 1024 floats: k = tid/16, co4 = (tid%16)*4
    const int b_k = tid / 16, b_c = (tid % 16) * 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        const int kglob = ks * BK;
        const int tap = kglob / p.Cin;
        const int ci0 = kglob % p.Cin;
        {
            const int t = t0 + a_row;
            const int s = t + (tap - p.taps / 2) * p.dil;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < p.L && s >= 0 && s < p.L) {
                v = *reinterpret_cast<const float4*>(inb + static_cast<long long>(s) * p.Cin + ci0 + a_k);
                v.x *= p.in_scale; v.y *= p.in_scale; v.z *= p.in_scale; v.w *= p.in_scale;
                if (paddb) {
                    const float4 e = *reinterpret_cast<const float4*>(paddb + ci0 + a_k);
                    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
                }
            }
            As[a_k + 0][a_row] = v.x; As[a_k + 1][a_row] = v.y; As[a_k + 2][a_row] = v.z; As[a_k + 3][a_row] = v.w;
        }
        {
            const float4 v = *reinterpret_cast<const float4*>(wb + static_cast<long long>(kglob + b_k) * p.ldw + n0 + b_c);
            *reinterpret_cast<float4*>(&Bs[b_k][b_c]) = v;
        }
        __syncthreads();
        // accumulate each 16-deep slice separately, then add it to the running sum: rounding error grows
        // with K/16 + 16 terms instead of K (keeps the 36-layer fp32 path inside 1e-5 of the reference)
        float part[4][4] = {};
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bb = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
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
    if (biasb) {
        const float4 bb = *reinterpret_cast<const float4*>(biasb + n0 + tx * 4);
        bias[0] = bb.x; bias[1] = bb.y; bias[2] = bb.z; bias[3] = bb.w;
    }
    float* outb = p.in_tab ? p.out_tab[b] : p.out + static_cast<long long>(b) * p.L * p.ldo;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t >= p.L) continue;
        float4 v = make_float4(acc[i][0] + bias[0], acc[i][1] + bias[1], acc[i][2] + bias[2], acc[i][3] + bias[3]);
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        *reinterpret_cast<float4*>(outb + static_cast<long long>(t) * p.ldo + n0 + tx * 4) = v;
    }
}

inline cudaError_t conv_cl_f32(const ConvF32Args& p, cudaStream_t stream) {
    // preconditions are checked by the caller (api.cu): Cin % 16 == 0, Cout % 64 == 0, 16B-aligned rows
    dim3 grid((p.L + 63) / 64, p.Cout / 64, p.nb);
    conv_cl_f32_kernel<<<grid, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Step embedding MLP (wavenet.py:88-92, :139-141): emb[b] = swish(fc2(swish(fc1([sin, cos](t f_j)))))
// One block of 512 threads per sample.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(512) embed_mlp_kernel(const float* __restrict__ c_noise, const float* __restrict__ w1,
                                                        const float* __restrict__ b1, const float* __restrict__ w2,
                                                        const float* __restrict__ b2, float* __restrict__ emb) {
    __shared__ float e0[128];
    __shared__ float e1[512];
    const int b = blockIdx.x, j = threadIdx.x;
    const float t = c_noise[b];
    if (j < 128) {
        const int jj = j % 64;
        const float arg = t * expf(-static_cast<float>(jj) * 4.0f / 63.0f);
        e0[j] = (j < 64) ? sinf(arg) : cosf(arg);
    }
    __syncthreads();
    float acc = b1[j];
    for (int k = 0; k < 128; ++k) acc = fmaf(w1[j * 128 + k], e0[k], acc);
    e1[j] = swish_f(acc);
    __syncthreads();
    acc = b2[j];
    for (int k = 0; k < 512; ++k) acc = fmaf(w2[j * 512 + k], e1[k], acc);
    emb[b * 512 + j] = swish_f(acc);
}

// p[layer][b][c] = bp[layer][c] + sum_k Wp[layer][c][k] * emb[b][k]   (wavenet.py:108)
// grid (B, layers), C threads.
__global__ void embed_proj_kernel(const float* __restrict__ emb, const float* const* __restrict__ wp,
                                  const float* const* __restrict__ bp, float* __restrict__ out, int B, int C) {
    __shared__ float e[512];
    const int b = blockIdx.x, layer = blockIdx.y;
    for (int k = threadIdx.x; k < 512; k += blockDim.x) e[k] = emb[b * 512 + k];
    __syncthreads();
    const float* w = wp[layer];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = bp[layer][c];
        const float4* row = reinterpret_cast<const float4*>(w + static_cast<long long>(c) * 512);
        for (int k = 0; k < 128; ++k) {
            const float4 v = row[k];
            acc = fmaf(v.x, e[4 * k], acc); acc = fmaf(v.y, e[4 * k + 1], acc);
            acc = fmaf(v.z, e[4 * k + 2], acc); acc = fmaf(v.w, e[4 * k + 3], acc);
        }
        out[(static_cast<long long>(layer) * B + b) * C + c] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Input projection (wavenet.py:172-174): h0[b][t][c] = relu(w_in[c] * (scale_b * x[b][t]) + b_in[c]).
// scale_b = c_in(sigma_b) folds the EDM input scaling (diffusion.py:50) into the same pass.
// OutT = float (fp32 path) or __nv_bfloat16 (tensor-core path). One thread per (b, t, 8 channels).
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256) in_proj_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                      int scale_stride, const float* __restrict__ w_in,
                                                      const float* __restrict__ b_in, void* __restrict__ out, int B, int L,
                                                      int C) {
    // thread -> fixed channel group g (its 8 weights / biases live in registers), rows strided; 32-bit index math only
    const int cg = C / 8;
    const int g = threadIdx.x % cg;
    const int rows_per_block = blockDim.x / cg;                  // C <= 2048: at least one row per block pass
    if (static_cast<int>(threadIdx.x) >= rows_per_block * cg) return;
    float wv[8], bv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { wv[k] = w_in[g * 8 + k]; bv[k] = b_in[g * 8 + k]; }
    const int total_rows = B * L;                                // < 2^31 (checked by the caller)
    for (int bt = blockIdx.x * rows_per_block + threadIdx.x / cg; bt < total_rows; bt += gridDim.x * rows_per_block) {
        const int b = bt / L;
        const float s = scale ? scale[b * scale_stride] : 1.0f;
        const float xv = __fmul_rn(s, x[bt]);
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fmaxf(fmaf(wv[k], xv, bv[k]), 0.f);
        const long long i = static_cast<long long>(bt) * cg + g;
        if constexpr (BF16) {
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            reinterpret_cast<uint4*>(out)[i] = o;
        } else {
            float4* o = reinterpret_cast<float4*>(out) + i * 2;
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

// z[b][t][c] = sigmoid(y[b][t][c]) * tanh(y[b][t][C + c])   (wavenet.py:111-112: gate = FIRST half)
__global__ void __launch_bounds__(256) gate_f32_kernel(const float* __restrict__ y, float* __restrict__ z, long long rows,
                                                       int C) {
    const long long total = rows * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / C;
        const int c = static_cast<int>(i % C);
        const float g = y[r * 2 * C + c], f = y[r * 2 * C + C + c];
        z[i] = (1.0f / (1.0f + expf(-g))) * tanhf(f);
    }
}

// h_out = (h_in + o[:, :C]) / sqrt(2) ; skip (+)= o[:, C:]   (wavenet.py:114-115, :149)
__global__ void __launch_bounds__(256) res_skip_f32_kernel(const float* __restrict__ h_in, const float* __restrict__ o,
                                                           float* __restrict__ h_out, float* __restrict__ skip,
                                                           long long rows, int C, int first) {
    const long long total = rows * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / C;
        const int c = static_cast<int>(i % C);
        const float res = o[r * 2 * C + c], sk = o[r * 2 * C + C + c];
        h_out[i] = __fdiv_rn(__fadd_rn(h_in[i], res), 1.41421356237309504880f);
        skip[i] = first ? sk : __fadd_rn(skip[i], sk);
    }
}

// Output projection (wavenet.py:179): out[b][t] = b_out + sum_c w_out[c] * s[b][t][c]; one warp per row.
__global__ void __launch_bounds__(256) out_proj_f32_kernel(const float* __restrict__ s, const float* __restrict__ w_out,
                                                           const float* __restrict__ b_out, float* __restrict__ out,
                                                           long long rows, int C) {
    const long long warp = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(w_out[c], s[warp * C + c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[warp] = acc + b_out[0];
}

inline unsigned grid_for(long long total, int block = 256, long long cap = 148LL * 32) {
    long long g = (total + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<unsigned>(g);
}

}  // namespace adb
