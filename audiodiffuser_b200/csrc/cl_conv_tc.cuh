// Tensor-core (tcgen05 / TMEM / TMA) version of the generic channels-last GEMM-convolution of cl_ops.cuh:
// bf16 activations [B][L][C], bf16 weights, fp32 accumulation in TMEM, fused bias / activation / residual epilogue.
// Used for every convolution and 1x1 / Linear layer of the U-Net whose K (= Cin per tap) is a multiple of 64
// (unet1d.py: ConvBlock1d.project :186-193, ResnetBlock1d.to_out :291-295, Downsample1d :214-225, Upsample1d
// :246-255, FeedForward1d :49-61, Attention projections attention_utils.py:95-110).
//
// Persistent CTA PAIRS (cta_group::2, one pair per two SMs) walk (m-tile pair, n-tile) groups: each CTA owns M = 128 rows of
// one sample, the pair runs them as ONE M = 256 MMA stream issued by rank 0; N tile = 64 / 128 / 256 columns. Per K-block (64
// input channels of one tap) every CTA's producer warp TMA-loads its activation box [128 rows][64 ch] at row offset
// off0 + tap*dil (out-of-range rows are zero-filled by TMA = the conv's zero padding) and only HALF of the weight box
// [NT / 2][64] (measured: the single-CTA form re-read every weight box per 128 rows and was bound by L2 / shared-memory
// bandwidth, not by the tensor pipe); 4 tcgen05.mma (K = 16 each) per K-block. Two TMEM accumulators alternate between
// groups so a tile's epilogue overlaps the next tile's MMAs. Warps: 0 .. EPI_WARPS-1 epilogue, then the TMA producer and the
// MMA issuer on the two highest warp ids (the scheduler arbitrates highest-warp-id-first).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "ptx.cuh"
#include "cl_ops.cuh"

namespace adb {

constexpr int CT_STAGES = 6;
constexpr int CT_A_BYTES = 128 * 64 * 2;            // 16 KB
constexpr int CT_B_BYTES = 128 * 64 * 2;            // 16 KB: this CTA's half of a weight box (NT = 256; smaller tiles use a prefix)
constexpr int CT_STAGE_BYTES = CT_A_BYTES + CT_B_BYTES;
constexpr int CT_THREADS = 192;                     // EPI_WARPS = 4 instantiation: warps 0..3 epilogue, warp 4 TMA, warp 5 MMA
constexpr int CT_SMEM_BYTES = CT_STAGES * CT_STAGE_BYTES + 1024 /*bias*/ + 16 * 8 + 16;

// Activation applied to a 32-column register chunk with the activation as a COMPILE-TIME constant: the switch is taken once
// per chunk, never per element (left to the optimiser, a per-element uniform branch on the runtime value can survive the
// unrolling and costs 30 % of the small-K convolutions).
template <int ACT>
__device__ __forceinline__ void ct_act32(float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        if constexpr (ACT == CL_ACT_RELU) v[i] = fmaxf(v[i], 0.f);
        else if constexpr (ACT == CL_ACT_SILU) v[i] = __fdividef(v[i], 1.0f + __expf(-v[i]));
        else if constexpr (ACT == CL_ACT_GELU) v[i] = 0.5f * v[i] * (1.0f + erff(v[i] * 0.70710678118654752f));
    }
}

enum CtWaitSite : uint32_t { SITE_CT_EMPTY = 20, SITE_CT_FULL = 21, SITE_CT_TEMPTY = 22, SITE_CT_TFULL = 23 };

struct ClConvTcParams {
    ClConvArgs a;
    int NT;                 // N tile (64, 128 or 256)
    int tiles_per_b;        // ceil(rows / 128)
    int tiles_m;            // B * tiles_per_b
    int tiles_n;            // N / NT
    int kb_per_tap;         // Cin / 64
    int epi_warps;          // 4 (block of 192 threads) or 8 (block of 320): the fused gate epilogues need the second set
    int nkb;                // K-blocks to run: taps * kb_per_tap, or fewer when the LAST tap's trailing blocks are all-zero weights (the
                            // strided convolution on the [L/f][f*C] view, unet1d.py:214-225: its third coarse tap holds one fine tap)
    int kb1;                // K-blocks per tap that come from the first input; the rest from the second (channel concatenation
                            // of two tensors, unet1d.py:552-556); = kb_per_tap for a single input
};

// EPI_WARPS = 4 (192 threads) for the plain / transposed / WAVdec stores; EPI_WARPS = 8 (320 threads) with FUSED = true adds
// the training step's gate-forward / gate-derivative epilogues, which are epilogue-bound with four warps.
// One thread's 32-column bf16 chunk of a row = 64 contiguous, 64-byte aligned bytes: two 256-bit accesses.
__device__ __forceinline__ void ct_store64(__nv_bfloat16* dst, const uint32_t (&v)[16]) {
    const uint32_t lo[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
    const uint32_t hi[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
    stg256(dst, lo);
    stg256(dst + 16, hi);
}
__device__ __forceinline__ void ct_load64(const __nv_bfloat16* src, uint32_t (&v)[16]) {
    uint32_t lo[8], hi[8];
    ldg256(src, lo);
    ldg256(src + 16, hi);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = lo[i]; v[8 + i] = hi[i]; }
}

template <int EPI_WARPS, bool FUSED>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
cl_conv_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_in2,
                  const __grid_constant__ CUtensorMap tm_w, const ClConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_bias = reinterpret_cast<float*>(smem + CT_STAGES * CT_STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CT_STAGES * CT_STAGE_BYTES + 1024);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + CT_STAGES;
    uint64_t* bar_tfull = bars + 2 * CT_STAGES;
    uint64_t* bar_tempty = bar_tfull + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ClConvArgs& a = p.a;
    constexpr int W_PROD = EPI_WARPS, W_MMA = EPI_WARPS + 1;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;

    if (warp == W_PROD && lane == 0) { tma_prefetch_desc(&tm_in); tma_prefetch_desc(&tm_in2); tma_prefetch_desc(&tm_w); }
    if (warp == W_MMA) {
        if (lane == 0) {
            // bar_full / bar_tempty: rank 0's copies are the live ones; bar_empty / bar_tfull: per CTA (multicast commits)
            for (int s = 0; s < CT_STAGES; ++s) { mbar_init(&bar_full[s], 2); mbar_init(&bar_empty[s], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 2 * EPI_WARPS); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair(s_tmem, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    pdl_wait();                    // programmatic dependent launch: the prologue above overlaps the previous kernel's tail
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;
    const int nkb = p.nkb;
    const int pair_id = static_cast<int>(blockIdx.x) >> 1, num_pairs = static_cast<int>(gridDim.x) >> 1;
    const int total_groups = ((p.tiles_m + 1) >> 1) * p.tiles_n;     // a group = two adjacent m-tiles (one per CTA) of one n-tile
    const uint32_t stage_tx = CT_A_BYTES + static_cast<uint32_t>(p.NT) * 64u;      // this CTA's bytes per stage

    if (warp == W_PROD) {
        uint32_t stage = 0, phase = 0;
        for (int grp = pair_id; grp < total_groups; grp += num_pairs) {
            const int tm = (grp / p.tiles_n) * 2 + rank, tn = grp % p.tiles_n;
            const int b = tm / p.tiles_per_b, t0 = (tm % p.tiles_per_b) * 128, n0 = tn * p.NT;   // b >= B for a padding tile: zero fill
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&bar_empty[stage], phase ^ 1, SITE_CT_EMPTY, stage);
                if (lane == 0) {
                    uint8_t* sa = smem + stage * CT_STAGE_BYTES;
                    const int tap = kb / p.kb_per_tap, cib = kb % p.kb_per_tap;
                    if (leader) mbar_arrive_expect_tx(&bar_full[stage], 2 * stage_tx);     // both CTAs' bytes
                    else        mbar_arrive_cluster(&bar_full[stage], 0);
                    if (cib < p.kb1) tma_load_3d_pair(sa, &tm_in, &bar_full[stage], cib * 64, t0 + a.off0 + tap * a.dil, b);
                    else             tma_load_3d_pair(sa, &tm_in2, &bar_full[stage], (cib - p.kb1) * 64, t0 + a.off0 + tap * a.dil, b);
                    tma_load_2d_pair(sa + CT_A_BYTES, &tm_w, &bar_full[stage], 0, kb * a.N + n0 + rank * (p.NT >> 1));
                }
                __syncwarp();
                if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == W_MMA) {
        uint32_t stage = 0, phase = 0, it = 0;
        const uint32_t idesc = umma_idesc_pair_bf16(static_cast<uint32_t>(p.NT));
        if (leader)
        for (int grp = pair_id; grp < total_groups; grp += num_pairs, ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            mbar_wait(&bar_tempty[buf], (use & 1) ^ 1, SITE_CT_TEMPTY, buf);
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_base + buf * 256;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&bar_full[stage], phase, SITE_CT_FULL, stage);
                tc_fence_after_sync();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * CT_STAGE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss_pair(d_tmem, umma_desc_sw128_kmajor(sa + k * 32), umma_desc_sw128_kmajor(sa + CT_A_BYTES + k * 32),
                                          idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_pair_mc(&bar_empty[stage], 3);
                    if (kb == nkb - 1) umma_commit_pair_mc(&bar_tfull[buf], 3);
                }
                __syncwarp();
                if (++stage == CT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        constexpr int nhalf = EPI_WARPS >> 2;          // 1: this warp handles every column chunk ; 2: half of them
        const int half = warp >> 2;                    // which half of the tile's column chunks this warp handles
        constexpr int epi_threads = 32 * EPI_WARPS;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a.out);
        const __nv_bfloat16* res = static_cast<const __nv_bfloat16*>(a.res);
        uint32_t it = 0;
        int cur_n0 = -1;
        for (int grp = pair_id; grp < total_groups; grp += num_pairs, ++it) {
            const int tm = (grp / p.tiles_n) * 2 + rank, tn = grp % p.tiles_n;
            const int b = tm / p.tiles_per_b, t = (tm % p.tiles_per_b) * 128 + row, n0 = tn * p.NT;
            const uint32_t buf = it & 1, use = it >> 1;
            if (n0 != cur_n0) {                     // bias slice of this n-tile (epilogue warps only: named barrier 1)
                asm volatile("bar.sync 1, %0;" ::"n"(epi_threads) : "memory");
                for (int i = threadIdx.x; i < p.NT; i += epi_threads) s_bias[i] = a.bias ? a.bias[n0 + i] : 0.f;
                asm volatile("bar.sync 1, %0;" ::"n"(epi_threads) : "memory");
                cur_n0 = n0;
            }
            mbar_wait(&bar_tfull[buf], use & 1, SITE_CT_TFULL, buf);
            tc_fence_after_sync();
            const bool row_ok = tm < p.tiles_m && t < a.rows;       // tm == tiles_m: the padding tile of an odd tile count
            const long long grow = static_cast<long long>(b) * a.rows + t;      // global row
            if (FUSED && a.mode == CL_MODE_GATE_FWD) {
                // tile = [128 gate | 128 filter] of channels 128 tn .. 128 tn + 127
                const int C = a.N / 2;
                __nv_bfloat16* yout = static_cast<__nv_bfloat16*>(a.aux_out);
                for (int cc = (nhalf == 2 ? 2 * half : 0); cc < (nhalf == 2 ? 2 * half + 2 : 4); ++cc) {
                    uint32_t rg[32], rf[32];
                    tmem_ld_32x32(t_lane + buf * 256 + cc * 32, rg);
                    tmem_ld_32x32(t_lane + buf * 256 + 128 + cc * 32, rf);
                    tmem_ld_wait();
                    const int c0 = 128 * tn + cc * 32;
                    uint32_t pg[16], pf[16], pz[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float g0 = __uint_as_float(rg[i]) + s_bias[cc * 32 + i], g1 = __uint_as_float(rg[i + 1]) + s_bias[cc * 32 + i + 1];
                        const float f0 = __uint_as_float(rf[i]) + s_bias[128 + cc * 32 + i], f1 = __uint_as_float(rf[i + 1]) + s_bias[128 + cc * 32 + i + 1];
                        pg[i >> 1] = pack_bf16x2(g0, g1);
                        pf[i >> 1] = pack_bf16x2(f0, f1);
                        pz[i >> 1] = pack_bf16x2(__fdividef(1.0f, 1.0f + __expf(-g0)) * tanh_fast(f0),
                                                 __fdividef(1.0f, 1.0f + __expf(-g1)) * tanh_fast(f1));
                    }
                    if (row_ok) {                   // 64 contiguous bytes per destination: two full-sector 256-bit stores each
                        ct_store64(yout + grow * a.N + c0, pg);
                        ct_store64(yout + grow * a.N + C + c0, pf);
                        ct_store64(out + grow * C + c0, pz);
                    }
                }
            } else if (FUSED && a.mode == CL_MODE_GATE_BWD) {
                const int C = a.N;
                const __nv_bfloat16* yin = static_cast<const __nv_bfloat16*>(a.aux_in);
                const int nch = p.NT / 32;
                const int c_lo = nhalf == 2 ? half * ((nch + 1) / 2) : 0, c_hi = (nhalf == 2 && !half) ? (nch + 1) / 2 : nch;
                for (int cc = c_lo; cc < c_hi; ++cc) {
                    uint32_t r[32];
                    tmem_ld_32x32(t_lane + buf * 256 + cc * 32, r);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    const int n = n0 + cc * 32;
                    uint32_t wg[16], wf[16], og[16], of[16];
                    ct_load64(yin + grow * 2 * C + n, wg);
                    ct_load64(yin + grow * 2 * C + C + n, wf);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        float dgv[2], dfv[2];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const float g = __uint_as_float(h ? (wg[e] & 0xFFFF0000u) : (wg[e] << 16));
                            const float f = __uint_as_float(h ? (wf[e] & 0xFFFF0000u) : (wf[e] << 16));
                            const float d = __uint_as_float(r[2 * e + h]);
                            const float sg = __fdividef(1.0f, 1.0f + __expf(-g)), th = tanh_fast(f);
                            dgv[h] = d * th * sg * (1.0f - sg);
                            dfv[h] = d * sg * (1.0f - th * th);
                        }
                        og[e] = pack_bf16x2(dgv[0], dgv[1]);
                        of[e] = pack_bf16x2(dfv[0], dfv[1]);
                    }
                    ct_store64(out + grow * 2 * C + n, og);
                    ct_store64(out + grow * 2 * C + C + n, of);
                }
            } else
            for (int cc = (nhalf == 2 ? half * ((p.NT / 32 + 1) / 2) : 0); cc < ((nhalf == 2 && !half) ? (p.NT / 32 + 1) / 2 : p.NT / 32); ++cc) {
                uint32_t r[32];
                tmem_ld_32x32(t_lane + buf * 256 + cc * 32, r);
                tmem_ld_wait();
                const int n = n0 + cc * 32;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + s_bias[cc * 32 + i];
                switch (a.act) {
                    case CL_ACT_RELU: ct_act32<CL_ACT_RELU>(v); break;
                    case CL_ACT_SILU: ct_act32<CL_ACT_SILU>(v); break;
                    case CL_ACT_GELU: ct_act32<CL_ACT_GELU>(v); break;
                    default: break;
                }
                long long o;
                bool ok = row_ok;
                if (a.cf_cout > 0) {
                    // WAVdec: columns [0, ups * cf_cout) of this row are output samples t*ups + phase - shift of cf_cout
                    // channels, written to the fp32 channels-first waveform (unet1d.py:596-622)
                    if (ok && n == 0 && a.cf_cout == 2 && a.ups == 16 && t * 16 - a.shift >= 0 && t * 16 - a.shift + 16 <= a.L_out &&
                        ((t * 16 - a.shift) & 7) == 0 && (a.L_out & 7) == 0) {
                        // stereo, stride 16 (config 4): this row's 16 output samples of a channel are 64 contiguous, 32-byte aligned
                        // bytes -> two 256-bit stores per channel instead of 16 scalar ones (the scalar form ran at 1.3 TB/s)
                        float* y = static_cast<float*>(a.out) + static_cast<long long>(b) * 2 * a.L_out + (t * 16 - a.shift);
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            uint32_t lo[8], hi[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) { lo[i] = __float_as_uint(v[2 * i + c]); hi[i] = __float_as_uint(v[2 * (8 + i) + c]); }
                            stg256(y + static_cast<long long>(c) * a.L_out, lo);
                            stg256(y + static_cast<long long>(c) * a.L_out + 8, hi);
                        }
                    } else if (ok && n < a.ups * a.cf_cout) {
                        float* y = static_cast<float*>(a.out);
                        const int ncols = a.ups * a.cf_cout - n;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i < ncols) {
                                const int col = n + i, ph = col / a.cf_cout, c = col % a.cf_cout;
                                const int to = t * a.ups + ph - a.shift;
                                if (to >= 0 && to < a.L_out) y[(static_cast<long long>(b) * a.cf_cout + c) * a.L_out + to] = v[i];
                            }
                        }
                    }
                    continue;
                }
                if (a.ups == 0) {
                    o = grow * (a.ldo ? a.ldo : a.N) + n;
                    if (res && ok) {
                        uint32_t w[16];
                        ct_load64(res + grow * (a.res_ld ? a.res_ld : a.N) + n, w);
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            v[2 * e] += __uint_as_float(w[e] << 16);
                            v[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
                        }
                    }
                    if (a.out_scale != 0.f) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] *= a.out_scale;
                    }
                } else {
                    const int cout = a.N / a.ups, ph = n / cout, c = n % cout;
                    const int to = t * a.ups + ph - a.shift;
                    ok = ok && to >= 0 && to < a.L_out;
                    o = (static_cast<long long>(b) * a.L_out + to) * cout + c;
                }
                if (ok) {
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
                    ct_store64(out + o, pk);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bar_tempty[buf], 0);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == W_MMA) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// WAVdec1d filter bank [F][Cout][W = 2S] (torch ConvTranspose1d layout) -> the packed B operand of the 2-tap
// GEMM-convolution Y[i][phase*Cout + c] = sum_f h[i][f] w[f][c][phase] + h[i-1][f] w[f][c][phase + S], N padded to 64
__global__ void cl_pack_wavdec_tc_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int F, int Cout, int W, int S) {
    const int kbt = F / 64;
    const int total = 2 * kbt * 64 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i & 63, n = (i >> 6) & 63, kb = i >> 12;
        const int tap = kb / kbt, f = (kb % kbt) * 64 + k;
        float v = 0.f;
        if (n < S * Cout) {
            const int ph = n / Cout, c = n % Cout;
            v = w[(static_cast<long long>(f) * Cout + c) * W + ph + tap * S];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// fp32 [taps][Cin][N] -> bf16 blocks [kb = tap * Cin/64 + cib][N][64] (K-major rows of 128 bytes, what one TMA box
// with the 128-byte swizzle loads)
template <typename OutT>
__global__ void cl_pack_conv_tc_kernel(const float* __restrict__ w, OutT* __restrict__ out, int Cin, int N, int taps) {
    const long long total = static_cast<long long>(taps) * Cin * N;
    const int kbt = Cin / 64;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int k = static_cast<int>(i & 63);
        const int n = static_cast<int>((i >> 6) % N);
        const int kb = static_cast<int>(i / (64LL * N));
        const int tap = kb / kbt, cib = kb % kbt;
        const float v = w[(static_cast<long long>(tap) * Cin + cib * 64 + k) * N + n];
        if constexpr (sizeof(OutT) == 2 && !std::is_same<OutT, __nv_bfloat16>::value) out[i] = __float2half_rn(v);      // fp16 blocks (cl_conv3_gn_tc_kernel)
        else out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace adb
