// Tensor-core weight-gradient GEMM of the training step (bf16 operands, fp32 accumulation in TMEM):
//
//   out[tap][i][j] += sum_{b, t} A[b][t + shift_tap][i] * G[b][t][j]        (rows of A outside [0, L) read as zero)
//
// A: layer input [B][L][Ca] bf16 (h + p for the dilated conv, z for the 1x1 conv), G: output gradient [B][L][Cg] bf16.
// This is the gradient of a channels-last convolution with respect to its weights (what autograd computes for
// wavenet.py:110 / :113): a GEMM whose K dimension is batch x time, so BOTH operands are "MN-major" in the tcgen05
// sense (the contiguous memory dimension is M / N, not K). TMA boxes of [64 channels][64 time steps] with the
// 128-byte swizzle land in exactly the canonical MN-major SWIZZLE_128B layout (64 K-rows of 128 bytes; descriptor
// stride-byte-offset 1024 between 8-row groups, leading-byte-offset = box size between 64-channel blocks), so no
// transpose pass is needed. The dilated tap is a row offset of the A box; TMA zero fill supplies the padding.
//
// One CTA = one (128 x 256) output tile of one tap and a K-slab (every `splits`-th 64-step chunk of the batch);
// 4-stage TMA ring, one MMA-issuing thread, single TMEM accumulator; the epilogue adds the tile to `out` with fp32
// atomics (each output element receives `splits` atomic adds per launch).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace adb {

constexpr int WG_STAGES = 4;
constexpr int WG_KT = 64;                                 // time steps per stage
constexpr int WG_A_BYTES = 2 * 64 * WG_KT * 2;            // 2 boxes [64 t][64 ch] bf16 = 16 KB (M = 128)
constexpr int WG_G_BYTES = 4 * 64 * WG_KT * 2;            // 4 boxes = 32 KB (N = 256)
constexpr int WG_BOX_BYTES = 64 * WG_KT * 2;              // 8 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_G_BYTES;
constexpr int WG_THREADS = 192;                           // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 16 * 8 + 16;

enum WgWaitSite : uint32_t { SITE_WG_EMPTY = 30, SITE_WG_FULL = 31, SITE_WG_DONE = 32 };

struct WgradTcParams {
    float* out;             // [taps][Ca][ldo] fp32, accumulated into
    int B, L, Ca, Cg, taps, dil;      // tap shift = (tap - taps / 2) * dil
    long long ldo;
    int tiles_m, tiles_n;   // Ca / 128, Cg / 256
    int splits;             // CTAs sharing one tile
    int chunks_per_b;       // ceil(L / 64)
    float scale;            // multiplies the result (e.g. the skip-sum scale of wavenet.py:151)
    // Column sums of G as a by-product (the CTAs of the first m-tile and the unshifted tap already stage every G tile of their K
    // slab; their four epilogue warps, idle during the main loop, add the rows up from shared memory):
    //   colsum_mode 1: colsum_out[b][n] += sum_t G[b][t][n]  for the first n-tile only (per-sample sums of the first 256 columns)
    //   colsum_mode 2: colsum_out[k][b][n] += sums over all rows (k = 0), rows t < colsum_d (k = 1), rows t >= L - colsum_d (k = 2)
    float* colsum_out;
    int colsum_mode, colsum_d;
};

// MN-major SWIZZLE_128B shared-memory descriptor: 64-element (128-byte) rows along M/N, K advances one row per element;
// SBO = 1024 B between 8-row K groups, LBO = bytes between consecutive 64-element blocks along M/N.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// kind::f16 instruction descriptor, bf16 x bf16 -> fp32, A and B both MN-major (bits 15 / 16)
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_mn(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_g, const WgradTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + WG_STAGES;
    uint64_t* bar_done = bars + 2 * WG_STAGES;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int job = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int tn = job % p.tiles_n, tm = (job / p.tiles_n) % p.tiles_m, tap = job / (p.tiles_n * p.tiles_m);
    // every CTA of an n-tile stages the same G tiles (whatever its m-tile and tap): they share the column sums, CTA (tm, tap) taking
    // every sum_ways-th chunk of the K slab, so that the adds stay far below the MMA time of a stage
    const bool do_sum = p.colsum_mode != 0 && (p.colsum_mode == 2 || tn == 0);
    const int sum_ways = p.tiles_m * p.taps, sum_me = tm * p.taps + tap;
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_g); }
    if (warp == 1) {
        if (lane == 0) {
            // a stage is free once its MMAs have completed AND (column-sum CTAs) the four epilogue warps have read its G tile
            for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], do_sum ? 5 : 1); }
            mbar_init(bar_done, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(s_tmem, 256);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;

    const int shift = (tap - p.taps / 2) * p.dil;
    const int m0 = tm * 128, n0 = tn * 256;
    const int total_chunks = p.B * p.chunks_per_b;
    const int my_chunks = (total_chunks - split + p.splits - 1) / p.splits;      // chunks split, split + splits, ...

    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int i = 0; i < my_chunks; ++i) {
            const int chunk = split + i * p.splits;
            const int b = chunk / p.chunks_per_b, t0 = (chunk % p.chunks_per_b) * WG_KT;
            mbar_wait(&bar_empty[stage], phase ^ 1, SITE_WG_EMPTY, stage);
            if (lane == 0) {
                uint8_t* sa = smem + stage * WG_STAGE_BYTES;
                uint8_t* sg = sa + WG_A_BYTES;
                mbar_arrive_expect_tx(&bar_full[stage], WG_STAGE_BYTES);
                for (int j = 0; j < 2; ++j) tma_load_3d(sa + j * WG_BOX_BYTES, &tm_a, &bar_full[stage], m0 + 64 * j, t0 + shift, b);
                for (int j = 0; j < 4; ++j) tma_load_3d(sg + j * WG_BOX_BYTES, &tm_g, &bar_full[stage], n0 + 64 * j, t0, b);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        uint32_t stage = 0, phase = 0;
        constexpr uint32_t IDESC = umma_idesc_bf16_f32_mn(128, 256);
        for (int i = 0; i < my_chunks; ++i) {
            mbar_wait(&bar_full[stage], phase, SITE_WG_FULL, stage);
            tc_fence_after_sync();
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
                const uint32_t sg = sa + WG_A_BYTES;
#pragma unroll
                for (int k = 0; k < WG_KT / 16; ++k)
                    umma_bf16_ss(tmem_base, umma_desc_sw128_mnmajor(sa + k * 2048, WG_BOX_BYTES),
                                 umma_desc_sw128_mnmajor(sg + k * 2048, WG_BOX_BYTES), IDESC, (i | k) != 0 ? 1u : 0u);
                umma_commit(&bar_empty[stage]);
                if (i == my_chunks - 1) umma_commit(bar_done);
            }
            __syncwarp();
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (my_chunks > 0) {
        if (do_sum) {
            // ---- column sums of G, straight from the staged boxes: warp w owns box w (64 channels); a lane reads 16 bytes (8
            //      channels) of a row, 8 lanes cover a 128-byte row and the warp 4 rows per instruction (conflict-free). Lanes
            //      l, l + 8, l + 16, l + 24 hold partial sums of the same channels and are combined when a sample ends ----
            const int w = warp - 2;
            const int cch = lane & 7, rsub = lane >> 3;          // 16-byte chunk of the row / row within a group of 4
            uint32_t stage = 0, phase = 0;
            float acc[3][8];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
            int cur_b = -1;
            auto flush = [&]() {
                if (cur_b >= 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (p.colsum_mode == 1 && k > 0) break;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float v = acc[k][e];
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            if (rsub == 0 && v != 0.f) {
                                const int n = n0 + 64 * w + 8 * cch + e;
                                float* o = p.colsum_mode == 1 ? p.colsum_out + static_cast<long long>(cur_b) * 256 + n
                                                              : p.colsum_out + (static_cast<long long>(k) * p.B + cur_b) * p.Cg + n;
                                atomicAdd(o, v);
                            }
                            acc[k][e] = 0.f;
                        }
                    }
                }
            };
            for (int i = 0; i < my_chunks; ++i) {
                const int chunk = split + i * p.splits;
                const int b = chunk / p.chunks_per_b, t0 = (chunk % p.chunks_per_b) * WG_KT;
                mbar_wait(&bar_full[stage], phase, SITE_WG_FULL, stage);       // also orders this arrival after the previous use of the stage
                if (i % sum_ways == sum_me) {
                    if (b != cur_b) { flush(); cur_b = b; }
                    const uint8_t* box = smem + stage * WG_STAGE_BYTES + WG_A_BYTES + w * WG_BOX_BYTES;
                    const bool edge = p.colsum_mode == 2 && (t0 < p.colsum_d || t0 + WG_KT > p.L - p.colsum_d);
#pragma unroll 4
                    for (int r4 = 0; r4 < WG_KT; r4 += 4) {
                        const int r = r4 + rsub;
                        const uint4 v = *reinterpret_cast<const uint4*>(box + r * 128 + ((cch ^ (r & 7)) << 4));
                        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { x[2 * e] = bf16_lo(u[e]); x[2 * e + 1] = bf16_hi(u[e]); }
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[0][e] += x[e];
                        if (edge) {
                            const int t = t0 + r;
                            const float lo = t < p.colsum_d ? 1.f : 0.f, hi = t >= p.L - p.colsum_d ? 1.f : 0.f;
#pragma unroll
                            for (int e = 0; e < 8; ++e) { acc[1][e] = fmaf(lo, x[e], acc[1][e]); acc[2][e] = fmaf(hi, x[e], acc[2][e]); }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[stage]);
                if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            }
            flush();
        }
        const int q = warp & 3;
        const int row = q * 32 + lane;                     // output row within the tile = input channel m0 + row
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        mbar_wait(bar_done, 0, SITE_WG_DONE, 0);
        tc_fence_after_sync();
        float* orow = p.out + (static_cast<long long>(tap) * p.Ca + m0 + row) * p.ldo + n0;
        for (int cc = 0; cc < 8; ++cc) {
            uint32_t r[32];
            tmem_ld_32x32(t_lane + cc * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(orow + cc * 32 + i, __uint_as_float(r[i]) * p.scale);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair form (cta_group::2) of the same GEMM for Ca % 256 == 0: the single-CTA kernel above is bound by L2 bandwidth,
// not by the tensor pipe (every 128 x 256 tile pulls 48 KB per 64-step chunk: 10 TB/s at the dilated-conv shape). A pair
// computes a 256 x 256 tile as ONE M = 256 MMA stream issued by rank 0: each CTA stages its own 128 channels of A and only
// HALF of the G tile (32 KB per chunk and CTA, -33 % L2 traffic), six stages. Both operands stay MN-major.
// Each CTA's TMA loads complete on its OWN barrier (its epilogue warps read the staged G boxes for the column sums); the
// peer's warp 5 forwards "my stage is full" to the leader. Warps: 0..3 epilogue / column sums, 4 TMA producer, 5 MMA issuer
// (rank 0) / forwarder (rank 1) — the single-thread roles on the highest warp ids (scheduler priority).
// ------------------------------------------------------------------------------------------------
constexpr int WP_STAGES = 6;
constexpr int WP_A_BYTES = 2 * 64 * WG_KT * 2;            // 16 KB: this CTA's 128 channels of A
constexpr int WP_G_BYTES = 2 * 64 * WG_KT * 2;            // 16 KB: this CTA's 128 of the tile's 256 G columns
constexpr int WP_STAGE_BYTES = WP_A_BYTES + WP_G_BYTES;
constexpr int WP_THREADS = 192;
constexpr int WP_SMEM_BYTES = WP_STAGES * WP_STAGE_BYTES + 32 * 8 + 16;

__global__ void __launch_bounds__(WP_THREADS, 1)
wgrad_tc_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_g, const WgradTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WP_STAGES * WP_STAGE_BYTES);
    uint64_t* bar_full = bars;                       // [WP_STAGES] per CTA: this CTA's boxes landed
    uint64_t* bar_pfull = bars + WP_STAGES;          // [WP_STAGES] rank 0: the peer's boxes landed
    uint64_t* bar_empty = bars + 2 * WP_STAGES;      // [WP_STAGES] per CTA
    uint64_t* bar_done = bars + 3 * WP_STAGES;       // per CTA
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 32);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;

    // p.tiles_m counts 256-channel pair tiles here
    const int pair = static_cast<int>(blockIdx.x) >> 1;
    const int job = pair / p.splits, split = pair % p.splits;
    const int tn = job % p.tiles_n, pm = (job / p.tiles_n) % p.tiles_m, tap = job / (p.tiles_n * p.tiles_m);
    // every CTA of an n-tile with this rank stages the same half of the G tiles (whatever its m-tile and tap): they share the sums
    const bool do_sum = p.colsum_mode != 0 && (p.colsum_mode == 2 || tn == 0);
    const int sum_ways = p.tiles_m * p.taps, sum_me = pm * p.taps + tap;
    if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_g); }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < WP_STAGES; ++s) {
                mbar_init(&bar_full[s], 1);
                mbar_init(&bar_pfull[s], 1);
                mbar_init(&bar_empty[s], do_sum ? 5 : 1);
            }
            mbar_init(bar_done, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair(s_tmem, 256);
        tmem_relinquish_pair();
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;

    const int shift = (tap - p.taps / 2) * p.dil;
    const int m0 = pm * 256 + rank * 128, n0 = tn * 256 + rank * 128;     // this CTA's A channels / G columns
    const int total_chunks = p.B * p.chunks_per_b;
    const int my_chunks = (total_chunks - split + p.splits - 1) / p.splits;

    if (warp == 4) {
        uint32_t stage = 0, phase = 0;
        for (int i = 0; i < my_chunks; ++i) {
            const int chunk = split + i * p.splits;
            const int b = chunk / p.chunks_per_b, t0 = (chunk % p.chunks_per_b) * WG_KT;
            mbar_wait(&bar_empty[stage], phase ^ 1, SITE_WG_EMPTY, stage);
            if (lane == 0) {
                uint8_t* sa = smem + stage * WP_STAGE_BYTES;
                uint8_t* sg = sa + WP_A_BYTES;
                mbar_arrive_expect_tx(&bar_full[stage], WP_STAGE_BYTES);
                for (int j = 0; j < 2; ++j) tma_load_3d(sa + j * WG_BOX_BYTES, &tm_a, &bar_full[stage], m0 + 64 * j, t0 + shift, b);
                for (int j = 0; j < 2; ++j) tma_load_3d(sg + j * WG_BOX_BYTES, &tm_g, &bar_full[stage], n0 + 64 * j, t0, b);
            }
            __syncwarp();
            if (++stage == WP_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 5) {
        uint32_t stage = 0, phase = 0;
        constexpr uint32_t IDESC = umma_idesc_bf16_f32_mn(256, 256);
        for (int i = 0; i < my_chunks; ++i) {
            mbar_wait(&bar_full[stage], phase, SITE_WG_FULL, stage);
            if (!leader) {
                if (lane == 0) mbar_arrive_cluster(&bar_pfull[stage], 0);        // forward: the peer's half of this stage is in place
            } else {
                mbar_wait(&bar_pfull[stage], phase, SITE_WG_FULL, stage + 8);
                tc_fence_after_sync();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * WP_STAGE_BYTES);
                    const uint32_t sg = sa + WP_A_BYTES;
#pragma unroll
                    for (int k = 0; k < WG_KT / 16; ++k)
                        umma_bf16_ss_pair(tmem_base, umma_desc_sw128_mnmajor(sa + k * 2048, WG_BOX_BYTES),
                                          umma_desc_sw128_mnmajor(sg + k * 2048, WG_BOX_BYTES), IDESC, (i | k) != 0 ? 1u : 0u);
                    umma_commit_pair_mc(&bar_empty[stage], 3);
                    if (i == my_chunks - 1) umma_commit_pair_mc(bar_done, 3);
                }
            }
            __syncwarp();
            if (++stage == WP_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        if (do_sum) {
            // column sums of this CTA's two G boxes: warp w takes box (w & 1), rows 32 (w >> 1) .. + 31; a lane reads 16 bytes (8
            // channels) of a row, 8 lanes cover a 128-byte row and the warp 4 rows per instruction (conflict-free)
            const int w = warp;
            const int cch = lane & 7, rsub = lane >> 3;
            uint32_t stage = 0, phase = 0;
            float acc[3][8];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
            int cur_b = -1;
            auto flush = [&]() {
                if (cur_b >= 0) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (p.colsum_mode == 1 && k > 0) break;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float v = acc[k][e];
                            v += __shfl_xor_sync(0xffffffffu, v, 8);
                            v += __shfl_xor_sync(0xffffffffu, v, 16);
                            if (rsub == 0 && v != 0.f) {
                                const int n = n0 + 64 * (w & 1) + 8 * cch + e;
                                float* o = p.colsum_mode == 1 ? p.colsum_out + static_cast<long long>(cur_b) * 256 + n
                                                              : p.colsum_out + (static_cast<long long>(k) * p.B + cur_b) * p.Cg + n;
                                atomicAdd(o, v);
                            }
                            acc[k][e] = 0.f;
                        }
                    }
                }
            };
            for (int i = 0; i < my_chunks; ++i) {
                const int chunk = split + i * p.splits;
                const int b = chunk / p.chunks_per_b, t0 = (chunk % p.chunks_per_b) * WG_KT;
                mbar_wait(&bar_full[stage], phase, SITE_WG_FULL, stage);
                if (i % sum_ways == sum_me) {
                    if (b != cur_b) { flush(); cur_b = b; }
                    const uint8_t* box = smem + stage * WP_STAGE_BYTES + WP_A_BYTES + (w & 1) * WG_BOX_BYTES;
                    const bool edge = p.colsum_mode == 2 && (t0 < p.colsum_d || t0 + WG_KT > p.L - p.colsum_d);
                    const int rlo = (w >> 1) * (WG_KT / 2);
#pragma unroll 4
                    for (int r4 = 0; r4 < WG_KT / 2; r4 += 4) {
                        const int r = rlo + r4 + rsub;
                        const uint4 v = *reinterpret_cast<const uint4*>(box + r * 128 + ((cch ^ (r & 7)) << 4));
                        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
                        float x[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) { x[2 * e] = bf16_lo(u[e]); x[2 * e + 1] = bf16_hi(u[e]); }
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[0][e] += x[e];
                        if (edge) {
                            const int t = t0 + r;
                            const float lo = t < p.colsum_d ? 1.f : 0.f, hi = t >= p.L - p.colsum_d ? 1.f : 0.f;
#pragma unroll
                            for (int e = 0; e < 8; ++e) { acc[1][e] = fmaf(lo, x[e], acc[1][e]); acc[2][e] = fmaf(hi, x[e], acc[2][e]); }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[stage]);
                if (++stage == WP_STAGES) { stage = 0; phase ^= 1; }
            }
            flush();
        }
        const int q = warp & 3;
        const int row = q * 32 + lane;                     // output row within this CTA's half = input channel m0 + row
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        mbar_wait(bar_done, 0, SITE_WG_DONE, 0);
        tc_fence_after_sync();
        float* orow = p.out + (static_cast<long long>(tap) * p.Ca + m0 + row) * p.ldo + tn * 256;
        for (int cc = 0; cc < 8; ++cc) {
            uint32_t r[32];
            tmem_ld_32x32(t_lane + cc * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) atomicAdd(orow + cc * 32 + i, __uint_as_float(r[i]) * p.scale);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 256);
    }
}

}  // namespace adb
