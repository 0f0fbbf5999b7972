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

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_g); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
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

    const int job = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
    const int tn = job % p.tiles_n, tm = (job / p.tiles_n) % p.tiles_m, tap = job / (p.tiles_n * p.tiles_m);
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

}  // namespace adb
