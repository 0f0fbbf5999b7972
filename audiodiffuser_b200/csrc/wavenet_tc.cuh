// Tensor-core (tcgen05 / TMEM / TMA) kernels for the DiffWave residual stack, C = 256, bf16 operands,
// fp32 accumulation. One launch per residual block computes, for every 128-sample time tile:
//
//   GEMM1 (implicit dilated conv, K = 3 taps x 256):  y = W1 * [x(t-d); x(t); x(t+d)]      (wavenet.py:110)
//   epilogue 1: + bias + boundary-corrected step-embedding term, z = sigmoid(gate) * tanh(filter)
//               written to shared memory as the next GEMM's A operand                       (wavenet.py:108-112)
//   GEMM2 (1x1 conv, K = 256): o = W2 * z                                                   (wavenet.py:113)
//   epilogue 2: h' = (h + o[:256] + b) / sqrt(2) -> bf16 ; skip += o[256:] + b (fp32)        (wavenet.py:114-115,149)
//
// Epilogue 2 never loads from global memory: the residual input h is added inside the tensor core
// (four extra N = 64 MMAs per tile multiply the centre-tap activation tile by a 64 x 64 identity, which
// is exact in fp32), and the skip sum is updated with a TMA reduce-add performed at L2. Outputs are
// transposed through shared memory (the z buffer, idle once GEMM2 has finished) and leave as TMA
// tensor stores, so every global transaction is a full 128-byte row segment.
//
// Data layout: activations channels-last bf16 [B][L][256], so both MMA operands are K-major and a
// dilated tap is a row offset of +-d in a 3-D TMA tensor map whose out-of-bounds zero fill IS the
// convolution's zero padding (wavenet.py:71). Weights are pre-packed (api.cu) into 32 KB blocks
// [256 n][64 k] bf16 that one TMA box loads with the 128-byte swizzle.
//
// The step-embedding add of wavenet.py:109 happens BEFORE zero padding in the reference, so it
// cannot be folded into a plain bias. It is folded exactly instead: with p = Linear(emb) (per
// sample, per layer), conv(x + p) = conv(x) + E1 + [t >= d] E0 + [t < L - d] E2 where
// E_tap = W1[tap] p (E1 also carries the conv bias). The E vectors are produced per network
// evaluation by one small fp32 GEMM and added in epilogue 1 in fp32.
//
// CTA = 10 warps: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2..9 epilogue
// (two warps per TMEM lane quarter, splitting the columns). TMEM: 2 accumulators x 256 columns.
// Per tile the MMA warp runs 4 jobs: G1a (gate/filter of channels 0..127), G1b (128..255),
// G2r (residual half), G2s (skip half); jobs alternate between the two accumulators so the
// epilogue of one job overlaps the MMAs of the next.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

namespace adb {

constexpr int TC_C = 256;
constexpr int TC_TILE_T = 128;
constexpr int TC_STAGES = 3;
constexpr int TC_A_BYTES = TC_TILE_T * 64 * 2;        // 16 KB: [128 t][64 ci] bf16
constexpr int TC_B_BYTES = 256 * 64 * 2;              // 32 KB: [256 n][64 k] bf16
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_Z_BYTES = 4 * TC_A_BYTES;            // gated activations: 4 K-blocks [128 t][64 c]
constexpr int TC_W_BLOCKS_PER_LAYER = 32;             // 2 x 12 (GEMM1) + 2 x 4 (GEMM2)
constexpr int TC_THREADS = 320;
constexpr int TC_EPI_THREADS = 256;

struct BlockTcSmem {
    // offsets from the 1024-aligned base
    static constexpr int stages = 0;
    static constexpr int z = TC_STAGES * TC_STAGE_BYTES;
    static constexpr int evec = z + TC_Z_BYTES;                 // 3 x 512 fp32
    static constexpr int b2 = evec + 3 * 512 * 4;               // 512 fp32
    static constexpr int ident = b2 + 512 * 4;                  // [64 n][64 k] bf16 identity, K-major SW128 (8 KB)
    static constexpr int esum = ident + 64 * 128;               // 512 fp32: E0 + E1 + E2 (gate half pre-scaled by 0.5)
    static constexpr int bars = esum + 512 * 4;                 // mbarriers
    static constexpr int tmem_ptr = bars + 16 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(BlockTcSmem::ident % 1024 == 0, "identity tile must sit on a swizzle-atom boundary");
constexpr int TC_BLOCK_SMEM_BYTES = BlockTcSmem::total;

struct BlockTcParams {
    const float* E;                 // [B][layers][3][512] epilogue-1 constants for this evaluation
    const float* b2;                // [512] output-projection bias of this layer
    const __nv_bfloat16* h_in;      // [B][L][256]
    __nv_bfloat16* h_out;           // [B][L][256]
    float* skip;                    // [B][L][256] fp32 running skip sum
    int B, L, layer, layers, dil;
    int tiles_per_b, num_tiles;
    int first_layer;                // 1: skip = value, 0: skip += value
    int write_h;                    // 0 on the last layer (its residual output is never used, wavenet.py:145-151)
    int dbg;                        // ADB_DEBUG builds only (ADB_DEBUG_FLAGS): 2 = in-kernel cycle accounting, other bits = timing experiments
    int cluster;                    // CTAs per cluster (1, 2 or 4): weight tiles are TMA-multicast across the cluster
    // pair kernel only — the fp32 skip sum is read-modify-written every SECOND layer:
    int skip_mode;                  // 0: skip += value (TMA reduce-add)  1: skip = value  2: stash the value as bf16 for the next layer
    int add_stash;                  // 1: the previous layer's stashed skip contribution is added to this layer's (identity MMA)
};

// Optional in-kernel cycle accounting (BlockTcParams::dbg & 2): where each role waits.
//  0 mma: wait accumulator-drained (G1 jobs)   1 mma: wait accumulator-drained (G2 jobs)   2 mma: wait z ready
//  3 mma: wait TMA stage full                   4 mma: total                                 5 producer: wait stage empty
//  6 producer: total                            7 epi(warp 2): wait accumulator full (G1)    8 epi: epilogue-1 work
//  9 epi: wait accumulator full (G2)           10 epi: epilogue-2 work                       11 epi: total
__device__ unsigned long long g_tc_cycles[16];
// Debug switches exist in -DADB_DEBUG builds only (libadb200_dbg.so, tools/): the release kernels carry no run-time predicate.
#ifdef ADB_DEBUG
#define TC_DBG_FLAGS(p) const int kdbg = (p).dbg
#else
#define TC_DBG_FLAGS(p) constexpr int kdbg = 0
#endif
#define TC_DBG_T0(var) long long var = 0; if (kdbg & 2) var = clock64()
#define TC_DBG_ACC(idx, var) if (kdbg & 2) dbg_acc[idx] += clock64() - var

enum TcWaitSite : uint32_t {
    SITE_PROD_EMPTY = 1, SITE_MMA_TEMPTY = 2, SITE_MMA_ZREADY = 3, SITE_MMA_FULL = 4, SITE_EPI_TFULL = 5,
    SITE_TAIL_W = 6, SITE_TAIL_MMA = 7, SITE_ML_FLAG = 8,
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
wavenet_block_tc_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                        const __grid_constant__ CUtensorMap tm_hout, const __grid_constant__ CUtensorMap tm_skip,
                        const BlockTcParams p) {
    TC_DBG_FLAGS(p);
    extern __shared__ __align__(1024) uint8_t smem[];      // SWIZZLE_128B tiles need 1024-byte alignment
    float* s_evec = reinterpret_cast<float*>(smem + BlockTcSmem::evec);
    float* s_b2 = reinterpret_cast<float*>(smem + BlockTcSmem::b2);
    float* s_esum = reinterpret_cast<float*>(smem + BlockTcSmem::esum);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BlockTcSmem::bars);
    uint64_t* bar_full = bars;                   // [TC_STAGES]  TMA -> MMA
    uint64_t* bar_empty = bars + TC_STAGES;      // [TC_STAGES]  MMA -> TMA
    uint64_t* bar_tfull = bars + 2 * TC_STAGES;  // [2] accumulator ready   (MMA -> epilogue)
    uint64_t* bar_tempty = bar_tfull + 2;        // [2] accumulator drained (epilogue -> MMA)
    uint64_t* bar_zready = bar_tempty + 2;       // [2] z K-blocks {2j, 2j+1} written (epilogue -> MMA)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + BlockTcSmem::tmem_ptr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_hout);
        tma_prefetch_desc(&tm_skip);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], p.cluster); }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_tfull[i], 1);
                mbar_init(&bar_tempty[i], TC_EPI_THREADS / 32);
                mbar_init(&bar_zready[i], TC_EPI_THREADS / 32);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(s_tmem, 512);
        tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < 512; i += TC_EPI_THREADS) s_b2[i] = p.b2[i];
        // identity operand: element (n, k) lives in 16-byte chunk (k / 8) ^ (n & 7) of row n
        uint4* id4 = reinterpret_cast<uint4*>(smem + BlockTcSmem::ident);
        for (int i = threadIdx.x - 64; i < 64 * 8; i += TC_EPI_THREADS) {
            const int n = i >> 3, phys = i & 7;
            const int chunk = phys ^ (n & 7);                  // logical chunk stored at this position
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if (chunk == (n >> 3)) {
                const int e = n & 7;                           // element inside the chunk
                w[e >> 1] = (e & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 in the high / low half
            }
            id4[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();      // every CTA's barriers exist before any remote arrive / multicast write
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    constexpr uint32_t IDESC = umma_idesc_bf16_f32(128, 256);
    constexpr uint32_t IDESC_N64 = umma_idesc_bf16_f32(128, 64);
    constexpr uint32_t IDESC_F16 = umma_idesc_f16_f32(128, 256);     // GEMM2: z and W2 are fp16
    // Tile schedule: clusters walk groups of `cluster` consecutive tiles; every CTA of a cluster runs the same
    // number of iterations (a CTA whose tile is past the end still moves its share of the weights and
    // consumes its stages, but stores nothing).
    const int crank = p.cluster > 1 ? static_cast<int>(cluster_ctarank()) : 0;
    const int cluster_id = blockIdx.x / p.cluster;
    const int num_clusters = gridDim.x / p.cluster;
    const int num_groups = (p.num_tiles + p.cluster - 1) / p.cluster;
    const uint16_t mc_mask = static_cast<uint16_t>((1u << p.cluster) - 1u);
    const int w_rows = 256 / p.cluster;           // rows of each weight tile this CTA fetches for the cluster

    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t stage = 0, phase = 0;
        long long dbg_acc[16] = {};
        TC_DBG_T0(tp_all);
        for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
            const int tile = grp * p.cluster + crank;
            const int b = tile / p.tiles_per_b;           // >= B for a padding tile: TMA zero-fills
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            for (int job = 0; job < 4; ++job) {
                if (job == 2 && !p.write_h) continue;
                const int nkb = job < 2 ? 12 : 4;
                for (int kb = 0; kb < nkb; ++kb) {
                    TC_DBG_T0(tw);
                    mbar_wait(&bar_empty[stage], phase ^ 1, SITE_PROD_EMPTY, stage);
                    TC_DBG_ACC(5, tw);
                    if (lane == 0) {
                        uint8_t* sa = smem + BlockTcSmem::stages + stage * TC_STAGE_BYTES;
                        uint8_t* sb = sa + TC_A_BYTES;
                        const int wblk = p.layer * TC_W_BLOCKS_PER_LAYER +
                                         (job < 2 ? job * 12 + kb : 24 + (job - 2) * 4 + kb);
                        if (job < 2) {
                            mbar_arrive_expect_tx(&bar_full[stage], TC_STAGE_BYTES);
                            const int tap = kb >> 2, cib = kb & 3;
                            tma_load_3d(sa, &tm_h, &bar_full[stage], cib * 64, t0 + (tap - 1) * p.dil, b);
                        } else if (job == 2) {
                            // residual job: also fetch h[t0.., 64 kb .. 64 kb + 63] for the identity MMA
                            mbar_arrive_expect_tx(&bar_full[stage], TC_STAGE_BYTES);
                            tma_load_3d(sa, &tm_h, &bar_full[stage], kb * 64, t0, b);
                        } else {
                            mbar_arrive_expect_tx(&bar_full[stage], TC_B_BYTES);
                        }
                        if (p.cluster == 1) tma_load_2d(sb, &tm_w, &bar_full[stage], 0, wblk * 256);
                        else tma_load_2d_mc(sb + crank * w_rows * 128, &tm_w, &bar_full[stage], 0, wblk * 256 + crank * w_rows,
                                            mc_mask);
                    }
                    __syncwarp();
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        TC_DBG_ACC(6, tp_all);
        if ((kdbg & 2) && lane == 0) { atomicAdd(&g_tc_cycles[5], dbg_acc[5]); atomicAdd(&g_tc_cycles[6], dbg_acc[6]); }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t stage = 0, phase = 0;
        uint32_t use0 = 0, use1 = 0;        // jobs issued so far into accumulator 0 / 1
        uint32_t it = 0;
        const uint32_t z_addr = smem_u32(smem + BlockTcSmem::z);
        long long dbg_acc[16] = {};
        TC_DBG_T0(tm_all);
        for (int grp = cluster_id; grp < num_groups; grp += num_clusters, ++it) {
            for (int job = 0; job < 4; ++job) {
                if (job == 2 && !p.write_h) continue;
                const int buf = job & 1;
                const uint32_t use = buf ? use1 : use0;
                TC_DBG_T0(tw0);
                mbar_wait(&bar_tempty[buf], (use & 1) ^ 1, SITE_MMA_TEMPTY, job);
                TC_DBG_ACC(job < 2 ? 0 : 1, tw0);
                const bool first_g2 = (job == 2 || (job == 3 && !p.write_h));
                if (first_g2) {
                    // z K-blocks 0,1 come from epilogue 1a, K-blocks 2,3 from epilogue 1b: start on the first
                    // half while 1b is still running (the second wait sits inside the K loop)
                    TC_DBG_T0(tw1);
                    mbar_wait(&bar_zready[0], it & 1, SITE_MMA_ZREADY, 0);
                    TC_DBG_ACC(2, tw1);
                }
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * 256;
                const int nkb = job < 2 ? 12 : 4;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (first_g2 && kb == 2) {
                        TC_DBG_T0(tw1);
                        mbar_wait(&bar_zready[1], it & 1, SITE_MMA_ZREADY, 1);
                        TC_DBG_ACC(2, tw1);
                    }
                    TC_DBG_T0(tw2);
                    mbar_wait(&bar_full[stage], phase, SITE_MMA_FULL, stage);
                    TC_DBG_ACC(3, tw2);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + BlockTcSmem::stages + stage * TC_STAGE_BYTES);
                        const uint32_t a_addr = job < 2 ? sa : z_addr + kb * TC_A_BYTES;
                        const uint32_t b_addr = sa + TC_A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_bf16_ss(d_tmem, umma_desc_sw128_kmajor(a_addr + k * 32),
                                         umma_desc_sw128_kmajor(b_addr + k * 32), job < 2 ? IDESC : IDESC_F16,
                                         (kb | k) != 0 ? 1u : 0u);
                        }
                        if (job == 2) {
                            // + h: D[:, 64 kb .. 64 kb + 63] += h_tile * I   (exact: bf16 x 1.0 in fp32)
                            const uint32_t id_addr = smem_u32(smem + BlockTcSmem::ident);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss(d_tmem + kb * 64, umma_desc_sw128_kmajor(sa + k * 32),
                                             umma_desc_sw128_kmajor(id_addr + k * 32), IDESC_N64, 1u);
                        }
                        if (p.cluster == 1) umma_commit(&bar_empty[stage]);
                        else umma_commit_mc(&bar_empty[stage], mc_mask);      // frees this stage in every CTA of the cluster
                        if (kb == nkb - 1) umma_commit(&bar_tfull[buf]);
                    }
                    __syncwarp();
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                if (buf) ++use1; else ++use0;
            }
        }
        TC_DBG_ACC(4, tm_all);
        if ((kdbg & 2) && lane == 0)
            for (int i = 0; i < 5; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
    } else {
        // ===================== epilogue warps =====================
        const int ew = warp - 2;              // 0..7
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int half = ew >> 2;             // which half of the columns this warp handles
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* zbase = smem + BlockTcSmem::z;
        uint32_t use0 = 0, use1 = 0;
        long long dbg_acc[16] = {};
        TC_DBG_T0(te_all);
        const float* Ec = s_evec + 512;       // tap 1 (centre, + conv bias)
        const float* E0 = s_evec;             // tap 0 (t - d)
        const float* E2 = s_evec + 1024;      // tap 2 (t + d)
        for (int grp = cluster_id; grp < num_groups; grp += num_clusters) {
            const int tile = grp * p.cluster + crank;
            const bool tile_valid = tile < p.num_tiles;
            const int b = tile / p.tiles_per_b;
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            const int t = t0 + row;
            named_bar_sync(1, TC_EPI_THREADS);
            {
                const float* src = p.E + (static_cast<long long>(tile_valid ? b : 0) * p.layers + p.layer) * 1536;
                for (int i = threadIdx.x - 64; i < 1536; i += TC_EPI_THREADS) s_evec[i] = src[i];
            }
            named_bar_sync(1, TC_EPI_THREADS);
            for (int i = threadIdx.x - 64; i < 512; i += TC_EPI_THREADS)
                s_esum[i] = (s_evec[i] + s_evec[512 + i] + s_evec[1024 + i]) * (i < 256 ? 0.5f : 1.0f);
            named_bar_sync(1, TC_EPI_THREADS);
            const float m_lo = (t >= p.dil) ? 1.0f : 0.0f;
            const float m_hi = (t < p.L - p.dil) ? 1.0f : 0.0f;
            // interior tile: every row sees all three taps, so the step-embedding term is one vector
            const bool interior = (t0 >= p.dil) && (t0 + TC_TILE_T - 1 < p.L - p.dil);

            // ---- epilogue 1: gate, two jobs (channels 128 j .. 128 j + 127) ----
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                TC_DBG_T0(tw);
                mbar_wait(&bar_tfull[j], (j ? use1 : use0) & 1, SITE_EPI_TFULL, j);
                TC_DBG_ACC(7, tw);
                TC_DBG_T0(tk);
                if (j) ++use1; else ++use0;
                tc_fence_after_sync();
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = half * 64 + cc * 32;          // column inside the job's gate half
                    uint32_t g[32], f[32];
                    tmem_ld_32x32(t_lane + j * 256 + col, g);
                    tmem_ld_32x32(t_lane + j * 256 + 128 + col, f);
                    tmem_ld_wait();
                    const int c0 = 128 * j + col;                 // original channel of g[0]
                    // z = sigmoid(g) * tanh(f) with sigmoid(g) = 0.5 tanh(g / 2) + 0.5, evaluated two channels at a
                    // time in fp16x2 (one MUFU per tanh pair); z in (-1, 1) is stored as fp16 for GEMM2
                    uint32_t packed[16];
                    if (interior) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float g0 = fmaf(__uint_as_float(g[i]), 0.5f, s_esum[c0 + i]);
                            const float g1 = fmaf(__uint_as_float(g[i + 1]), 0.5f, s_esum[c0 + i + 1]);
                            const float f0 = __uint_as_float(f[i]) + s_esum[256 + c0 + i];
                            const float f1 = __uint_as_float(f[i + 1]) + s_esum[256 + c0 + i + 1];
                            const uint32_t tg = tanh_f16x2(pack_f16x2(g0, g1));
                            const uint32_t tf = tanh_f16x2(pack_f16x2(f0, f1));
                            packed[i >> 1] = hmul2(hfma2(tg, 0x38003800u, 0x38003800u), tf);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float gv[2], fv[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int c = c0 + i + u;
                                gv[u] = __uint_as_float(g[i + u]) + Ec[c];
                                fv[u] = __uint_as_float(f[i + u]) + Ec[256 + c];
                                gv[u] = fmaf(m_lo, E0[c], gv[u]);
                                fv[u] = fmaf(m_lo, E0[256 + c], fv[u]);
                                gv[u] = 0.5f * fmaf(m_hi, E2[c], gv[u]);
                                fv[u] = fmaf(m_hi, E2[256 + c], fv[u]);
                            }
                            const uint32_t tg = tanh_f16x2(pack_f16x2(gv[0], gv[1]));
                            const uint32_t tf = tanh_f16x2(pack_f16x2(fv[0], fv[1]));
                            packed[i >> 1] = hmul2(hfma2(tg, 0x38003800u, 0x38003800u), tf);
                        }
                    }
                    // z K-block (64 channels) = 2 j + half ; 16-byte chunk inside the 128-byte row = 4 cc + m
                    uint8_t* zrow = zbase + (2 * j + half) * TC_A_BYTES + row * 128;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int chunk = (4 * cc + m) ^ (row & 7);
                        *reinterpret_cast<uint4*>(zrow + chunk * 16) =
                            make_uint4(packed[4 * m], packed[4 * m + 1], packed[4 * m + 2], packed[4 * m + 3]);
                    }
                }
                fence_proxy_async_smem();       // generic-proxy writes of z -> visible to the MMA (async proxy)
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bar_tempty[j]);
                    mbar_arrive(&bar_zready[j]);
                }
                TC_DBG_ACC(8, tk);
            }

            // ---- epilogue 2r: residual half -> h_out, stored straight from registers as soon as the residual
            //      GEMM is done (the skip GEMM is still reading z, so no staging buffer is free yet). Each thread
            //      owns one time row and writes 64 contiguous bytes per chunk.
            const long long rowoff = (static_cast<long long>(b) * p.L + t) * TC_C;
            if (p.write_h) {
                TC_DBG_T0(tw3);
                mbar_wait(&bar_tfull[0], use0 & 1, SITE_EPI_TFULL, 2);
                TC_DBG_ACC(9, tw3);
                TC_DBG_T0(tk3);
                ++use0;
                tc_fence_after_sync();
                const bool st_ok = tile_valid && (t < p.L) && !(kdbg & 1);
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = half * 128 + cc * 32;     // residual channel of r[0]
                    uint32_t r[32];
                    tmem_ld_32x32(t_lane + col, r);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float v0 = (__uint_as_float(r[i]) + s_b2[col + i]) * 0.70710678118654752f;
                        const float v1 = (__uint_as_float(r[i + 1]) + s_b2[col + i + 1]) * 0.70710678118654752f;
                        pk[i >> 1] = pack_bf16x2(v0, v1);
                    }
                    if (st_ok) {
                        uint4* dst = reinterpret_cast<uint4*>(p.h_out + rowoff + col);
#pragma unroll
                        for (int m = 0; m < 4; ++m) dst[m] = make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_tempty[0]);
                TC_DBG_ACC(10, tk3);
            }

            // ---- epilogue 2s: skip half, after the last MMAs reading z have finished. Each warp owns 8 KB of
            //      the now idle z buffer as two 4 KB transposition boxes (32 rows x 128 B, 128-byte swizzle) that
            //      leave through TMA reduce-add.
            TC_DBG_T0(tw2);
            mbar_wait(&bar_tfull[1], use1 & 1, SITE_EPI_TFULL, 3);
            TC_DBG_ACC(9, tw2);
            TC_DBG_T0(tk2);
            ++use1;
            uint8_t* stg = zbase + ew * 8192;
            const int trow = t0 + q * 32;                 // first time step of this warp's 32 rows
            {
                tc_fence_after_sync();
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = half * 128 + cc * 32;     // skip channel of r[0]
                    uint32_t r[32];
                    tmem_ld_32x32(t_lane + 256 + col, r);
                    tmem_ld_wait();
                    uint8_t* brow = stg + (cc & 1) * 4096 + lane * 128;
                    // box (cc & 1) was handed to TMA two commits ago (by the residual half or by chunk cc - 2)
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        float4 v;
                        v.x = __uint_as_float(r[4 * m + 0]) + s_b2[256 + col + 4 * m + 0];
                        v.y = __uint_as_float(r[4 * m + 1]) + s_b2[256 + col + 4 * m + 1];
                        v.z = __uint_as_float(r[4 * m + 2]) + s_b2[256 + col + 4 * m + 2];
                        v.w = __uint_as_float(r[4 * m + 3]) + s_b2[256 + col + 4 * m + 3];
                        const int chunk = m ^ (lane & 7);
                        *reinterpret_cast<float4*>(brow + chunk * 16) = v;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && tile_valid && !(kdbg & 1)) {
                        if (p.first_layer) tma_store_3d(&tm_skip, stg + (cc & 1) * 4096, col, trow, b);
                        else               tma_reduce_add_3d(&tm_skip, stg + (cc & 1) * 4096, col, trow, b);
                        tma_store_commit();
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bar_tempty[1]);
                    tma_store_wait_read<0>();          // z is rewritten by the next tile's epilogue 1
                }
                __syncwarp();
            }
            TC_DBG_ACC(10, tk2);
        }
        TC_DBG_ACC(11, te_all);
        if ((kdbg & 2) && warp == 2 && lane == 0)
            for (int i = 7; i < 12; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores fully performed before exit
    }

    tc_fence_before_sync();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();      // no CTA leaves while a peer may still signal its barriers
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// Tail: F[b][t] = b_out + sum_c w_out[c] * relu(b_sp[c] + sum_k Wsp[c][k] * (skip[b][t][k] * scale))
// (wavenet.py:151, :177-179). One 128-row tile at a time: the CTA converts the fp32 skip sum to a
// swizzled bf16 A operand in shared memory, one thread issues the 16 MMAs (K = 256, N = 256), and
// the epilogue reduces the 256 columns against w_out. The skip-projection weights (128 KB) stay
// resident in shared memory for the CTA's lifetime.
// ------------------------------------------------------------------------------------------------
struct TailTcSmem {
    static constexpr int w = 0;                              // 4 x 32 KB
    static constexpr int a = 4 * TC_B_BYTES;                 // 4 x 16 KB
    static constexpr int bsp = a + TC_Z_BYTES;               // 256 fp32
    static constexpr int wout = bsp + 1024;                  // 256 fp32
    static constexpr int part = wout + 1024;                 // 128 fp32 partial sums (half 1)
    static constexpr int bars = part + 512;
    static constexpr int tmem_ptr = bars + 32;
    static constexpr int total = tmem_ptr + 16;
};
constexpr int TC_TAIL_SMEM_BYTES = TailTcSmem::total;

struct TailTcParams {
    const float* skip;      // [B][L][256]
    const float* b_sp;      // [256]
    const float* w_out;     // [256]
    const float* b_out;     // [1]
    float* out;             // [B][L]
    float scale;            // sqrt(1 / layers)
    int B, L, tiles_per_b, num_tiles;
};

__global__ void __launch_bounds__(256, 1)
wavenet_tail_tc_kernel(const __grid_constant__ CUtensorMap tm_wsp, const TailTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_bsp = reinterpret_cast<float*>(smem + TailTcSmem::bsp);
    float* s_wout = reinterpret_cast<float*>(smem + TailTcSmem::wout);
    float* s_part = reinterpret_cast<float*>(smem + TailTcSmem::part);
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + TailTcSmem::bars);
    uint64_t* bar_mma = bar_w + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TailTcSmem::tmem_ptr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tm_wsp);
            mbar_init(bar_w, 1);
            mbar_init(bar_mma, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(s_tmem, 256);
        tmem_relinquish();
    }
    s_bsp[threadIdx.x] = p.b_sp[threadIdx.x];
    s_wout[threadIdx.x] = p.w_out[threadIdx.x];
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    if (warp == 0 && lane == 0) {
        mbar_arrive_expect_tx(bar_w, 4 * TC_B_BYTES);
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(smem + TailTcSmem::w + kb * TC_B_BYTES, &tm_wsp, bar_w, 0, kb * 256);
    }
    mbar_wait(bar_w, 0, SITE_TAIL_W);

    constexpr uint32_t IDESC = umma_idesc_bf16_f32(128, 256);
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float b_out = p.b_out[0];
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int b = tile / p.tiles_per_b;
        const int t = (tile % p.tiles_per_b) * TC_TILE_T + row;
        const bool valid = t < p.L;
        // ---- A operand: warp w converts rows 16 w .. 16 w + 15; a warp-wide load covers one whole row (1 KB
        //      contiguous), lane l owning channels 8 l .. 8 l + 7 = 16-byte chunk (l & 7) of K-block (l >> 3)
        {
            const int tile_t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            const float* src0 = p.skip + (static_cast<long long>(b) * p.L + tile_t0) * TC_C + lane * 8;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = warp * 16 + i;
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                if (tile_t0 + r < p.L) {
                    const float4* src = reinterpret_cast<const float4*>(src0 + static_cast<long long>(r) * TC_C);
                    v0 = src[0]; v1 = src[1];
                }
                const uint4 o = make_uint4(pack_bf16x2(v0.x * p.scale, v0.y * p.scale), pack_bf16x2(v0.z * p.scale, v0.w * p.scale),
                                           pack_bf16x2(v1.x * p.scale, v1.y * p.scale), pack_bf16x2(v1.z * p.scale, v1.w * p.scale));
                const int kb = lane >> 3, chunk = (lane & 7) ^ (r & 7);
                *reinterpret_cast<uint4*>(smem + TailTcSmem::a + kb * TC_A_BYTES + r * 128 + chunk * 16) = o;
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after_sync();
            if (lane == 0) {
                const uint32_t a0 = smem_u32(smem + TailTcSmem::a), w0 = smem_u32(smem + TailTcSmem::w);
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss(tmem_base, umma_desc_sw128_kmajor(a0 + kb * TC_A_BYTES + k * 32),
                                     umma_desc_sw128_kmajor(w0 + kb * TC_B_BYTES + k * 32), IDESC, (kb | k) != 0 ? 1u : 0u);
                umma_commit(bar_mma);
            }
            __syncwarp();
        }
        mbar_wait(bar_mma, it & 1, SITE_TAIL_MMA);
        tc_fence_after_sync();
        float acc = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            const int col = half * 128 + cc * 32;
            uint32_t r[32];
            tmem_ld_32x32(t_lane + col, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
                acc = fmaf(fmaxf(__uint_as_float(r[i]) + s_bsp[col + i], 0.f), s_wout[col + i], acc);
        }
        if (half == 1) s_part[row] = acc;
        tc_fence_before_sync();
        __syncthreads();
        if (half == 0 && valid) p.out[static_cast<long long>(b) * p.L + t] = acc + s_part[row] + b_out;
        // the next iteration's __syncthreads (after the A-operand stores) orders s_part reuse
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------
// Weight packing for the tensor-core path. The per-block TMA blocks are written by refold_layers_kernel (train_kernels.cuh):
//   GEMM1 blocks (job j in {0,1}, kb = tap * 4 + cib): row n < 128 -> gate channel 128 j + n,
//   row n >= 128 -> filter channel 256 + 128 j + (n - 128)   (gate = FIRST half, wavenet.py:111)
//   GEMM2 blocks (job j, kb): row n -> output row 256 j + n (j = 0 residual, 1 skip, wavenet.py:114), fp16 bit patterns
// ------------------------------------------------------------------------------------------------
// skip projection: out[kb][n][k] = wspf[kb*64 + k][n]
__global__ void pack_tc_tail_kernel(const float* __restrict__ wspf, __nv_bfloat16* __restrict__ out) {
    const int total = 4 * 256 * 64;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i & 63, n = (i >> 6) & 255, kb = i >> 14;
        out[i] = __float2bfloat16_rn(wspf[static_cast<long long>(kb * 64 + k) * 256 + n]);
    }
}

}  // namespace adb
