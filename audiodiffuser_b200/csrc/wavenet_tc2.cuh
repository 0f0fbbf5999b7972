// CTA-pair version of the fused DiffWave residual-block kernel (see wavenet_tc.cuh for the algorithm and
// the reference citations). Two CTAs of a cluster (one TPC) process two adjacent 128-sample time tiles as
// ONE M = 256 tcgen05.mma.cta_group::2 stream:
//   * rank 0 issues every MMA; each CTA's TMEM receives the accumulator rows of its own tile;
//   * each CTA loads its own activation taps (A) but only HALF of every weight tile (B): per-CTA shared
//     memory traffic and L2->SM weight traffic both drop, and a pipeline stage shrinks from 48 KB to 32 KB,
//     which buys a fourth stage (the 3-stage single-CTA pipeline was TMA-latency bound) plus a dedicated
//     transposition buffer for the residual epilogue;
//   * TMA completions of both CTAs are counted on rank 0's "full" barrier; tcgen05.commit multicasts the
//     "stage free" / "accumulator ready" arrivals to both CTAs; both CTAs' epilogue warps arrive on rank 0's
//     "accumulator drained" / "z ready" barriers.
// Skip accumulation (wavenet.py:145-149) touches the fp32 running sum only every second layer: an even layer writes its
// skip contribution as fp16 into a stash buffer (2 bytes per element, plain TMA store), the following odd layer loads that
// tile through TMA and adds it to its own accumulator with the same identity-MMA trick that adds the residual input
// (exact: fp16 x 1.0 accumulated in fp32), then does ONE fp32 reduce-add for both layers. HBM traffic per pair of layers
// drops from 2 x 8 bytes of read-modify-write per skip element to 8 + 2 + 2, and the L2 reductions are halved. The stashed
// contribution is rounded to fp16 once (relative 2^-12 of one layer's skip term; saturating at +-65504).
#pragma once
#include "wavenet_tc.cuh"

namespace adb {

constexpr int T2_STAGES = 4;
constexpr int T2_A_BYTES = 128 * 64 * 2;            // 16 KB: this CTA's [128 t][64 ci] activation block
constexpr int T2_B_BYTES = 128 * 64 * 2;            // 16 KB: this CTA's half [128 n][64 k] of a weight tile
constexpr int T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES;
constexpr int T2_STG_BYTES = 2048;                  // per-warp transposition buffer (32 rows x 64 B)

struct Tc2Smem {
    static constexpr int stages = 0;
    static constexpr int z = T2_STAGES * T2_STAGE_BYTES;            // 64 KB gated activations (GEMM2 A operand)
    static constexpr int stg = z + TC_Z_BYTES;                      // 8 x 2 KB
    static constexpr int ident = stg + 8 * T2_STG_BYTES;            // [32 n][64 k] bf16 half identity (4 KB)
    static constexpr int ident16 = ident + 32 * 128;                // the same identity in fp16 (for the fp16 skip stash)
    static constexpr int evec = ident16 + 32 * 128;                 // 3 x 512 fp32
    static constexpr int esum = evec + 3 * 512 * 4;                 // 512 fp32
    static constexpr int b2 = esum + 512 * 4;                       // 512 fp32
    static constexpr int bars = b2 + 512 * 4;
    static constexpr int tmem_ptr = bars + 16 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(Tc2Smem::ident % 1024 == 0 && Tc2Smem::z % 1024 == 0 && Tc2Smem::stg % 1024 == 0, "swizzle alignment");
static_assert(Tc2Smem::total <= 232448, "shared memory budget");
constexpr int TC2_SMEM_BYTES = Tc2Smem::total;

__global__ void __launch_bounds__(TC_THREADS, 1)
wavenet_block_pair_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                          const __grid_constant__ CUtensorMap tm_skip, const __grid_constant__ CUtensorMap tm_hout,
                          const __grid_constant__ CUtensorMap tm_stash_ld, const __grid_constant__ CUtensorMap tm_stash_st,
                          const BlockTcParams p) {
    TC_DBG_FLAGS(p);
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_evec = reinterpret_cast<float*>(smem + Tc2Smem::evec);
    float* s_esum = reinterpret_cast<float*>(smem + Tc2Smem::esum);
    float* s_b2 = reinterpret_cast<float*>(smem + Tc2Smem::b2);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Tc2Smem::bars);
    uint64_t* bar_full = bars;                   // [T2_STAGES] (rank 0's copy is the live one)
    uint64_t* bar_empty = bars + T2_STAGES;      // [T2_STAGES] per CTA
    uint64_t* bar_tfull = bars + 2 * T2_STAGES;  // [2] per CTA
    uint64_t* bar_tempty = bar_tfull + 2;        // [2] rank 0
    uint64_t* bar_zready = bar_tempty + 2;       // [2] rank 0
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Tc2Smem::tmem_ptr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_skip);
        tma_prefetch_desc(&tm_hout);
        tma_prefetch_desc(&tm_stash_ld);
        tma_prefetch_desc(&tm_stash_st);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < T2_STAGES; ++s) { mbar_init(&bar_full[s], 2); mbar_init(&bar_empty[s], 1); }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_tfull[i], 1);
                mbar_init(&bar_tempty[i], 2 * (TC_EPI_THREADS / 32));
                mbar_init(&bar_zready[i], 2 * (TC_EPI_THREADS / 32));
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair(s_tmem, 512);
        tmem_relinquish_pair();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < 512; i += TC_EPI_THREADS) s_b2[i] = p.b2[i];
        // this CTA's half of the 64 x 64 identity: local row i is n = 32 rank + i; element (n, k = n) sits in
        // 16-byte chunk (n / 8) ^ (i & 7) of row i (128-byte swizzle)
        uint4* id4 = reinterpret_cast<uint4*>(smem + Tc2Smem::ident);
        uint4* id16 = reinterpret_cast<uint4*>(smem + Tc2Smem::ident16);
        for (int i = threadIdx.x - 64; i < 32 * 8; i += TC_EPI_THREADS) {
            const int r = i >> 3, phys = i & 7;
            const int n = 32 * rank + r;
            const int chunk = phys ^ (r & 7);
            uint32_t w[4] = {0u, 0u, 0u, 0u}, w16[4] = {0u, 0u, 0u, 0u};
            if (chunk == (n >> 3)) {
                const int e = n & 7;
                w[e >> 1] = (e & 1) ? 0x3F800000u : 0x00003F80u;
                w16[e >> 1] = (e & 1) ? 0x3C000000u : 0x00003C00u;
            }
            id4[i] = make_uint4(w[0], w[1], w[2], w[3]);
            id16[i] = make_uint4(w16[0], w16[1], w16[2], w16[3]);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    constexpr uint32_t IDESC = umma_idesc_pair_bf16(256);
    constexpr uint32_t IDESC_N64 = umma_idesc_pair_bf16(64);
    constexpr uint32_t IDESC_F16 = umma_idesc_pair_f16(256);
    constexpr uint32_t IDESC_N64_F16 = umma_idesc_pair_f16(64);

    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_groups = (p.num_tiles + 1) >> 1;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        uint32_t stage = 0, phase = 0;
        long long dbg_acc[16] = {};
        TC_DBG_T0(tp_all);
        for (int grp = pair_id; grp < num_groups; grp += num_pairs) {
            const int tile = grp * 2 + rank;
            const int b_real = tile / p.tiles_per_b;      // >= B for a padding tile: TMA zero-fills
            // 2048 (timing experiment): every activation / skip access confined to a window of (dbg >> 16) samples (L2-resident)
            const int b = ((kdbg & 2048) && b_real < p.B) ? b_real % (kdbg >> 16) : b_real;
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            if ((kdbg & 4) && lane < 4 && grp + num_pairs < num_groups) {
                // warm L2 with the centre-tap rows of this CTA's NEXT tile (first touch of those rows: the dilated
                // taps re-read rows that neighbouring tiles of the same wave already pulled in)
                const int ntile = (grp + num_pairs) * 2 + rank;
                tma_prefetch_l2_3d(&tm_h, lane * 64, (ntile % p.tiles_per_b) * TC_TILE_T, ntile / p.tiles_per_b);
            }
            for (int job = 0; job < 4; ++job) {
                if (job == 2 && !p.write_h) continue;
                const int nkb = job < 2 ? 12 : 4;
                for (int kb = 0; kb < nkb; ++kb) {
                    TC_DBG_T0(tw);
                    mbar_wait(&bar_empty[stage], phase ^ 1, SITE_PROD_EMPTY, stage);
                    TC_DBG_ACC(5, tw);
                    if (lane == 0) {
                        uint8_t* sa = smem + Tc2Smem::stages + stage * T2_STAGE_BYTES;
                        uint8_t* sb = sa + T2_A_BYTES;
                        const int wblk = p.layer * TC_W_BLOCKS_PER_LAYER +
                                         (job < 2 ? job * 12 + kb : 24 + (job - 2) * 4 + kb);
                        const bool ident_job = job == 2 || (job == 3 && p.add_stash);            // the A slot carries a tile for the identity MMA
                        const uint32_t bytes = (job < 2 || ident_job) ? T2_STAGE_BYTES : T2_B_BYTES;
                        if (leader) mbar_arrive_expect_tx(&bar_full[stage], 2 * bytes);   // both CTAs' bytes
                        else        mbar_arrive_cluster(&bar_full[stage], 0);
                        if (job < 2) {
                            const int tap = kb >> 2, cib = kb & 3;
                            tma_load_3d_pair(sa, &tm_h, &bar_full[stage], cib * 64, t0 + (tap - 1) * p.dil, b);
                        } else if (job == 2) {
                            tma_load_3d_pair(sa, &tm_h, &bar_full[stage], kb * 64, t0, b);   // h for the identity MMA
                        } else if (p.add_stash) {
                            tma_load_3d_pair(sa, &tm_stash_ld, &bar_full[stage], kb * 64, t0, b);   // previous layer's skip term
                        }
                        tma_load_2d_pair(sb, &tm_w, &bar_full[stage], 0, wblk * 256 + rank * 128);
                    }
                    __syncwarp();
                    if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        TC_DBG_ACC(6, tp_all);
        if ((kdbg & 2) && lane == 0 && leader) { atomicAdd(&g_tc_cycles[5], dbg_acc[5]); atomicAdd(&g_tc_cycles[6], dbg_acc[6]); }
    } else if (warp == 1) {
        if (leader) {
            // ===================== MMA issuer (rank 0 only) =====================
            uint32_t stage = 0, phase = 0;
            uint32_t use0 = 0, use1 = 0;
            uint32_t it = 0;
            const uint32_t z_addr = smem_u32(smem + Tc2Smem::z);
            const uint32_t id_addr = smem_u32(smem + Tc2Smem::ident);
            const uint32_t id16_addr = smem_u32(smem + Tc2Smem::ident16);
            long long dbg_acc[16] = {};
            TC_DBG_T0(tm_all);
            for (int grp = pair_id; grp < num_groups; grp += num_pairs, ++it) {
                for (int job = 0; job < 4; ++job) {
                    if (job == 2 && !p.write_h) continue;
                    const int buf = job & 1;
                    const uint32_t use = buf ? use1 : use0;
                    TC_DBG_T0(tw0);
                    mbar_wait(&bar_tempty[buf], (use & 1) ^ 1, SITE_MMA_TEMPTY, job);
                    TC_DBG_ACC(job < 2 ? 0 : 1, tw0);
                    const bool first_g2 = (job == 2 || (job == 3 && !p.write_h));
                    if (first_g2) {
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
                            const uint32_t sa = smem_u32(smem + Tc2Smem::stages + stage * T2_STAGE_BYTES);
                            const uint32_t a_addr = job < 2 ? sa : z_addr + kb * TC_A_BYTES;
                            const uint32_t b_addr = sa + T2_A_BYTES;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss_pair(d_tmem, umma_desc_sw128_kmajor(a_addr + k * 32),
                                                  umma_desc_sw128_kmajor(b_addr + k * 32), job < 2 ? IDESC : IDESC_F16,
                                                  (kb | k) != 0 ? 1u : 0u);
                            if (job == 2) {
                                // + h: D[:, 64 kb .. 64 kb + 63] += h_tile * I   (exact: bf16 x 1.0 in fp32)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ss_pair(d_tmem + kb * 64, umma_desc_sw128_kmajor(sa + k * 32),
                                                      umma_desc_sw128_kmajor(id_addr + k * 32), IDESC_N64, 1u);
                            } else if (job == 3 && p.add_stash) {
                                // + the previous layer's stashed skip term (fp16 x 1.0 in fp32)
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ss_pair(d_tmem + kb * 64, umma_desc_sw128_kmajor(sa + k * 32),
                                                      umma_desc_sw128_kmajor(id16_addr + k * 32), IDESC_N64_F16, 1u);
                            }
                            umma_commit_pair_mc(&bar_empty[stage], 3);           // stage free in both CTAs
                            if (kb == nkb - 1) umma_commit_pair_mc(&bar_tfull[buf], 3);   // accumulator ready in both CTAs
                        }
                        __syncwarp();
                        if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (buf) ++use1; else ++use0;
                }
            }
            TC_DBG_ACC(4, tm_all);
            if ((kdbg & 2) && lane == 0)
                for (int i = 0; i < 5; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
        }
    } else {
        // ===================== epilogue warps (both CTAs) =====================
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* zbase = smem + Tc2Smem::z;
        uint8_t* tbuf = smem + Tc2Smem::stg + ew * T2_STG_BYTES;
        uint32_t use0 = 0, use1 = 0;
        long long dbg_acc[16] = {};
        TC_DBG_T0(te_all);
        const float* Ec = s_evec + 512;
        const float* E0 = s_evec;
        const float* E2 = s_evec + 1024;
        for (int grp = pair_id; grp < num_groups; grp += num_pairs) {
            const int tile = grp * 2 + rank;
            const bool tile_valid = tile < p.num_tiles;
            const int b = (kdbg & 2048) ? (tile / p.tiles_per_b) % (kdbg >> 16) : tile / p.tiles_per_b;
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
            const bool interior = (t0 >= p.dil) && (t0 + TC_TILE_T - 1 < p.L - p.dil);

            // ---- epilogue 1: gate ----
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
                    const int col = half * 64 + cc * 32;
                    uint32_t g[32], f[32];
                    tmem_ld_32x32(t_lane + j * 256 + col, g);
                    tmem_ld_32x32(t_lane + j * 256 + 128 + col, f);
                    tmem_ld_wait();
                    const int c0 = 128 * j + col;
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
                    uint8_t* zrow = zbase + (2 * j + half) * TC_A_BYTES + row * 128;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int chunk = (4 * cc + m) ^ (row & 7);
                        *reinterpret_cast<uint4*>(zrow + chunk * 16) =
                            make_uint4(packed[4 * m], packed[4 * m + 1], packed[4 * m + 2], packed[4 * m + 3]);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(&bar_tempty[j], 0);
                    mbar_arrive_cluster(&bar_zready[j], 0);
                }
                TC_DBG_ACC(8, tk);
            }

            // ---- epilogue 2r: residual half -> h_out. Each 32 x 32-channel chunk is transposed through this
            //      warp's 2 KB buffer so that every global store instruction writes 8 rows x 64 contiguous bytes.
            if (p.write_h) {
                TC_DBG_T0(tw3);
                mbar_wait(&bar_tfull[0], use0 & 1, SITE_EPI_TFULL, 2);
                TC_DBG_ACC(9, tw3);
                TC_DBG_T0(tk3);
                ++use0;
                tc_fence_after_sync();
                const int sw_w = (lane >> 1) & 3;                 // write swizzle of this lane's own row
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = half * 128 + cc * 32;
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
                    if (!(kdbg & 64)) {
                        // the 32 x 32-channel chunk leaves as ONE asynchronous TMA tensor store from this warp's 2 KB buffer
                        // (rows of 64 bytes, 64-byte swizzle); rows past L are clipped by the hardware
                        if (lane == 0) tma_store_wait_read<0>();       // previous chunk's store has read the buffer
                        __syncwarp();
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            *reinterpret_cast<uint4*>(tbuf + lane * 64 + ((m ^ sw_w) << 4)) =
                                make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && tile_valid && !(kdbg & 17)) {
                            tma_store_3d(&tm_hout, tbuf, col, t0 + q * 32, b);
                            tma_store_commit();
                        }
                        continue;
                    }
                    __syncwarp();                                  // previous chunk fully read back
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        *reinterpret_cast<uint4*>(tbuf + lane * 64 + ((m ^ sw_w) << 4)) =
                            make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int rr = 8 * m + (lane >> 2);        // row of the warp's 32
                        const int c = lane & 3;                    // 16-byte chunk of the 64-byte row segment
                        const uint4 v = *reinterpret_cast<const uint4*>(tbuf + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
                        const int tt = t0 + q * 32 + rr;
                        if (tile_valid && tt < p.L && !(kdbg & 17))
                            *reinterpret_cast<uint4*>(p.h_out + (static_cast<long long>(b) * p.L + tt) * TC_C + col + c * 8) = v;
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&bar_tempty[0], 0);
                TC_DBG_ACC(10, tk3);
            }

            // ---- epilogue 2s: skip half via the idle z buffer + TMA reduce-add ----
            TC_DBG_T0(tw2);
            mbar_wait(&bar_tfull[1], use1 & 1, SITE_EPI_TFULL, 3);
            TC_DBG_ACC(9, tw2);
            TC_DBG_T0(tk2);
            ++use1;
            uint8_t* stg = zbase + ew * 8192;
            const int trow = t0 + q * 32;
            {
                tc_fence_after_sync();
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = half * 128 + cc * 32;
                    uint32_t r[32];
                    tmem_ld_32x32(t_lane + 256 + col, r);
                    tmem_ld_wait();
                    uint8_t* cbuf = stg + (cc & 1) * 4096;
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                    if (p.skip_mode == 2) {
                        // stash: this layer's skip term leaves as fp16 (32 rows x 64 bytes, 64-byte swizzle) for the next layer
                        const int sw_w = (lane >> 1) & 3;
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            pk[i >> 1] = pack_f16x2_sat(__uint_as_float(r[i]) + s_b2[256 + col + i], __uint_as_float(r[i + 1]) + s_b2[256 + col + i + 1]);
#pragma unroll
                        for (int m = 0; m < 4; ++m)
                            *reinterpret_cast<uint4*>(cbuf + lane * 64 + ((m ^ sw_w) << 4)) =
                                make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && tile_valid && !(kdbg & 33)) {
                            tma_store_3d(&tm_stash_st, cbuf, col, trow, b);
                            tma_store_commit();
                        }
                        continue;
                    }
                    uint8_t* brow = cbuf + lane * 128;
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
                    if (lane == 0 && tile_valid && !(kdbg & 33)) {
                        if (p.skip_mode == 1) tma_store_3d(&tm_skip, cbuf, col, trow, b);
                        else                  tma_reduce_add_3d(&tm_skip, cbuf, col, trow, b);
                        tma_store_commit();
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(&bar_tempty[1], 0);
                    tma_store_wait_read<0>();
                }
                __syncwarp();
            }
            TC_DBG_ACC(10, tk2);
        }
        TC_DBG_ACC(11, te_all);
        if ((kdbg & 2) && warp == 2 && lane == 0 && leader)
            for (int i = 7; i < 12; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace adb
