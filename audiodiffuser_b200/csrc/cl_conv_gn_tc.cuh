// ConvBlock1d of the U-Net with the GroupNorm *apply* pass fused into the convolution's operand path
// (unet1d.py:160-195: GroupNorm -> optional x * (scale + 1) + shift -> SiLU -> Conv1d k = 3 "same"; ResnetBlock1d :297-316
// adds the residual). bf16 activations [B][L][C] channels-last, fp32 accumulation in TMEM.
//
//   Y[b][t][n] = bias[n] (+ res[b][t][n]) + sum_tap sum_ci Xn[b][t + tap - 1][ci] W[tap][ci][n]
//   Xn[b][r][ci] = 0 <= r < L ? fp16(SiLU(ca[b][ci] X[b][r][ci] + cb[b][ci])) : 0      (zero padding AFTER the activation;
//                                fp16, not bf16: the operand never reaches HBM, so its format is free, W is packed in fp16 too)
//
// The normalised tensor never exists in HBM (the unfused path writes and re-reads it: 17 % of a config-4 evaluation).
// A naive operand-path transform would run once per tap and box; here each K-block [64 ch] of a tile is staged ONCE with its
// halo (130 rows: t0 - 1 .. t0 + 128, one TMA box, zero fill outside the sample), eight transform warps rewrite it in place,
// and the three taps are three shared-memory descriptors into the same box, each one row (128 bytes) further down: the
// 128-byte swizzle is a function of the absolute shared-memory address bits, so a row-shifted start address keeps reading
// what TMA wrote. Weights stream through their own ring, one box per (K-block, tap).
//
// CTA pair, cta_group::2 (measured: the single-CTA form of this kernel and cl_conv_tc_kernel are bound by shared-memory
// bandwidth, not by the tensor pipe: an M = 128 x N = 256 tile reads and TMA-writes 96 KB of weights per K-block against 48 KB of
// activations): two CTAs of a cluster run adjacent m-tiles as ONE M = 256 MMA stream issued by rank 0; each CTA stages its own
// activation box but only HALF of every weight box.
//
// The input may be the channel concatenation of TWO tensors (UpsampleBlock1d, unet1d.py:552-556: cat(x, skip * 2^-1/2)):
// K-blocks below kb1 come from the first, the rest from the second, whose constant pre-scale is folded into ca (and into its
// group statistics), so the concatenated tensor is never written either.
//
// Per-(sample, channel) coefficients come from the statistics kernel's last block (GnCoefArgs, cl_ops.cuh; rebuilding them in
// this kernel whenever a tile's sample changed cost 1.5 k cycles per K-block on the critical path); a transform thread always
// works on the same 8 channels of a K-block, so it keeps them in registers. SiLU(y) = h tanh(h) + h, h = y / 2 (one MUFU).
//
// Warps: 0..7 epilogue, 8..15 transform, 16 TMA producer, 17 MMA issuer (rank 0). The two single-thread roles sit at the HIGHEST
// warp ids on purpose: the scheduler arbitrates highest-warp-id-first, and a producer / issuer that waits behind eight busy
// warps delays every TMA and MMA. N tile <= 256, two TMEM accumulators (a tile's epilogue overlaps the next tile's MMAs).
#pragma once
#include "cl_conv_tc.cuh"

namespace adb {

constexpr int GC_SA = 6;                        // activation slots
constexpr int GC_SB = 6;                        // weight slots
constexpr int GC_A_ROWS = 130;
constexpr int GC_A_TX = GC_A_ROWS * 128;        // bytes one activation box delivers
constexpr int GC_A_BYTES = 17 * 1024;           // slot pitch (1024-byte aligned for the swizzle)
constexpr int GC_B_BYTES = 128 * 128;           // 16 KB: this CTA's half [NT / 2][64] of a weight box (NT = 256; smaller tiles use a prefix)
constexpr int GC_A_AHEAD = 3;                   // activation boxes requested this many K-blocks before their weights
constexpr int GC_XF_WARPS = 8;                  // transform warps (two per scheduler: one hides the other's MUFU latency)
constexpr int GC_XF_THREADS = 32 * GC_XF_WARPS;
constexpr int GC_EPI_WARPS = 8;
constexpr int GC_WARP_PROD = GC_EPI_WARPS + GC_XF_WARPS;        // 16
constexpr int GC_WARP_MMA = GC_WARP_PROD + 1;                   // 17
constexpr int GC_THREADS = 64 + 32 * GC_EPI_WARPS + GC_XF_THREADS;

struct GcSmem {
    static constexpr int a = 0;
    static constexpr int b = GC_SA * GC_A_BYTES;
    static constexpr int bias = b + GC_SB * GC_B_BYTES;
    static constexpr int bars = bias + 256 * 4;
    static constexpr int tmem_ptr = bars + 40 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(GcSmem::b % 1024 == 0, "swizzle alignment");
static_assert(GC_SA % 2 == 0 && GC_XF_WARPS == 8, "two transform groups of four warps walk the slots of one parity each");
static_assert(GcSmem::total <= 232448, "shared memory budget");
constexpr int GC_SMEM_BYTES = GcSmem::total;

enum GcWaitSite : uint32_t { SITE_GC_AEMPTY = 30, SITE_GC_BEMPTY = 31, SITE_GC_ARAW = 32, SITE_GC_AREADY = 33, SITE_GC_BFULL = 34,
                             SITE_GC_TEMPTY = 35, SITE_GC_TFULL = 36 };

struct GnConvParams {
    const float* coef;              // [2][B][Cin]: halved slopes, then halved offsets (GnCoefArgs, cl_gn_stats_vec_kernel)
    const float* bias;              // [N] or nullptr
    const __nv_bfloat16* res;       // [B][L][N] or nullptr
    __nv_bfloat16* out;             // [B][L][N]
    int B, L, Cin, N;
    int NT, tiles_per_b, tiles_m, tiles_n, kb_total, kb1;
    int dbg;                        // ADB_DEBUG builds only: 2 = in-kernel cycle accounting into g_tc_cycles (tools/time_gnconv.py)
};

__global__ void __launch_bounds__(GC_THREADS, 1)
cl_conv3_gn_tc_kernel(const __grid_constant__ CUtensorMap tm_in1, const __grid_constant__ CUtensorMap tm_in2,
                      const __grid_constant__ CUtensorMap tm_w, const GnConvParams p) {
    TC_DBG_FLAGS(p);
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_bias = reinterpret_cast<float*>(smem + GcSmem::bias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GcSmem::bars);
    uint64_t* bar_araw = bars;                       // [GC_SA] per CTA: TMA landed the raw box
    uint64_t* bar_aready = bars + GC_SA;             // [GC_SA] rank 0: both CTAs transformed their boxes in place
    uint64_t* bar_aempty = bars + 2 * GC_SA;         // [GC_SA] per CTA: all three taps' MMAs have read it
    uint64_t* bar_bfull = bars + 3 * GC_SA;          // [GC_SB] rank 0: both halves of the weight box landed
    uint64_t* bar_bempty = bar_bfull + GC_SB;        // [GC_SB] per CTA
    uint64_t* bar_tfull = bar_bempty + GC_SB;        // [2] per CTA
    uint64_t* bar_tempty = bar_tfull + 2;            // [2] rank 0
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + GcSmem::tmem_ptr);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;

    if (warp == GC_WARP_PROD && lane == 0) { tma_prefetch_desc(&tm_in1); tma_prefetch_desc(&tm_in2); tma_prefetch_desc(&tm_w); }
    if (warp == GC_WARP_MMA) {
        if (lane == 0) {
            for (int s = 0; s < GC_SA; ++s) { mbar_init(&bar_araw[s], 1); mbar_init(&bar_aready[s], GC_XF_WARPS /* 2 CTAs x the 4 warps of one transform group */); mbar_init(&bar_aempty[s], 1); }
            for (int s = 0; s < GC_SB; ++s) { mbar_init(&bar_bfull[s], 2); mbar_init(&bar_bempty[s], 1); }
            for (int i = 0; i < 2; ++i) { mbar_init(&bar_tfull[i], 1); mbar_init(&bar_tempty[i], 2 * GC_EPI_WARPS); }
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
    // everything above touched only this CTA's shared memory / TMEM: under programmatic dependent launch it overlaps the tail of
    // the previous kernel; global memory is first read (and written) below
    pdl_wait();
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;
    const int pair_id = static_cast<int>(blockIdx.x) >> 1, num_pairs = static_cast<int>(gridDim.x) >> 1;
    const int groups_m = (p.tiles_m + 1) >> 1;       // a group = two adjacent m-tiles (one per CTA) of one n-tile
    const int total_groups = groups_m * p.tiles_n;
    const int my_groups = (total_groups - pair_id + num_pairs - 1) / num_pairs;
    const int items = my_groups * p.kb_total;        // (group, K-block) pairs this CTA pair walks, in order
    const uint32_t b_tx = static_cast<uint32_t>(p.NT) * 64u;      // this CTA's half of a weight box
    // group index of this pair's gi-th group -> this CTA's m-tile and the n-tile
    auto group_at = [&](int gi, int& tm, int& n0) {
        const int g = pair_id + gi * num_pairs;
        tm = (g / p.tiles_n) * 2 + rank;
        n0 = (g % p.tiles_n) * p.NT;
    };

    if (warp == GC_WARP_PROD) {
        // ===================== TMA producer (both CTAs) =====================
        uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
        int a_next = 0;                               // next activation item to request
        long long dbg_acc[16] = {};
        TC_DBG_T0(tp_all);
        auto issue_a = [&]() {
            const int gi = a_next / p.kb_total, kb = a_next - gi * p.kb_total;
            int tm, n0;
            group_at(gi, tm, n0);
            const int b = tm / p.tiles_per_b, t0 = (tm % p.tiles_per_b) * 128;     // b >= B for a padding tile: TMA zero-fills
            TC_DBG_T0(tw);
            mbar_wait(&bar_aempty[a_slot], a_phase ^ 1, SITE_GC_AEMPTY, a_slot);
            TC_DBG_ACC(4, tw);
            if (lane == 0) {
                uint8_t* sa = smem + GcSmem::a + a_slot * GC_A_BYTES;
                mbar_arrive_expect_tx(&bar_araw[a_slot], GC_A_TX);
                if (kb < p.kb1) tma_load_3d(sa, &tm_in1, &bar_araw[a_slot], kb * 64, t0 - 1, b);
                else            tma_load_3d(sa, &tm_in2, &bar_araw[a_slot], (kb - p.kb1) * 64, t0 - 1, b);
            }
            __syncwarp();
            if (++a_slot == GC_SA) { a_slot = 0; a_phase ^= 1; }
            ++a_next;
        };
        for (int i = 0; i < GC_A_AHEAD && a_next < items; ++i) issue_a();
        for (int it = 0; it < items; ++it) {
            if (a_next < items) issue_a();
            const int gi = it / p.kb_total, kb = it - gi * p.kb_total;
            int tm, n0;
            group_at(gi, tm, n0);
            for (int tap = 0; tap < 3; ++tap) {
                TC_DBG_T0(tw);
                mbar_wait(&bar_bempty[b_slot], b_phase ^ 1, SITE_GC_BEMPTY, b_slot);
                TC_DBG_ACC(5, tw);
                if (lane == 0) {
                    if (leader) mbar_arrive_expect_tx(&bar_bfull[b_slot], 2 * b_tx);     // both CTAs' bytes
                    else        mbar_arrive_cluster(&bar_bfull[b_slot], 0);
                    tma_load_2d_pair(smem + GcSmem::b + b_slot * GC_B_BYTES, &tm_w, &bar_bfull[b_slot], 0,
                                     (tap * p.kb_total + kb) * p.N + n0 + rank * (p.NT >> 1));
                }
                __syncwarp();
                if (++b_slot == GC_SB) { b_slot = 0; b_phase ^= 1; }
            }
        }
        TC_DBG_ACC(6, tp_all);
        if ((kdbg & 2) && lane == 0 && leader) for (int i = 4; i < 7; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
    } else if (warp == GC_WARP_MMA) {
        if (leader) {
            // ===================== MMA issuer (rank 0 only) =====================
            uint32_t a_slot = 0, a_phase = 0, b_slot = 0, b_phase = 0;
            const uint32_t idesc = umma_idesc_pair_f16(static_cast<uint32_t>(p.NT));     // fp16 x fp16 -> fp32
            long long dbg_acc[16] = {};
            TC_DBG_T0(tm_all);
            for (int gi = 0; gi < my_groups; ++gi) {
                const uint32_t buf = gi & 1, use = gi >> 1;
                TC_DBG_T0(tw0);
                mbar_wait(&bar_tempty[buf], (use & 1) ^ 1, SITE_GC_TEMPTY, buf);
                TC_DBG_ACC(0, tw0);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kb = 0; kb < p.kb_total; ++kb) {
                    TC_DBG_T0(tw1);
                    mbar_wait(&bar_aready[a_slot], a_phase, SITE_GC_AREADY, a_slot);
                    TC_DBG_ACC(1, tw1);
                    tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + GcSmem::a + a_slot * GC_A_BYTES);
                    for (int tap = 0; tap < 3; ++tap) {
                        TC_DBG_T0(tw2);
                        mbar_wait(&bar_bfull[b_slot], b_phase, SITE_GC_BFULL, b_slot);
                        TC_DBG_ACC(2, tw2);
                        tc_fence_after_sync();
                        if (lane == 0) {
                            const uint32_t sb = smem_u32(smem + GcSmem::b + b_slot * GC_B_BYTES);
                            // rows tap .. tap + 127 of the staged box: a start address 128 bytes further down. Measured on B200: the
                            // descriptor's "matrix base offset" field must stay 0 for this (the swizzle follows the absolute address
                            // bits [7,10), exactly as TMA wrote it); setting it to the row shift reads the wrong 16-byte chunks.
                            const uint32_t arow = sa + tap * 128;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_ss_pair(d_tmem, umma_desc_sw128_kmajor(arow + k * 32), umma_desc_sw128_kmajor(sb + k * 32),
                                                  idesc, (kb | tap | k) != 0 ? 1u : 0u);
                            umma_commit_pair_mc(&bar_bempty[b_slot], 3);
                            if (tap == 2) umma_commit_pair_mc(&bar_aempty[a_slot], 3);
                            if (tap == 2 && kb == p.kb_total - 1) umma_commit_pair_mc(&bar_tfull[buf], 3);
                        }
                        __syncwarp();
                        if (++b_slot == GC_SB) { b_slot = 0; b_phase ^= 1; }
                    }
                    if (++a_slot == GC_SA) { a_slot = 0; a_phase ^= 1; }
                }
            }
            TC_DBG_ACC(3, tm_all);
            if ((kdbg & 2) && lane == 0) for (int i = 0; i < 4; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
        }
    } else if (warp < GC_EPI_WARPS) {
        // ===================== epilogue (both CTAs): + bias (+ residual) -> bf16. Two warps per TMEM lane quarter, half of the
        //                       columns each; the residual rows are fetched BEFORE the accumulator wait
        const int q = warp & 3;
        const int half = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const int nch = p.NT / 32;                    // 2, 4 or 8 column chunks per tile
        const int c_lo = half * (nch >> 1), c_n = nch >> 1;      // this warp's chunks: c_lo .. c_lo + c_n - 1 (c_n <= 4)
        int cur_n0 = -1;
        long long dbg_acc[16] = {};
        TC_DBG_T0(te_all);
        for (int gi = 0; gi < my_groups; ++gi) {
            int tm, n0;
            group_at(gi, tm, n0);
            const int b = tm / p.tiles_per_b, t = (tm % p.tiles_per_b) * 128 + row;
            const uint32_t buf = gi & 1, use = gi >> 1;
            if (n0 != cur_n0) {
                named_bar_sync(1, 32 * GC_EPI_WARPS);
                for (int i = threadIdx.x; i < p.NT; i += 32 * GC_EPI_WARPS) s_bias[i] = p.bias ? p.bias[n0 + i] : 0.f;
                named_bar_sync(1, 32 * GC_EPI_WARPS);
                cur_n0 = n0;
            }
            const bool row_ok = tm < p.tiles_m && t < p.L;
            const long long grow = static_cast<long long>(b) * p.L + t;
            const bool has_res = p.res != nullptr && row_ok;
            uint32_t rres[2][16];                     // residual chunks k and k + 1 (two in flight)
            if (has_res) {
                ct_load64(p.res + grow * p.N + n0 + c_lo * 32, rres[0]);
                if (c_n > 1) ct_load64(p.res + grow * p.N + n0 + (c_lo + 1) * 32, rres[1]);
            }
            TC_DBG_T0(tw);
            mbar_wait(&bar_tfull[buf], use & 1, SITE_GC_TFULL, buf);
            TC_DBG_ACC(10, tw);
            TC_DBG_T0(tk);
            tc_fence_after_sync();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k >= c_n) break;
                const int cc = c_lo + k;
                uint32_t r[32];
                tmem_ld_32x32(t_lane + buf * 256 + cc * 32, r);
                tmem_ld_wait();
                if (!row_ok) continue;
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + cc * 32);
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = b4[i >> 2];
                    float v0 = __uint_as_float(r[i]) + bb.x, v1 = __uint_as_float(r[i + 1]) + bb.y;
                    float v2 = __uint_as_float(r[i + 2]) + bb.z, v3 = __uint_as_float(r[i + 3]) + bb.w;
                    if (has_res) {
                        const uint32_t w0 = rres[k & 1][i >> 1], w1 = rres[k & 1][(i >> 1) + 1];
                        v0 += __uint_as_float(w0 << 16); v1 += __uint_as_float(w0 & 0xFFFF0000u);
                        v2 += __uint_as_float(w1 << 16); v3 += __uint_as_float(w1 & 0xFFFF0000u);
                    }
                    pk[i >> 1] = pack_bf16x2(v0, v1);
                    pk[(i >> 1) + 1] = pack_bf16x2(v2, v3);
                }
                if (has_res && k + 2 < c_n) ct_load64(p.res + grow * p.N + n0 + (cc + 2) * 32, rres[k & 1]);
                ct_store64(p.out + grow * p.N + n0 + cc * 32, pk);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bar_tempty[buf], 0);
            TC_DBG_ACC(11, tk);
        }
        TC_DBG_ACC(12, te_all);
        if ((kdbg & 2) && warp == 0 && lane == 0 && leader) for (int i = 10; i < 13; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
    } else {
        // ===================== transform warps (both CTAs): raw box -> SiLU(ca x + cb), rows outside the sample -> 0 =====================
        const int tt = threadIdx.x - 32 * GC_EPI_WARPS;               // 0 .. GC_XF_THREADS - 1
        // The eight warps work as TWO groups of four on alternate K-blocks: a K-block's transform is a latency chain (coefficients
        // from L2, box from shared memory, MUFU, stores, the proxy fence: ~1.7 k cycles whatever the arithmetic costs — measured),
        // so two boxes in flight double the rate. A thread always works on the same 16-byte chunk column and the same row
        // phase (128 threads = 16 rows x 8 chunks per pass), so the 8 channels it transforms are fixed within a K-block: their
        // coefficients live in registers.
        const int xg = tt >> 7, tg = tt & 127;
        const int pc = tg & 7, r_base = tg >> 3;      // r_base 0..15
        const int lc8 = (pc ^ (r_base & 7)) << 3;     // 128-byte swizzle: logical chunk = physical chunk ^ (row % 8)
        uint32_t a_slot = xg, a_phase = 0;            // GC_SA is even: group g walks the slots of parity g
        long long dbg_acc[16] = {};
        TC_DBG_T0(tx_all);
        for (int item = xg; item < items; item += 2) {
            const int gi = item / p.kb_total, kb = item - gi * p.kb_total;
            int tm, n0;
            group_at(gi, tm, n0);
            const bool tile_ok = tm < p.tiles_m;
            const int b = tile_ok ? tm / p.tiles_per_b : 0, t0 = (tm % p.tiles_per_b) * 128;
            {
                // this K-block's coefficients: 64 bytes per thread from L2, requested before the wait for the box
                const float4* cap = reinterpret_cast<const float4*>(p.coef + static_cast<long long>(b) * p.Cin + kb * 64 + lc8);
                const float4* cbp = reinterpret_cast<const float4*>(p.coef + (static_cast<long long>(p.B) + b) * p.Cin + kb * 64 + lc8);
                const float4 a0 = __ldg(cap), a1 = __ldg(cap + 1), b0 = __ldg(cbp), b1 = __ldg(cbp + 1);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                TC_DBG_T0(tw);
                mbar_wait(&bar_araw[a_slot], a_phase, SITE_GC_ARAW, a_slot);
                TC_DBG_ACC(7, tw);
                TC_DBG_T0(tk);
                uint8_t* col = smem + GcSmem::a + a_slot * GC_A_BYTES + pc * 16 + r_base * 128;
                // bf16 in, fp16 out (the MMA's A operand): y = slope x + offset in fp32 (the mean subtraction needs it), then
                // SiLU(y) = h tanh(h) + h, h = y / 2, on packed fp16 pairs: one MUFU and one HFMA2 per TWO elements
                auto silu8 = [&](const uint4& xin, bool valid) {
                    const uint32_t xw[4] = {xin.x, xin.y, xin.z, xin.w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float h0 = fmaf(__uint_as_float(xw[e] << 16), av[2 * e], bv[2 * e]);
                        const float h1 = fmaf(__uint_as_float(xw[e] & 0xFFFF0000u), av[2 * e + 1], bv[2 * e + 1]);
                        const uint32_t hp = pack_f16x2_sat(h0, h1);
                        const uint32_t tp = (kdbg & 4) ? hp : tanh_f16x2(hp);     // ADB_DEBUG timing experiment 4 (results wrong): no MUFU
                        o[e] = valid ? hfma2(hp, tp, hp) : 0u;
                    }
                    return make_uint4(o[0], o[1], o[2], o[3]);
                };
                const int gr0 = t0 - 1 + r_base;
                if (!(kdbg & 16)) {               // ADB_DEBUG timing experiment 16 (results wrong): the transform does nothing
                // rows r_base + 16 i, i = 0..7 (every thread) in two batches of four (loads first, arithmetic interleaved, stores
                // last: branch-free), then rows 128 + r_base for r_base < 2. (r % 8) == (r_base % 8) in every pass.
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    uint4 x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = *reinterpret_cast<const uint4*>(col + (4 * hb + i) * 2048);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int gr = gr0 + 16 * (4 * hb + i);
                        x[i] = silu8(x[i], tile_ok && gr >= 0 && gr < p.L);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(col + (4 * hb + i) * 2048) = x[i];
                }
                if (r_base < GC_A_ROWS - 128) {
                    uint4 xt = *reinterpret_cast<const uint4*>(col + 8 * 2048);
                    xt = silu8(xt, tile_ok && gr0 + 128 < p.L);
                    *reinterpret_cast<uint4*>(col + 8 * 2048) = xt;
                }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&bar_aready[a_slot], 0);
                TC_DBG_ACC(8, tk);
                a_slot += 2;
                if (a_slot >= GC_SA) { a_slot -= GC_SA; a_phase ^= 1; }
            }
        }
        TC_DBG_ACC(9, tx_all);
        if ((kdbg & 2) && tt == 0 && leader) for (int i = 7; i < 10; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == GC_WARP_MMA) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace adb
