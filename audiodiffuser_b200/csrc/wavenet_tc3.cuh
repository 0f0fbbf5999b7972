// Residual-block kernel, third generation ("z-stash"), and the skip GEMM that goes with it.
//
// wavenet.py:94-115 per block:  y = DilatedConv(h + p) ; z = sigmoid(y_gate) * tanh(y_filter) ; o = W2 z + b2 ;
//                               h' = (h + o[:C]) / sqrt(2) ; skip_l = o[C:]
// wavenet.py:145-151:           skip = sum_l skip_l ; tail(skip / sqrt(layers))
//
// The sum over blocks of the skip halves is ONE contraction with K = layers x 256:
//     skip[t][n] = sum_l sum_k z_l[t][k] W2s_l[n][k] + sum_l b2s_l[n]
// so the block kernel does not compute its skip half at all. It only leaves its gated activations z_l (fp16, 2 bytes per
// element, already sitting in shared memory as GEMM2's A operand in exactly the layout a TMA tensor store wants) in a
// per-layer stash, and `wavenet_skip_gemm_kernel` contracts the whole stash against the stacked skip weights afterwards
// with the accumulator resident in TMEM across all layers: the fp32 skip sum is written once per evaluation instead of
// being read-modify-written by every (second) block. Relative to the pair kernel of wavenet_tc2.cuh a block launch loses
//   * the skip GEMM job (1/8 of its MMA work, 128 KB of weight loads and 64 KB of stash loads per tile),
//   * epilogue 2s (128 KB of TMEM reads, 128 KB of staging writes, the fp32 TMA reduce-adds at L2),
//   * the identity MMAs and the 64 KB activation re-load that added the residual input: the epilogue reads its own 64
//     bytes of h per chunk straight from global memory (L2 hits: the tile was just loaded as the centre tap),
// which is what the shared-memory port (MMA operand reads + TMA writes + epilogue LSU traffic share 128 B/clk) and the
// L2 write path were saturated with. Everything else (CTA pair, cta_group::2 M = 256 MMAs, 4 x 32 KB stages, exact
// step-embedding fold, fp16 GEMM2) is the design of wavenet_tc2.cuh.
//
// TMEM: 2 accumulators x 256 columns; the jobs of a tile group (G1a, G1b, G2r) simply alternate between them, so the
// accumulator a job uses is (running job counter & 1).
#pragma once
#include "wavenet_tc.cuh"

namespace adb {

constexpr int T3_STAGES = 4;
constexpr int T3_A_BYTES = 128 * 64 * 2;            // 16 KB: this CTA's [128 t][64 ci] activation block
constexpr int T3_B_BYTES = 128 * 64 * 2;            // 16 KB: this CTA's half [128 n][64 k] of a weight tile
constexpr int T3_STAGE_BYTES = T3_A_BYTES + T3_B_BYTES;
constexpr int T3_STG_BYTES = 4096;                  // per epilogue warp: two [32 t][32 ch] bf16 boxes (64-byte rows) for h'

struct Tc3Smem {
    static constexpr int stages = 0;
    static constexpr int z = T3_STAGES * T3_STAGE_BYTES;            // 64 KB gated activations (GEMM2 A operand, stash source)
    static constexpr int stg = z + TC_Z_BYTES;                      // 8 x 4 KB h' staging
    static constexpr int esum = stg + 8 * T3_STG_BYTES;             // 512 fp32
    static constexpr int bars = esum + 512 * 4;
    static constexpr int tmem_ptr = bars + 16 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(Tc3Smem::z % 1024 == 0 && Tc3Smem::stg % 1024 == 0, "swizzle alignment");
static_assert(Tc3Smem::total <= 232448, "shared memory budget");
constexpr int TC3_SMEM_BYTES = Tc3Smem::total;

struct BlockZsParams {
    const float* E;                 // [B][layers][3][512] epilogue-1 constants for this evaluation
    const float* b2;                // [512] output-projection bias of this layer (first 256 used here)
    const __nv_bfloat16* h_in;      // [B][L][256] (residual input, read by epilogue 2)
    __nv_bfloat16* h_out_dbg;       // ADB_DEBUG builds only: h' for the LSU-store timing experiment
    __nv_bfloat16* y_out;           // training forward: pre-gate activations of this block, [B][L][512] bf16 as [gate 0..255 | filter
                                    // 0..255] (what the backward's gate-derivative pass reads), or nullptr (sampling)
    int B, L, layer, layers, dil;
    int tiles_per_b, num_tiles;
    int write_h;                    // 0 on the last layer (its residual output is never used, wavenet.py:145-151)
    __nv_bfloat16* zb_out;          // training forward: this block's stash slot [B][L][256], written as bf16 from registers instead of
                                    // the fp16 TMA store (the backward's weight-gradient GEMM and the training skip GEMM take bf16)
    int zrow0;                      // first "batch" coordinate of this layer's slot in the stash map (slot * B)
    int hi_roles;                   // 1: producer / MMA issuer on warps 8 / 9 (highest scheduler priority), epilogue on 0..7
    int e_uniform;                  // 1: every sample has the same noise level (sampling: sigma is a scalar), E holds ONE sample's constants
    // multi-layer launch (ML = true): all blocks of the chunk in wavefront order
    const __nv_bfloat16* h_in2;     // pong buffer (residual input of odd layers)
    const float* const* b2_tab;     // per-layer output-projection biases
    unsigned int* ml_flags;         // [layers][num_tiles] completion counters of the h' stores (zero at launch; a tile is done at 8)
    int cycle;                      // dilation = 1 << (layer % cycle)
    int ml_S;                       // samples per sub-pass
    int ml_items, ml_items_per_sp;  // (sub-pass, layer, group) items of the launch / of one full sub-pass
    int dbg;                        // ADB_DEBUG builds only: 2 = in-kernel cycle accounting; timing experiments (results wrong):
                                    // 4 no A re-load for G1b, 8 no stash stores, 16 no h' stores, 32 no gate math, 64 no residual read,
                                    // 1024 all stores to a fixed L2-resident tile per CTA
};

// Job order of one CTA pair over its n tile groups. Types: 0 = G1a (gate / filter of channels 0..127), 1 = G1b (128..255),
// 2 = G2r (residual projection). PIPE = false: G1a(i) G1b(i) G2r(i) per group. PIPE = true (software-pipelined):
//     G1a(0) G1b(0) | G1a(1) G2r(0) G1b(1) | G1a(2) G2r(1) G1b(2) | ... | G2r(n-1)
// G2r(i) needs the gated activations of BOTH halves, i.e. epilogue 1b of G1b(i), which takes ~2 k cycles after G1b(i)
// completes; issuing the next group's G1a in between gives the tensor pipe 6 k cycles of independent work instead of a
// bubble. The producer, the MMA issuer and the epilogue warps all walk this sequence; slot s of the pipelined order that
// has no job (i + 1 == n) returns false.
template <bool PIPE>
__host__ __device__ __forceinline__ int zs_num_slots(int n, int write_h) {
    if (!write_h) return 2 * n;
    return PIPE ? 2 + 3 * n : 3 * n;
}
template <bool PIPE>
__host__ __device__ __forceinline__ bool zs_job_at(int s, int n, int write_h, int& type, int& i) {
    if (!write_h) { i = s >> 1; type = s & 1; return true; }
    if (!PIPE) { i = s / 3; type = s - 3 * i; return true; }
    if (s < 2) { i = 0; type = s; return true; }
    const int u = s - 2, g = u / 3, r = u - 3 * g;
    if (r == 1) { i = g; type = 2; return true; }
    i = g + 1; type = r == 0 ? 0 : 1;
    return i < n;
}

// Wavefront order of the multi-layer launch, shared by the kernel and the host-side check of its dependency structure
// (adb_debug_ml_order): item n -> its block (layer), the first tile of its group (the pair's rank-0 tile; rank 1 takes the next one)
// and the end of its sub-pass's tile range. Sub-passes of S samples, then layers, then tile groups.
__host__ __device__ __forceinline__ void ml_item_decode(int n, int items_per_sp, int S, int tiles_per_b, int num_tiles, int& layer,
                                                        int& tile0, int& t_end) {
    const int sp = n / items_per_sp, r = n - sp * items_per_sp;
    const int t_begin = sp * S * tiles_per_b;
    t_end = t_begin + S * tiles_per_b;
    if (t_end > num_tiles) t_end = num_tiles;
    const int groups = (t_end - t_begin + 1) >> 1;
    layer = r / groups;
    tile0 = t_begin + 2 * (r - layer * groups);
}
// Tiles of block layer - 1 that block `layer` of `tile` must see finished: everything it reads (+-dil rows) and everything whose
// readers its h' store would overrun (+-previous dilation: ping / pong are reused every second block), inside the tile's sample.
__host__ __device__ __forceinline__ void ml_dep_range(int layer, int tile, int cycle, int tiles_per_b, int& lo, int& hi) {
    const int dil = 1 << (layer % cycle), dprev = 1 << ((layer - 1) % cycle);
    const int dmax = dil > dprev ? dil : dprev;
    const int kt = (dmax + TC_TILE_T - 1) / TC_TILE_T;
    const int tb0 = (tile / tiles_per_b) * tiles_per_b, tb1 = tb0 + tiles_per_b - 1;
    lo = tile - kt < tb0 ? tb0 : tile - kt;
    hi = tile + kt > tb1 ? tb1 : tile + kt;
}

// ML = true: ONE launch runs ALL blocks of a 256-sample chunk as a wavefront over (sub-pass of S samples, layer, tile group): block
// l + 1 of a tile starts as soon as block l has written the +-d neighbourhood it reads (per-tile completion flags in global
// memory; the groups are dealt to the CTA pairs round-robin in wavefront order, so a pair only ever waits for EARLIER items and
// all pairs are co-resident: no deadlock). With S = 2..4 the residual stream of a sub-pass (ping + pong, 2 x S x 8.2 MB) stays in
// L2 across the 36 blocks: h is read from and h' overwritten in L2, only the z stash streams to HBM. Layer l reads ping for even
// l and pong for odd l (tm_h / tm_h2, p.h_in / p.h_in2) and writes the other (tm_hout / tm_hout2).
template <bool PIPE, bool ML>
__global__ void __launch_bounds__(TC_THREADS, 1)
wavenet_block_zs_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                        const __grid_constant__ CUtensorMap tm_hout, const __grid_constant__ CUtensorMap tm_zst,
                        const __grid_constant__ CUtensorMap tm_h2, const __grid_constant__ CUtensorMap tm_hout2,
                        const BlockZsParams p) {
    TC_DBG_FLAGS(p);
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_esum = reinterpret_cast<float*>(smem + Tc3Smem::esum);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Tc3Smem::bars);
    uint64_t* bar_full = bars;                   // [T3_STAGES] (rank 0's copy is the live one)
    uint64_t* bar_empty = bars + T3_STAGES;      // [T3_STAGES] per CTA
    uint64_t* bar_tfull = bars + 2 * T3_STAGES;  // [2] per CTA
    uint64_t* bar_tempty = bar_tfull + 2;        // [2] rank 0
    uint64_t* bar_zready = bar_tempty + 2;       // [2] rank 0
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + Tc3Smem::tmem_ptr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;
    // warp roles: the scheduler arbitrates highest-warp-id-first, so with hi_roles the two single-thread roles (TMA producer, MMA
    // issuer) take warps 8 / 9 and never wait behind the eight epilogue warps (0..7); otherwise warps 0 / 1 and epilogue 2..9
    const int w_prod = p.hi_roles ? 8 : 0, w_mma = p.hi_roles ? 9 : 1, ew0 = p.hi_roles ? 0 : 2;

    if (warp == w_prod && lane == 0) {
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_hout);
        tma_prefetch_desc(&tm_zst);
    }
    if (warp == w_mma) {
        if (lane == 0) {
            for (int s = 0; s < T3_STAGES; ++s) { mbar_init(&bar_full[s], 2); mbar_init(&bar_empty[s], 1); }
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
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    constexpr uint32_t IDESC = umma_idesc_pair_bf16(256);
    constexpr uint32_t IDESC_F16 = umma_idesc_pair_f16(256);

    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_groups = ML ? p.ml_items : (p.num_tiles + 1) >> 1;           // ML: (sub-pass, layer, group) items in wavefront order
    const int n_mine = (num_groups - pair_id + num_pairs - 1) / num_pairs;     // tile groups of this CTA pair (>= 1)
    const int n_slots = zs_num_slots<PIPE>(n_mine, p.write_h);
    // this CTA's tile of the pair's gi-th group, and (ML) the block it belongs to
    auto item_at = [&](int gi, int& layer, int& tile, bool& valid) {
        if constexpr (!ML) {
            layer = p.layer;
            tile = (pair_id + gi * num_pairs) * 2 + rank;
            valid = tile < p.num_tiles;
        } else {
            int tile0, t_end;
            ml_item_decode(pair_id + gi * num_pairs, p.ml_items_per_sp, p.ml_S, p.tiles_per_b, p.num_tiles, layer, tile0, t_end);
            tile = tile0 + rank;
            valid = tile < t_end;
        }
    };

    if (warp == w_prod) {
        // ===================== TMA producer (both CTAs) =====================
        uint32_t stage = 0, phase = 0;
        long long dbg_acc[12] = {};
        TC_DBG_T0(tp_all);
        for (int s = 0; s < n_slots; ++s) {
            int job, gi;
            if (!zs_job_at<PIPE>(s, n_mine, p.write_h, job, gi)) continue;
            int layer, tile;
            bool tvalid;
            item_at(gi, layer, tile, tvalid);
            const int b = tvalid ? tile / p.tiles_per_b : p.B;      // >= B for a padding tile: TMA zero-fills
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            const int dil = ML ? 1 << (layer % p.cycle) : p.dil;
            const int nkb = job < 2 ? 12 : 4;
            if (ML && job == 0 && layer > 0 && tvalid) {
                // wavefront dependency: the previous block must have finished every tile this one reads (+-dil rows) and every tile whose
                // readers this one's h' store would overrun (+-previous dilation: ping / pong are reused every second block)
                int lo, hi;
                ml_dep_range(layer, tile, p.cycle, p.tiles_per_b, lo, hi);
                const unsigned int* fl = p.ml_flags + static_cast<long long>(layer - 1) * p.num_tiles;
                TC_DBG_T0(tf);
                for (int j = lo + lane; j <= hi; j += 32) {
                    unsigned int polls = 0;
                    while (ld_acquire_gpu_u32(fl + j) < TC_EPI_THREADS / 32) {
                        __nanosleep(64);
                        if ((++polls & 0xFFF) == 0 && *reinterpret_cast<volatile unsigned int*>(&g_spin_guard.abort_flag) != 0) break;
                        if (polls > (1u << 24)) {               // ~2 s: give up loudly instead of hanging the GPU
                            if (atomicCAS(&g_spin_guard.abort_flag, 0u, 1u) == 0u) {
                                g_spin_guard.site = SITE_ML_FLAG; g_spin_guard.block = blockIdx.x; g_spin_guard.aux = layer;
                                __threadfence();
                            }
                            break;
                        }
                    }
                }
                __syncwarp();
                fence_proxy_async_all();            // the neighbours' TMA stores (async proxy) before this CTA's TMA loads
                TC_DBG_ACC(7, tf);
            }
            const CUtensorMap* tmh = (ML && (layer & 1)) ? &tm_h2 : &tm_h;
            for (int kb = 0; kb < nkb; ++kb) {
                TC_DBG_T0(tw);
                mbar_wait(&bar_empty[stage], phase ^ 1, SITE_PROD_EMPTY, stage);
                TC_DBG_ACC(5, tw);
                if (lane == 0) {
                    uint8_t* sa = smem + Tc3Smem::stages + stage * T3_STAGE_BYTES;
                    uint8_t* sb = sa + T3_A_BYTES;
                    const int wblk = layer * TC_W_BLOCKS_PER_LAYER + (job < 2 ? job * 12 + kb : 24 + kb);
                    const bool load_a = job < 2 && !((kdbg & 4) && job == 1);     // timing experiment: G1b re-uses stale A tiles
                    const uint32_t bytes = load_a ? T3_STAGE_BYTES : T3_B_BYTES;
                    if (leader) mbar_arrive_expect_tx(&bar_full[stage], 2 * bytes);   // both CTAs' bytes
                    else        mbar_arrive_cluster(&bar_full[stage], 0);
                    if (load_a) {
                        const int tap = kb >> 2, cib = kb & 3;
                        // 2048 (timing experiment): activations confined to a window of (dbg >> 16) samples, i.e. L2-resident
                        const int lb = (kdbg & 2048) ? b % (kdbg >> 16) : b;
                        tma_load_3d_pair(sa, tmh, &bar_full[stage], cib * 64, t0 + (tap - 1) * dil, lb);
                    }
                    tma_load_2d_pair(sb, &tm_w, &bar_full[stage], 0, wblk * 256 + rank * 128);
                }
                __syncwarp();
                if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        TC_DBG_ACC(6, tp_all);
        if ((kdbg & 2) && lane == 0 && leader) { atomicAdd(&g_tc_cycles[5], dbg_acc[5]); atomicAdd(&g_tc_cycles[6], dbg_acc[6]); atomicAdd(&g_tc_cycles[12], dbg_acc[7]); }
    } else if (warp == w_mma) {
        if (leader) {
            // ===================== MMA issuer (rank 0 only) =====================
            uint32_t stage = 0, phase = 0;
            uint32_t jg = 0;                      // jobs issued so far: accumulator = jg & 1, its use count = jg >> 1
            const uint32_t z_addr = smem_u32(smem + Tc3Smem::z);
            long long dbg_acc[12] = {};
            TC_DBG_T0(tm_all);
            for (int s = 0; s < n_slots; ++s) {
                int job, gi;
                if (!zs_job_at<PIPE>(s, n_mine, p.write_h, job, gi)) continue;
                const uint32_t buf = jg & 1;
                TC_DBG_T0(tw0);
                mbar_wait(&bar_tempty[buf], ((jg >> 1) & 1) ^ 1, SITE_MMA_TEMPTY, job);
                TC_DBG_ACC(job < 2 ? 0 : 1, tw0);
                if (job == 2) {
                    // z K-blocks 0,1 come from epilogue 1a, K-blocks 2,3 from epilogue 1b (second wait inside the K loop)
                    TC_DBG_T0(tw1);
                    mbar_wait(&bar_zready[0], gi & 1, SITE_MMA_ZREADY, 0);
                    TC_DBG_ACC(2, tw1);
                }
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * 256;
                const int nkb = job < 2 ? 12 : 4;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (job == 2 && kb == 2) {
                        TC_DBG_T0(tw1);
                        mbar_wait(&bar_zready[1], gi & 1, SITE_MMA_ZREADY, 1);
                        TC_DBG_ACC(2, tw1);
                        tc_fence_after_sync();
                    }
                    TC_DBG_T0(tw2);
                    mbar_wait(&bar_full[stage], phase, SITE_MMA_FULL, stage);
                    TC_DBG_ACC(3, tw2);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + Tc3Smem::stages + stage * T3_STAGE_BYTES);
                        const uint32_t a_addr = job < 2 ? sa : z_addr + kb * TC_A_BYTES;
                        const uint32_t b_addr = sa + T3_A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(d_tmem, umma_desc_sw128_kmajor(a_addr + k * 32),
                                              umma_desc_sw128_kmajor(b_addr + k * 32), job < 2 ? IDESC : IDESC_F16,
                                              (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair_mc(&bar_empty[stage], 3);           // stage free in both CTAs
                        if (kb == nkb - 1) umma_commit_pair_mc(&bar_tfull[buf], 3);   // accumulator ready in both CTAs
                    }
                    __syncwarp();
                    if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                }
                ++jg;
            }
            TC_DBG_ACC(4, tm_all);
            if ((kdbg & 2) && lane == 0)
                for (int i = 0; i < 5; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
        }
    } else {
        // ===================== epilogue warps (both CTAs) =====================
        const int ew = warp - ew0;
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int half = ew >> 2;             // which half of the columns this warp handles
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* zbase = smem + Tc3Smem::z;
        // the two 4 KB pieces of the z buffer this warp owns: rows 32 q .. 32 q + 31 of K-blocks (half) and (2 + half).
        // It writes its z values there and hands the same bytes to TMA for the stash store.
        uint8_t* own0 = zbase + half * TC_A_BYTES + q * 4096;
        uint8_t* own1 = zbase + (2 + half) * TC_A_BYTES + q * 4096;
        uint8_t* tbuf = smem + Tc3Smem::stg + ew * T3_STG_BYTES;
        uint32_t jg = 0;
        int flag_tile = -1, flag_layer = 0;   // ML: tile whose h' stores are in flight; its completion flag is raised once they have landed
        // raise the pending completion flag: this warp's h' stores (and everything else it committed) are performed, then one
        // release-add per warp; a tile is complete when its 8 epilogue warps have added
        auto raise_flag = [&]() {
            if (ML && flag_tile >= 0) {
                if (lane == 0) {
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    fence_proxy_async_all();
                    __threadfence();
                    atomicAdd(p.ml_flags + static_cast<long long>(flag_layer) * p.num_tiles + flag_tile, 1u);
                }
                flag_tile = -1;
            }
        };
        uint32_t zp[2][16];                   // gated activations of a G1a job whose z write waits for the previous G2r (PIPE)
        bool z_pending = false;
        int pend_t0 = 0, pend_b = 0, pend_zrow = 0;
        bool pend_valid = false;
        long long dbg_acc[12] = {};
        TC_DBG_T0(te_all);

        // write this warp's 32 x 64-channel piece of z (from zp) into K-block 2 j + half, publish it to the MMA issuer and send it
        // to the stash
        auto publish_z = [&](int j, int zt0, int zb, bool zvalid, int zrow0) {
            uint8_t* own = j ? own1 : own0;
            if (lane == 0) tma_store_wait_read<0>();      // the previous stash store from these bytes has been read
            __syncwarp();
            uint8_t* zrow = own + lane * 128;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int chunk = (4 * cc + m) ^ (lane & 7);
                    *reinterpret_cast<uint4*>(zrow + chunk * 16) = make_uint4(zp[cc][4 * m], zp[cc][4 * m + 1], zp[cc][4 * m + 2], zp[cc][4 * m + 3]);
                }
            fence_proxy_async_smem();
            __syncwarp();
            if (p.zb_out) {
                // 32 columns (2 j + half) * 64 + 32 cc .. + 31 of this thread's row: 64 contiguous bytes per cc
                const int zt = zt0 + q * 32 + lane;
                if (zvalid && zt < p.L) {
                    __nv_bfloat16* zr = p.zb_out + (static_cast<long long>(zb) * p.L + zt) * TC_C + (2 * j + half) * 64;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t lo[8], hi[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { lo[i] = f16x2_to_bf16x2(zp[cc][i]); hi[i] = f16x2_to_bf16x2(zp[cc][8 + i]); }
                        stg256(zr + cc * 32, lo);
                        stg256(zr + cc * 32 + 16, hi);
                    }
                }
            }
            if (lane == 0) {
                mbar_arrive_cluster(&bar_zready[j], 0);
                if (zvalid && !p.zb_out && !(kdbg & 8)) {
                    // 1024 (timing experiment): every CTA stores to its own fixed tile, so the stores never leave L2
                    const int st0 = (kdbg & 1024) ? (static_cast<int>(blockIdx.x) % p.tiles_per_b) * TC_TILE_T : zt0;
                    const int sb = (kdbg & 1024) ? static_cast<int>(blockIdx.x) / p.tiles_per_b : zb;
                    tma_store_3d(&tm_zst, own, (2 * j + half) * 64, st0 + q * 32, zrow0 + sb);
                    tma_store_commit();
                }
            }
            __syncwarp();
        };

        for (int s = 0; s < n_slots; ++s) {
            int job, gi;
            if (!zs_job_at<PIPE>(s, n_mine, p.write_h, job, gi)) continue;
            int layer, tile;
            bool tile_valid;
            item_at(gi, layer, tile, tile_valid);
            const int b = tile / p.tiles_per_b;
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            const int t = t0 + row;
            const int dil = ML ? 1 << (layer % p.cycle) : p.dil;
            const int zrow0 = ML ? layer * p.B : p.zrow0;
            const uint32_t buf = jg & 1;
            const uint32_t par = (jg >> 1) & 1;
            ++jg;
            if (job < 2) {
                // ---- epilogue 1: gate ----
                const float* Eg = p.E + (static_cast<long long>((tile_valid && !p.e_uniform) ? b : 0) * p.layers + layer) * 1536;
                if (job == 0) {
                    // interior tiles (every row sees all three taps) add ONE vector: E0 + E1 + E2, gate half pre-scaled by 1/2
                    named_bar_sync(1, TC_EPI_THREADS);            // everyone is done with the previous group's vector
                    for (int i = threadIdx.x - 32 * ew0; i < 512; i += TC_EPI_THREADS)
                        s_esum[i] = (Eg[i] + Eg[512 + i] + Eg[1024 + i]) * (i < 256 ? 0.5f : 1.0f);
                    named_bar_sync(1, TC_EPI_THREADS);
                }
                const float m_lo = (t >= dil) ? 1.0f : 0.0f;
                const float m_hi = (t < p.L - dil) ? 1.0f : 0.0f;
                const bool interior = (t0 >= dil) && (t0 + TC_TILE_T - 1 < p.L - dil);
                const int j = job;
                raise_flag();                  // the previous group's h' stores have had an MMA job's time to land
                TC_DBG_T0(tw);
                mbar_wait(&bar_tfull[buf], par, SITE_EPI_TFULL, j);
                TC_DBG_ACC(7, tw);
                TC_DBG_T0(tk);
                tc_fence_after_sync();
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col = half * 64 + cc * 32;
                    uint32_t g[32], f[32];
                    tmem_ld_32x32(t_lane + buf * 256 + col, g);
                    tmem_ld_32x32(t_lane + buf * 256 + 128 + col, f);
                    tmem_ld_wait();
                    const int c0 = 128 * j + col;
                    if (kdbg & 32) {                               // timing experiment: no gate math
#pragma unroll
                        for (int i = 0; i < 32; i += 2) zp[cc][i >> 1] = pack_f16x2(__uint_as_float(g[i]), __uint_as_float(f[i + 1]));
                    } else if (interior) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float g0 = fmaf(__uint_as_float(g[i]), 0.5f, s_esum[c0 + i]);
                            const float g1 = fmaf(__uint_as_float(g[i + 1]), 0.5f, s_esum[c0 + i + 1]);
                            const float f0 = __uint_as_float(f[i]) + s_esum[256 + c0 + i];
                            const float f1 = __uint_as_float(f[i + 1]) + s_esum[256 + c0 + i + 1];
                            const uint32_t tg = tanh_f16x2(pack_f16x2(g0, g1));
                            const uint32_t tf = tanh_f16x2(pack_f16x2(f0, f1));
                            zp[cc][i >> 1] = hmul2(hfma2(tg, 0x38003800u, 0x38003800u), tf);
                            if (p.y_out) { g[i >> 1] = pack_bf16x2(2.0f * g0, 2.0f * g1); f[i >> 1] = pack_bf16x2(f0, f1); }   // g0 = y_gate / 2
                        }
                    } else {
                        // boundary tile (~4 % of the tiles): per-row tap masks, constants straight from global memory
                        const float* Ec = Eg + 512;
                        const float* E0 = Eg;
                        const float* E2 = Eg + 1024;
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            float gv[2], fv[2];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int c = c0 + i + u;
                                gv[u] = __uint_as_float(g[i + u]) + __ldg(Ec + c);
                                fv[u] = __uint_as_float(f[i + u]) + __ldg(Ec + 256 + c);
                                gv[u] = fmaf(m_lo, __ldg(E0 + c), gv[u]);
                                fv[u] = fmaf(m_lo, __ldg(E0 + 256 + c), fv[u]);
                                gv[u] = 0.5f * fmaf(m_hi, __ldg(E2 + c), gv[u]);
                                fv[u] = fmaf(m_hi, __ldg(E2 + 256 + c), fv[u]);
                            }
                            const uint32_t tg = tanh_f16x2(pack_f16x2(gv[0], gv[1]));
                            const uint32_t tf = tanh_f16x2(pack_f16x2(fv[0], fv[1]));
                            zp[cc][i >> 1] = hmul2(hfma2(tg, 0x38003800u, 0x38003800u), tf);
                            if (p.y_out) { g[i >> 1] = pack_bf16x2(2.0f * gv[0], 2.0f * gv[1]); f[i >> 1] = pack_bf16x2(fv[0], fv[1]); }
                        }
                    }
                    if (p.y_out && tile_valid && t < p.L) {
                        // training: keep y for the backward (64 contiguous bytes per half: two full-sector 256-bit stores each)
                        __nv_bfloat16* yrow = p.y_out + (static_cast<long long>(b) * p.L + t) * 512 + c0;
                        uint32_t lo[8], hi[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { lo[i] = g[i]; hi[i] = g[8 + i]; }
                        stg256(yrow, lo); stg256(yrow + 16, hi);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { lo[i] = f[i]; hi[i] = f[8 + i]; }
                        stg256(yrow + 256, lo); stg256(yrow + 256 + 16, hi);
                    }
                }
                // the accumulator is drained: hand it back before touching shared memory
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&bar_tempty[buf], 0);
                // K-blocks 0,1 of z are still being read by the PREVIOUS group's G2r when that job was issued after this G1a
                // (pipelined order): keep the values in registers until its accumulator-ready event, which is next in line.
                // G1b is always issued after the previous G2r, so its completion already implies that z is free.
                if (PIPE && p.write_h && j == 0 && gi > 0) {
                    z_pending = true; pend_t0 = t0; pend_b = b; pend_valid = tile_valid; pend_zrow = zrow0;
                } else {
                    publish_z(j, t0, b, tile_valid, zrow0);
                }
                TC_DBG_ACC(8, tk);
            } else {
                // ---- epilogue 2: h' = (h + W2r z + b) / sqrt(2) -> bf16. Each 32 x 32-channel chunk is transposed through one of
                //      this warp's two 2 KB boxes (64-byte rows, 64-byte swizzle) and leaves as an asynchronous TMA tensor store.
                uint32_t hres[4][16];         // residual input: this thread's row, 128 channels, fetched before the accumulator wait
                {
                    const bool ok = tile_valid && t < p.L && !(kdbg & 64);     // 64: timing experiment without the residual read
                    const int rb = (kdbg & 2048) ? b % (kdbg >> 16) : b;
                    const __nv_bfloat16* hsrc = (ML && (layer & 1)) ? p.h_in2 : p.h_in;
                    const __nv_bfloat16* hrow = hsrc + (static_cast<long long>(ok ? rb : 0) * p.L + (ok ? t : 0)) * TC_C + half * 128;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (ok) {
                            uint32_t lo[8], hi[8];
                            if constexpr (ML) {       // written by other CTAs earlier in this launch: read at L2, never from a stale L1 line
                                ldg256_cg(hrow + cc * 32, lo);
                                ldg256_cg(hrow + cc * 32 + 16, hi);
                            } else {
                                ldg256(hrow + cc * 32, lo);
                                ldg256(hrow + cc * 32 + 16, hi);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) { hres[cc][i] = lo[i]; hres[cc][8 + i] = hi[i]; }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i) hres[cc][i] = 0u;
                        }
                    }
                }
                TC_DBG_T0(tw3);
                mbar_wait(&bar_tfull[buf], par, SITE_EPI_TFULL, 2);
                TC_DBG_ACC(9, tw3);
                TC_DBG_T0(tk3);
                tc_fence_after_sync();
                if (z_pending) {              // this G2r has finished reading z: the next group's K-blocks 0,1 can land
                    publish_z(0, pend_t0, pend_b, pend_valid, pend_zrow);
                    z_pending = false;
                }
                const float* b2g = (ML ? p.b2_tab[layer] : p.b2) + half * 128;
                const bool store_h = !ML || layer + 1 < p.layers;        // the last block's residual output is never used
                const CUtensorMap* tmo = (ML && (layer & 1)) ? &tm_hout2 : &tm_hout;
                const int sw_w = (lane >> 1) & 3;                 // write swizzle of this lane's own row
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int col = half * 128 + cc * 32;
                    uint32_t r[32];
                    tmem_ld_32x32(t_lane + buf * 256 + col, r);
                    tmem_ld_wait();
                    uint32_t pk[16];
                    const float4* b4 = reinterpret_cast<const float4*>(b2g + cc * 32);      // 128-byte aligned: 8 vector loads per chunk
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 bb = __ldg(b4 + (i >> 2));
                        const uint32_t h0 = hres[cc][i >> 1], h1 = hres[cc][(i >> 1) + 1];
                        const float v0 = (__uint_as_float(r[i]) + bf16_lo(h0) + bb.x) * 0.70710678118654752f;
                        const float v1 = (__uint_as_float(r[i + 1]) + bf16_hi(h0) + bb.y) * 0.70710678118654752f;
                        const float v2 = (__uint_as_float(r[i + 2]) + bf16_lo(h1) + bb.z) * 0.70710678118654752f;
                        const float v3 = (__uint_as_float(r[i + 3]) + bf16_hi(h1) + bb.w) * 0.70710678118654752f;
                        pk[i >> 1] = pack_bf16x2(v0, v1);
                        pk[(i >> 1) + 1] = pack_bf16x2(v2, v3);
                    }
                    uint8_t* box = tbuf + (cc & 1) * 2048;
                    if (lane == 0) tma_store_wait_read<1>();       // the store that last used this box (two commits ago) has read it
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        *reinterpret_cast<uint4*>(box + lane * 64 + ((m ^ sw_w) << 4)) = make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && tile_valid && store_h && !(kdbg & 16)) {
                        const int st0 = (kdbg & 1024) ? (static_cast<int>(blockIdx.x) % p.tiles_per_b) * TC_TILE_T : t0;
                        const int sb = (kdbg & 1024) ? static_cast<int>(blockIdx.x) / p.tiles_per_b : (kdbg & 2048) ? b % (kdbg >> 16) : b;
                        tma_store_3d(tmo, box, col, st0 + q * 32, sb);
                        tma_store_commit();
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&bar_tempty[buf], 0);
                if (ML && tile_valid && store_h) { flag_tile = tile; flag_layer = layer; }
                TC_DBG_ACC(10, tk3);
            }
        }
        TC_DBG_ACC(11, te_all);
        if ((kdbg & 2) && ew == 0 && lane == 0 && leader)
            for (int i = 7; i < 12; ++i) atomicAdd(&g_tc_cycles[i], dbg_acc[i]);
        raise_flag();
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores fully performed before exit
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == w_mma) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// Skip GEMM: skip[b][t][n] (=|+=) bias[n] + sum_{l < G} sum_k z_l[b][t][k] W2s_{layer0 + l}[n][k]
// (wavenet.py:114 skip half summed over blocks, :145-149). Persistent CTA pairs, one M = 256 x N = 256 fp16 MMA stream
// per tile group with K = 256 G, two TMEM accumulators so a group's epilogue overlaps the next group's MMAs, 5 x 32 KB
// stages. The epilogue leaves fp32 [32 t][32 ch] boxes through TMA (plain store for the first layer group, reduce-add
// for later ones).
// ------------------------------------------------------------------------------------------------
constexpr int SG_STAGES = 5;
struct SkipGemmSmem {
    static constexpr int stages = 0;
    static constexpr int stg = SG_STAGES * T3_STAGE_BYTES;          // 8 warps x 2 x 4 KB fp32 transposition boxes
    static constexpr int bias = stg + 8 * 8192;                     // 256 fp32
    static constexpr int bars = bias + 1024;
    static constexpr int tmem_ptr = bars + 16 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(SkipGemmSmem::stg % 1024 == 0, "swizzle alignment");
static_assert(SkipGemmSmem::total <= 232448, "shared memory budget");
constexpr int SKIP_GEMM_SMEM_BYTES = SkipGemmSmem::total;

struct SkipGemmParams {
    const float* bias;              // [256] sum over ALL layers of the skip half of b2, or nullptr (later layer groups)
    int B, L, tiles_per_b, num_tiles;
    int G;                          // layers in this group (stash slots 0 .. G-1)
    int layer0;                     // first layer of the group (weight blocks)
    int accumulate;                 // 0: skip = value   1: skip += value (TMA reduce-add)
    int w_bf16;                     // 1 (training): the stash and the weights are bf16; tm_w is the [layers][4][256][64] bf16 copy of the
                                    // skip half of W2 instead of the fp16 blocks 28..31 of every layer
};

__global__ void __launch_bounds__(TC_THREADS, 1)
wavenet_skip_gemm_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_w,
                         const __grid_constant__ CUtensorMap tm_skip, const SkipGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_bias = reinterpret_cast<float*>(smem + SkipGemmSmem::bias);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SkipGemmSmem::bars);
    uint64_t* bar_full = bars;                   // [SG_STAGES] rank 0
    uint64_t* bar_empty = bars + SG_STAGES;      // [SG_STAGES] per CTA
    uint64_t* bar_tfull = bars + 2 * SG_STAGES;  // [2] per CTA
    uint64_t* bar_tempty = bar_tfull + 2;        // [2] rank 0
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + SkipGemmSmem::tmem_ptr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_z);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_skip);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < SG_STAGES; ++s) { mbar_init(&bar_full[s], 2); mbar_init(&bar_empty[s], 1); }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_tfull[i], 1);
                mbar_init(&bar_tempty[i], 2 * (TC_EPI_THREADS / 32));
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc_pair(s_tmem, 512);
        tmem_relinquish_pair();
    }
    if (warp >= 2)
        for (int i = threadIdx.x - 64; i < 256; i += TC_EPI_THREADS) s_bias[i] = p.bias ? p.bias[i] : 0.f;
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t IDESC_F16 = p.w_bf16 ? umma_idesc_pair_bf16(256) : umma_idesc_pair_f16(256);     // operand format of stash and weights

    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_groups = (p.num_tiles + 1) >> 1;
    const int nkb = 4 * p.G;

    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int grp = pair_id; grp < num_groups; grp += num_pairs) {
            const int tile = grp * 2 + rank;
            const bool tile_valid = tile < p.num_tiles;
            const int b = tile_valid ? tile / p.tiles_per_b : 0;     // a padding tile re-reads sample 0 (never stored)
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&bar_empty[stage], phase ^ 1, SITE_PROD_EMPTY, stage);
                if (lane == 0) {
                    uint8_t* sa = smem + SkipGemmSmem::stages + stage * T3_STAGE_BYTES;
                    uint8_t* sb = sa + T3_A_BYTES;
                    const int l = kb >> 2, cib = kb & 3;
                    const int wblk = p.w_bf16 ? (p.layer0 + l) * 4 + cib : (p.layer0 + l) * TC_W_BLOCKS_PER_LAYER + 28 + cib;
                    if (leader) mbar_arrive_expect_tx(&bar_full[stage], 2 * T3_STAGE_BYTES);
                    else        mbar_arrive_cluster(&bar_full[stage], 0);
                    tma_load_3d_pair(sa, &tm_z, &bar_full[stage], cib * 64, t0, l * p.B + b);
                    tma_load_2d_pair(sb, &tm_w, &bar_full[stage], 0, wblk * 256 + rank * 128);
                }
                __syncwarp();
                if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            uint32_t stage = 0, phase = 0, it = 0;
            for (int grp = pair_id; grp < num_groups; grp += num_pairs, ++it) {
                const uint32_t buf = it & 1;
                mbar_wait(&bar_tempty[buf], ((it >> 1) & 1) ^ 1, SITE_MMA_TEMPTY, 0);
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar_full[stage], phase, SITE_MMA_FULL, stage);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + SkipGemmSmem::stages + stage * T3_STAGE_BYTES);
                        const uint32_t sb = sa + T3_A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(d_tmem, umma_desc_sw128_kmajor(sa + k * 32), umma_desc_sw128_kmajor(sb + k * 32),
                                              IDESC_F16, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair_mc(&bar_empty[stage], 3);
                        if (kb == nkb - 1) umma_commit_pair_mc(&bar_tfull[buf], 3);
                    }
                    __syncwarp();
                    if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* stg = smem + SkipGemmSmem::stg + ew * 8192;
        uint32_t it = 0;
        for (int grp = pair_id; grp < num_groups; grp += num_pairs, ++it) {
            const int tile = grp * 2 + rank;
            const bool tile_valid = tile < p.num_tiles;
            const int b = tile / p.tiles_per_b;
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            const int trow = t0 + q * 32;
            const uint32_t buf = it & 1;
            mbar_wait(&bar_tfull[buf], (it >> 1) & 1, SITE_EPI_TFULL, 0);
            tc_fence_after_sync();
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col = half * 128 + cc * 32;
                uint32_t r[32];
                tmem_ld_32x32(t_lane + buf * 256 + col, r);
                tmem_ld_wait();
                uint8_t* cbuf = stg + (cc & 1) * 4096;
                if (lane == 0) tma_store_wait_read<1>();      // box (cc & 1) was handed to TMA two commits ago
                __syncwarp();
                uint8_t* brow = cbuf + lane * 128;
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    float4 v;
                    v.x = __uint_as_float(r[4 * m + 0]) + s_bias[col + 4 * m + 0];
                    v.y = __uint_as_float(r[4 * m + 1]) + s_bias[col + 4 * m + 1];
                    v.z = __uint_as_float(r[4 * m + 2]) + s_bias[col + 4 * m + 2];
                    v.w = __uint_as_float(r[4 * m + 3]) + s_bias[col + 4 * m + 3];
                    const int chunk = m ^ (lane & 7);
                    *reinterpret_cast<float4*>(brow + chunk * 16) = v;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0 && tile_valid) {
                    if (p.accumulate) tma_reduce_add_3d(&tm_skip, cbuf, col, trow, b);
                    else              tma_store_3d(&tm_skip, cbuf, col, trow, b);
                    tma_store_commit();
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bar_tempty[buf], 0);
        }
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

// ------------------------------------------------------------------------------------------------
// Skip GEMM with the network's tail fused into its epilogue (all `layers` blocks in one K loop):
//   skip = bias + sum_l W2s_l z_l                                         (wavenet.py:114, :145-149)
//   F[b][t] = b_out + sum_c w_out[c] relu(b_sp[c] + sum_k Wsp[c][k] skip[k] / sqrt(layers))   (wavenet.py:151, :177-179)
// The fp32 skip sum never reaches HBM (the unfused pair of kernels writes and re-reads 4.2 GB per 256-sample evaluation).
// Per tile group: job S (K = 256 * layers, accumulator 0) -> epilogue a: (acc + bias) * scale -> bf16 A operand in shared
// memory -> job T (K = 256 against the skip-projection weights streamed through the same stage ring, accumulator 1) -> epilogue
// b: relu, dot with w_out over the 256 columns (two warps per row quarter, halves combined through shared memory).
// Job S of the next group only needs accumulator 0, which epilogue a has drained before job T starts.
// ------------------------------------------------------------------------------------------------
struct SkipTailSmem {
    static constexpr int stages = 0;
    static constexpr int a2 = SG_STAGES * T3_STAGE_BYTES;           // 4 K-blocks [128 t][64 ch] bf16, 128-byte swizzle (64 KB)
    static constexpr int part = a2 + TC_Z_BYTES;                    // 128 fp32 partial row sums
    static constexpr int bars = part + 512;
    static constexpr int tmem_ptr = bars + 16 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(SkipTailSmem::a2 % 1024 == 0, "swizzle alignment");
static_assert(SkipTailSmem::total <= 232448, "shared memory budget");
constexpr int SKIP_TAIL_SMEM_BYTES = SkipTailSmem::total;

struct SkipTailParams {
    const float* bias;              // [256] sum over all layers of the skip half of b2
    const float* b_sp;              // [256]
    const float* w_out;             // [256]
    const float* b_out;             // [1]
    float* out;                     // [B][L]
    float scale;                    // sqrt(1 / layers)
    int B, L, tiles_per_b, num_tiles;
    int G;                          // = layers
};

__global__ void __launch_bounds__(TC_THREADS, 1)
wavenet_skip_tail_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_w,
                         const __grid_constant__ CUtensorMap tm_wsp, const SkipTailParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* s_part = reinterpret_cast<float*>(smem + SkipTailSmem::part);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SkipTailSmem::bars);
    uint64_t* bar_full = bars;                   // [SG_STAGES] rank 0
    uint64_t* bar_empty = bars + SG_STAGES;      // [SG_STAGES] per CTA
    uint64_t* bar_tfull = bars + 2 * SG_STAGES;  // [2] per CTA: accumulator 0 (job S) / 1 (job T) ready
    uint64_t* bar_tempty = bar_tfull + 2;        // [2] rank 0
    uint64_t* bar_a2 = bar_tempty + 2;           // [1] rank 0: the A operand of job T is written
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + SkipTailSmem::tmem_ptr);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_z);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_wsp);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < SG_STAGES; ++s) { mbar_init(&bar_full[s], 2); mbar_init(&bar_empty[s], 1); }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bar_tfull[i], 1);
                mbar_init(&bar_tempty[i], 2 * (TC_EPI_THREADS / 32));
            }
            mbar_init(bar_a2, 2 * (TC_EPI_THREADS / 32));
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
    const uint32_t tmem_base = *s_tmem;
    constexpr uint32_t IDESC_F16 = umma_idesc_pair_f16(256);
    constexpr uint32_t IDESC_BF16 = umma_idesc_pair_bf16(256);

    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_groups = (p.num_tiles + 1) >> 1;
    const int nkb = 4 * p.G;

    if (warp == 0) {
        uint32_t stage = 0, phase = 0;
        for (int grp = pair_id; grp < num_groups; grp += num_pairs) {
            const int tile = grp * 2 + rank;
            const bool tile_valid = tile < p.num_tiles;
            const int b = tile_valid ? tile / p.tiles_per_b : 0;     // a padding tile re-reads sample 0 (never stored)
            const int t0 = (tile % p.tiles_per_b) * TC_TILE_T;
            for (int kb = 0; kb < nkb + 4; ++kb) {
                mbar_wait(&bar_empty[stage], phase ^ 1, SITE_PROD_EMPTY, stage);
                if (lane == 0) {
                    uint8_t* sa = smem + SkipTailSmem::stages + stage * T3_STAGE_BYTES;
                    uint8_t* sb = sa + T3_A_BYTES;
                    if (kb < nkb) {
                        const int l = kb >> 2, cib = kb & 3;
                        const int wblk = l * TC_W_BLOCKS_PER_LAYER + 28 + cib;
                        if (leader) mbar_arrive_expect_tx(&bar_full[stage], 2 * T3_STAGE_BYTES);
                        else        mbar_arrive_cluster(&bar_full[stage], 0);
                        tma_load_3d_pair(sa, &tm_z, &bar_full[stage], cib * 64, t0, l * p.B + b);
                        tma_load_2d_pair(sb, &tm_w, &bar_full[stage], 0, wblk * 256 + rank * 128);
                    } else {
                        // job T: this CTA's half of K-block kb - nkb of the skip-projection weights
                        if (leader) mbar_arrive_expect_tx(&bar_full[stage], 2 * T3_B_BYTES);
                        else        mbar_arrive_cluster(&bar_full[stage], 0);
                        tma_load_2d_pair(sb, &tm_wsp, &bar_full[stage], 0, (kb - nkb) * 256 + rank * 128);
                    }
                }
                __syncwarp();
                if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            uint32_t stage = 0, phase = 0, it = 0;
            const uint32_t a2_addr = smem_u32(smem + SkipTailSmem::a2);
            for (int grp = pair_id; grp < num_groups; grp += num_pairs, ++it) {
                // ---- job S -> accumulator 0 ----
                mbar_wait(&bar_tempty[0], (it & 1) ^ 1, SITE_MMA_TEMPTY, 0);
                tc_fence_after_sync();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar_full[stage], phase, SITE_MMA_FULL, stage);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + SkipTailSmem::stages + stage * T3_STAGE_BYTES);
                        const uint32_t sb = sa + T3_A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(tmem_base, umma_desc_sw128_kmajor(sa + k * 32), umma_desc_sw128_kmajor(sb + k * 32),
                                              IDESC_F16, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair_mc(&bar_empty[stage], 3);
                        if (kb == nkb - 1) umma_commit_pair_mc(&bar_tfull[0], 3);
                    }
                    __syncwarp();
                    if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
                }
                // ---- job T -> accumulator 1 (A = the scaled skip sum written by epilogue a) ----
                mbar_wait(&bar_tempty[1], (it & 1) ^ 1, SITE_MMA_TEMPTY, 1);
                mbar_wait(bar_a2, it & 1, SITE_MMA_ZREADY, 0);
                tc_fence_after_sync();
                for (int kb = 0; kb < 4; ++kb) {
                    mbar_wait(&bar_full[stage], phase, SITE_MMA_FULL, stage);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t sb = smem_u32(smem + SkipTailSmem::stages + stage * T3_STAGE_BYTES) + T3_A_BYTES;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(tmem_base + 256, umma_desc_sw128_kmajor(a2_addr + kb * TC_A_BYTES + k * 32),
                                              umma_desc_sw128_kmajor(sb + k * 32), IDESC_BF16, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair_mc(&bar_empty[stage], 3);
                        if (kb == 3) umma_commit_pair_mc(&bar_tfull[1], 3);
                    }
                    __syncwarp();
                    if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* a2 = smem + SkipTailSmem::a2;
        const float b_out = __ldg(p.b_out);
        uint32_t it = 0;
        for (int grp = pair_id; grp < num_groups; grp += num_pairs, ++it) {
            const int tile = grp * 2 + rank;
            const bool tile_valid = tile < p.num_tiles;
            const int b = tile / p.tiles_per_b;
            const int t = (tile % p.tiles_per_b) * TC_TILE_T + row;
            // ---- epilogue a: (skip + bias) / sqrt(layers) -> bf16 A operand of the skip projection ----
            mbar_wait(&bar_tfull[0], it & 1, SITE_EPI_TFULL, 0);
            tc_fence_after_sync();
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col = half * 128 + cc * 32;
                uint32_t r[32];
                tmem_ld_32x32(t_lane + col, r);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2)
                    pk[i >> 1] = pack_bf16x2((__uint_as_float(r[i]) + __ldg(p.bias + col + i)) * p.scale,
                                             (__uint_as_float(r[i + 1]) + __ldg(p.bias + col + i + 1)) * p.scale);
                uint8_t* arow = a2 + (col >> 6) * TC_A_BYTES + row * 128;
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int chunk = (((col & 63) >> 3) + m) ^ (row & 7);
                    *reinterpret_cast<uint4*>(arow + chunk * 16) = make_uint4(pk[4 * m], pk[4 * m + 1], pk[4 * m + 2], pk[4 * m + 3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(&bar_tempty[0], 0);
                mbar_arrive_cluster(bar_a2, 0);
            }
            // ---- epilogue b: F = b_out + sum_c w_out[c] relu(s2[c] + b_sp[c]) ----
            mbar_wait(&bar_tfull[1], it & 1, SITE_EPI_TFULL, 1);
            tc_fence_after_sync();
            float acc = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
                const int col = half * 128 + cc * 32;
                uint32_t r[32];
                tmem_ld_32x32(t_lane + 256 + col, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    acc = fmaf(fmaxf(__uint_as_float(r[i]) + __ldg(p.b_sp + col + i), 0.f), __ldg(p.w_out + col + i), acc);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&bar_tempty[1], 0);
            named_bar_sync(1, TC_EPI_THREADS);                   // the previous group's partial sums have been consumed
            if (half == 1) s_part[row] = acc;
            named_bar_sync(1, TC_EPI_THREADS);
            if (half == 0 && tile_valid && t < p.L) p.out[static_cast<long long>(b) * p.L + t] = acc + s_part[row] + b_out;
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// bias[n] = sum_l b2_l[256 + n]  (the skip half of every block's output-projection bias, wavenet.py:114)
__global__ void skip_bias_sum_kernel(const float* const* __restrict__ b2_tab, int layers, float* __restrict__ out) {
    const int n = threadIdx.x;
    float s = 0.f;
    for (int l = 0; l < layers; ++l) s += b2_tab[l][256 + n];
    out[n] = s;
}

}  // namespace adb
