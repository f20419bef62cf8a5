// 3x3 / 1x1 stride-1 "same" convolution forward (and, with transposed+rotated weights, dgrad) as a
// PERSISTENT implicit GEMM on the 5th-generation tensor cores:
//
//   D[128 pixels, BN channels] (fp32, TMEM)  +=  A[128 pixels, BK] (bf16, smem)  x  W[BN, BK]^T (bf16, smem)
//
// * A is never materialised as an im2col matrix.  The activation tensor is NHWC bf16; a pixel tile is a
//   BW x BH x BNI box of (x, y, n) with BW*BH*BNI = 128, and the tile of filter tap (ky, kx) is the SAME box
//   shifted by (kx-1, ky-1): one 4-D TMA tiled load with signed coordinates, whose out-of-bounds elements
//   (the zero padding, and the ragged edge of the last tile) are zero-filled by the TMA unit.  The box
//   lands in shared memory as 128 rows x BK channels with the 128/64/32-byte swizzle tcgen05 expects
//   (K-major canonical layout), so there is no register staging at all.
// * W is the packed weight matrix [Cout_pad, k*k*Cin] (K contiguous), a 2-D TMA box [BN, BK].
// * An optional second operand pair (a2, w2) appends the K slices of a 1x1 convolution of another tensor
//   (the ResBlock shortcut, diff_cifar/model.py:145-148,:167), and the epilogue adds bias, a per-sample
//   per-channel row (the time-embedding projection, model.py:164) and a residual tensor.
// * One CTA per SM walks the (pixel tile, channel tile) list.  Warp roles: warp 0 = TMA producer,
//   warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue.  Three pipelines:
//   the smem ring of (A, W) stages (full/empty mbarriers, slots released by tcgen05.commit), TWO TMEM
//   accumulators (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1, and the
//   epilogue's own double-buffered staging tile: tcgen05.ld -> +bias/row/residual -> bf16 -> swizzled smem
//   -> TMA tiled store (which also clips ragged tiles), or direct fp32 NCHW stores for the narrow tails.
//
// Replaces nn.Conv2d forward/backward-data at diff_cifar/model.py:69,:133,:143,:146,:396;
// diff_mnist/torch_ddpm/ddpm/models/unet/layers.py:286,:300,:305-312; pdearena twod_unetbase.py:19-24.
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace {
using namespace ub;
using namespace ub::tc;

struct FpropParams {
    int N, H, W, Cout;
    int BW, BH, BNI, tiles_w, tiles_h;
    int BN, n_tiles, num_tiles, m_tiles;
    int msub;                         // pixel tiles per CTA tile (2 when Cout <= 128: both share one weight tile)
    int Cin, cblocks, taps, kb_main, kb_extra, stages;
    uint32_t a_stage_bytes, b_stage_bytes, tmem_cols, tbl_bytes;
    int cluster_tiles;                // (channel tile, pixel-tile group) units walked by one cluster
    int cluster;                      // CTAs per cluster: 1, or 2 with weight multicast / as a cta_group::2 pair
    int pair;                         // 1: the two CTAs of a cluster run ONE tcgen05.mma.cta_group::2 per k-step (M = 256: each CTA
                                      // its own 128-pixel tile, HALF of the weight tile staged per CTA)
    int S;                            // convolution stride (1 or 2): input pixel = S * output pixel + tap offset
    const float *bias, *bias2, *rowadd;
    const __nv_bfloat16 *residual; int64_t ld_res;
    int has_out;                      // bf16 NHWC output through the TMA store
    float *out_nchw;
};

constexpr int kThreads = 320;                         // TMA warp, MMA warp, two epilogue groups of four warps
constexpr int kEpiThreads = 128;                      // one epilogue group: one warp per TMEM lane quadrant
constexpr uint32_t kStagingBytes = 128 * 64 * 2;      // one [128 pixels][64 channels] bf16 box
constexpr int kStagingBufs = 4;                       // two per epilogue group
constexpr int kTblRows = 8;                           // addend table: up to 8 samples per pixel tile
// Optional CTA pair (cluster of 2): the weight tile is fetched once from L2 and multicast to both CTAs.  Measured
// neutral-to-negative on B200 for these shapes: the kernel is bound by shared-memory bandwidth (tensor-core operand
// reads + TMA writes), which multicast does not reduce; kept selectable (UB200_FPROP_CLUSTER=2) for the 2-CTA MMA work.

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// FAST: the common epilogue (bf16 NHWC output through TMA, Cout and BN multiples of 64) with every per-launch option
// resolved at compile time (RES: residual tensor, TBL: bias / per-sample row through the smem table) and both
// epilogue groups working on alternate 64-channel chunks.  The generic path keeps all runtime options and uses
// one group.  The epilogue is what bounds the 1x1 and short-K convolutions, so its instruction count matters.
// PAIR is a template parameter, not a run-time flag: a kernel that contains cta_group::2 instructions can only be launched
// with an even cluster size (cudaErrorInvalidClusterSize otherwise, measured), so the single-CTA kernels must not contain any.
template <int BK, bool FAST, bool RES, bool TBL, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1) conv_fprop_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                const __grid_constant__ CUtensorMap tm_w,
                                                                const __grid_constant__ CUtensorMap tm_a2,
                                                                const __grid_constant__ CUtensorMap tm_w2,
                                                                const __grid_constant__ CUtensorMap tm_out,
                                                                const FpropParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [2 x output staging][stages x A][stages x B][full][empty][tmem_full x2][tmem_empty x2][tmem ptr]
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_out = smem;
    float *tbl = reinterpret_cast<float *>(smem_out + kStagingBufs * kStagingBytes);   // bias + per-sample rows
    uint8_t *smem_a = smem_out + kStagingBufs * kStagingBytes + p.tbl_bytes;
    uint8_t *smem_b = smem_a + (size_t)p.stages * p.a_stage_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_b + (size_t)p.stages * p.b_stage_bytes);
    uint64_t *empty = full + p.stages;
    uint64_t *tmem_full = empty + p.stages;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

    // warp index through a shuffle: the compiler then knows it is warp-uniform, so the role branches are uniform
    // control flow and the producer / MMA loops below keep their state in uniform registers
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int num_kb = p.kb_main + p.kb_extra;
    pdl_trigger();                     // the next kernel in the stream may set itself up behind this one

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_w);
        if (p.kb_extra) { prefetch_tmap(&tm_a2); prefetch_tmap(&tm_w2); }
        if (p.has_out) prefetch_tmap(&tm_out);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, PAIR ? 1 : p.cluster); }
        // pair: the leader's MMA may overwrite an accumulator only when the epilogue warps of BOTH CTAs have drained it
        for (int a = 0; a < 2; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, PAIR ? 16 : 8); }
        fence_barrier_init();
    }
    if (warp == 1) { if constexpr (PAIR) tmem_alloc_pair(tmem_slot, p.tmem_cols); else tmem_alloc(tmem_slot, p.tmem_cols); }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();   // the peer multicasts into our smem and arrives on our barriers
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();                        // everything above overlapped the previous kernel's tail; its data is needed now
    const int kCluster = p.cluster;
    const int crank = kCluster > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / kCluster, num_clusters = gridDim.x / kCluster;
    // this CTA's m unit of cluster tile ct:  (ct / n_tiles) * kCluster + crank   (may lie past the end: zero tile)

    if (warp == 0) {
        // ===================== TMA producer.  The whole warp walks the loop (uniform trip counts, so every address and
        // coordinate is warp-uniform and lives in uniform registers: TMA / mbarrier operands need no per-lane
        // "waterfall"); only the instructions with side effects are issued by lane 0.
        {
            const bool leader = lane == 0;
            constexpr bool pair = PAIR;
            // pair: the leader's barrier counts the bytes of both CTAs (each: its A tiles + half of the weight tile)
            const uint32_t tx_bytes = pair ? 2u * (uint32_t)p.msub * 128u * BK * 2u + (uint32_t)p.BN * BK * 2u
                                           : (uint32_t)p.msub * 128u * BK * 2u + (uint32_t)p.BN * BK * 2u;
            const int half = p.BN / kCluster;                 // weight rows this CTA fetches for the pair
            const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b) + (pair ? 0u : (uint32_t)(crank * half * (BK * 2)));
            const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
            uint32_t full_lead = full0;                                       // cluster address of the leader's `full` barriers
            if constexpr (PAIR) full_lead = mapa_rank(full0, 0);
            const int cow = crank * half;
            int s = 0; uint32_t ph = 0;                       // ring position, carried across tiles
            for (int ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters) {
                const int nt = ct % p.n_tiles, co0 = nt * p.BN + cow;
                int xa, ya, na, xb = 0, yb = 0, nb = 0;
                {
                    int mt = ((ct / p.n_tiles) * kCluster + crank) * p.msub;
                    xa = (mt % p.tiles_w) * p.BW; mt /= p.tiles_w;
                    ya = (mt % p.tiles_h) * p.BH; mt /= p.tiles_h;
                    na = mt * p.BNI;                          // past the batch for a padded trailing tile: TMA zero-fills
                    if (p.msub == 2) {
                        mt = ((ct / p.n_tiles) * kCluster + crank) * p.msub + 1;
                        xb = (mt % p.tiles_w) * p.BW; mt /= p.tiles_w;
                        yb = (mt % p.tiles_h) * p.BH; mt /= p.tiles_h;
                        nb = mt * p.BNI;
                    }
                }
                const int kdim = p.taps == 9 ? 3 : 1;
                int wk = 0;                                   // K coordinate into the packed weights: tap * Cin + cb * BK
                for (int ky = 0; ky < kdim; ++ky)
                    for (int kx = 0; kx < kdim; ++kx) {
                        const int dy = kdim == 3 ? ky - 1 : 0, dx = kdim == 3 ? kx - 1 : 0;
                        for (int c = 0; c < p.Cin; c += BK, wk += BK) {
                            const uint32_t fb = full0 + 8u * s, sa = a0 + (uint32_t)s * p.a_stage_bytes;
                            mbar_wait_a(empty0 + 8u * s, ph ^ 1u);
                            if constexpr (PAIR) {
                                if (leader) {
                                    const uint32_t fl = full_lead + 8u * s;
                                    if (crank == 0) mbar_arrive_expect_tx_a(fb, tx_bytes);
                                    tma_load_4d_pair_a(sa, &tm_a, fl, c, p.S * xa + dx, p.S * ya + dy, na);
                                    if (p.msub == 2) tma_load_4d_pair_a(sa + 128 * BK * 2, &tm_a, fl, c, p.S * xb + dx, p.S * yb + dy, nb);
                                    tma_load_2d_pair_a(b0 + (uint32_t)s * p.b_stage_bytes, &tm_w, fl, wk, co0);
                                }
                            } else if (leader) {
                                mbar_arrive_expect_tx_a(fb, tx_bytes);
                                tma_load_4d_a(sa, &tm_a, fb, c, p.S * xa + dx, p.S * ya + dy, na);
                                if (p.msub == 2) tma_load_4d_a(sa + 128 * BK * 2, &tm_a, fb, c, p.S * xb + dx, p.S * yb + dy, nb);
                                if (kCluster > 1) tma_load_2d_mcast_a(b0 + (uint32_t)s * p.b_stage_bytes, &tm_w, fb, wk, co0, (uint16_t)0x3);
                                else tma_load_2d_a(b0 + (uint32_t)s * p.b_stage_bytes, &tm_w, fb, wk, co0);
                            }
                            if (++s == p.stages) { s = 0; ph ^= 1u; }
                        }
                    }
                for (int c = 0; c < p.kb_extra * BK; c += BK) {   // K slices of the fused 1x1 term
                    const uint32_t fb = full0 + 8u * s, sa = a0 + (uint32_t)s * p.a_stage_bytes;
                    mbar_wait_a(empty0 + 8u * s, ph ^ 1u);
                    if constexpr (PAIR) {
                        if (leader) {
                            const uint32_t fl = full_lead + 8u * s;
                            if (crank == 0) mbar_arrive_expect_tx_a(fb, tx_bytes);
                            tma_load_4d_pair_a(sa, &tm_a2, fl, c, xa, ya, na);
                            if (p.msub == 2) tma_load_4d_pair_a(sa + 128 * BK * 2, &tm_a2, fl, c, xb, yb, nb);
                            tma_load_2d_pair_a(b0 + (uint32_t)s * p.b_stage_bytes, &tm_w2, fl, c, co0);
                        }
                    } else if (leader) {
                        mbar_arrive_expect_tx_a(fb, tx_bytes);
                        tma_load_4d_a(sa, &tm_a2, fb, c, xa, ya, na);
                        if (p.msub == 2) tma_load_4d_a(sa + 128 * BK * 2, &tm_a2, fb, c, xb, yb, nb);
                        if (kCluster > 1) tma_load_2d_mcast_a(b0 + (uint32_t)s * p.b_stage_bytes, &tm_w2, fb, c, co0, (uint16_t)0x3);
                        else tma_load_2d_a(b0 + (uint32_t)s * p.b_stage_bytes, &tm_w2, fb, c, co0);
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: warp-uniform loop, tcgen05.mma / commit issued by lane 0 =====================
        // (pair: only the leader CTA issues; its instruction drives the tensor cores of both SMs)
        if (!(PAIR && crank != 0)) {
            const bool leader = lane == 0;
            constexpr bool pair = PAIR;
            const uint32_t idesc = make_idesc(pair ? 256 : 128, p.BN, 0, 0);
            constexpr uint32_t swz = swizzle_code(BK * 2);
            constexpr uint32_t sbo = 8u * BK * 2u;
            // descriptors differ between stages / k-steps only in the 14-bit start-address field (units of 16 bytes)
            const uint64_t da0 = make_smem_desc(smem_u32(smem_a), 16, sbo, swz);
            const uint64_t db0 = make_smem_desc(smem_u32(smem_b), 16, sbo, swz);
            const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
            const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty);
            int s = 0; uint32_t ph = 0, local = 0;
            for (int ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, ++local) {
                const uint32_t acc = local & 1u, use = local >> 1;
                mbar_wait(tmem_empty + acc, (use & 1u) ^ 1u);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)(p.msub * p.BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait_a(full0 + 8u * s, ph);
                    tc_fence_after();
                    const uint64_t da = da0 + (uint64_t)(s * a_step), db = db0 + (uint64_t)(s * b_step);
                    if constexpr (PAIR) {
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                            if (p.msub == 2) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_bf16_pair(d_tmem + (uint32_t)p.BN, da + (128 * BK * 2 / 16) + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                            }
                            umma_commit_pair_a(empty0 + 8u * s, (uint16_t)0x3);                  // frees the slot in both CTAs
                        }
                    } else if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        if (p.msub == 2) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_bf16(d_tmem + (uint32_t)p.BN, da + (128 * BK * 2 / 16) + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        }
                        if (kCluster > 1) umma_commit_mcast_a(empty0 + 8u * s, (uint16_t)0x3);   // both CTAs' MMAs must have read it
                        else umma_commit_a(empty0 + 8u * s);                                     // slot reusable once read
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                if (leader) {                                          // accumulator complete (pair: in both CTAs)
                    if constexpr (PAIR) umma_commit_pair_a(smem_u32(tmem_full + acc), (uint16_t)0x3);
                    else umma_commit(tmem_full + acc);
                }
            }
        }
    } else {
        // ===================== epilogue: two groups of 4 warps, one TMEM lane quadrant per warp =====================
        const int grp = (warp - 2) >> 2;              // 0: warps 2..5, 1: warps 6..9
        // "this accumulator is drained": on the own barrier, or (pair, follower CTA) on the leader's, whose MMA thread waits
        auto arrive_tmem_empty = [&](uint32_t a) {
            if constexpr (PAIR) {
                if (crank != 0) { mbar_arrive_remote(mapa_rank(smem_u32(tmem_empty + a), 0)); return; }
            }
            mbar_arrive(tmem_empty + a);
        };
        if constexpr (FAST) {
            const int qd = warp & 3;
            const int r = qd * 32 + lane;             // row of the tile == TMEM lane
            const int et = threadIdx.x - 64;          // 0..255 within the epilogue warps
            const int wi = r % p.BW, hi = (r / p.BW) % p.BH, ni = r / (p.BW * p.BH);
            const bool issuer = ((warp == 2 || warp == 6) && lane == 0);
            const int bar_id = 1 + grp;
            const int nchunks = p.BN >> 6, total = p.msub * nchunks;
            const uint32_t stage0 = smem_u32(smem_out) + (uint32_t)grp * 2u * kStagingBytes;
            const uint32_t tbl0 = smem_u32(tbl);
            const int tbl_elems = p.msub * p.BNI * p.BN;                     // per tile parity
            const uint32_t row_off = (uint32_t)r * 128u, sw = (uint32_t)(r & 7);
            uint32_t local = 0, lc = 0;
            for (int ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, ++local) {
                const int nt = ct % p.n_tiles, co0 = nt * p.BN;
                const uint32_t acc = local & 1u, use = local >> 1;
                int xa, ya, na, xb = 0, yb = 0, nb = 0;
                {
                    int mt = ((ct / p.n_tiles) * kCluster + crank) * p.msub;
                    xa = (mt % p.tiles_w) * p.BW; mt /= p.tiles_w;
                    ya = (mt % p.tiles_h) * p.BH; mt /= p.tiles_h;
                    na = mt * p.BNI;
                    if (p.msub == 2) {
                        mt = ((ct / p.n_tiles) * kCluster + crank) * p.msub + 1;
                        xb = (mt % p.tiles_w) * p.BW; mt /= p.tiles_w;
                        yb = (mt % p.tiles_h) * p.BH; mt /= p.tiles_h;
                        nb = mt * p.BNI;
                    }
                }
                if (TBL) {
                    // tbl[parity][sub][s][c] = bias[co0+c] + rowadd[n0(sub)+s][co0+c].  The all-epilogue barrier below
                    // also proves every reader of this parity's previous contents (tile local-2) has finished.
                    float *tb = tbl + (local & 1u) * tbl_elems;
                    for (int i = et; i < tbl_elems; i += 2 * kEpiThreads) {
                        const int sub = i / (p.BNI * p.BN), rem = i - sub * (p.BNI * p.BN);
                        const int sidx = rem / p.BN, c = rem - sidx * p.BN, co = co0 + c;
                        const int ns = (sub ? nb : na) + sidx;
                        float v = 0.f;
                        if (co < p.Cout) {
                            if (p.bias) v = __ldg(p.bias + co);
                            if (p.bias2) v += __ldg(p.bias2 + co);
                            if (p.rowadd && ns < p.N) v += __ldg(p.rowadd + (int64_t)ns * p.Cout + co);
                        }
                        tb[i] = v;
                    }
                    named_barrier_sync(3, 2 * kEpiThreads);
                }
                mbar_wait(tmem_full + acc, use & 1u);
                tc_fence_after();
                bool arrived = false;
                for (int cc = grp; cc < total; cc += 2, ++lc) {
                    const int sub = cc >= nchunks ? 1 : 0, ch = cc - sub * nchunks;
                    const int x0 = sub ? xb : xa, y0 = sub ? yb : ya, n0 = sub ? nb : na;
                    const uint32_t stage = stage0 + (lc & 1u) * kStagingBytes;
                    if (issuer) tma_store_wait_read<1>();           // the store that last used this buffer has read it
                    named_barrier_sync(bar_id, kEpiThreads);
                    uint4 rq[8];
                    bool valid = false;
                    if (RES) {
                        const int x = x0 + wi, y = y0 + hi, n = n0 + ni;
                        valid = x < p.W && y < p.H && n < p.N;
                        if (valid) {
                            const uint4 *rp = reinterpret_cast<const uint4 *>(
                                p.residual + (((int64_t)n * p.H + y) * p.W + x) * p.ld_res + co0 + ch * 64);
#pragma unroll
                            for (int j = 0; j < 8; ++j) rq[j] = __ldg(rp + j);
                        }
                    }
                    float v[64];
                    tmem_ld64(tmem_base + ((uint32_t)(qd * 32) << 16) + acc * (uint32_t)(p.msub * p.BN) + (uint32_t)(sub * p.BN + ch * 64), v);
                    if (cc + 2 >= total) {                           // this group's last tcgen05.ld of the accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_tmem_empty(acc);
                        arrived = true;
                    }
                    if (TBL) {
                        const uint32_t ta = tbl0 + 4u * (uint32_t)((local & 1u) * tbl_elems + (sub * p.BNI + ni) * p.BN + ch * 64);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float4 t = lds_f4(ta + 16u * i);
                            v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                        }
                    }
                    if (RES && valid) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float f[8];
                            unpack8(rq[j], f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[8 * j + i] += f[i];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                    // 16-byte pieces of row r, 128-byte swizzle
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) f[i] = v[8 * j + i];
                        sts_u4(stage + row_off + (((uint32_t)j ^ sw) << 4), pack8(f));
                    }
                    fence_proxy_async();
                    named_barrier_sync(bar_id, kEpiThreads);
                    if (issuer) {
                        tma_store_4d(&tm_out, smem_out + (stage - smem_u32(smem_out)), co0 + ch * 64, x0, y0, n0);
                        tma_store_commit();
                    }
                }
                if (!arrived) {                                      // no chunk of this tile was ours: stay in lock step
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_tmem_empty(acc);
                }
            }
            if (issuer) tma_store_wait_all();
        } else if (grp == 1) {
            // generic path: the second group only keeps the accumulator hand-shake in lock step
            uint32_t local = 0;
            for (int ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, ++local) {
                const uint32_t acc = local & 1u, use = local >> 1;
                mbar_wait(tmem_full + acc, use & 1u);
                tc_fence_after();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_tmem_empty(acc);
            }
        } else {
        const int qd = warp & 3;
        const int r = qd * 32 + lane;                 // row of the tile == TMEM lane
        const int et = threadIdx.x - 64;              // 0..127 within the epilogue group
        const int wi = r % p.BW, hi = (r / p.BW) % p.BH, ni = r / (p.BW * p.BH);
        const bool issuer = (warp == 2 && lane == 0);
        const int64_t hw = (int64_t)p.H * p.W;
        const bool use_tbl = (p.bias || p.bias2 || p.rowadd) && p.BNI <= kTblRows;
        uint32_t local = 0, chunk_ctr = 0;
        for (int ct = cluster_id; ct < p.cluster_tiles; ct += num_clusters, ++local) {
            const int nt = ct % p.n_tiles, co0 = nt * p.BN;
            const uint32_t acc = local & 1u, use = local >> 1;
            mbar_wait(tmem_full + acc, use & 1u);
            tc_fence_after();
            for (int sub = 0; sub < p.msub; ++sub) {
                int mt = ((ct / p.n_tiles) * kCluster + crank) * p.msub + sub;
                const int tw = mt % p.tiles_w; mt /= p.tiles_w;
                const int th = mt % p.tiles_h; mt /= p.tiles_h;
                const int x0 = tw * p.BW, y0 = th * p.BH, n0 = mt * p.BNI;
                const int x = x0 + wi, y = y0 + hi, n = n0 + ni;
                const bool valid = x < p.W && y < p.H && n < p.N;
                const int64_t pix = ((int64_t)n * p.H + y) * p.W + x;
                if (use_tbl) {
                    // tbl[s][c] = bias[co0+c] + rowadd[n0+s][co0+c]: one smem row per sample of the pixel tile
                    named_barrier_sync(1, kEpiThreads);             // previous readers of the table are done
                    for (int i = et; i < p.BNI * p.BN; i += kEpiThreads) {
                        const int sidx = i / p.BN, c = i - sidx * p.BN, co = co0 + c;
                        float v = 0.f;
                        if (co < p.Cout) {
                            if (p.bias) v = __ldg(p.bias + co);
                            if (p.bias2) v += __ldg(p.bias2 + co);
                            if (p.rowadd && n0 + sidx < p.N) v += __ldg(p.rowadd + (int64_t)(n0 + sidx) * p.Cout + co);
                        }
                        tbl[i] = v;
                    }
                    named_barrier_sync(1, kEpiThreads);
                }
                const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16) + acc * (uint32_t)(p.msub * p.BN) + sub * (uint32_t)p.BN;
                const int nchunks = (p.BN + 63) / 64;
                for (int ch = 0; ch < nchunks; ++ch, ++chunk_ctr) {
                    uint8_t *stage = smem_out + (chunk_ctr & 1u) * kStagingBytes;
                    if (p.has_out) {
                        if (issuer) tma_store_wait_read<1>();       // the store that last used this buffer has read it
                        named_barrier_sync(1, kEpiThreads);
                    }
                    const int groups = (p.BN - ch * 64) >= 64 ? 4 : (p.BN - ch * 64) / 16;
                    float v[64];
                    if (groups == 4) {
                        tmem_ld64(trow + ch * 64, v);
                    } else {
#pragma unroll
                        for (int cg = 0; cg < 3; ++cg) {
                            if (cg < groups) {                   // warp-uniform
                                float t16[16];
                                tmem_ld16(trow + ch * 64 + cg * 16, t16);
#pragma unroll
                                for (int i = 0; i < 16; ++i) v[cg * 16 + i] = t16[i];
                            }
                        }
                    }
                    if (ch == nchunks - 1 && sub == p.msub - 1) {   // all tcgen05.ld of this accumulator are done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_tmem_empty(acc);
                    }
#pragma unroll
                    for (int cg = 0; cg < 4; ++cg) {
                        if (cg >= groups) continue;
                        float *vv = v + cg * 16;
                        const int co = co0 + ch * 64 + cg * 16;
                        const int nv = p.Cout - co < 16 ? p.Cout - co : 16;    // may be <= 0 in the padded tail
                        if (valid && nv > 0) {
                            if (use_tbl) {
                                const float4 *tp = reinterpret_cast<const float4 *>(tbl + ni * p.BN + ch * 64 + cg * 16);
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float4 t = tp[i];
                                    vv[4 * i] += t.x; vv[4 * i + 1] += t.y; vv[4 * i + 2] += t.z; vv[4 * i + 3] += t.w;
                                }
                            } else {
                                if (p.bias) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) if (i < nv) vv[i] += __ldg(p.bias + co + i);
                                }
                                if (p.bias2) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) if (i < nv) vv[i] += __ldg(p.bias2 + co + i);
                                }
                                if (p.rowadd) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) if (i < nv) vv[i] += __ldg(p.rowadd + (int64_t)n * p.Cout + co + i);
                                }
                            }
                            if (p.residual) {
                                const __nv_bfloat16 *rp = p.residual + pix * p.ld_res + co;
                                if (nv == 16) {
                                    float f[8];
                                    unpack8(*reinterpret_cast<const uint4 *>(rp), f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) vv[i] += f[i];
                                    unpack8(*reinterpret_cast<const uint4 *>(rp + 8), f);
#pragma unroll
                                    for (int i = 0; i < 8; ++i) vv[8 + i] += f[i];
                                } else {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) if (i < nv) vv[i] += __bfloat162float(rp[i]);
                                }
                            }
                            if (p.out_nchw) {
                                float *op = p.out_nchw + ((int64_t)n * p.Cout + co) * hw + (int64_t)y * p.W + x;
#pragma unroll
                                for (int i = 0; i < 16; ++i) if (i < nv) op[i * hw] = vv[i];
                            }
                        }
                        if (p.has_out) {
                            // two 16-byte pieces of row r, 128-byte swizzle: piece index XOR (row & 7)
                            float f[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = vv[i];
                            *reinterpret_cast<uint4 *>(stage + r * 128 + (((2 * cg) ^ (r & 7)) << 4)) = pack8(f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) f[i] = vv[8 + i];
                            *reinterpret_cast<uint4 *>(stage + r * 128 + (((2 * cg + 1) ^ (r & 7)) << 4)) = pack8(f);
                        }
                    }
                    if (p.has_out) {
                        fence_proxy_async();
                        named_barrier_sync(1, kEpiThreads);
                        if (issuer) {
                            tma_store_4d(&tm_out, stage, co0 + ch * 64, x0, y0, n0);
                            tma_store_commit();
                        }
                    }
                }
            }
        }
        if (issuer && p.has_out) tma_store_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();   // nobody exits while the peer may still write our smem / barriers
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// --------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------
struct PixelTile { int BW, BH, BNI; };

// box of 128 pixels (powers of two per axis) that wastes the fewest rows on ragged edges
PixelTile pick_pixel_tile(int64_t N, int64_t H, int64_t W) {
    PixelTile best{1, 1, 128};
    double best_cost = 1e300;
    for (int bw = 1; bw <= 128; bw *= 2)
        for (int bh = 1; bw * bh <= 128; bh *= 2) {
            const int bn = 128 / (bw * bh);
            const double cost = (double)((W + bw - 1) / bw * bw) * (double)((H + bh - 1) / bh * bh) *
                                (double)((N + bn - 1) / bn * bn);
            // ties: prefer wide boxes (longer contiguous runs per TMA row), then tall ones (fewer samples per tile: the
            // epilogue's per-sample addend table stays within kTblRows rows)
            if (cost < best_cost - 0.5 || (cost < best_cost + 0.5 && (bw > best.BW || (bw == best.BW && bh > best.BH)))) {
                best_cost = cost; best = {bw, bh, bn};
            }
        }
    return best;
}

int pick_bk(int64_t Cin, int64_t Cin2) {
    for (int bk : {64, 32, 16})
        if (Cin % bk == 0 && (Cin2 == 0 || Cin2 % bk == 0)) return bk;
    return 0;
}

uint32_t pow2_at_least(uint32_t v, uint32_t lo) { uint32_t r = lo; while (r < v) r <<= 1; return r; }

template <int BK, bool FAST, bool RES, bool TBL, bool PAIR = false>
int launch_fprop(const CUtensorMap &ta, const CUtensorMap &tw, const CUtensorMap &ta2, const CUtensorMap &tw2,
                 const CUtensorMap &tout, const FpropParams &p, int grid, size_t smem, cudaStream_t s) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_fprop_kernel<BK, FAST, RES, TBL, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        227 * 1024);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    pdl_attr(attr[1]);
    cfg.attrs = attr; cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_fprop_kernel<BK, FAST, RES, TBL, PAIR>, ta, tw, ta2, tw2, tout, p);
    if (e != cudaSuccess) return (int)e;
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // namespace

extern "C" int ub200_conv_fprop(const ub200_conv_args *a, void *stream) {
    UB_REQUIRE(a && a->a && a->w && (a->out || a->out_f32_nchw), UB200_E_BADARG);
    UB_REQUIRE(a->N > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0, UB200_E_BADARG);
    UB_REQUIRE(a->ksize == 1 || a->ksize == 3, UB200_E_UNSUPPORTED);
    const bool extra = a->a2 != nullptr;
    UB_REQUIRE(!extra || (a->w2 && a->Cin2 > 0), UB200_E_BADARG);
    const int bk = pick_bk(a->Cin, extra ? a->Cin2 : 0);
    UB_REQUIRE(bk != 0, UB200_E_UNSUPPORTED);
    UB_REQUIRE(a->ld_a % 8 == 0 && a->ld_a >= a->Cin && ub::aligned16(a->a), UB200_E_UNSUPPORTED);
    UB_REQUIRE(!extra || (a->ld_a2 % 8 == 0 && a->ld_a2 >= a->Cin2 && ub::aligned16(a->a2)), UB200_E_UNSUPPORTED);
    UB_REQUIRE(!a->out || (a->Cout % 8 == 0 && a->ld_out % 8 == 0 && a->ld_out >= a->Cout && ub::aligned16(a->out)),
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(!a->residual || (a->ld_res % 8 == 0 && a->ld_res >= a->Cout && ub::aligned16(a->residual) && a->Cout % 8 == 0),
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(a->N < (1 << 24) && a->H < (1 << 15) && a->W < (1 << 15) && a->Cin <= 16384 && a->Cout <= 16384,
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(!a->gn_partial, UB200_E_UNSUPPORTED);   // epilogue statistics: not in this build yet
    const int S = a->stride == 0 ? 1 : a->stride;
    UB_REQUIRE(S == 1 || S == 2, UB200_E_UNSUPPORTED);
    // a->H, a->W are the INPUT extents; with padding k/2 the output extents are ceil(H / S), ceil(W / S)
    const int64_t Ho = (a->H + S - 1) / S, Wo = (a->W + S - 1) / S;

    const int64_t cout_pad = (a->Cout + 15) / 16 * 16;
    FpropParams p{};
    p.N = (int)a->N; p.H = (int)Ho; p.W = (int)Wo; p.Cout = (int)a->Cout; p.S = S;
    const PixelTile pt = pick_pixel_tile(a->N, Ho, Wo);
    p.BW = pt.BW; p.BH = pt.BH; p.BNI = pt.BNI;
    p.tiles_w = (p.W + p.BW - 1) / p.BW;
    p.tiles_h = (p.H + p.BH - 1) / p.BH;
    const int m_tiles = p.tiles_w * p.tiles_h * ((p.N + p.BNI - 1) / p.BNI);
    // channel tile: as wide as possible (up to 256 TMEM columns per accumulator) while the tile list still fills
    // most of the 148 SMs; narrow tiles re-read A and make the single-thread MMA issue rate the limit
    static const int env_bn = [] { const char *e = getenv("UB200_FPROP_BN"); return e ? atoi(e) : 0; }();
    static const int env_stages = [] { const char *e = getenv("UB200_FPROP_STAGES"); return e ? atoi(e) : 0; }();
    int bn_max = 256;
    if (env_bn) bn_max = env_bn;
    else while (bn_max > 64 && (int64_t)m_tiles * ((cout_pad + bn_max - 1) / bn_max) < 120) bn_max >>= 1;
    p.n_tiles = (int)((cout_pad + bn_max - 1) / bn_max);
    // several channel tiles: keep BN a multiple of the 64-channel store box so no tile writes columns it did not
    // compute; a single tile may be any multiple of 16 (columns past Cout fall outside the tensor and are clipped)
    p.BN = p.n_tiles > 1 ? (int)(((cout_pad + p.n_tiles - 1) / p.n_tiles + 63) / 64 * 64) : (int)cout_pad;
    p.m_tiles = m_tiles;
    // narrow outputs (one channel tile of <= 128): two pixel tiles share each weight tile (halves the weight traffic
    // per MAC and the k-block count per MAC); 4 accumulators x BN <= 512 TMEM columns
    static const int env_cluster = [] { const char *e = getenv("UB200_FPROP_CLUSTER"); return e ? atoi(e) : 1; }();
    static const int env_msub = [] { const char *e = getenv("UB200_FPROP_MSUB"); return e ? atoi(e) : 2; }();
    // CTA pairs (cta_group::2; UB200_FPROP_PAIR=0 switches them off).  Each CTA stages half of the weight tile, so the per-SM
    // shared-memory fill per MMA cycle drops by a third on the 256-channel layers (the L2 -> SM ingest, ~68 B/clk measured,
    // is what caps cta_group::1 at ~72 % tensor-active there).  Measured on B200: 256->256 @ 32x32 1410 -> 1490 TFLOP/s,
    // 512->256 @ 8x8 820 -> 897; the 1x1 convolutions (2-12 k-blocks, epilogue-bound) LOSE 10-25 % to the pair's extra
    // hand-shakes, so only 3x3 convolutions with a long K loop are paired.
    static const int env_pair = [] { const char *e = getenv("UB200_FPROP_PAIR"); return e ? atoi(e) : 1; }();
    static const int env_fast = [] { const char *e = getenv("UB200_FPROP_FAST_EPI"); return e ? atoi(e) : 1; }();
    const bool has_tbl = (a->bias || a->bias2 || a->rowadd);
    const bool fast = env_fast && bk == 64 && a->out && !a->out_f32_nchw && a->Cout % 64 == 0 && p.BN % 64 == 0 &&
                      (!has_tbl || p.BNI <= kTblRows);
    p.pair = (env_pair == 2 || (env_pair && a->ksize == 3 && a->Cin / bk * 9 >= 16)) && fast && m_tiles >= 2 ? 1 : 0;   // fast epilogue only
    p.cluster = (env_cluster == 2 || p.pair) ? 2 : 1;
    const int kCluster = p.cluster;
    p.msub = (env_msub == 2 && p.n_tiles == 1 && p.BN <= 128 && m_tiles >= 2 * ub::kSMs) ? 2 : 1;
    const int m_units = (m_tiles + p.msub - 1) / p.msub;
    p.num_tiles = m_units * p.n_tiles;
    p.cluster_tiles = ((m_units + kCluster - 1) / kCluster) * p.n_tiles;
    // generic: [BNI][BN]; fast: [2 tile parities][msub][BNI][BN]
    p.tbl_bytes = ((uint32_t)((fast ? 2 * p.msub : 1) * (p.BNI <= kTblRows ? p.BNI : 1) * p.BN * 4) + 1023u) & ~1023u;
    p.Cin = (int)a->Cin;
    p.cblocks = (int)(a->Cin / bk);
    p.taps = a->ksize * a->ksize;
    p.kb_main = p.taps * p.cblocks;
    p.kb_extra = extra ? (int)(a->Cin2 / bk) : 0;
    p.a_stage_bytes = (uint32_t)p.msub * 128u * bk * 2u;
    p.b_stage_bytes = ((uint32_t)(p.pair ? p.BN / 2 : p.BN) * bk * 2u + 1023u) & ~1023u;
    const uint32_t stage = p.a_stage_bytes + p.b_stage_bytes;
    const uint32_t budget = 227u * 1024u - 1024u - kStagingBufs * kStagingBytes - p.tbl_bytes - 256u;
    int stages = (int)(budget / stage);
    if (stages > 8) stages = 8;
    if (env_stages && stages > env_stages) stages = env_stages;
    if (stages < 2) stages = 2;
    p.stages = stages;
    p.tmem_cols = pow2_at_least(2u * (uint32_t)(p.msub * p.BN), 32);
    p.bias = a->bias; p.bias2 = a->bias2; p.rowadd = a->rowadd;
    p.residual = reinterpret_cast<const __nv_bfloat16 *>(a->residual); p.ld_res = a->ld_res;
    p.has_out = a->out != nullptr;
    p.out_nchw = a->out_f32_nchw;

    // tensor maps
    CUtensorMap ta, tw, ta2, tw2, tout;
    {
        const int64_t dims[4] = {a->Cin, a->W, a->H, a->N};
        const int64_t str[3] = {a->ld_a, a->ld_a * a->W, a->ld_a * a->W * a->H};
        const int box[4] = {bk, p.BW, p.BH, p.BNI};
        const int es[4] = {1, S, S, 1};
        int rc = encode_bf16_tensor_map(&ta, a->a, 4, dims, str, box, es);
        if (rc) return rc;
        const int64_t wd[2] = {(int64_t)p.taps * a->Cin, cout_pad};
        const int64_t ws[1] = {(int64_t)p.taps * a->Cin};
        const int wb[2] = {bk, p.BN / kCluster};            // each CTA of the pair fetches half the rows and multicasts
        rc = encode_bf16_tensor_map(&tw, a->w, 2, wd, ws, wb);
        if (rc) return rc;
    }
    if (extra) {
        const int64_t dims[4] = {a->Cin2, Wo, Ho, a->N};          // the 1x1 term lives at the output resolution
        const int64_t str[3] = {a->ld_a2, a->ld_a2 * Wo, a->ld_a2 * Wo * Ho};
        const int box[4] = {bk, p.BW, p.BH, p.BNI};
        int rc = encode_bf16_tensor_map(&ta2, a->a2, 4, dims, str, box);
        if (rc) return rc;
        const int64_t wd[2] = {a->Cin2, cout_pad};
        const int64_t ws[1] = {a->Cin2};
        const int wb[2] = {bk, p.BN / kCluster};
        rc = encode_bf16_tensor_map(&tw2, a->w2, 2, wd, ws, wb);
        if (rc) return rc;
    } else {
        ta2 = ta; tw2 = tw;
    }
    if (a->out) {
        const int64_t dims[4] = {a->Cout, Wo, Ho, a->N};
        const int64_t str[3] = {a->ld_out, a->ld_out * Wo, a->ld_out * Wo * Ho};
        const int box[4] = {64, p.BW, p.BH, p.BNI};
        int rc = encode_bf16_tensor_map(&tout, a->out, 4, dims, str, box);
        if (rc) return rc;
    } else {
        tout = ta;
    }

    const int grid = kCluster * (p.cluster_tiles < ub::kSMs / kCluster ? p.cluster_tiles : ub::kSMs / kCluster);
    const size_t smem = 1024 + kStagingBufs * kStagingBytes + p.tbl_bytes + (size_t)stages * stage + (2 * stages + 4) * sizeof(uint64_t) + 16;
    cudaStream_t s = ub::as_stream(stream);
    if (fast && p.pair) {
        const bool res = a->residual != nullptr;
        if (res) return has_tbl ? launch_fprop<64, true, true, true, true>(ta, tw, ta2, tw2, tout, p, grid, smem, s)
                                : launch_fprop<64, true, true, false, true>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
        return has_tbl ? launch_fprop<64, true, false, true, true>(ta, tw, ta2, tw2, tout, p, grid, smem, s)
                       : launch_fprop<64, true, false, false, true>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
    }
    if (fast) {
        const bool res = a->residual != nullptr;
        if (res) return has_tbl ? launch_fprop<64, true, true, true>(ta, tw, ta2, tw2, tout, p, grid, smem, s)
                                : launch_fprop<64, true, true, false>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
        return has_tbl ? launch_fprop<64, true, false, true>(ta, tw, ta2, tw2, tout, p, grid, smem, s)
                       : launch_fprop<64, true, false, false>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
    }
    switch (bk) {
        case 64: return launch_fprop<64, false, false, false>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
        case 32: return launch_fprop<32, false, false, false>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
        default: return launch_fprop<16, false, false, false>(ta, tw, ta2, tw2, tout, p, grid, smem, s);
    }
}
