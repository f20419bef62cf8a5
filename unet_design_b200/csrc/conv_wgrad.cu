// Weight gradient of the 3x3 / 1x1 stride-1 "same" convolution on tcgen05 tensor cores.
//
//   dW[co, (ky,kx), ci] += sum over pixels  G[pix, co] * A[pix + (ky-1, kx-1), ci]
//
// is a GEMM whose reduction dimension is the PIXEL index, so both operands are "MN-major" for the tensor
// core: a TMA box of 64 pixels x CW channels of an NHWC tensor lands in shared memory as 64 rows (K) of
// CW contiguous channels (M or N) -- exactly the canonical MN-major swizzled layout -- and is consumed
// in place.  One CTA owns a [128 co] x [<=128 ci] x [3 taps of one filter row] slab of dW: three fp32
// accumulators in TMEM (3 x 128 columns), fed by one G tile and three shifted A tiles per 64-pixel stage
// (the shifts are TMA coordinates; zero padding is the TMA out-of-bounds fill; in halo mode the three taps are
// row offsets into ONE haloed box).  The pixel range is split across CTAs (split-K) and the partial slabs are
// added into dW -- kept in the packed [Cout, k, k, Cin] layout (= the channels_last memory of the torch weight)
// -- as TMA reduce-add boxes (cp.reduce.async.bulk.tensor; per-thread 128-bit fp32 atomics only where a ci tile
// is not whole 32-column boxes).  Layers with an even number of whole co tiles run as CTA pairs (cta_group::2).
//
// Replaces the wgrad half of nn.Conv2d backward at the sites listed in conv_fprop.cu.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace {
using namespace ub;
using namespace ub::tc;

constexpr int kThreads = 192;
constexpr int kPix = 64;   // pixels (GEMM K) per pipeline stage

struct WgradParams {
    int N, H, W, Cin, Cout, ksize;
    int BW, BH, BNI, tiles_w, tiles_h, pix_tiles, tiles_per_split;
    int ci_tiles, ky_groups, ntaps;     // ntaps = taps handled by one CTA (3 or 1)
    int co_tiles, pair;
    int cw_g, cw_a;                     // channel chunk width (elements) of the G and A boxes
    int stages;
    uint32_t g_stage_bytes, a_tap_bytes, tmem_cols;
    // halo mode (3x3, 128-byte channel chunks, BW % 16 == 0): the three kx taps of a stage are ONE TMA box of
    // (BW + 2) x BH x BNI pixels; tap kx of pixel (x, y) is row y * (BW + 2) + x + kx of that box, so each 16-pixel
    // MMA slab is a run of 16 consecutive 128-byte rows starting at an arbitrary row (not an 8-row swizzle atom)
    int halo, bo_mode;
    int S;                              // conv stride: A pixel = S * G pixel + tap offset (TMA traversal stride S)
    int k1wide;                         // 1x1 conv: the CTA's `ntaps` accumulators are consecutive 128-channel ci tiles, so one
                                        // G tile feeds 2-3 of them (a lone 128x128 accumulator needs 128 B/clk of operand fill)
    uint32_t a_chunk_stride, a_halo_bytes, ring_bytes;
    int dbg;                            // diagnostics (UB200_WGRAD_DEBUG): 1 = no atomics, 2 = no TMEM loads either
    int tma_red;                        // split-K partials leave through TMA reduce-add boxes instead of per-thread atomics
    float *dw;
};

// PAIR (template, not a run-time flag: a kernel that contains cta_group::2 instructions only launches with an even cluster
// size): two CTAs with ADJACENT co tiles, the same ci tile / filter row / pixel range form a cluster; ONE tcgen05.mma
// .cta_group::2 of M = 256 drives both.  Each CTA stages its own 128-channel G tile and HALF of the activation halo box (64 of
// the 128 ci), so per MMA cycle an SM reads 96 instead of 128 bytes of operands and fills 33 instead of 45: the M = N = 128
// single-CTA shape sits at the 128 B/clk shared-memory port.
template <bool PAIR>
__global__ void __launch_bounds__(kThreads) conv_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                             const __grid_constant__ CUtensorMap tm_a,
                                                             const __grid_constant__ CUtensorMap tm_dw,
                                                             const WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = p.g_stage_bytes + (p.halo ? p.a_halo_bytes : p.ntaps * p.a_tap_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + p.ring_bytes);        // behind the ring (>= stages * stage_bytes)
    uint64_t *empty = full + p.stages;
    uint64_t *tmem_full = empty + p.stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    // shuffled warp index: provably warp-uniform, so the role branches are uniform control flow (see conv_fprop.cu)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    pdl_trigger();

    // blockIdx.x -> (ci tile, ky group, co tile); blockIdx.y -> pixel split.  PAIR: the co tile is the fastest index, so the two
    // CTAs of a cluster (consecutive blockIdx.x) hold co tiles 2j and 2j + 1
    int t = blockIdx.x;
    int cit, kyg, cot;
    if constexpr (PAIR) {
        cot = t % p.co_tiles; t /= p.co_tiles;
        kyg = t % p.ky_groups; cit = t / p.ky_groups;
    } else {
        cit = t % p.ci_tiles; t /= p.ci_tiles;
        kyg = t % p.ky_groups; cot = t / p.ky_groups;
    }
    const int crank = PAIR ? (int)cluster_ctarank() : 0;
    const int co0 = cot * 128, ci0 = cit * 128 * (p.k1wide ? p.ntaps : 1);
    const int ncols = (p.Cin - ci0 < 128 ? p.Cin - ci0 : 128);          // multiple of 16
    const int mrows = (p.Cout - co0 < 128 ? p.Cout - co0 : 128);
    const int chunks_g = (mrows + p.cw_g - 1) / p.cw_g, chunks_a = (ncols + p.cw_a - 1) / p.cw_a;
    const int pt0 = blockIdx.y * p.tiles_per_split;
    int pt1 = pt0 + p.tiles_per_split;
    if (pt1 > p.pix_tiles) pt1 = p.pix_tiles;
    const int iters = pt1 - pt0;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_g);
        prefetch_tmap(&tm_a);
        if (p.tma_red) prefetch_tmap(&tm_dw);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) { if constexpr (PAIR) tmem_alloc_pair(tmem_slot, p.tmem_cols); else tmem_alloc(tmem_slot, p.tmem_cols); }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // the peer's TMA loads and commits arrive on our barriers
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();

    const uint32_t g_chunk_bytes = kPix * p.cw_g * 2, a_chunk_bytes = kPix * p.cw_a * 2;

    if (warp == 0) {
        // ===================== TMA producer: warp-uniform loop, lane 0 issues (no div/mod per stage) =====================
        {
            const bool leader = lane == 0;
            const uint32_t halo_box_bytes = (uint32_t)((p.BW + 2) * p.BH * p.BNI) * 128u;
            // pair: the leader's barrier counts the bytes of both CTAs (each: its G tile + ONE 64-channel chunk of the halo box)
            const uint32_t tx_bytes = PAIR ? 2u * (chunks_g * g_chunk_bytes + (p.halo ? halo_box_bytes : p.ntaps * a_chunk_bytes))
                                           : chunks_g * g_chunk_bytes +
                                                 (p.halo ? chunks_a * halo_box_bytes : p.ntaps * chunks_a * a_chunk_bytes);
            const uint32_t base = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
            uint32_t full_lead = full0;                                       // cluster address of the leader's `full` barriers
            if constexpr (PAIR) full_lead = mapa_rank(full0, 0);
            // first pixel tile of this split, then advance (x fastest, then y, then image group) incrementally
            int pt = pt0;
            int tw = pt % p.tiles_w; pt /= p.tiles_w;
            int th = pt % p.tiles_h; pt /= p.tiles_h;
            int tn = pt;
            const int ky = p.ksize == 3 ? kyg - 1 : 0;
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < iters; ++it) {
                const int x0 = tw * p.BW, y0 = th * p.BH, n0 = tn * p.BNI;
                const uint32_t fb = full0 + 8u * s, sg = base + (uint32_t)s * stage_bytes;
                mbar_wait_a(empty0 + 8u * s, ph ^ 1u);
                if constexpr (PAIR) {
                    if (leader) {
                        const uint32_t fl = full_lead + 8u * s;
                        if (crank == 0) mbar_arrive_expect_tx_a(fb, tx_bytes);
                        for (int c = 0; c < chunks_g; ++c)
                            tma_load_4d_pair_a(sg + c * g_chunk_bytes, &tm_g, fl, co0 + c * p.cw_g, x0, y0, n0);
                        if (p.halo) {
                            tma_load_4d_pair_a(sg + p.g_stage_bytes, &tm_a, fl, ci0 + crank * 64, x0 - 1, y0 + ky, n0);
                        } else {
                            for (int tp = 0; tp < p.ntaps; ++tp) {            // one 64-channel chunk per tap and CTA
                                const int kx = p.ksize == 3 ? tp - 1 : 0;
                                tma_load_4d_pair_a(sg + p.g_stage_bytes + tp * p.a_tap_bytes, &tm_a, fl, ci0 + crank * 64,
                                                   p.S * x0 + kx, p.S * y0 + ky, n0);
                            }
                        }
                    }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                    if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
                    continue;
                }
                if (leader) mbar_arrive_expect_tx_a(fb, tx_bytes);
                for (int c = 0; c < chunks_g; ++c)
                    if (leader) tma_load_4d_a(sg + c * g_chunk_bytes, &tm_g, fb, co0 + c * p.cw_g, x0, y0, n0);
                if (p.halo) {
                    for (int c = 0; c < chunks_a; ++c)
                        if (leader) tma_load_4d_a(sg + p.g_stage_bytes + c * p.a_chunk_stride, &tm_a, fb, ci0 + c * 64, x0 - 1, y0 + ky, n0);
                } else {
                    for (int tp = 0; tp < p.ntaps; ++tp) {
                        const int kx = p.ksize == 3 ? tp - 1 : 0;
                        const uint32_t sa = sg + p.g_stage_bytes + tp * p.a_tap_bytes;
                        for (int c = 0; c < chunks_a; ++c)
                            if (leader) tma_load_4d_a(sa + c * a_chunk_bytes, &tm_a, fb, ci0 + c * p.cw_a + (p.k1wide ? tp * 128 : 0), p.S * x0 + kx, p.S * y0 + ky, n0);
                    }
                }
                if (++s == p.stages) { s = 0; ph ^= 1u; }
                if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++tn; } }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: warp-uniform loop, lane 0 issues =====================
        // (pair: only the leader CTA issues; its instruction drives the tensor cores of both SMs)
        if (!(PAIR && crank != 0)) {
            const bool leader = lane == 0;
            const uint32_t idesc = make_idesc(PAIR ? 256 : 128, ncols, 1, 1);
            const uint32_t swz_g = swizzle_code(p.cw_g * 2), swz_a = swizzle_code(p.cw_a * 2);
            const uint32_t row_g = p.cw_g * 2, row_a = p.cw_a * 2;
            const uint32_t base = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
            const uint64_t dg0 = make_smem_desc(base, g_chunk_bytes, 8 * row_g, swz_g);
            const uint64_t da0 = make_smem_desc(base + p.g_stage_bytes, a_chunk_bytes, 8 * row_a, swz_a);
            const uint32_t kstep_g = (16 * row_g) >> 4, kstep_a = (16 * row_a) >> 4, tap_step = p.a_tap_bytes >> 4;
            const uint32_t stage_step = stage_bytes >> 4;
            // halo mode: first box row of each 16-pixel slab (tap kx adds kx rows); rows are 128 bytes = 8 descriptor units
            uint32_t slab_row[kPix / 16];
#pragma unroll
            for (int k = 0; k < kPix / 16; ++k) slab_row[k] = (uint32_t)(((16 * k) / p.BW) * (p.BW + 2) + (16 * k) % p.BW);
            const uint64_t dh0 = make_smem_desc(base + p.g_stage_bytes, p.a_chunk_stride, 1024, 2u);
            int s = 0; uint32_t ph = 0, accum = 0;
            for (int it = 0; it < iters; ++it) {
                mbar_wait_a(full0 + 8u * s, ph);
                tc_fence_after();
                const uint64_t dg = dg0 + (uint64_t)(s * stage_step), da = da0 + (uint64_t)(s * stage_step);
                if (PAIR && !p.halo) {
                    if constexpr (PAIR) {
                        for (int tp = 0; tp < p.ntaps; ++tp) {
                            if (leader) {
#pragma unroll
                                for (int k = 0; k < kPix / 16; ++k)
                                    umma_bf16_pair(tmem_base + tp * 128, dg + k * kstep_g, da + tp * tap_step + k * kstep_a, idesc,
                                                   (accum | (uint32_t)k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                } else if (p.halo) {
                    const uint64_t dh = dh0 + (uint64_t)(s * stage_step);
                    for (int tp = 0; tp < 3; ++tp) {
#pragma unroll
                        for (int k = 0; k < kPix / 16; ++k) {
                            const uint32_t row = slab_row[k] + (uint32_t)tp;
                            // swizzle phase of a start address that is not 1024-byte aligned (descriptor bits 49..51)
                            const uint32_t bo = p.bo_mode == 0 ? 0u : p.bo_mode == 1 ? (row & 7u) : ((8u - (row & 7u)) & 7u);
                            const uint64_t db = (dh + (uint64_t)(row * 8u)) | ((uint64_t)bo << 49);
                            if constexpr (PAIR) {
                                if (leader) umma_bf16_pair(tmem_base + tp * 128, dg + k * kstep_g, db, idesc, (accum | (uint32_t)k) != 0 ? 1u : 0u);
                            } else {
                                if (leader) umma_bf16(tmem_base + tp * 128, dg + k * kstep_g, db, idesc, (accum | (uint32_t)k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                } else if constexpr (!PAIR) {
                for (int tp = 0; tp < p.ntaps; ++tp) {
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < kPix / 16; ++k)
                            umma_bf16(tmem_base + tp * 128, dg + k * kstep_g, da + tp * tap_step + k * kstep_a, idesc,
                                      (accum | (uint32_t)k) != 0 ? 1u : 0u);
                    }
                }
                }
                accum = 1;
                if constexpr (PAIR) { if (leader) umma_commit_pair_a(empty0 + 8u * s, (uint16_t)0x3); }   // frees the slot in both CTAs
                else if (leader) umma_commit_a(empty0 + 8u * s);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            if (iters > 0 && leader) {
                if constexpr (PAIR) umma_commit_pair_a(smem_u32(tmem_full), (uint16_t)0x3);     // accumulators complete in both CTAs
                else umma_commit(tmem_full);
            }
        }
    } else if (iters > 0) {
        const int qd = warp & 3;
        const int co = co0 + qd * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
        const int taps_total = p.ksize * p.ksize;
        if (p.tma_red) {
            // Split-K partials through the TMA unit: each warp stages its 32 rows x 32 fp32 columns (4 KB, 128-byte swizzle) in
            // the drained pipeline ring and lane 0 issues ONE reduce-add box; the L2 does the adds.  The per-thread path below
            // issues 96 16-byte reds per thread, and an SM retires ~0.8 red lanes per clock: 12288 lane-ops = 8 us per CTA.
            constexpr int kBufs = 4;
            const uint32_t ring = smem_u32(smem) + (uint32_t)(warp - 2) * (kBufs * 4096u);
            const uint32_t row_off = (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7);
            int c = 0;
            for (int tp = 0; tp < p.ntaps; ++tp) {
                const int tap = p.ksize == 3 ? kyg * 3 + tp : 0;
                const int col0 = tap * p.Cin + ci0 + (p.k1wide ? tp * 128 : 0);
                for (int cg = 0; cg < ncols / 32; ++cg, ++c) {
                    const uint32_t buf = ring + (uint32_t)(c % kBufs) * 4096u;
                    if (c >= kBufs) {                       // the box that last used this buffer has been read
                        if (lane == 0) tma_store_wait_read<kBufs - 1>();
                        __syncwarp();
                    }
                    float v[32];
                    tmem_ld32(trow + tp * 128 + cg * 32, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(buf + row_off + (((uint32_t)j ^ sw) << 4)),
                                     "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tm_dw, buf, col0 + cg * 32, co0 + qd * 32);
                        tma_store_commit();
                    }
                }
            }
            if (lane == 0) tma_store_wait_all();
            __syncwarp();
        } else
        for (int tp = 0; tp < p.ntaps; ++tp) {
            const int tap = p.ksize == 3 ? kyg * 3 + tp : 0;
            for (int cg = 0; cg < ncols / 16; ++cg) {
                float v[16];
                if (p.dbg == 2) continue;
                tmem_ld16(trow + tp * 128 + cg * 16, v);
                if (p.dbg == 1) { if (v[0] == 12345.678f) p.dw[0] = v[1]; continue; }
                if (co < p.Cout) {
                    float4 *dst = reinterpret_cast<float4 *>(p.dw + ((int64_t)co * taps_total + tap) * p.Cin + ci0 + (p.k1wide ? tp * 128 : 0) + cg * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        atomicAdd(dst + i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();      // nobody exits while the leader's MMAs may still read the peer's shared memory
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, p.tmem_cols); else tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// per-channel sums of an NHWC bf16 tensor.  Pass 1: every CTA streams a pixel range of one sample (16 bytes per
// thread), reduces across its threads in shared memory and issues ONE atomic per channel into per_sample[n, c].
// Pass 2 (tiny): total[c] += sum_n per_sample[n, c].
__global__ void __launch_bounds__(256) chansum_kernel(const __nv_bfloat16 *__restrict__ x, int64_t ld, int64_t HW, int C,
                                                     int chunks, int rows, int64_t pix_per_cta,
                                                     float *__restrict__ per_sample) {
    extern __shared__ float part[];   // [256 threads][8]: per-thread partial sums, combined without atomics
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % chunks, r = threadIdx.x / chunks;
    if (r < rows) {
        float s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = 0.f;
        const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta;
        int64_t p1 = p0 + pix_per_cta;
        if (p1 > HW) p1 = HW;
        const __nv_bfloat16 *xb = x + n * HW * ld + 8 * q;
        int64_t pp = p0 + r;
        for (; pp + rows < p1; pp += 2 * rows) {
            const uint4 a = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + pp * ld));
            const uint4 b = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + (pp + rows) * ld));
            float f[8], g[8];
            unpack8(a, f); unpack8(b, g);
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += f[u] + g[u];
        }
        for (; pp < p1; pp += rows) {
            float f[8];
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(xb + pp * ld)), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += f[u];
        }
        float4 *dst = reinterpret_cast<float4 *>(part + (size_t)threadIdx.x * 8);
        dst[0] = make_float4(s[0], s[1], s[2], s[3]); dst[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int rr = 0; rr < rows; ++rr) acc += part[(size_t)(rr * chunks + (c >> 3)) * 8 + (c & 7)];
        atomicAdd(per_sample + n * C + c, acc);
    }
}

// One thread-block cluster per sample (no atomics, no zero-fill of per_sample): every CTA sums its pixel range, rank 0
// combines the cluster through DSMEM and writes per_sample[n, :].  The column sums over n (bias gradient) stay a second,
// tiny launch: doing them in the last-finishing cluster (device counter) was measured SLOWER (+0.3 ms/step), one CTA
// walking N rows is a serial chain of L2 round trips.

__global__ void __launch_bounds__(256) chansum_cluster_kernel(const __nv_bfloat16 *__restrict__ x, int64_t ld, int64_t HW, int C,
                                                             int chunks, int rows, int64_t pix_per_cta, int cs,
                                                             float *__restrict__ per_sample) {
    extern __shared__ float csm[];                 // part [256][8], then chan [C]
    float *part = csm, *chan = csm + 256 * 8;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int64_t n = blockIdx.x / cs;
    const int q = threadIdx.x % chunks, r = threadIdx.x / chunks;
    pdl_trigger();
    pdl_wait();
    float s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] = 0.f;
    if (r < rows) {
        const int64_t p0 = (int64_t)rank * pix_per_cta;
        int64_t p1 = p0 + pix_per_cta;
        if (p1 > HW) p1 = HW;
        const int64_t step = (int64_t)rows * ld;
        const __nv_bfloat16 *xp = x + (n * HW + p0 + r) * ld + 8 * q;
        int64_t pp = p0 + r;
        for (; pp + 3 * rows < p1; pp += 4 * rows, xp += 4 * step) {          // four 16-byte loads in flight per thread
            const uint4 a = ld_stream_u4(reinterpret_cast<const uint4 *>(xp));
            const uint4 b = ld_stream_u4(reinterpret_cast<const uint4 *>(xp + step));
            const uint4 c = ld_stream_u4(reinterpret_cast<const uint4 *>(xp + 2 * step));
            const uint4 d = ld_stream_u4(reinterpret_cast<const uint4 *>(xp + 3 * step));
            float f[8], g[8], h[8], k[8];
            unpack8(a, f); unpack8(b, g); unpack8(c, h); unpack8(d, k);
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += (f[u] + g[u]) + (h[u] + k[u]);
        }
        for (; pp < p1; pp += rows, xp += step) {
            float f[8];
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(xp)), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += f[u];
        }
    }
    {
        float4 *dst = reinterpret_cast<float4 *>(part + (size_t)threadIdx.x * 8);
        dst[0] = make_float4(s[0], s[1], s[2], s[3]); dst[1] = make_float4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float acc = 0.f;
        for (int rr = 0; rr < rows; ++rr) acc += part[(size_t)(rr * chunks + (c >> 3)) * 8 + (c & 7)];
        chan[c] = acc;
    }
    cluster.sync();
    if (rank == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float acc = 0.f;
            for (int rk = 0; rk < cs; ++rk) acc += cluster.map_shared_rank(chan, rk)[c];
            per_sample[n * C + c] = acc;
        }
    }
    cluster.sync();                                // peers stay alive until rank 0 has read their sums
}

__global__ void __launch_bounds__(256) colsum_rows_kernel(const float *__restrict__ per_sample, int64_t N, int C,
                                                         float *__restrict__ total) {
    // 32 channels x 8 row-lanes per CTA; rows strided by 8, combined through shared memory
    __shared__ float part[8][33];
    pdl_trigger();
    pdl_wait();
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float s = 0.f;
    if (c < C)
        for (int64_t n = rl; n < N; n += 8) s += __ldg(per_sample + n * C + c);
    part[rl][cl] = s;
    __syncthreads();
    if (rl == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][cl];
        total[c] += t;
    }
}

// fp32 weights with arbitrary strides -> packed bf16 [rows_pad, k, k, cols] (rows padded to 16 with zeros)
__global__ void __launch_bounds__(256) pack_weight_kernel(const float *__restrict__ w, int Cout, int Cin, int k,
                                                         int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx,
                                                         int transpose_flip, __nv_bfloat16 *__restrict__ out,
                                                         int64_t total) {
    const int rows = transpose_flip ? Cin : Cout, cols = transpose_flip ? Cout : Cin;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        int64_t t = i / cols;
        const int kx = (int)(t % k); t /= k;
        const int ky = (int)(t % k);
        const int r = (int)(t / k);
        float v = 0.f;
        if (r < rows) {
            if (!transpose_flip) v = __ldg(w + r * s_co + c * s_ci + ky * s_ky + kx * s_kx);
            else v = __ldg(w + c * s_co + r * s_ci + (k - 1 - ky) * s_ky + (k - 1 - kx) * s_kx);
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

void pick_pixel_tile64(int64_t N, int64_t H, int64_t W, int &BW, int &BH, int &BNI) {
    double best_cost = 1e300;
    BW = 1; BH = 1; BNI = kPix;
    for (int bw = 1; bw <= kPix; bw *= 2)
        for (int bh = 1; bw * bh <= kPix; bh *= 2) {
            const int bn = kPix / (bw * bh);
            const double cost = (double)((W + bw - 1) / bw * bw) * (double)((H + bh - 1) / bh * bh) *
                                (double)((N + bn - 1) / bn * bn);
            if (cost < best_cost - 0.5 || (cost < best_cost + 0.5 && bw > BW)) { best_cost = cost; BW = bw; BH = bh; BNI = bn; }
        }
}

int chunk_width(int64_t C) { return C % 64 == 0 ? 64 : C % 32 == 0 ? 32 : C % 16 == 0 ? 16 : 0; }

}  // namespace

extern "C" {

int ub200_conv_wgrad(const void *gout, int64_t ld_g, const void *a, int64_t ld_a, int64_t N, int64_t H, int64_t W,
                     int64_t Cin, int64_t Cout, int ksize, float *dw, void *stream) {
    return ub200_conv_wgrad_strided(gout, ld_g, a, ld_a, N, H, W, Cin, Cout, ksize, 1, dw, stream);
}

int ub200_conv_wgrad_strided(const void *gout, int64_t ld_g, const void *a, int64_t ld_a, int64_t N, int64_t Hin, int64_t Win,
                             int64_t Cin, int64_t Cout, int ksize, int stride, float *dw, void *stream) {
    UB_REQUIRE(gout && a && dw && N > 0 && Hin > 0 && Win > 0 && Cin > 0 && Cout > 0, UB200_E_BADARG);
    UB_REQUIRE(stride == 1 || stride == 2, UB200_E_UNSUPPORTED);
    const int64_t H = (Hin + stride - 1) / stride, W = (Win + stride - 1) / stride;      // extents of gout: the pixel tiling
    UB_REQUIRE(ksize == 1 || ksize == 3, UB200_E_UNSUPPORTED);
    WgradParams p{};
    p.cw_g = chunk_width(Cout); p.cw_a = chunk_width(Cin);
    UB_REQUIRE(p.cw_g && p.cw_a, UB200_E_UNSUPPORTED);
    UB_REQUIRE(ld_g % 8 == 0 && ld_a % 8 == 0 && ld_g >= Cout && ld_a >= Cin && ub::aligned16(gout) && ub::aligned16(a) &&
                   ub::aligned16(dw),
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(N < (1 << 24) && H < (1 << 15) && W < (1 << 15) && Cin <= 16384 && Cout <= 16384, UB200_E_UNSUPPORTED);
    p.N = (int)N; p.H = (int)H; p.W = (int)W; p.Cin = (int)Cin; p.Cout = (int)Cout; p.ksize = ksize;
    pick_pixel_tile64(N, H, W, p.BW, p.BH, p.BNI);
    p.tiles_w = (p.W + p.BW - 1) / p.BW;
    p.tiles_h = (p.H + p.BH - 1) / p.BH;
    p.pix_tiles = p.tiles_w * p.tiles_h * ((p.N + p.BNI - 1) / p.BNI);
    p.ci_tiles = (int)((Cin + 127) / 128);
    p.ky_groups = ksize == 3 ? 3 : 1;
    p.ntaps = ksize == 3 ? 3 : 1;
    // measured on B200: 256->768 @ 16x16 27.3 -> 24.5 us, but 384->128 @ 32x32 29.3 -> 38.0 us (one output tile: 148 pixel
    // splits, each paying a 3x larger atomic epilogue) and the config-2 step 6.42 -> 6.45 ms: off unless asked for
    static const int env_k1wide = [] { const char *e = getenv("UB200_WGRAD_K1WIDE"); return e ? atoi(e) : 0; }();
    if (env_k1wide && ksize == 1 && Cin % 128 == 0 && Cin >= 256) {
        p.k1wide = 1;
        p.ntaps = (p.ci_tiles % 3 == 0) ? 3 : (p.ci_tiles % 2 == 0 ? 2 : 1);
        if (p.ntaps == 1) p.k1wide = 0;
        p.ci_tiles /= p.ntaps;
    }
    const int co_tiles = (int)((Cout + 127) / 128);
    p.co_tiles = co_tiles;
    const int out_tiles = p.ci_tiles * p.ky_groups * co_tiles;
    // one CTA per SM (the smem ring takes the whole SM): keep the grid within ONE wave of 148 CTAs so there is no
    // tail wave, and give every split the same number of pixel tiles
    int splits = ub::kSMs / out_tiles;
    if (splits > p.pix_tiles) splits = p.pix_tiles;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (p.pix_tiles + splits - 1) / splits;
    // every split pays the full epilogue (3 x 128 x 128 fp32 atomics): keep at least this many 64-pixel stages per split
    // (6 measured marginally best on config 2: 7.07 -> 7.04 ms/step; 12 is worse)
    static const int min_tiles = [] { const char *e = getenv("UB200_WGRAD_MIN_TILES"); return e ? atoi(e) : 6; }();
    if (p.tiles_per_split < min_tiles) p.tiles_per_split = min_tiles < p.pix_tiles ? min_tiles : p.pix_tiles;
    splits = (p.pix_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    const int ncols_max = (int)(Cin < 128 ? Cin : 128);
    p.g_stage_bytes = kPix * 128 * 2;                       // room for a full 128-channel G tile
    p.a_tap_bytes = (uint32_t)((kPix * ncols_max * 2 + 1023) & ~1023);
    static const int env_halo = [] { const char *e = getenv("UB200_WGRAD_HALO"); return e ? atoi(e) : 1; }();
    static const int env_bo = [] { const char *e = getenv("UB200_WGRAD_BO"); return e ? atoi(e) : 0; }();
    p.S = stride;
    p.halo = (env_halo && stride == 1 && ksize == 3 && p.cw_a == 64 && p.BW % 16 == 0) ? 1 : 0;
    p.bo_mode = env_bo;
    p.a_chunk_stride = (uint32_t)(((p.BW + 2) * p.BH * p.BNI * 128 + 1023) & ~1023);
    p.a_halo_bytes = (uint32_t)((ncols_max + 63) / 64) * p.a_chunk_stride;
    // CTA pairs (cta_group::2, see the kernel): halo mode, whole 128-channel ci tiles, an even number of whole co tiles.
    // Measured on B200 (profiles/r02_wgrad_pair_probe.txt): 256->256 @ 32x32 1,200 -> 1,437 TFLOP/s, 256->256 @ 16x16 1,000 -> 1,180,
    // 512->256 @ 16x16 1,122 -> 1,358; config-2 step 6.05 -> 5.98 ms.  UB200_WGRAD_PAIR=0 switches them off.
    static const int env_pair = [] { const char *e = getenv("UB200_WGRAD_PAIR"); return e ? atoi(e) : 2; }();   // 1: halo layers only, 2: every eligible layer
    p.pair = (env_pair && (p.halo || env_pair >= 2) && p.cw_a == 64 && p.cw_g == 64 && !p.k1wide && Cin % 128 == 0 && Cout % 256 == 0) ? 1 : 0;
    if (p.pair) {           // each CTA stages ONE 64-channel chunk of the halo box / of every tap tile
        p.a_halo_bytes = p.a_chunk_stride;
        p.a_tap_bytes = (uint32_t)(kPix * 64 * 2);
    }
    const uint32_t stage = p.g_stage_bytes + (p.halo ? p.a_halo_bytes : p.ntaps * p.a_tap_bytes);
    static const int env_stages = [] { const char *e = getenv("UB200_WGRAD_STAGES"); return e ? atoi(e) : 6; }();
    int stages = (int)((200u * 1024u) / stage);
    if (stages > env_stages) stages = env_stages;
    if (stages > p.tiles_per_split) stages = p.tiles_per_split;
    if (stages < 1) stages = 1;
    p.stages = stages;
    p.tmem_cols = p.ntaps == 3 ? 512 : (p.ntaps == 2 ? 256 : 128);
    p.dw = dw;
    static const int env_dbg = [] { const char *e = getenv("UB200_WGRAD_DEBUG"); return e ? atoi(e) : 0; }();
    p.dbg = env_dbg;

    CUtensorMap tg, ta;
    {
        const int64_t dims[4] = {Cout, W, H, N};
        const int64_t str[3] = {ld_g, ld_g * W, ld_g * W * H};
        const int box[4] = {p.cw_g, p.BW, p.BH, p.BNI};
        int rc = encode_bf16_tensor_map(&tg, gout, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        const int64_t dims[4] = {Cin, Win, Hin, N};
        const int64_t str[3] = {ld_a, ld_a * Win, ld_a * Win * Hin};
        const int box[4] = {p.cw_a, p.halo ? p.BW + 2 : p.BW, p.BH, p.BNI};
        const int es[4] = {1, stride, stride, 1};
        int rc = encode_bf16_tensor_map(&ta, a, 4, dims, str, box, es);
        if (rc) return rc;
    }
    // TMA reduce-add epilogue: every 128-channel ci tile must be whole 32-column boxes (a box past the tile would land in the next tap)
    static const int env_tma_red = [] { const char *e = getenv("UB200_WGRAD_TMA_RED"); return e ? atoi(e) : 1; }();
    p.tma_red = (env_tma_red && Cin % 32 == 0 && !env_dbg) ? 1 : 0;
    CUtensorMap tdw = tg;
    if (p.tma_red) {
        const int64_t taps_total = (int64_t)ksize * ksize;
        const int64_t dims[2] = {taps_total * Cin, Cout};
        const int64_t str[1] = {taps_total * Cin};
        const int box[2] = {32, 32};
        int rc = encode_f32_tensor_map(&tdw, dw, 2, dims, str, box);
        if (rc) return rc;
    }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    // (the epilogue's staging ring -- 4 warps x 4 boxes x 4 KB -- re-uses the pipeline ring, which must be at least that large;
    // the mbarriers behind the ring are not touched)
    size_t ring_bytes = (size_t)stages * stage;
    if (p.tma_red && ring_bytes < 64 * 1024) ring_bytes = 64 * 1024;
    p.ring_bytes = (uint32_t)ring_bytes;
    const size_t smem = 1024 + ring_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16;
    if (env_dbg == 3) {                 // how many clusters of 2 / 4 / 8 one-per-SM CTAs can be resident (printed once)
        static std::once_flag once2;
        std::call_once(once2, [&] {
            for (int cs : {1, 2, 4, 8}) {
                cudaLaunchConfig_t q{};
                q.gridDim = dim3(8, 8, 1); q.blockDim = dim3(kThreads, 1, 1); q.dynamicSmemBytes = 200 * 1024;
                cudaLaunchAttribute qa[1];
                qa[0].id = cudaLaunchAttributeClusterDimension;
                qa[0].val.clusterDim.x = 1; qa[0].val.clusterDim.y = (unsigned)cs; qa[0].val.clusterDim.z = 1;
                q.attrs = qa; q.numAttrs = 1;
                int nc = -1;
                cudaError_t qe = cudaOccupancyMaxActiveClusters(&nc, conv_wgrad_kernel<false>, &q);
                fprintf(stderr, "[ub200] max active clusters of %d CTAs (200 KB smem): %d (%s)\n", cs, nc, cudaGetErrorString(qe));
            }
        });
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)out_tiles, (unsigned)splits, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ub::as_stream(stream);
    cudaLaunchAttribute attr[2];
    ub::pdl_attr(attr[0]);
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = p.pair ? 2 : 1;
    ub::prefer_max_smem_carveout(reinterpret_cast<const void *>(conv_wgrad_kernel<false>));
    ub::prefer_max_smem_carveout(reinterpret_cast<const void *>(conv_wgrad_kernel<true>));
    cudaError_t le = p.pair ? cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<true>, tg, ta, tdw, p)
                            : cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<false>, tg, ta, tdw, p);
    if (le != cudaSuccess) return (int)le;
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_chansum_nhwc_bf16(const void *x, int64_t ld, int64_t N, int64_t HW, int64_t C, float *per_sample, float *total,
                            void *stream) {
    UB_REQUIRE(x && per_sample && N > 0 && HW > 0 && C > 0, UB200_E_BADARG);
    UB_REQUIRE(C % 8 == 0 && C <= 2048 && ld % 8 == 0 && ld >= C && ub::aligned16(x) && N <= 65535, UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const int chunks = (int)(C / 8), rows = 256 / chunks;
    static const bool use_cluster = [] { const char *e = getenv("UB200_CHANSUM_CLUSTER"); return !(e && e[0] == '0'); }();
    if (use_cluster) {
        // cluster size: enough CTAs to fill the machine, at least 4 passes of the CTA's threads per CTA
        int cs = 1;
        while (cs < 8 && N * cs < 2 * ub::kSMs && HW / (2 * cs) >= 4 * rows) cs *= 2;
        while (cs < 8 && HW * C * 2 / cs > 256 * 1024 && HW / (2 * cs) >= 4 * rows) cs *= 2;
        const int64_t ppc = (HW + cs - 1) / cs;
        if ((int64_t)(cs - 1) * ppc < HW) {           // no empty CTA in the cluster
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(N * cs), 1, 1);
            cfg.blockDim = dim3(256, 1, 1);
            cfg.dynamicSmemBytes = (256 * 8 + C) * sizeof(float);
            cfg.stream = s;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            ub::pdl_attr(attr[1]);
            cfg.attrs = attr; cfg.numAttrs = 2;
            ub::prefer_max_smem_carveout(reinterpret_cast<const void *>(chansum_cluster_kernel));   // runs next to the conv kernels
            ub::prefer_max_smem_carveout(reinterpret_cast<const void *>(colsum_rows_kernel));
            cudaError_t e = cudaLaunchKernelEx(&cfg, chansum_cluster_kernel, reinterpret_cast<const __nv_bfloat16 *>(x), ld, HW,
                                               (int)C, chunks, rows, ppc, cs, per_sample);
            if (e != cudaSuccess) return (int)e;
            UB_LAUNCH_CHECK();
            if (total) {
                cudaLaunchConfig_t c2{};
                c2.gridDim = dim3((unsigned)((C + 31) / 32), 1, 1);
                c2.blockDim = dim3(256, 1, 1);
                c2.stream = s;
                cudaLaunchAttribute a2[1];
                ub::pdl_attr(a2[0]);
                c2.attrs = a2; c2.numAttrs = 1;
                e = cudaLaunchKernelEx(&c2, colsum_rows_kernel, (const float *)per_sample, N, (int)C, total);
                if (e != cudaSuccess) return (int)e;
                UB_LAUNCH_CHECK();
            }
            return UB200_OK;
        }
    }
    cudaError_t e = cudaMemsetAsync(per_sample, 0, sizeof(float) * N * C, s);
    if (e != cudaSuccess) return (int)e;
    int64_t splits = (148 * 8 + N - 1) / N;
    const int64_t max_splits = (HW + rows - 1) / rows;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int64_t ppc = (HW + splits - 1) / splits;
    splits = (HW + ppc - 1) / ppc;
    dim3 grid((unsigned)splits, (unsigned)N, 1);
    ub::prefer_max_smem_carveout(reinterpret_cast<const void *>(chansum_kernel));
    ub::prefer_max_smem_carveout(reinterpret_cast<const void *>(colsum_rows_kernel));
    chansum_kernel<<<grid, 256, 256 * 8 * sizeof(float), s>>>(reinterpret_cast<const __nv_bfloat16 *>(x), ld, HW, (int)C, chunks,
                                                       rows, ppc, per_sample);
    UB_LAUNCH_CHECK();
    if (total) {
        colsum_rows_kernel<<<(unsigned)((C + 31) / 32), 256, 0, s>>>(per_sample, N, (int)C, total);
        UB_LAUNCH_CHECK();
    }
    return UB200_OK;
}

int ub200_colsum_rows_f32(const float *rows, int64_t N, int64_t C, float *total, void *stream) {
    UB_REQUIRE(rows && total && N > 0 && C > 0 && C < (1 << 24), UB200_E_BADARG);
    cudaLaunchConfig_t c2{};
    c2.gridDim = dim3((unsigned)((C + 31) / 32), 1, 1);
    c2.blockDim = dim3(256, 1, 1);
    c2.stream = ub::as_stream(stream);
    cudaLaunchAttribute a2[1];
    ub::pdl_attr(a2[0]);
    c2.attrs = a2; c2.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&c2, colsum_rows_kernel, rows, N, (int)C, total);
    if (e != cudaSuccess) return (int)e;
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_pack_conv_weight(const float *w, int64_t Cout, int64_t Cin, int ksize, int64_t s_co, int64_t s_ci,
                           int64_t s_ky, int64_t s_kx, int transpose_flip, void *out_bf16, void *stream) {
    UB_REQUIRE(w && out_bf16 && Cout > 0 && Cin > 0 && (ksize == 1 || ksize == 3), UB200_E_BADARG);
    const int64_t rows = transpose_flip ? Cin : Cout, cols = transpose_flip ? Cout : Cin;
    const int64_t rows_pad = (rows + 15) / 16 * 16;
    const int64_t total = rows_pad * ksize * ksize * cols;
    int grid = ub::grid_for(total, 256, 8);
    pack_weight_kernel<<<grid, 256, 0, ub::as_stream(stream)>>>(w, (int)Cout, (int)Cin, ksize, s_co, s_ci, s_ky, s_kx,
                                                               transpose_flip, reinterpret_cast<__nv_bfloat16 *>(out_bf16),
                                                               total);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // extern "C"
