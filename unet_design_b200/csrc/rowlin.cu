// Batched small fp32 "row" linears of the time-embedding path:  y_i = act(x_i) @ w_i^T + bias_i  for a LIST of
// (x, w, bias) items in ONE launch, and the matching backward (weight / bias gradients accumulated in place, input
// gradient summed over the items that share an input).
//
// Why: every ResBlock of the reference projects the time embedding with its own Swish + Linear(tdim, Cout)
// (diff_cifar/model.py:134-137, applied at :164) and every level has its own TimeEmbedding MLP (model.py:29-36).
// As separate framework ops that is ~500 launches of 2-3 us per training step for < 0.1 % of the FLOPs: 1.4 ms of an
// 8.6 ms step on B200.  Here the whole path is 3 forward and 6 backward launches.
//
// Plain fp32 FMA tiles (64 x 64 per CTA, 4 x 4 per thread, 16-deep slices through shared memory); the matrices are
// tiny (batch x 512 x Cout), so the tensor cores would buy nothing and fp32 keeps the reference's arithmetic.
// The item list travels BY VALUE in the kernel parameters: no device-side table, valid under CUDA-graph replay as
// long as the tensors keep their addresses.
#include "common.cuh"

namespace {
using namespace ub;

constexpr int kMaxItems = 24;
constexpr int kMaxGroups = 8;
constexpr int kTile = 64, kSlice = 16, kPad = 4, kThreads = 256;
constexpr int kAhead = 4;      // operand slices in flight per thread: one slice of FMAs (~0.15 us) does not cover an L2 round trip (~0.8 us)

struct Item {
    const float *x, *w, *bias, *gy;
    float *y, *gw, *gb;
    int cout, tile0, group;
};
struct Group { const float *x; float *gx; int accumulate; };
struct Batch {
    Item it[kMaxItems];
    Group gr[kMaxGroups];
    int n_items, n_groups, N, K, silu, tiles_total;
};

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
__device__ __forceinline__ float dsilu_f(float v) {
    const float s = 1.0f / (1.0f + expf(-v));
    return s * (1.0f + v * (1.0f - s));
}

// acc[4][4] += As[kk][ty*4 + i] * Bs[kk][tx*4 + j] over one 16-deep slice
__device__ __forceinline__ void fma_slice(const float (*As)[kTile + kPad], const float (*Bs)[kTile + kPad], int ty, int tx,
                                          float (&acc)[4][4]) {
#pragma unroll
    for (int kk = 0; kk < kSlice; ++kk) {
        const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// Operand fetches for one 16-deep slice, split into the global read (issued one slice ahead, so the L2 latency is
// covered by the FMAs of the current slice) and the shared-memory store.
// "T": 64 rows x 16 columns of a row-major matrix (contiguous along the 16), stored transposed as S[col][row];
// "D": 16 rows x 64 columns (contiguous along the 64), stored as S[row][col].
// Rows >= rows_valid and columns >= cols_valid read as zero.  ACT applies SiLU at the shared-memory store (not at the
// fetch, which would make the prefetched value a dependency of the current slice).
__device__ __forceinline__ float4 fetch_T(const float *src, int64_t ld, int row0, int rows_valid, int col0, int cols_valid) {
    const int r = threadIdx.x >> 2, c4 = (threadIdx.x & 3) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows_valid && col0 + c4 < cols_valid) v = __ldg(reinterpret_cast<const float4 *>(src + (int64_t)(row0 + r) * ld + col0 + c4));
    return v;
}
template <bool ACT>
__device__ __forceinline__ void put_T(float (*S)[kTile + kPad], float4 v) {
    const int r = threadIdx.x >> 2, c4 = (threadIdx.x & 3) * 4;
    if (ACT) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    S[c4][r] = v.x; S[c4 + 1][r] = v.y; S[c4 + 2][r] = v.z; S[c4 + 3][r] = v.w;
}
__device__ __forceinline__ float4 fetch_D(const float *src, int64_t ld, int row0, int rows_valid, int col0, int cols_valid) {
    const int r = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < rows_valid && col0 + c4 < cols_valid) v = __ldg(reinterpret_cast<const float4 *>(src + (int64_t)(row0 + r) * ld + col0 + c4));
    return v;
}
template <bool ACT>
__device__ __forceinline__ void put_D(float (*S)[kTile + kPad], float4 v) {
    const int r = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;
    if (ACT) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    *reinterpret_cast<float4 *>(&S[r][c4]) = v;
}

__device__ __forceinline__ int find_item(const Batch &b, int tile) {
    int i = 0;
    while (i + 1 < b.n_items && b.it[i + 1].tile0 <= tile) ++i;
    return i;
}

// y[n, c] = bias[c] + sum_k act(x[n, k]) w[c, k].   grid: (cout tiles of all items, batch tiles)
template <bool ACT>
__global__ void __launch_bounds__(kThreads) rowlin_fwd_kernel(const __grid_constant__ Batch b) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[2][kSlice][kTile + kPad], Bs[2][kSlice][kTile + kPad];
    const int ii = find_item(b, blockIdx.x);
    const Item &it = b.it[ii];
    const int c0 = (blockIdx.x - it.tile0) * kTile, n0 = blockIdx.y * kTile;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[4][4] = {};
    const int ns = (b.K + kSlice - 1) / kSlice;
    float4 ra[kAhead], rb[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
        ra[d] = fetch_T(it.x, b.K, n0, b.N, d * kSlice, b.K);                 // columns past K read as zero
        rb[d] = fetch_T(it.w, b.K, c0, it.cout, d * kSlice, b.K);
    }
    for (int s0 = 0; s0 < ns; s0 += kAhead) {
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            const int s = s0 + d;
            if (s < ns) {                                                       // uniform across the CTA
                put_T<ACT>(As[s & 1], ra[d]); put_T<false>(Bs[s & 1], rb[d]);
                __syncthreads();
                ra[d] = fetch_T(it.x, b.K, n0, b.N, (s + kAhead) * kSlice, b.K);
                rb[d] = fetch_T(it.w, b.K, c0, it.cout, (s + kAhead) * kSlice, b.K);
                fma_slice(As[s & 1], Bs[s & 1], ty, tx, acc);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= b.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + tx * 4 + j;
            if (c < it.cout) it.y[(int64_t)n * it.cout + c] = acc[i][j] + (it.bias ? __ldg(it.bias + c) : 0.f);
        }
    }
}

// gw[c, k] += sum_n gy[n, c] act(x[n, k]);  gb[c] += sum_n gy[n, c].   grid: (cout tiles of all items, K tiles)
template <bool ACT>
__global__ void __launch_bounds__(kThreads) rowlin_wgrad_kernel(const __grid_constant__ Batch b) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[2][kSlice][kTile + kPad], Bs[2][kSlice][kTile + kPad];
    const int ii = find_item(b, blockIdx.x);
    const Item &it = b.it[ii];
    const int c0 = (blockIdx.x - it.tile0) * kTile, k0 = blockIdx.y * kTile;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[4][4] = {};
    float colsum[4] = {0.f, 0.f, 0.f, 0.f};
    const int ns = (b.N + kSlice - 1) / kSlice;
    float4 ra[kAhead], rb[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
        ra[d] = fetch_D(it.gy, it.cout, d * kSlice, b.N, c0, it.cout);        // rows past N read as zero
        rb[d] = fetch_D(it.x, b.K, d * kSlice, b.N, k0, b.K);
    }
    for (int s0 = 0; s0 < ns; s0 += kAhead) {
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            const int s = s0 + d;
            if (s < ns) {
                put_D<false>(As[s & 1], ra[d]); put_D<ACT>(Bs[s & 1], rb[d]);
                __syncthreads();
                ra[d] = fetch_D(it.gy, it.cout, (s + kAhead) * kSlice, b.N, c0, it.cout);
                rb[d] = fetch_D(it.x, b.K, (s + kAhead) * kSlice, b.N, k0, b.K);
                fma_slice(As[s & 1], Bs[s & 1], ty, tx, acc);
                if (tx == 0 && blockIdx.y == 0) {
#pragma unroll
                    for (int kk = 0; kk < kSlice; ++kk)
#pragma unroll
                        for (int i = 0; i < 4; ++i) colsum[i] += As[s & 1][kk][ty * 4 + i];
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty * 4 + i;
        if (c >= it.cout) continue;
        if (it.gw) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + tx * 4 + j;
                if (k < b.K) it.gw[(int64_t)c * b.K + k] += acc[i][j];      // this CTA owns the element: plain read-modify-write
            }
        }
        if (it.gb && tx == 0 && blockIdx.y == 0) it.gb[c] += colsum[i];
    }
}

// gx[n, k] += act'(x[n, k]) * sum_c gy[n, c] w[c, k] for ONE item per CTA (fp32 atomics: the items that share an input
// run in parallel; gx is zeroed by the host function first).   grid: (K tiles, batch tiles, items)
template <bool ACT>
__global__ void __launch_bounds__(kThreads) rowlin_xgrad_kernel(const __grid_constant__ Batch b) {
    pdl_trigger();
    pdl_wait();
    __shared__ __align__(16) float As[2][kSlice][kTile + kPad], Bs[2][kSlice][kTile + kPad];
    const Item &it = b.it[blockIdx.z];
    if (it.group < 0) return;
    const Group &g = b.gr[it.group];
    const int k0 = blockIdx.x * kTile, n0 = blockIdx.y * kTile;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[4][4] = {};
    const int ns = (it.cout + kSlice - 1) / kSlice;
    float4 ra[kAhead], rb[kAhead];
#pragma unroll
    for (int d = 0; d < kAhead; ++d) {
        ra[d] = fetch_T(it.gy, it.cout, n0, b.N, d * kSlice, it.cout);
        rb[d] = fetch_D(it.w, b.K, d * kSlice, it.cout, k0, b.K);
    }
    for (int s0 = 0; s0 < ns; s0 += kAhead) {
#pragma unroll
        for (int d = 0; d < kAhead; ++d) {
            const int s = s0 + d;
            if (s < ns) {
                put_T<false>(As[s & 1], ra[d]); put_D<false>(Bs[s & 1], rb[d]);
                __syncthreads();
                ra[d] = fetch_T(it.gy, it.cout, n0, b.N, (s + kAhead) * kSlice, it.cout);
                rb[d] = fetch_D(it.w, b.K, (s + kAhead) * kSlice, it.cout, k0, b.K);
                fma_slice(As[s & 1], Bs[s & 1], ty, tx, acc);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= b.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k >= b.K) continue;
            const int64_t o = (int64_t)n * b.K + k;
            float v = acc[i][j];
            if (ACT) v *= dsilu_f(__ldg(g.x + o));
            atomicAdd(g.gx + o, v);
        }
    }
}

int fill_batch(Batch &b, const ub200_rowlin_item *items, int n, int64_t N, int64_t K, int silu) {
    b.n_items = n; b.N = (int)N; b.K = (int)K; b.silu = silu; b.n_groups = 0;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const ub200_rowlin_item &s = items[i];
        if (!s.x || !s.w || s.cout <= 0 || s.cout % 4 != 0 || s.cout > (1 << 20)) return UB200_E_BADARG;
        Item &d = b.it[i];
        d.x = s.x; d.w = s.w; d.bias = s.bias; d.gy = s.gy; d.y = s.y; d.gw = s.gw; d.gb = s.gbias;
        d.cout = (int)s.cout; d.tile0 = tiles; d.group = -1;
        tiles += (int)((s.cout + kTile - 1) / kTile);
    }
    b.tiles_total = tiles;
    return UB200_OK;
}

}  // namespace

extern "C" {

int ub200_rowlin_fwd(const ub200_rowlin_item *items, int n_items, int64_t N, int64_t K, int silu, void *stream) {
    UB_REQUIRE(items && n_items > 0 && N > 0 && K > 0, UB200_E_BADARG);
    UB_REQUIRE(K % 4 == 0 && N < (1 << 24) && K < (1 << 24), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    for (int i0 = 0; i0 < n_items; i0 += kMaxItems) {
        Batch b;
        const int n = n_items - i0 < kMaxItems ? n_items - i0 : kMaxItems;
        int rc = fill_batch(b, items + i0, n, N, K, silu);
        if (rc) return rc;
        for (int i = 0; i < n; ++i) UB_REQUIRE(items[i0 + i].y, UB200_E_BADARG);
        dim3 grid((unsigned)b.tiles_total, (unsigned)((N + kTile - 1) / kTile), 1);
        cudaError_t le = silu ? launch_pdl(rowlin_fwd_kernel<true>, grid, dim3(kThreads), 0, s, b)
                              : launch_pdl(rowlin_fwd_kernel<false>, grid, dim3(kThreads), 0, s, b);
        if (le != cudaSuccess) return (int)le;
        UB_LAUNCH_CHECK();
    }
    return UB200_OK;
}

int ub200_rowlin_bwd(const ub200_rowlin_item *items, int n_items, int64_t N, int64_t K, int silu, void *stream) {
    UB_REQUIRE(items && n_items > 0 && N > 0 && K > 0, UB200_E_BADARG);
    UB_REQUIRE(K % 4 == 0 && N < (1 << 24) && K < (1 << 24), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    // gx pointers already written by an earlier chunk of this call: later chunks accumulate into them
    float *seen[kMaxGroups * 8];
    int n_seen = 0;
    for (int i0 = 0; i0 < n_items; i0 += kMaxItems) {
        Batch b;
        const int n = n_items - i0 < kMaxItems ? n_items - i0 : kMaxItems;
        int rc = fill_batch(b, items + i0, n, N, K, silu);
        if (rc) return rc;
        bool any_w = false;
        for (int i = 0; i < n; ++i) {
            const ub200_rowlin_item &it = items[i0 + i];
            UB_REQUIRE(it.gy, UB200_E_BADARG);
            any_w = any_w || it.gw || it.gbias;
            if (!it.gx) continue;
            int g = 0;
            while (g < b.n_groups && b.gr[g].gx != it.gx) ++g;
            if (g == b.n_groups) {
                UB_REQUIRE(b.n_groups < kMaxGroups, UB200_E_UNSUPPORTED);
                bool was = false;
                for (int q = 0; q < n_seen; ++q) was = was || seen[q] == it.gx;
                b.gr[g].x = it.x; b.gr[g].gx = it.gx; b.gr[g].accumulate = was ? 1 : 0;
                if (!was) {
                    UB_REQUIRE(n_seen < kMaxGroups * 8, UB200_E_UNSUPPORTED);
                    seen[n_seen++] = it.gx;
                }
                ++b.n_groups;
            }
            UB_REQUIRE(b.gr[g].x == it.x, UB200_E_BADARG);      // one gx belongs to one x
            b.it[i].group = g;
        }
        if (any_w) {
            dim3 grid((unsigned)b.tiles_total, (unsigned)((K + kTile - 1) / kTile), 1);
            cudaError_t le = silu ? launch_pdl(rowlin_wgrad_kernel<true>, grid, dim3(kThreads), 0, s, b)
                                  : launch_pdl(rowlin_wgrad_kernel<false>, grid, dim3(kThreads), 0, s, b);
            if (le != cudaSuccess) return (int)le;
            UB_LAUNCH_CHECK();
        }
        if (b.n_groups > 0) {
            for (int g = 0; g < b.n_groups; ++g)
                if (!b.gr[g].accumulate) {
                    cudaError_t e = cudaMemsetAsync(b.gr[g].gx, 0, (size_t)N * K * sizeof(float), s);
                    if (e != cudaSuccess) return (int)e;
                }
            dim3 grid((unsigned)((K + kTile - 1) / kTile), (unsigned)((N + kTile - 1) / kTile), (unsigned)n);
            cudaError_t le = silu ? launch_pdl(rowlin_xgrad_kernel<true>, grid, dim3(kThreads), 0, s, b)
                                  : launch_pdl(rowlin_xgrad_kernel<false>, grid, dim3(kThreads), 0, s, b);
            if (le != cudaSuccess) return (int)le;
            UB_LAUNCH_CHECK();
        }
    }
    return UB200_OK;
}

}  // extern "C"
