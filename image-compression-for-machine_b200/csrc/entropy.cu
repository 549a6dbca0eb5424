// Entropy-model step on the device (rows E1-E6, E9 of SURVEY.md §8a): fused, HBM-bound elementwise
// kernels.  One pass reads (y, mu, scale) and writes everything the next stage needs -- int32 symbols and
// CDF indexes already in the coder's stream order, ŷ in fp32, and ŷ in bf16 straight into the channel
// slots of the context-model support buffers -- instead of the reference's ~70 elementwise launches per
// slice (63 compare/subtract passes in build_indexes alone, entropy_models.py:661-666).
//
// All tensors are strided [B, C, P] views so the same kernel serves NCHW (the reference's layout, used by
// the drop-in nn.Module API) and channels-last (the layout of the conv kernels).  A CTA owns a 32-channel
// x 32-pixel tile; every tensor is touched with its own unit-stride dimension on the fast thread index
// and layout changes happen through a padded shared-memory tile, so loads and stores are 128-byte
// coalesced on both sides.
#include "common.cuh"

#include <stdlib.h>

#include <type_traits>

namespace icm {

constexpr int TILE = 32;
constexpr int THREADS = 256;
constexpr int PER_THREAD = TILE * TILE / THREADS; // 4

struct TileCoord {
    int b, c0;
    long long p0;
    int nc, np; // valid extent of this tile
};

// element k (0..3) of thread t inside the tile, for a tensor whose fast dimension is channels / pixels
__device__ __forceinline__ void elem_cfast(int t, int k, int &ci, int &pi) { ci = t & 31; pi = (t >> 5) + 8 * k; }
__device__ __forceinline__ void elem_pfast(int t, int k, int &ci, int &pi) { pi = t & 31; ci = (t >> 5) + 8 * k; }

// A thread's four tile cells differ by 8 in the slow index, so its addresses are one base (the only 64-bit multiplies)
// plus k * step.  Recomputing (c0 + ci) * sc + (p0 + pi) * sp per cell and per tensor cost ~260 instructions per
// element and made these kernels issue-bound at a fifth of the HBM roofline (ncu: IPC 2.98, 15 % DRAM utilisation).
struct CellWalk {
    int ci0, pi0, dci, dpi; // cell k = (ci0 + k * dci, pi0 + k * dpi)
    long long first, step;  // element offset of cell 0 and between cells
};

__device__ __forceinline__ CellWalk cell_walk(const View &v, const TileCoord &tc)
{
    CellWalk w;
    const int t = threadIdx.x;
    if (v.sc == 1) { w.ci0 = t & 31; w.pi0 = t >> 5; w.dci = 0; w.dpi = 8; w.step = 8 * v.sp; }  // channels fastest
    else           { w.pi0 = t & 31; w.ci0 = t >> 5; w.dci = 8; w.dpi = 0; w.step = 8 * v.sc; }  // pixels fastest
    w.first = (long long)tc.b * v.sb + (long long)(tc.c0 + w.ci0) * v.sc + (tc.p0 + w.pi0) * v.sp;
    return w;
}

template <typename T>
__device__ __forceinline__ void load_tile(const View &v, const TileCoord &tc, float (*tile)[TILE + 1])
{
    if (!v.ptr) { // absent optional input (e.g. no means): reads as zero
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) { int ci, pi; elem_pfast(threadIdx.x, k, ci, pi); tile[ci][pi] = 0.f; }
        return;
    }
    const CellWalk w = cell_walk(v, tc);
    const T *src = reinterpret_cast<const T *>(v.ptr) + w.first;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k, src += w.step) {
        const int ci = w.ci0 + k * w.dci, pi = w.pi0 + k * w.dpi;
        if (ci < tc.nc && pi < tc.np) {
            const T val = *src;
            if constexpr (std::is_same<T, int32_t>::value) tile[ci][pi] = __int_as_float(val);
            else tile[ci][pi] = (float)val;
        }
    }
}

template <typename T>
__device__ __forceinline__ void store_tile(const View &v, const TileCoord &tc, float (*tile)[TILE + 1])
{
    if (!v.ptr) return;
    const CellWalk w = cell_walk(v, tc);
    T *dst = reinterpret_cast<T *>(v.ptr) + w.first;
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k, dst += w.step) {
        const int ci = w.ci0 + k * w.dci, pi = w.pi0 + k * w.dpi;
        if (ci < tc.nc && pi < tc.np) {
            const float f = tile[ci][pi];
            if constexpr (std::is_same<T, int32_t>::value) *dst = __float_as_int(f);
            else if constexpr (std::is_same<T, __nv_bfloat16>::value) *dst = __float2bfloat16_rn(f);
            else *dst = f;
        }
    }
}

__device__ __forceinline__ TileCoord tile_coord(int C, long long P)
{
    TileCoord tc;
    tc.p0 = (long long)blockIdx.x * TILE;
    tc.c0 = blockIdx.y * TILE;
    tc.b = blockIdx.z;
    tc.nc = min(TILE, C - tc.c0);
    tc.np = (int)min((long long)TILE, P - tc.p0);
    return tc;
}

// idx = (n-1) - #{j < n-1 : s <= table[j]} == first j in [0, n-1) with s <= table[j], else n-1
// (entropy_models.py:661-666); a NaN scale compares false everywhere and lands on n-1 like the reference.
__device__ __forceinline__ int bucket_index(float s, const float *table, int n)
{
    int lo = 0, hi = n - 1; // answer in [lo, hi]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s <= table[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Two-probe form of bucket_index for the vectorised kernels.  A CTA first builds, in shared memory, a table over the float
// bit pattern of s (exponent + 4 mantissa bits = 16 bins per octave): lut[k] = bucket_index(lower edge of bin k).  When no
// bin holds more than one table value (the reference's geometric scale table spaces its 64 levels by a factor 1.131; a
// bin spans 1.044-1.0625) the answer is lut[k] or lut[k] + 1, decided by ONE compare against table[lut[k]] -- 8 instructions
// instead of a 6-step dependent binary search (~40), which made build_indexes issue-bound at 0.3 of the HBM roofline.
// Any other table (too many bins, or two values in one bin) clears `ok` and the kernels keep the binary search.
constexpr int LUT_SHIFT = 19;  // 23 mantissa bits - 4
constexpr int LUT_MAX = 256;

struct IndexLut {
    int key0, nkeys, ok;
};

__device__ __forceinline__ IndexLut build_index_lut(const float *s_table, int n, int *s_lut, int *s_flag)
{
    IndexLut L;
    // bins from the one holding table[0] to the one holding table[n-2] (the last value bucket_index compares with)
    const int last = n >= 2 ? n - 2 : 0;
    L.key0 = (int)(__float_as_uint(s_table[0]) >> LUT_SHIFT);
    L.nkeys = (int)(__float_as_uint(s_table[last]) >> LUT_SHIFT) - L.key0 + 1;
    const bool sane = n >= 2 && s_table[0] > 0.f && L.nkeys >= 1 && L.nkeys <= LUT_MAX;
    if (threadIdx.x == 0) *s_flag = sane ? 1 : 0;
    __syncthreads();
    if (sane) {
        for (int k = threadIdx.x; k < L.nkeys; k += blockDim.x) {
            const float lo = __uint_as_float((uint32_t)(L.key0 + k) << LUT_SHIFT), hi = __uint_as_float((uint32_t)(L.key0 + k + 1) << LUT_SHIFT);
            const int cand = bucket_index(lo, s_table, n);
            s_lut[k] = cand;
            // a second table value below the bin's upper edge would need a second compare; so would an unsorted table
            if (cand + 1 <= last && !(s_table[cand + 1] >= hi)) *s_flag = 0;
            if (cand <= last && cand > 0 && !(s_table[cand - 1] < lo)) *s_flag = 0;
        }
    }
    __syncthreads();
    L.ok = *s_flag;
    return L;
}

__device__ __forceinline__ int bucket_index_fast(float s, const float *s_table, int n, const int *s_lut, const IndexLut &L)
{
    if (!L.ok) return bucket_index(s, s_table, n);
    int k = (int)(__float_as_uint(s) >> LUT_SHIFT) - L.key0; // negative / NaN patterns clamp to the ends
    k = min(max(k, 0), L.nkeys - 1);
    if (s < s_table[0]) k = 0; // negative values have large bit patterns (false for NaN, which must land on n - 1)
    const int cand = s_lut[k];
    return (cand < n - 1 && !(s <= s_table[cand])) ? cand + 1 : cand;
}

__device__ __forceinline__ float lower_bound_f(float x, float b)
{ // torch.max(x, bound): NaN propagates
    return (x != x) ? x : fmaxf(x, b);
}

__device__ __forceinline__ float std_cumulative(float v)
{ // entropy_models.py:578-582
    return 0.5f * erfcf(-0.70710678118654752440f * v);
}

enum GcMode { GC_QUANT = 0, GC_INDEX = 1, GC_DEQUANT = 2, GC_LIK = 3, GC_ADD = 4 };

struct GcArgs {
    View y, mu, scale;          // fp32 inputs (GC_DEQUANT: y is the int32 symbol view; GC_ADD: y = lrp)
    View sym, idx;              // int32 outputs
    View yhat, lik, bf_a, bf_b; // fp32, fp32, bf16, bf16 outputs (GC_ADD: yhat is read-modify-write)
    const float *table;
    int n_levels;
    float scale_bound, lik_bound;
    int C;
    long long P;
};

template <int MODE>
__global__ void __launch_bounds__(THREADS) gc_kernel(GcArgs a)
{
    __shared__ float t0[TILE][TILE + 1], t1[TILE][TILE + 1], t2[TILE][TILE + 1];
    __shared__ float s_table[256];
    const TileCoord tc = tile_coord(a.C, a.P);
    if ((MODE == GC_QUANT || MODE == GC_INDEX) && a.table)
        for (int i = threadIdx.x; i < a.n_levels; i += THREADS) s_table[i] = a.table[i];
    if (MODE == GC_QUANT || MODE == GC_LIK) { load_tile<float>(a.y, tc, t0); load_tile<float>(a.mu, tc, t1); load_tile<float>(a.scale, tc, t2); }
    if (MODE == GC_INDEX) load_tile<float>(a.scale, tc, t2);
    if (MODE == GC_DEQUANT) { load_tile<int32_t>(a.y, tc, t0); load_tile<float>(a.mu, tc, t1); }
    if (MODE == GC_ADD) { load_tile<float>(a.yhat, tc, t0); load_tile<float>(a.y, tc, t1); }
    __syncthreads();
    // every thread transforms the 4 tile cells it owns (cell ownership is arbitrary once in smem)
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        int ci, pi;
        elem_pfast(threadIdx.x, k, ci, pi);
        if (MODE == GC_QUANT) {
            const float y = t0[ci][pi], mu = t1[ci][pi], sc = t2[ci][pi];
            const float r = rintf(y - mu); // torch.round: half to even
            const int q = (int)r;
            const int id = a.table ? bucket_index(lower_bound_f(sc, a.scale_bound), s_table, a.n_levels) : 0;
            t0[ci][pi] = __int_as_float(q);
            t2[ci][pi] = __int_as_float(id);
            t1[ci][pi] = (float)q + mu; // y_q_slice + mu (stf.py:716)
        } else if (MODE == GC_INDEX) {
            t2[ci][pi] = __int_as_float(bucket_index(lower_bound_f(t2[ci][pi], a.scale_bound), s_table, a.n_levels));
        } else if (MODE == GC_DEQUANT) {
            t1[ci][pi] = (float)__float_as_int(t0[ci][pi]) + t1[ci][pi];
        } else if (MODE == GC_LIK) {
            const float y = t0[ci][pi], mu = t1[ci][pi];
            const float s = lower_bound_f(t2[ci][pi], a.scale_bound);
            const float yh = rintf(y - mu) + mu;         // quantize(..., "dequantize", means)
            const float v = fabsf(yh - mu);              // values = |outputs - means|
            const float up = std_cumulative((0.5f - v) / s);
            const float lo = std_cumulative((-0.5f - v) / s);
            t1[ci][pi] = yh;
            t0[ci][pi] = lower_bound_f(up - lo, a.lik_bound);
        } else if (MODE == GC_ADD) {
            t1[ci][pi] = t0[ci][pi] + t1[ci][pi];
        }
    }
    __syncthreads();
    if (MODE == GC_QUANT) { store_tile<int32_t>(a.sym, tc, t0); store_tile<int32_t>(a.idx, tc, t2); }
    if (MODE == GC_INDEX) store_tile<int32_t>(a.idx, tc, t2);
    if (MODE == GC_LIK) store_tile<float>(a.lik, tc, t0);
    store_tile<float>(a.yhat, tc, t1);
    store_tile<__nv_bfloat16>(a.bf_a, tc, t1);
    store_tile<__nv_bfloat16>(a.bf_b, tc, t1);
}

// ------------------------------------------------------------------------------------------------
// The codec's own layouts: activations channels-last (channel stride 1), symbols / indexes in stream order (pixel
// stride 1).  Inputs and y_hat outputs then share a layout, so each thread loads its cells straight into registers
// (lane = channel: 128-byte rows), computes, and stores y_hat directly; only the int32 symbol / index planes go through
// a shared-memory transpose.  The strided-view kernel above stages every tensor through shared memory and carries both
// layouts' code (~1 000 SASS instructions per thread of four cells: issue-bound at a fifth of the HBM roofline).
__device__ __forceinline__ bool chan_fast(const View &v) { return !v.ptr || v.sc == 1; }
__device__ __forceinline__ bool pix_fast(const View &v) { return !v.ptr || v.sp == 1; }

template <int MODE>
__global__ void __launch_bounds__(THREADS) gc_cl_kernel(GcArgs a)
{
    __shared__ float t0[TILE][TILE + 1], t2[TILE][TILE + 1];
    __shared__ float s_table[256];
    const TileCoord tc = tile_coord(a.C, a.P);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if ((MODE == GC_QUANT || MODE == GC_INDEX) && a.table)
        for (int i = threadIdx.x; i < a.n_levels; i += THREADS) s_table[i] = a.table[i];
    // channels-last cell k of this thread: channel `lane`, pixel w + 8k (channel stride 1); stream-order cell k: pixel
    // `lane`, channel w + 8k (pixel stride 1).  One 64-bit base per view, then k * step.
    auto cl_ptr = [&](const View &v, int esize) -> char * { return v.ptr ? v.ptr + ((long long)tc.b * v.sb + (tc.c0 + lane) + (tc.p0 + w) * v.sp) * esize : nullptr; };
    auto so_ptr = [&](const View &v) -> char * { return v.ptr ? v.ptr + ((long long)tc.b * v.sb + (long long)(tc.c0 + w) * v.sc + (tc.p0 + lane)) * 4 : nullptr; };
    const char *py = (MODE == GC_DEQUANT) ? so_ptr(a.y) : cl_ptr(a.y, 4), *pmu = cl_ptr(a.mu, 4), *psc = cl_ptr(a.scale, 4);
    char *pyh = cl_ptr(a.yhat, 4), *pba = cl_ptr(a.bf_a, 2), *pbb = cl_ptr(a.bf_b, 2), *psym = so_ptr(a.sym), *pidx = so_ptr(a.idx);
    const long long sy = 8 * ((MODE == GC_DEQUANT) ? a.y.sc : a.y.sp) * 4, smu = 8 * a.mu.sp * 4, ssc = 8 * a.scale.sp * 4, syh = 8 * a.yhat.sp * 4,
                    sba = 8 * a.bf_a.sp * 2, sbb = 8 * a.bf_b.sp * 2, ssym = 8 * a.sym.sc * 4, sidx = 8 * a.idx.sc * 4;
    auto put_yhat = [&](int k, float yh) {
        if (pyh) *reinterpret_cast<float *>(pyh + k * syh) = yh;
        if (pba) *reinterpret_cast<__nv_bfloat16 *>(pba + k * sba) = __float2bfloat16_rn(yh);
        if (pbb) *reinterpret_cast<__nv_bfloat16 *>(pbb + k * sbb) = __float2bfloat16_rn(yh);
    };
    const bool c_ok = lane < tc.nc;
    if (MODE == GC_DEQUANT) { // symbols arrive in stream order: transpose them first
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k)
            if (w + 8 * k < tc.nc && lane < tc.np) t0[w + 8 * k][lane] = __int_as_float(*reinterpret_cast<const int32_t *>(py + k * sy));
        __syncthreads();
    } else if (MODE == GC_QUANT || MODE == GC_INDEX) {
        __syncthreads(); // s_table
    }
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        const int pi = w + 8 * k;
        if (!(c_ok && pi < tc.np)) continue;
        if (MODE == GC_QUANT) {
            const float y = *reinterpret_cast<const float *>(py + k * sy);
            const float mu = pmu ? *reinterpret_cast<const float *>(pmu + k * smu) : 0.f;
            const float r = rintf(y - mu); // torch.round: half to even
            const int q = (int)r;
            int id = 0;
            if (a.table) id = bucket_index(lower_bound_f(*reinterpret_cast<const float *>(psc + k * ssc), a.scale_bound), s_table, a.n_levels);
            t0[lane][pi] = __int_as_float(q);
            t2[lane][pi] = __int_as_float(id);
            put_yhat(k, (float)q + mu); // y_q_slice + mu (stf.py:716)
        } else if (MODE == GC_INDEX) {
            t2[lane][pi] = __int_as_float(bucket_index(lower_bound_f(*reinterpret_cast<const float *>(psc + k * ssc), a.scale_bound), s_table, a.n_levels));
        } else if (MODE == GC_DEQUANT) {
            const float mu = pmu ? *reinterpret_cast<const float *>(pmu + k * smu) : 0.f;
            put_yhat(k, (float)__float_as_int(t0[lane][pi]) + mu);
        } else if (MODE == GC_ADD) {
            put_yhat(k, *reinterpret_cast<const float *>(pyh + k * syh) + *reinterpret_cast<const float *>(py + k * sy));
        }
    }
    if (MODE == GC_QUANT || MODE == GC_INDEX) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const int ci = w + 8 * k;
            if (ci < tc.nc && lane < tc.np) {
                if (MODE == GC_QUANT && psym) *reinterpret_cast<int32_t *>(psym + k * ssym) = __float_as_int(t0[ci][lane]);
                if (pidx) *reinterpret_cast<int32_t *>(pidx + k * sidx) = __float_as_int(t2[ci][lane]);
            }
        }
    }
}

// Vectorised form of the codec-layout kernel (round 2): a thread owns FOUR consecutive channels of one pixel, so every
// channels-last tensor moves as one 16-byte (fp32) or 8-byte (bf16) access per thread -- 3 loads and 2-3 stores per four
// elements instead of 12 and 8-12 scalar ones, and one 64-bit address per tensor instead of four.  ncu on the scalar form:
// 169 instructions per element, SM throughput 65 %, DRAM 40 %: issue-bound, not HBM-bound.  A warp covers 4 pixels x 32
// channels (lane = 8 * pixel + channel group): four full 128-byte rows per load instruction.  The int32 symbol / index
// planes still pass through the padded shared-memory tile; with this mapping both its writes (bank = 4 cg + pl + const) and
// its row reads are conflict-free.  Requirements beyond codec_layout(): full 32-channel tiles, 16-byte aligned rows.
// pixel tiles per CTA: only build_indexes (8 B per element) gains from amortising the table / lookup set-up over several tiles;
// the heavier modes lose memory-level parallelism (quantise x6: 85 -> 102 us with 4)
template <int MODE> struct V4Tiles { static constexpr int value = MODE == 1 /* GC_INDEX */ ? 8 : 1; };

template <int MODE>
__global__ void __launch_bounds__(THREADS) gc_v4_kernel(GcArgs a)
{
    __shared__ float t0[TILE][TILE + 1], t2[TILE][TILE + 1];
    __shared__ float s_table[256];
    TileCoord tc = tile_coord(a.C, a.P);
    constexpr int V4_TILES = V4Tiles<MODE>::value;
    tc.p0 *= V4_TILES; // a CTA walks V4_TILES consecutive pixel tiles: the scale table and the index lookup are set up once
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cg = lane & 7, pi = 4 * w + (lane >> 3); // channel group (4 channels) and pixel of this thread inside the tile
    __shared__ int s_lut[LUT_MAX];
    __shared__ int s_flag;
    IndexLut lut{0, 0, 0};
    if ((MODE == GC_QUANT || MODE == GC_INDEX) && a.table) {
        for (int i = threadIdx.x; i < a.n_levels; i += THREADS) s_table[i] = a.table[i];
        __syncthreads();
        lut = build_index_lut(s_table, a.n_levels, s_lut, &s_flag);
    }
  for (int it = 0; it < V4_TILES && tc.p0 < a.P; ++it, tc.p0 += TILE) {
    tc.np = (int)min((long long)TILE, a.P - tc.p0);
    auto cl_ptr = [&](const View &v, int esize) -> char * { return v.ptr ? v.ptr + ((long long)tc.b * v.sb + (tc.c0 + 4 * cg) + (tc.p0 + pi) * v.sp) * esize : nullptr; };
    auto so_ptr = [&](const View &v) -> char * { return v.ptr ? v.ptr + ((long long)tc.b * v.sb + (long long)(tc.c0 + w) * v.sc + (tc.p0 + lane)) * 4 : nullptr; };
    const bool p_ok = pi < tc.np;
    if (MODE == GC_DEQUANT) { // symbols arrive in stream order: rows of 32 pixels per channel -> tile
        const char *ps = so_ptr(a.y);
        const long long ss = 8 * a.y.sc * 4;
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k)
            if (lane < tc.np) t0[w + 8 * k][lane] = __int_as_float(*reinterpret_cast<const int32_t *>(ps + k * ss));
    }
    __syncthreads(); // s_table / symbol tile
    if (p_ok) {
        float4 yh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == GC_QUANT) {
            const float4 y = *reinterpret_cast<const float4 *>(cl_ptr(a.y, 4));
            const float4 mu = a.mu.ptr ? *reinterpret_cast<const float4 *>(cl_ptr(a.mu, 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float ya[4] = {y.x, y.y, y.z, y.w}, ma[4] = {mu.x, mu.y, mu.z, mu.w};
            float sa[4] = {0.f, 0.f, 0.f, 0.f};
            if (a.table) { const float4 sc = *reinterpret_cast<const float4 *>(cl_ptr(a.scale, 4)); sa[0] = sc.x; sa[1] = sc.y; sa[2] = sc.z; sa[3] = sc.w; }
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int q = (int)rintf(ya[j] - ma[j]); // torch.round: half to even
                const int id = a.table ? bucket_index_fast(lower_bound_f(sa[j], a.scale_bound), s_table, a.n_levels, s_lut, lut) : 0;
                t0[4 * cg + j][pi] = __int_as_float(q);
                t2[4 * cg + j][pi] = __int_as_float(id);
                o[j] = (float)q + ma[j]; // y_q_slice + mu (stf.py:716)
            }
            yh = make_float4(o[0], o[1], o[2], o[3]);
        } else if (MODE == GC_INDEX) {
            const float4 sc = *reinterpret_cast<const float4 *>(cl_ptr(a.scale, 4));
            const float sa[4] = {sc.x, sc.y, sc.z, sc.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) t2[4 * cg + j][pi] = __int_as_float(bucket_index_fast(lower_bound_f(sa[j], a.scale_bound), s_table, a.n_levels, s_lut, lut));
        } else if (MODE == GC_DEQUANT) {
            const float4 mu = a.mu.ptr ? *reinterpret_cast<const float4 *>(cl_ptr(a.mu, 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            yh = make_float4((float)__float_as_int(t0[4 * cg][pi]) + mu.x, (float)__float_as_int(t0[4 * cg + 1][pi]) + mu.y,
                             (float)__float_as_int(t0[4 * cg + 2][pi]) + mu.z, (float)__float_as_int(t0[4 * cg + 3][pi]) + mu.w);
        } else if (MODE == GC_ADD) {
            const float4 h = *reinterpret_cast<const float4 *>(cl_ptr(a.yhat, 4)), l = *reinterpret_cast<const float4 *>(cl_ptr(a.y, 4));
            yh = make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
        }
        if (MODE != GC_INDEX) {
            if (a.yhat.ptr) *reinterpret_cast<float4 *>(cl_ptr(a.yhat, 4)) = yh;
            if (a.bf_a.ptr || a.bf_b.ptr) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(yh.x, yh.y), hi = __floats2bfloat162_rn(yh.z, yh.w);
                const uint2 pk = make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
                if (a.bf_a.ptr) *reinterpret_cast<uint2 *>(cl_ptr(a.bf_a, 2)) = pk;
                if (a.bf_b.ptr) *reinterpret_cast<uint2 *>(cl_ptr(a.bf_b, 2)) = pk;
            }
        }
    }
    if (MODE == GC_QUANT || MODE == GC_INDEX) {
        __syncthreads();
        char *psym = so_ptr(a.sym), *pidx = so_ptr(a.idx);
        const long long ssym = 8 * a.sym.sc * 4, sidx = 8 * a.idx.sc * 4;
        if (lane < tc.np) {
#pragma unroll
            for (int k = 0; k < PER_THREAD; ++k) {
                if (MODE == GC_QUANT && psym) *reinterpret_cast<int32_t *>(psym + k * ssym) = __float_as_int(t0[w + 8 * k][lane]);
                if (pidx) *reinterpret_cast<int32_t *>(pidx + k * sidx) = __float_as_int(t2[w + 8 * k][lane]);
            }
        }
    }
    __syncthreads(); // the tiles are rewritten by the next iteration
  }
}

// 16-byte (fp32) / 8-byte (bf16) access of four channels: base, batch stride, pixel stride and channel offset aligned
static bool vec4_view(const View &v, int esize)
{
    if (!v.ptr) return true;
    return ((uintptr_t)v.ptr % (esize == 4 ? 16 : 8) == 0) && (v.sp % 4 == 0) && (v.sb % 4 == 0);
}

template <int MODE>
static bool vec4_layout(const GcArgs &a)
{
    if (a.C % TILE) return false;
    const bool f32s = vec4_view(a.mu, 4) && vec4_view(a.yhat, 4) && vec4_view(a.bf_a, 2) && vec4_view(a.bf_b, 2);
    if (MODE == GC_DEQUANT) return f32s;                       // a.y is the int32 symbol view (stream order)
    if (MODE == GC_INDEX) return vec4_view(a.scale, 4);
    if (MODE == GC_ADD) return f32s && vec4_view(a.y, 4);
    return f32s && vec4_view(a.y, 4) && vec4_view(a.scale, 4); // GC_QUANT
}

template <int MODE>
static bool codec_layout(const GcArgs &a)
{
    if (MODE == GC_LIK) return false;
    if (MODE == GC_DEQUANT) return a.y.ptr && a.y.sp == 1 && (!a.mu.ptr || a.mu.sc == 1) && (!a.yhat.ptr || a.yhat.sc == 1) && (!a.bf_a.ptr || a.bf_a.sc == 1) && (!a.bf_b.ptr || a.bf_b.sc == 1);
    if (MODE == GC_ADD) return a.y.ptr && a.y.sc == 1 && a.yhat.ptr && a.yhat.sc == 1 && (!a.bf_a.ptr || a.bf_a.sc == 1) && (!a.bf_b.ptr || a.bf_b.sc == 1);
    const bool outs = (!a.sym.ptr || a.sym.sp == 1) && (!a.idx.ptr || a.idx.sp == 1) && (!a.yhat.ptr || a.yhat.sc == 1) && (!a.bf_a.ptr || a.bf_a.sc == 1) && (!a.bf_b.ptr || a.bf_b.sc == 1);
    if (MODE == GC_INDEX) return outs && a.scale.ptr && a.scale.sc == 1;
    return outs && a.y.ptr && a.y.sc == 1 && (!a.mu.ptr || a.mu.sc == 1) && (!a.table || (a.scale.ptr && a.scale.sc == 1)); // GC_QUANT
}

template <int MODE>
static int launch_gc(const GcArgs &a, int B, void *stream)
{
    ICM_CHECK_ARG(B > 0 && a.C > 0 && a.P > 0, "entropy kernel: empty tensor (B=%d C=%d P=%lld)", B, a.C, a.P);
    ICM_CHECK_ARG(B <= 65535 && (a.C + TILE - 1) / TILE <= 65535, "entropy kernel: B or C too large");
    dim3 grid((unsigned)((a.P + TILE - 1) / TILE), (a.C + TILE - 1) / TILE, B);
    dim3 grid4((unsigned)((a.P + TILE * V4Tiles<MODE>::value - 1) / (TILE * V4Tiles<MODE>::value)), (a.C + TILE - 1) / TILE, B);
    static const bool force_generic = getenv("ICM_GC_GENERIC") != nullptr; // A/B switches for tools/gc_one.py
    static const bool force_scalar = getenv("ICM_GC_SCALAR") != nullptr;
    if (MODE != GC_LIK && !force_generic && !force_scalar && codec_layout<MODE>(a) && vec4_layout<MODE>(a)) gc_v4_kernel<MODE><<<grid4, THREADS, 0, as_stream(stream)>>>(a);
    else if (MODE != GC_LIK && !force_generic && codec_layout<MODE>(a)) gc_cl_kernel<MODE><<<grid, THREADS, 0, as_stream(stream)>>>(a);
    else gc_kernel<MODE><<<grid, THREADS, 0, as_stream(stream)>>>(a);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

static View stream_view(int32_t *p, long long stream_stride, long long stream_offset, long long P)
{
    return View{(char *)(p ? p + stream_offset : nullptr), stream_stride, P, 1};
}

// ------------------------------------------------------------------------------------------------
// EntropyBottleneck: per-channel 1-3-3-3-3-1 scalar network (entropy_models.py:400-433)
constexpr int EBP = ICM_EB_PARAMS_PER_CHANNEL;

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// p: transformed parameters of one channel (softplus(matrix), bias, tanh(factor))
__device__ __forceinline__ float eb_logits(const float *p, float v)
{
    float h[3], g[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        h[j] = p[j] * v + p[3 + j];
        h[j] += p[6 + j] * tanhf(h[j]);
    }
    const float *q = p + 9;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float acc = q[j * 3 + 0] * h[0];
            acc += q[j * 3 + 1] * h[1];
            acc += q[j * 3 + 2] * h[2];
            acc += q[9 + j];
            g[j] = acc + q[12 + j] * tanhf(acc);
        }
        h[0] = g[0]; h[1] = g[1]; h[2] = g[2];
        q += 15;
    }
    float out = q[0] * h[0];
    out += q[1] * h[1];
    out += q[2] * h[2];
    return out + q[3];
}

struct EbArgs {
    View z, sym, idx, zhat, zhat_bf, lik;
    const float *params;
    float lik_bound;
    int C;
    long long P;
};

template <int MODE>
__global__ void __launch_bounds__(THREADS) eb_kernel(EbArgs a)
{
    __shared__ float t0[TILE][TILE + 1], t1[TILE][TILE + 1];
    __shared__ float s_par[TILE][EBP + 1];
    const TileCoord tc = tile_coord(a.C, a.P);
    for (int i = threadIdx.x; i < tc.nc * EBP; i += THREADS) {
        const int c = i / EBP, k = i % EBP;
        float v = a.params[(size_t)(tc.c0 + c) * EBP + k];
        // positions of matrices / factors inside the 59-float record
        const bool is_matrix = (k < 3) || (k >= 9 && k < 18) || (k >= 24 && k < 33) || (k >= 39 && k < 48) || (k >= 54 && k < 57);
        const bool is_factor = (k >= 6 && k < 9) || (k >= 21 && k < 24) || (k >= 36 && k < 39) || (k >= 51 && k < 54);
        if (MODE == 1) { if (is_matrix) v = softplus_f(v); else if (is_factor) v = tanhf(v); }
        s_par[c][k] = v;
    }
    if (MODE == 2) load_tile<int32_t>(a.sym, tc, t0); else load_tile<float>(a.z, tc, t0);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER_THREAD; ++k) {
        int ci, pi;
        elem_pfast(threadIdx.x, k, ci, pi);
        if (ci >= tc.nc) continue;
        const float med = s_par[ci][EBP - 1];
        if (MODE == 0) {
            const int q = (int)rintf(t0[ci][pi] - med);
            t0[ci][pi] = __int_as_float(q);
            t1[ci][pi] = (float)q + med;
        } else if (MODE == 2) {
            t1[ci][pi] = (float)__float_as_int(t0[ci][pi]) + med;
        } else {
            const float zh = rintf(t0[ci][pi] - med) + med;
            const float lo = eb_logits(s_par[ci], zh - 0.5f);
            const float up = eb_logits(s_par[ci], zh + 0.5f);
            const float sum = lo + up;
            const float sign = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f); // -torch.sign(lo + up)
            const float l = fabsf(sigmoid_f(sign * up) - sigmoid_f(sign * lo));
            t1[ci][pi] = zh;
            t0[ci][pi] = lower_bound_f(l, a.lik_bound);
        }
    }
    __syncthreads();
    if (MODE == 0) {
        store_tile<int32_t>(a.sym, tc, t0);
        if (a.idx.ptr) { // index = channel id (entropy_models.py:492-502)
            __syncthreads();
#pragma unroll
            for (int k = 0; k < PER_THREAD; ++k) {
                int ci, pi;
                elem_pfast(threadIdx.x, k, ci, pi);
                t0[ci][pi] = __int_as_float(tc.c0 + ci);
            }
            __syncthreads();
            store_tile<int32_t>(a.idx, tc, t0);
        }
    }
    if (MODE == 1) store_tile<float>(a.lik, tc, t0);
    store_tile<float>(a.zhat, tc, t1);
    store_tile<__nv_bfloat16>(a.zhat_bf, tc, t1);
}

}  // namespace icm

using namespace icm;

extern "C" int icm_gc_quantize_index(icm_view y, icm_view mu, icm_view scale, int B, int C, int64_t P,
                                     const float *d_scale_table, int n_levels, float scale_bound,
                                     int32_t *d_symbols, int32_t *d_indexes, int64_t stream_stride, int64_t stream_offset,
                                     icm_view y_hat_f32, icm_view y_hat_bf16_a, icm_view y_hat_bf16_b, void *stream)
{
    ICM_CHECK_ARG(y.ptr && d_symbols, "icm_gc_quantize_index: null argument");
    ICM_CHECK_ARG(!scale.ptr == !d_indexes, "icm_gc_quantize_index: scale and indexes go together");
    ICM_CHECK_ARG(!scale.ptr || (d_scale_table && n_levels >= 1 && n_levels <= 256), "icm_gc_quantize_index: n_levels=%d outside [1,256]", n_levels);
    GcArgs a{};
    a.y = as_view(y); a.mu = as_view(mu); a.scale = as_view(scale);
    a.sym = stream_view(d_symbols, stream_stride, stream_offset, P);
    a.idx = stream_view(d_indexes, stream_stride, stream_offset, P);
    a.yhat = as_view(y_hat_f32); a.bf_a = as_view(y_hat_bf16_a); a.bf_b = as_view(y_hat_bf16_b);
    a.table = d_scale_table; a.n_levels = n_levels; a.scale_bound = scale_bound; a.C = C; a.P = P;
    return launch_gc<GC_QUANT>(a, B, stream);
}

extern "C" int icm_gc_build_indexes(icm_view scale, int B, int C, int64_t P, const float *d_scale_table, int n_levels,
                                    float scale_bound, int32_t *d_indexes, int64_t stream_stride, int64_t stream_offset,
                                    void *stream)
{
    ICM_CHECK_ARG(scale.ptr && d_scale_table && d_indexes, "icm_gc_build_indexes: null argument");
    ICM_CHECK_ARG(n_levels >= 1 && n_levels <= 256, "icm_gc_build_indexes: n_levels=%d outside [1,256]", n_levels);
    GcArgs a{};
    a.scale = as_view(scale);
    a.idx = stream_view(d_indexes, stream_stride, stream_offset, P);
    a.table = d_scale_table; a.n_levels = n_levels; a.scale_bound = scale_bound; a.C = C; a.P = P;
    return launch_gc<GC_INDEX>(a, B, stream);
}

extern "C" int icm_gc_dequantize(const int32_t *d_symbols, int64_t stream_stride, int64_t stream_offset, icm_view mu,
                                 int B, int C, int64_t P, icm_view y_hat_f32, icm_view y_hat_bf16_a,
                                 icm_view y_hat_bf16_b, void *stream)
{
    ICM_CHECK_ARG(d_symbols, "icm_gc_dequantize: null argument");
    GcArgs a{};
    a.y = stream_view(const_cast<int32_t *>(d_symbols), stream_stride, stream_offset, P);
    a.mu = as_view(mu);
    a.yhat = as_view(y_hat_f32); a.bf_a = as_view(y_hat_bf16_a); a.bf_b = as_view(y_hat_bf16_b);
    a.C = C; a.P = P;
    return launch_gc<GC_DEQUANT>(a, B, stream);
}

extern "C" int icm_gc_likelihood(icm_view y, icm_view mu, icm_view scale, int B, int C, int64_t P, float scale_bound,
                                 float likelihood_bound, icm_view y_hat_f32, icm_view likelihood, icm_view y_hat_bf16_a,
                                 icm_view y_hat_bf16_b, void *stream)
{
    ICM_CHECK_ARG(y.ptr && mu.ptr && scale.ptr && likelihood.ptr, "icm_gc_likelihood: null argument");
    GcArgs a{};
    a.y = as_view(y); a.mu = as_view(mu); a.scale = as_view(scale);
    a.yhat = as_view(y_hat_f32); a.lik = as_view(likelihood); a.bf_a = as_view(y_hat_bf16_a); a.bf_b = as_view(y_hat_bf16_b);
    a.scale_bound = scale_bound; a.lik_bound = likelihood_bound; a.C = C; a.P = P;
    return launch_gc<GC_LIK>(a, B, stream);
}

extern "C" int icm_add_lrp(icm_view y_hat_f32, icm_view lrp, int B, int C, int64_t P, icm_view y_hat_bf16_a,
                           icm_view y_hat_bf16_b, void *stream)
{
    ICM_CHECK_ARG(y_hat_f32.ptr && lrp.ptr, "icm_add_lrp: null argument");
    GcArgs a{};
    a.y = as_view(lrp);
    a.yhat = as_view(y_hat_f32); a.bf_a = as_view(y_hat_bf16_a); a.bf_b = as_view(y_hat_bf16_b);
    a.C = C; a.P = P;
    return launch_gc<GC_ADD>(a, B, stream);
}

extern "C" int icm_eb_process(int mode, icm_view z, int B, int C, int64_t P, const float *d_params,
                              float likelihood_bound, int32_t *d_symbols, int32_t *d_indexes, icm_view z_hat_f32,
                              icm_view z_hat_bf16, icm_view likelihood, void *stream)
{
    ICM_CHECK_ARG(mode >= 0 && mode <= 2 && d_params, "icm_eb_process: bad mode or null parameters");
    ICM_CHECK_ARG(B > 0 && C > 0 && P > 0 && B <= 65535, "icm_eb_process: empty tensor");
    ICM_CHECK_ARG(mode == 2 || z.ptr, "icm_eb_process: null z");
    ICM_CHECK_ARG(mode == 1 || d_symbols, "icm_eb_process: null symbols");
    ICM_CHECK_ARG(mode != 1 || likelihood.ptr, "icm_eb_process: null likelihood");
    EbArgs a{};
    a.z = as_view(z);
    a.sym = stream_view(d_symbols, (long long)C * P, 0, P);
    a.idx = stream_view(d_indexes, (long long)C * P, 0, P);
    a.zhat = as_view(z_hat_f32); a.zhat_bf = as_view(z_hat_bf16); a.lik = as_view(likelihood);
    a.params = d_params; a.lik_bound = likelihood_bound; a.C = C; a.P = P;
    dim3 grid((unsigned)((P + TILE - 1) / TILE), (C + TILE - 1) / TILE, B);
    cudaStream_t st = as_stream(stream);
    if (mode == 0) eb_kernel<0><<<grid, THREADS, 0, st>>>(a);
    else if (mode == 1) eb_kernel<1><<<grid, THREADS, 0, st>>>(a);
    else eb_kernel<2><<<grid, THREADS, 0, st>>>(a);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}
