// rANS coder on the GPU: the reference's 64-bit-state, 32-bit-renormalising coder, one state per
// stream, bit-exact with compressai.ans (R1-R4 of SURVEY.md §8a; ryg rans64.h:59-142).
//
// Encoder = three kernels
//   (1) rans_records_kernel   embarrassingly parallel: symbol -> {reciprocal of freq, bias, freq, bypass
//                             payload}.  The 64-bit division of the reference's Rans64EncPut is replaced
//                             by Alverson reciprocal multiplication (rans64.h:167-278 shows the exact
//                             equivalence), so no division is left on the serial path.
//   (2) rans_encode_kernel    one warp per stream walks the records back to front.  All 32 lanes carry
//                             the state redundantly (warp-uniform control flow); records are fetched
//                             32 at a time with one coalesced 16-byte load per lane and handed round by
//                             shuffles; emitted words are parked one per lane and stored 128 B at a time.
//   (3) rans_pack_kernel      moves every stream's bytes to its final offset in one packed buffer.
// Decoder = one kernel, one warp per stream, CDF rows (uint16, ragged) and a per-table 2^k-entry
//   "cum >> (16-k) -> first candidate symbol" table staged in shared memory; the symbol search is a
//   single ballot over 32 consecutive CDF entries (the reference scans linearly from entry 0).
#include "common.cuh"

#include <new>
#include <vector>

namespace icm {

constexpr int kPrecision = 16;
constexpr uint64_t kRansL = 1ull << 31;

struct TablesDev {
    int n_cdf, stride, lut_bits, total16;
    const int32_t *cdf32;   // [n_cdf][stride]
    const int32_t *sizes;   // [n_cdf]
    const int32_t *offsets; // [n_cdf]
    const uint16_t *cdf16;  // ragged rows, entry "size-1" (== 65536) is implied
    const int32_t *base;    // [n_cdf] first entry of row t inside cdf16
    const uint16_t *lut;    // [n_cdf << lut_bits]
};

}  // namespace icm

struct icm_tables {
    icm::TablesDev dev;
    void *d_blob;
    size_t smem_bytes;
    int device;
};

struct icm_rans_decoder {
    int n_streams;
    uint64_t *d_state;   // [n_streams] rANS state
    int64_t *d_pos;      // [n_streams] next word index, -1 = not initialised
    int64_t *d_word_off; // [n_streams] first word of the stream inside the byte buffer
    int64_t *d_nwords;   // [n_streams]
    const uint32_t *d_words;
    int32_t *d_status;
};

namespace icm {

// ------------------------------------------------------------------------------------------------
// (1) records.  One 16-byte record per symbol drives a branch-free state update
//        if (hi32(x) >= thi) { emit lo32(x); x >>= 32; }          thi = freq << 15  (x_max = freq << 47)
//        q = mulhi64(x, rcp) >> shift;  x += bias + q * cmpl;      cmpl = 2^16 - freq
//     which equals Rans64EncPut's  x = (x / freq << 16) + x % freq + start  (rans64.h:167-278).
struct __align__(16) Record {
    uint32_t rcp_lo, rcp_hi; // fixed-point reciprocal of freq (rans64.h:223-241)
    uint32_t meta;           // cmpl[0..16] | rcp_shift[17..20] | bypass[21] | nibbles[24..27]
    uint32_t bias;           // start (freq >= 2) or start + 65535 (freq == 1)
};
constexpr uint32_t kRecBypass = 1u << 21;

__global__ void __launch_bounds__(256) rans_records_kernel(TablesDev T, const int32_t *__restrict__ sym,
                                                           const int32_t *__restrict__ idx, long long n_total,
                                                           long long n_per_stream, Record *__restrict__ rec,
                                                           uint32_t *__restrict__ raw_out,
                                                           int32_t *__restrict__ status)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total;
         i += (long long)gridDim.x * blockDim.x) {
        int t = idx[i];
        if (t < 0 || t >= T.n_cdf) { // the reference has only a compiled-out assert here (UB)
            status[i / n_per_stream] = ICM_ERR_BAD_INDEX;
            t = 0;
        }
        const int32_t *cdf = T.cdf32 + (size_t)t * T.stride;
        const int max_value = T.sizes[t] - 2;
        int v = sym[i] - T.offsets[t];
        uint32_t raw = 0;
        if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = max_value; }
        else if (v >= max_value) { raw = (uint32_t)(2 * (v - max_value)); v = max_value; }
        const uint32_t start = (uint16_t)cdf[v];
        const uint32_t freq = (uint16_t)(cdf[v + 1] - cdf[v]);
        Record r;
        uint32_t shift_m1 = 0;
        if (freq < 2) { // rans64.h:192-221
            r.rcp_lo = 0xFFFFFFFFu; r.rcp_hi = 0xFFFFFFFFu;
            r.bias = start + (1u << kPrecision) - 1;
        } else {
            const uint32_t shift = 32 - __clz(freq - 1); // ceil(log2(freq))
            const uint64_t x1 = 1ull << (shift + 31);
            const uint64_t t1 = x1 / freq;
            const uint64_t x0 = (uint64_t)(freq - 1) + ((x1 % freq) << 32);
            const uint64_t t0 = x0 / freq;
            const uint64_t rcp = t0 + (t1 << 32);
            r.rcp_lo = (uint32_t)rcp; r.rcp_hi = (uint32_t)(rcp >> 32);
            shift_m1 = shift - 1;
            r.bias = start;
        }
        r.meta = ((1u << kPrecision) - freq) | (shift_m1 << 17);
        if (v == max_value) {
            uint32_t nb = 0;
            while (nb < 8 && (raw >> (nb * 4)) != 0) ++nb;
            r.meta |= kRecBypass | (nb << 24);
        }
        rec[i] = r;
        raw_out[i] = raw;
    }
}

// ------------------------------------------------------------------------------------------------
// (2) serial walk, one warp per stream.  Every lane carries the state (warp-uniform control flow); the
// records of 32 symbols are staged in shared memory (double buffered, next chunk's global loads in flight)
// and read back with one broadcast LDS.128 per symbol, issued one symbol ahead of its use.  Emitted words
// go to a 128-word shared ring (one predicated store, no branch) that is drained 128 B at a time between
// chunks; a symbol emits at most two words (<= 52 bits of payload), so a chunk emits at most 65.
constexpr int kRing = 128;

struct Emitter {
    uint32_t *ring;
    uint32_t count, flushed; // words emitted / words already stored to global (warp-uniform)
    uint32_t *top;           // one past the last word of the stream's scratch area
    uint32_t capacity, overflow, lane;
    __device__ __forceinline__ void put_if(bool emit, uint32_t w)
    {
        if (emit) ring[count & (kRing - 1)] = w; // every lane stores the same word: uniform, one transaction
        count += emit ? 1u : 0u;
    }
    __device__ __forceinline__ void drain(bool all)
    {
        __syncwarp();
        while (count - flushed >= 32u || (all && count != flushed)) {
            const uint32_t e = flushed + lane;
            if (e < count) {
                if (e < capacity) *(top - 1 - e) = ring[e & (kRing - 1)]; // descending addresses, 128 B per pass
                else overflow = 1;
            }
            flushed = min(flushed + 32u, count);
        }
        __syncwarp();
    }
};

__device__ __forceinline__ void put_bits4(uint64_t &x, uint32_t val, Emitter &e)
{ // Rans64EncPutBits with nbits = 4: freq = 2^12, x_max = 2^59
    const bool emit = (uint32_t)(x >> 32) >= (1u << 27);
    e.put_if(emit, (uint32_t)x);
    x = emit ? (x >> 32) : x;
    x = (x << 4) | val;
}

__global__ void __launch_bounds__(32) rans_encode_kernel(const Record *__restrict__ rec,
                                                         const uint32_t *__restrict__ raw_in,
                                                         long long n_per_stream, uint32_t *__restrict__ words,
                                                         long long cap_words, int32_t *__restrict__ sizes,
                                                         const int32_t *__restrict__ status)
{
    __shared__ uint4 s_rec[2][32];
    __shared__ uint32_t s_raw[2][32];
    __shared__ uint32_t s_ring[kRing];
    const int s = blockIdx.x, lane = threadIdx.x;
    const uint4 *R = reinterpret_cast<const uint4 *>(rec + (size_t)s * n_per_stream);
    const uint32_t *RAW = raw_in + (size_t)s * n_per_stream;
    Emitter e;
    e.ring = s_ring; e.count = 0; e.flushed = 0; e.lane = lane; e.overflow = 0;
    e.top = words + (size_t)(s + 1) * cap_words;
    e.capacity = (uint32_t)cap_words;
    uint64_t x = kRansL;

    const long long n_chunks = (n_per_stream + 31) / 32;
    uint4 g = make_uint4(0, 0, 0, 0);
    uint32_t graw = 0;
    if (n_chunks > 0) {
        const long long j = (n_chunks - 1) * 32 + lane;
        if (j < n_per_stream) { g = __ldg(R + j); graw = __ldg(RAW + j); }
    }
    for (long long c = n_chunks - 1; c >= 0; --c) {
        const int buf = (int)(c & 1);
        s_rec[buf][lane] = g;
        s_raw[buf][lane] = graw;
        if (c > 0) { // the next (earlier) chunk's loads fly while this one is coded
            const long long j = (c - 1) * 32 + lane;
            g = __ldg(R + j);
            graw = __ldg(RAW + j);
        }
        e.drain(false); // also orders the staging stores above before the reads below
        const int valid = (int)min(32LL, n_per_stream - c * 32);
        uint4 cur = s_rec[buf][valid - 1];
#pragma unroll 1
        for (int k = valid - 1; k >= 0; --k) {
            const uint4 nxt = s_rec[buf][k > 0 ? k - 1 : 0];
            if (cur.z & kRecBypass) {
                // records of an escaped symbol, in push order: main, count(nb), nibble_0..nibble_{nb-1};
                // drained back to front.  nb <= 8 < 15, so the count is a single nibble.
                const uint32_t raw = s_raw[buf][k];
                const int nb = (int)((cur.z >> 24) & 15u);
#pragma unroll 1
                for (int j = nb - 1; j >= 0; --j) put_bits4(x, (raw >> (j * 4)) & 15u, e);
                put_bits4(x, (uint32_t)nb, e);
            }
            const uint32_t cmpl = cur.z & 0x1FFFFu;
            const uint32_t shift = (cur.z >> 17) & 15u;
            const uint32_t thi = ((1u << kPrecision) - cmpl) << 15;
            const bool emit = (uint32_t)(x >> 32) >= thi;
            e.put_if(emit, (uint32_t)x);
            x = emit ? (x >> 32) : x;
            const uint64_t rcp = ((uint64_t)cur.y << 32) | cur.x;
            const uint64_t q = __umul64hi(x, rcp) >> shift;
            x = x + cur.w + q * (uint64_t)cmpl;
            cur = nxt;
        }
    }
    // Rans64EncFlush: ptr -= 2; ptr[0] = lo; ptr[1] = hi  => hi is the "earlier" emitted word
    e.drain(false);
    e.put_if(true, (uint32_t)(x >> 32));
    e.put_if(true, (uint32_t)x);
    e.drain(true);
    if (lane == 0) {
        const int32_t st = status[s];
        sizes[s] = st < 0 ? st : (e.overflow ? ICM_ERR_CAPACITY : (int32_t)(e.count * 4));
    }
}

// device-side set_streams: word offsets / lengths from the encoder's int32 byte sizes (exclusive scan, one warp)
__global__ void rans_set_streams_kernel(const int32_t *__restrict__ sizes, int n_streams, int64_t *__restrict__ pos,
                                        int64_t *__restrict__ word_off, int64_t *__restrict__ nwords, int32_t *__restrict__ status)
{
    long long run = 0;
    for (int base = 0; base < n_streams; base += 32) {
        const int i = base + threadIdx.x;
        const long long v = (i < n_streams && sizes[i] > 0) ? sizes[i] / 4 : 0;
        long long incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)threadIdx.x >= d) incl += o;
        }
        if (i < n_streams) {
            word_off[i] = run + incl - v;
            nwords[i] = v;
            pos[i] = -1;
            status[i] = sizes[i] < 0 ? sizes[i] : 0;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// (3) pack: exclusive scan of sizes (one warp) + copy
__global__ void rans_scan_kernel(const int32_t *__restrict__ sizes, int n_streams, long long *__restrict__ offsets,
                                 int32_t *__restrict__ total_out)
{
    long long run = 0;
    for (int base = 0; base < n_streams; base += 32) {
        int i = base + threadIdx.x;
        long long v = (i < n_streams && sizes[i] > 0) ? sizes[i] : 0;
        long long incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)threadIdx.x >= d) incl += o;
        }
        if (i < n_streams) offsets[i] = run + incl - v;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) *total_out = (int32_t)min(run, (long long)INT32_MAX);
}

__global__ void __launch_bounds__(256) rans_pack_kernel(const uint32_t *__restrict__ words, long long cap_words,
                                                        int32_t *sizes, const long long *__restrict__ offsets,
                                                        uint32_t *__restrict__ packed, long long packed_cap_words)
{
    const int s = blockIdx.y;
    const int nb = sizes[s];
    if (nb <= 0) return;
    const long long nw = nb / 4, off = offsets[s] / 4;
    if (off + nw > packed_cap_words) {
        if (blockIdx.x == 0 && threadIdx.x == 0) sizes[s] = ICM_ERR_CAPACITY;
        return;
    }
    const uint32_t *src = words + (size_t)(s + 1) * cap_words - nw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nw; i += (long long)gridDim.x * blockDim.x)
        packed[off + i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// decoder: one warp per stream.  Per-symbol critical path (all lanes carry x):
//     cum = x & 0xFFFF -> 32 lanes compare one CDF entry each -> ballot -> ffs -> every lane has already
//     formed freq * (x >> 16) + cum - start for ITS entry; the winner's 64-bit result is shuffled out -> renorm.
// Everything else is taken off that path: the table of symbol k+1 is known in advance (indexes are an
// input), so its metadata and a speculative 32-entry window of its CDF row (the whole row for tables with
// <= 32 entries, else the 32 entries around the distribution's centre) are loaded from shared memory while
// symbol k is decoded.  Only when cum falls outside that window does the coder take the general route:
// the per-table "cum >> (16-k) -> first candidate" table, then 32-entry windows until one brackets cum.
__device__ __forceinline__ uint32_t cdf_window(const uint16_t *s_cdf, uint32_t base, int size, int s0, int lane)
{
    const int cand = s0 + lane;
    return (cand >= size - 1) ? 0x10000u : (uint32_t)s_cdf[base + cand]; // entry size-1 is 65536 by definition
}

__global__ void __launch_bounds__(32) rans_decode_kernel(TablesDev T, const uint32_t *__restrict__ words,
                                                         const int64_t *__restrict__ word_off,
                                                         const int64_t *__restrict__ nwords_arr,
                                                         uint64_t *__restrict__ state, int64_t *__restrict__ pos_arr,
                                                         const int32_t *__restrict__ idx, long long n_per_stream,
                                                         int32_t *__restrict__ out, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint16_t *s_cdf = reinterpret_cast<uint16_t *>(smem_raw);
    uint16_t *s_lut = s_cdf + ((T.total16 + 7) & ~7);
    uint4 *s_meta = reinterpret_cast<uint4 *>(s_lut + ((size_t)T.n_cdf << T.lut_bits)); // {base, size | t << 16, offset, s0}
    __shared__ uint4 s_sym[2][32];
    const int lane = threadIdx.x;
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(T.cdf16);
        uint4 *d = reinterpret_cast<uint4 *>(s_cdf);
        for (int i = lane; i < (T.total16 + 7) / 8; i += 32) d[i] = g[i];
        const uint4 *gl = reinterpret_cast<const uint4 *>(T.lut);
        uint4 *dl = reinterpret_cast<uint4 *>(s_lut);
        for (int i = lane; i < (int)(((size_t)T.n_cdf << T.lut_bits) / 8); i += 32) dl[i] = gl[i];
        for (int i = lane; i < T.n_cdf; i += 32) {
            const int size = T.sizes[i], off = T.offsets[i];
            int s0 = 0;
            if (size > 32) { s0 = -off - 15; s0 = max(0, min(s0, size - 32)); }
            s_meta[i] = make_uint4((uint32_t)T.base[i], (uint32_t)size | ((uint32_t)i << 16), (uint32_t)off, (uint32_t)s0);
        }
    }
    __syncwarp();

    const int s = blockIdx.x;
    const uint32_t *W = words + word_off[s];
    const long long nwords = nwords_arr[s];
    uint64_t x;
    long long pos = pos_arr[s];
    if (pos < 0) { // set_stream: Rans64DecInit
        const uint32_t w0 = nwords > 0 ? W[0] : 0u, w1 = nwords > 1 ? W[1] : 0u;
        x = (uint64_t)w0 | ((uint64_t)w1 << 32);
        pos = 2;
    } else {
        x = state[s];
    }
    // word window: lane l of wcur holds word (wblock*32 + l); wnxt is the following block
    long long wblock = pos >> 5;
    auto load_block = [&](long long b) -> uint32_t {
        const long long j = b * 32 + lane;
        return j < nwords ? __ldg(W + j) : 0u; // past-the-end reads are UB in the reference; we feed zeros
    };
    uint32_t wcur = load_block(wblock), wnxt = load_block(wblock + 1);
    uint32_t wv = __shfl_sync(0xffffffffu, wcur, (int)(pos & 31)); // the next stream word, always kept ready
    // renormalise (Rans64DecAdvance tail / Rans64DecGetBits tail): predicated, plus a rare window refill
    auto renorm = [&]() {
        const bool rn = x < kRansL;
        x = rn ? ((x << 32) | wv) : x;
        pos += rn ? 1 : 0;
        if (rn && (pos & 31) == 0) { wcur = wnxt; ++wblock; wnxt = load_block(wblock + 1); }
        wv = __shfl_sync(0xffffffffu, wcur, (int)(pos & 31));
    };
    auto get4 = [&]() -> int {
        const int val = (int)(x & 15u);
        x >>= 4;
        renorm();
        return val;
    };

    const int32_t *I = idx + (size_t)s * n_per_stream;
    int32_t *O = out + (size_t)s * n_per_stream;
    const int lut_shift = kPrecision - T.lut_bits;
    const long long n = n_per_stream;
    const long long n_chunks = (n + 31) / 32;
    bool bad = false;
    auto meta_of = [&](int t) -> uint4 {
        if (t < 0 || t >= T.n_cdf) { bad = true; t = 0; }
        return s_meta[t];
    };
    auto sym_at = [&](long long g) -> uint4 {
        g = min(g, n - 1);
        return s_sym[(g >> 5) & 1][g & 31];
    };
    int ireg = (lane < n) ? __ldg(I + lane) : 0;
    s_sym[0][lane] = meta_of(ireg);
    ireg = (32 + lane < n) ? __ldg(I + 32 + lane) : 0;
    __syncwarp();
    uint4 meta0 = sym_at(0), meta1 = sym_at(1);
    uint32_t v0 = cdf_window(s_cdf, meta0.x, (int)(meta0.y & 0xFFFFu), (int)meta0.w, lane);
    uint32_t f0 = __shfl_down_sync(0xffffffffu, v0, 1) - v0;

    for (long long c = 0; c < n_chunks; ++c) {
        __syncwarp();
        if (c + 1 < n_chunks) { // stage the next chunk's per-symbol metadata; its indexes were fetched a chunk ago
            s_sym[(c + 1) & 1][lane] = meta_of(ireg);
            const long long j = (c + 2) * 32 + lane;
            ireg = (j < n) ? __ldg(I + j) : 0;
        }
        __syncwarp();
        const int valid = (int)min(32LL, n - c * 32);
        int result = 0;
#pragma unroll 1
        for (int k = 0; k < valid; ++k) {
            const long long gk = c * 32 + k;
            // ---- off the critical path: operands of the next two symbols, the next stream word
            const uint32_t v1 = cdf_window(s_cdf, meta1.x, (int)(meta1.y & 0xFFFFu), (int)meta1.w, lane);
            const uint4 meta2 = sym_at(gk + 2);
            const int size = (int)(meta0.y & 0xFFFFu);
            // ---- critical path
            const uint32_t cum = (uint32_t)x & 0xFFFFu;
            uint32_t vv = v0, ff = f0;
            int s0 = (int)meta0.w;
            uint32_t m = __ballot_sync(0xffffffffu, vv > cum);
            if ((m & 1u) | (m == 0u)) { // cum is outside the speculative window
                const int t = (int)(meta0.y >> 16);
                s0 = s_lut[(t << T.lut_bits) + (cum >> lut_shift)];
                while (true) {
                    vv = cdf_window(s_cdf, meta0.x, size, s0, lane);
                    m = __ballot_sync(0xffffffffu, vv > cum);
                    if (m != 0u) break;
                    s0 += 31; // more than 31 symbols share this bucket
                }
                ff = __shfl_down_sync(0xffffffffu, vv, 1) - vv;
            }
            const int p = __ffs(m) - 1; // >= 1: entry s0 is <= cum
            // Rans64DecAdvance, formed by every lane for its own entry; lane p-1 holds the real one
            const uint64_t nx = (uint64_t)ff * (x >> kPrecision) + (uint64_t)(cum - vv);
            const uint32_t nlo = __shfl_sync(0xffffffffu, (uint32_t)nx, p - 1);
            const uint32_t nhi = __shfl_sync(0xffffffffu, (uint32_t)(nx >> 32), p - 1);
            x = ((uint64_t)nhi << 32) | nlo;
            renorm();
            const int symbol = s0 + p - 1;
            int value = symbol;
            if (symbol == size - 2) { // bypass escape
                const int max_value = size - 2;
                int val = get4();
                int nb = val;
#pragma unroll 1
                while (val == 15) { val = get4(); nb += val; }
                int raw = 0;
#pragma unroll 1
                for (int j = 0; j < nb; ++j) { val = get4(); raw |= val << ((j * 4) & 31); }
                value = raw >> 1;
                if (raw & 1) value = -value - 1; else value += max_value;
            }
            value += (int)meta0.z;
            if (lane == k) result = value;
            // ---- rotate the pipeline
            meta0 = meta1; meta1 = meta2;
            v0 = v1;
            f0 = __shfl_down_sync(0xffffffffu, v1, 1) - v1;
        }
        if (lane < valid) O[c * 32 + lane] = result;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        state[s] = x;
        pos_arr[s] = pos;
        if (bad) status[s] = ICM_ERR_BAD_INDEX;
    }
}

}  // namespace icm

// =================================================================================================
// C ABI
using namespace icm;

extern "C" int icm_tables_create(const int32_t *h_cdfs, int n_cdf, int stride, const int32_t *h_sizes,
                                 const int32_t *h_offsets, icm_tables **out)
{
    ICM_CHECK_ARG(h_cdfs && h_sizes && h_offsets && out, "icm_tables_create: null argument");
    ICM_CHECK_ARG(n_cdf > 0 && stride >= 3, "icm_tables_create: bad shape n_cdf=%d stride=%d", n_cdf, stride);
    std::vector<int32_t> base(n_cdf);
    int total = 0;
    for (int t = 0; t < n_cdf; ++t) {
        const int size = h_sizes[t];
        ICM_CHECK_ARG(size >= 3 && size <= stride, "icm_tables_create: cdf_size[%d]=%d outside [3,%d]", t, size, stride);
        const int32_t *c = h_cdfs + (size_t)t * stride;
        ICM_CHECK_ARG(c[0] == 0 && c[size - 1] == (1 << kPrecision), "icm_tables_create: row %d is not a 16-bit CDF", t);
        for (int j = 0; j + 1 < size; ++j)
            ICM_CHECK_ARG(c[j] < c[j + 1], "icm_tables_create: row %d not strictly increasing at %d", t, j);
        base[t] = total;
        total += (size + 1) & ~1; // keep rows 4-byte aligned
    }
    int lut_bits = 8;
    auto smem_need = [&](int bits) {
        return (size_t)((total + 7) & ~7) * 2 + ((size_t)n_cdf << bits) * 2 + (size_t)n_cdf * 16;
    };
    while (lut_bits > 3 && smem_need(lut_bits) > 200 * 1024) --lut_bits;
    ICM_CHECK_ARG(smem_need(lut_bits) <= 200 * 1024, "icm_tables_create: tables too large for shared memory");
    const size_t lut_n = (size_t)n_cdf << lut_bits;
    std::vector<uint16_t> cdf16(((size_t)total + 7) & ~(size_t)7, 0), lut(lut_n);
    for (int t = 0; t < n_cdf; ++t) {
        const int size = h_sizes[t];
        const int32_t *c = h_cdfs + (size_t)t * stride;
        for (int j = 0; j < size; ++j) cdf16[base[t] + j] = (uint16_t)c[j]; // 65536 wraps to 0; implied by position
        int sidx = 0;
        for (int b = 0; b < (1 << lut_bits); ++b) {
            const int32_t lo = b << (kPrecision - lut_bits);
            while (sidx + 1 <= size - 2 && c[sidx + 1] <= lo) ++sidx;
            lut[((size_t)t << lut_bits) + b] = (uint16_t)sidx;
        }
    }
    // one device blob: cdf32 | sizes | offsets | base | cdf16 | lut
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_cdf32 = 0, o_sizes = al(o_cdf32 + (size_t)n_cdf * stride * 4), o_offsets = al(o_sizes + n_cdf * 4),
                 o_base = al(o_offsets + n_cdf * 4), o_cdf16 = al(o_base + n_cdf * 4),
                 o_lut = al(o_cdf16 + cdf16.size() * 2), blob_bytes = al(o_lut + lut_n * 2);
    std::vector<unsigned char> host(blob_bytes, 0);
    memcpy(&host[o_cdf32], h_cdfs, (size_t)n_cdf * stride * 4);
    memcpy(&host[o_sizes], h_sizes, n_cdf * 4);
    memcpy(&host[o_offsets], h_offsets, n_cdf * 4);
    memcpy(&host[o_base], base.data(), n_cdf * 4);
    memcpy(&host[o_cdf16], cdf16.data(), cdf16.size() * 2);
    memcpy(&host[o_lut], lut.data(), lut_n * 2);
    icm_tables *T = new (std::nothrow) icm_tables();
    ICM_CHECK_ARG(T, "icm_tables_create: out of host memory");
    if (cudaGetDevice(&T->device) != cudaSuccess) { delete T; set_error("icm_tables_create: no CUDA device"); return ICM_ERR_NO_DEVICE; }
    cudaError_t e = cudaMalloc(&T->d_blob, blob_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(T->d_blob, host.data(), blob_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("icm_tables_create: %s", cudaGetErrorString(e)); if (T->d_blob) cudaFree(T->d_blob); delete T; return ICM_ERR_CUDA; }
    char *b = (char *)T->d_blob;
    T->dev = TablesDev{n_cdf, stride, lut_bits, total,
                       (const int32_t *)(b + o_cdf32), (const int32_t *)(b + o_sizes), (const int32_t *)(b + o_offsets),
                       (const uint16_t *)(b + o_cdf16), (const int32_t *)(b + o_base), (const uint16_t *)(b + o_lut)};
    T->smem_bytes = smem_need(lut_bits);
    *out = T;
    return ICM_OK;
}

extern "C" void icm_tables_destroy(icm_tables *t)
{
    if (!t) return;
    cudaFree(t->d_blob);
    delete t;
}

static inline long long enc_cap_words(long long n)
{
    // worst case per symbol: 16 bits + escape (unary nibble + 8 payload nibbles) = 52 bits; + final state
    long long w = (n * 52 + 31) / 32 + 2;
    return (w + 31) & ~31LL;
}

struct EncLayout { size_t rec, raw, words, offs, status, total; long long cap_words; };
static EncLayout enc_layout(int n_streams, long long n)
{
    EncLayout L;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t nt = (size_t)n_streams * n;
    L.cap_words = enc_cap_words(n);
    L.rec = 0;
    L.raw = al(L.rec + nt * sizeof(Record));
    L.words = al(L.raw + nt * 4);
    L.offs = al(L.words + (size_t)n_streams * L.cap_words * 4);
    L.status = al(L.offs + (size_t)n_streams * 8);
    L.total = al(L.status + (size_t)n_streams * 4);
    return L;
}

extern "C" int64_t icm_rans_encode_workspace_bytes(int n_streams, int64_t n_per_stream)
{
    if (n_streams <= 0 || n_per_stream < 0) return ICM_ERR_INVALID_ARG;
    return (int64_t)enc_layout(n_streams, n_per_stream).total;
}

extern "C" int icm_rans_encode_batch(const icm_tables *t, const int32_t *d_symbols, const int32_t *d_indexes,
                                     int n_streams, int64_t n_per_stream, void *d_work, uint8_t *d_packed,
                                     int64_t packed_capacity, int32_t *d_sizes, void *stream)
{
    ICM_CHECK_ARG(t && d_work && d_packed && d_sizes, "icm_rans_encode_batch: null argument");
    ICM_CHECK_ARG(n_streams > 0 && n_per_stream >= 0, "icm_rans_encode_batch: bad sizes");
    ICM_CHECK_ARG(n_per_stream == 0 || (d_symbols && d_indexes), "icm_rans_encode_batch: null symbols");
    ICM_CHECK_ARG(((uintptr_t)d_packed & 3) == 0 && ((uintptr_t)d_work & 255) == 0, "icm_rans_encode_batch: misaligned buffers");
    cudaStream_t st = as_stream(stream);
    const EncLayout L = enc_layout(n_streams, n_per_stream);
    char *w = (char *)d_work;
    Record *rec = (Record *)(w + L.rec);
    uint32_t *raw = (uint32_t *)(w + L.raw);
    uint32_t *words = (uint32_t *)(w + L.words);
    long long *offs = (long long *)(w + L.offs);
    int32_t *status = (int32_t *)(w + L.status);
    ICM_CUDA(cudaMemsetAsync(status, 0, (size_t)n_streams * 4, st));
    const long long nt = (long long)n_streams * n_per_stream;
    if (nt > 0) {
        const int grid = (int)min((nt + 255) / 256, (long long)sm_count() * 16);
        rans_records_kernel<<<grid, 256, 0, st>>>(t->dev, d_symbols, d_indexes, nt, n_per_stream, rec, raw, status);
        ICM_LAUNCH_CHECK();
    }
    rans_encode_kernel<<<n_streams, 32, 0, st>>>(rec, raw, n_per_stream, words, L.cap_words, d_sizes, status);
    ICM_LAUNCH_CHECK();
    rans_scan_kernel<<<1, 32, 0, st>>>(d_sizes, n_streams, offs, d_sizes + n_streams);
    ICM_LAUNCH_CHECK();
    dim3 grid((unsigned)max(1LL, min(64LL, (L.cap_words + 4095) / 4096)), n_streams);
    rans_pack_kernel<<<grid, 256, 0, st>>>(words, L.cap_words, d_sizes, offs, (uint32_t *)d_packed, packed_capacity / 4);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_rans_decoder_create(int n_streams, icm_rans_decoder **out)
{
    ICM_CHECK_ARG(out && n_streams > 0, "icm_rans_decoder_create: bad arguments");
    icm_rans_decoder *d = new (std::nothrow) icm_rans_decoder();
    ICM_CHECK_ARG(d, "icm_rans_decoder_create: out of host memory");
    d->n_streams = n_streams;
    d->d_words = nullptr;
    char *blob = nullptr;
    const size_t per = 8 + 8 + 8 + 8 + 4;
    cudaError_t e = cudaMalloc(&blob, per * n_streams + 64);
    if (e != cudaSuccess) { delete d; set_error("icm_rans_decoder_create: %s", cudaGetErrorString(e)); return ICM_ERR_CUDA; }
    d->d_state = (uint64_t *)blob;
    d->d_pos = (int64_t *)(blob + 8 * (size_t)n_streams);
    d->d_word_off = (int64_t *)(blob + 16 * (size_t)n_streams);
    d->d_nwords = (int64_t *)(blob + 24 * (size_t)n_streams);
    d->d_status = (int32_t *)(blob + 32 * (size_t)n_streams);
    *out = d;
    return ICM_OK;
}

extern "C" void icm_rans_decoder_destroy(icm_rans_decoder *d)
{
    if (!d) return;
    cudaFree(d->d_state);
    delete d;
}

extern "C" int icm_rans_decoder_set_streams(icm_rans_decoder *d, const uint8_t *d_bytes, const int64_t *h_offsets,
                                            const int64_t *h_sizes, void *stream)
{
    ICM_CHECK_ARG(d && d_bytes && h_offsets && h_sizes, "icm_rans_decoder_set_streams: null argument");
    ICM_CHECK_ARG(((uintptr_t)d_bytes & 3) == 0, "icm_rans_decoder_set_streams: byte buffer must be 4-byte aligned");
    const int n = d->n_streams;
    std::vector<int64_t> host(3 * (size_t)n);
    for (int s = 0; s < n; ++s) {
        ICM_CHECK_ARG(h_offsets[s] % 4 == 0 && h_sizes[s] % 4 == 0 && h_sizes[s] >= 0,
                      "icm_rans_decoder_set_streams: stream %d offset/size not a multiple of 4", s);
        host[s] = -1;                       // pos: uninitialised
        host[n + s] = h_offsets[s] / 4;     // word_off
        host[2 * (size_t)n + s] = h_sizes[s] / 4;
    }
    cudaStream_t st = as_stream(stream);
    // d_pos, d_word_off, d_nwords are contiguous
    ICM_CUDA(cudaMemcpyAsync(d->d_pos, host.data(), host.size() * 8, cudaMemcpyHostToDevice, st));
    ICM_CUDA(cudaMemsetAsync(d->d_status, 0, (size_t)n * 4, st));
    ICM_CUDA(cudaStreamSynchronize(st)); // `host` is pageable and dies with this frame
    d->d_words = (const uint32_t *)d_bytes;
    return ICM_OK;
}

extern "C" int icm_rans_decoder_set_streams_device(icm_rans_decoder *d, const uint8_t *d_bytes, const int32_t *d_sizes, void *stream)
{
    ICM_CHECK_ARG(d && d_bytes && d_sizes, "icm_rans_decoder_set_streams_device: null argument");
    ICM_CHECK_ARG(((uintptr_t)d_bytes & 3) == 0, "icm_rans_decoder_set_streams_device: byte buffer must be 4-byte aligned");
    rans_set_streams_kernel<<<1, 32, 0, as_stream(stream)>>>(d_sizes, d->n_streams, d->d_pos, d->d_word_off, d->d_nwords, d->d_status);
    ICM_LAUNCH_CHECK();
    d->d_words = (const uint32_t *)d_bytes;
    return ICM_OK;
}

extern "C" int icm_rans_decoder_step(icm_rans_decoder *d, const icm_tables *t, const int32_t *d_indexes,
                                     int64_t n_per_stream, int32_t *d_out, void *stream)
{
    ICM_CHECK_ARG(d && t && d->d_words, "icm_rans_decoder_step: decoder has no stream (call set_streams first)");
    ICM_CHECK_ARG(n_per_stream >= 0 && (n_per_stream == 0 || (d_indexes && d_out)), "icm_rans_decoder_step: bad arguments");
    if (n_per_stream == 0) return ICM_OK;
    static thread_local size_t configured = 0;
    if (t->smem_bytes > configured) {
        ICM_CUDA(cudaFuncSetAttribute(rans_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem_bytes));
        configured = t->smem_bytes;
    }
    rans_decode_kernel<<<d->n_streams, 32, t->smem_bytes, as_stream(stream)>>>(
        t->dev, d->d_words, d->d_word_off, d->d_nwords, d->d_state, d->d_pos, d_indexes, n_per_stream, d_out, d->d_status);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_rans_decoder_status(icm_rans_decoder *d, int32_t *h_status, void *stream)
{
    ICM_CHECK_ARG(d && h_status, "icm_rans_decoder_status: null argument");
    ICM_CUDA(cudaMemcpyAsync(h_status, d->d_status, (size_t)d->n_streams * 4, cudaMemcpyDeviceToHost, as_stream(stream)));
    ICM_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return ICM_OK;
}
