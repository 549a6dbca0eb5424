// rANS coder on the GPU: the reference's 64-bit-state, 32-bit-renormalising coder, one state per
// stream, bit-exact with compressai.ans (R1-R4 of SURVEY.md §8a; ryg rans64.h:59-142).
//
// A stream is a strictly serial recurrence carried by ONE WARP, and a lone warp on B200 pays ~3-5 cycles per
// instruction it issues (ncu, round 2: 5.1 cycles per issued instruction, of which 1.5 fixed-latency dependency
// waits, 0.8 global-memory scoreboard, 0.6 branch resolution, 0.4 instruction fetch), so a stream's speed is set by
// the NUMBER of instructions on the serial path.  Both coders therefore keep the per-symbol path short and straight,
// and move everything that does not depend on the coder state into parallel kernels or into bulk work of the other
// 31 lanes.
//
// Encoder = three kernels
//   (1) rans_records_kernel   embarrassingly parallel: symbol -> 32-byte record {reciprocal of freq, renorm
//                             threshold, 2^16 - freq, shift, bias, escape payload}.  The 64-bit division of
//                             the reference's Rans64EncPut becomes an Alverson reciprocal multiplication
//                             (rans64.h:167-278 proves the equivalence), so no division is left on the
//                             serial path.
//   (2) rans_encode_kernel    one warp per stream walks the records back to front: two broadcast LDS.128 per
//                             symbol from a double-buffered shared-memory stage, one predicated store per
//                             emitted word into a shared buffer that is drained coalesced between chunks.
//   (3) rans_pack_kernel      moves every stream's bytes to its final offset in one packed buffer.
// Decoder = one kernel per step, one warp per stream, state persists across steps (decode_stream is called 12 times
//   per image):
//   rans_decode_bucket_kernel per-table bucket tables in shared memory resolve a symbol from ONE LDS.128 and a few
//                             compares / selects in registers (csrc/rans_lane.cuh); crowded buckets and escapes take
//                             one out-of-line path.
//   rans_decode_warp_kernel   round-1 kernel (speculative 32-entry window + ballot search), the fallback for table
//                             sets that do not fit the bucket image or contain a symbol of frequency 65535.
#include "common.cuh"
#include <string.h>

#include "rans_lane.cuh"

#include <stdlib.h>

#include <new>
#include <vector>

namespace icm {

constexpr int kPrecision = 16;
constexpr uint64_t kRansL = 1ull << 31;
constexpr uint32_t kSentinel = 0x10000u; // == cdf[last]; also pads every row in the fallback decoder's table
constexpr int kRowPad = 31;

struct TablesDev {
    int n_cdf, stride, lut_bits, total_pad;
    const int32_t *cdf32;    // [n_cdf][stride]
    const int32_t *sizes;    // [n_cdf]
    const int32_t *offsets;  // [n_cdf]
    const uint32_t *cdf_pad; // fallback decoder rows: size entries + 31 sentinels each
    const int32_t *base;     // [n_cdf] first entry of row t inside cdf_pad
    const uint16_t *lut;     // [n_cdf << lut_bits]
    // bucket decoder
    const uint4 *image;      // shared-memory image (rans_lane.cuh), nullptr if the tables do not fit
    uint32_t image_bytes, meta_off, row_off, rs, M;
};

}  // namespace icm

struct icm_tables {
    icm::TablesDev dev;
    void *d_blob;
    size_t smem_bytes;      // fallback decoder
    size_t image_smem_bytes; // bucket decoder: image + alignment slack (per-warp areas come on top), 0 = not available
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
    int device;
};

namespace icm {

// shared-memory accessors on 32-bit shared-window addresses (keeps the address arithmetic out of the loops)
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    // laundered through an opaque move so that the compiler keeps the address in a register instead of
    // re-deriving the shared window (S2UR SR_CgaCtaId + ULEA) at every use inside the serial loops
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


// ------------------------------------------------------------------------------------------------
// (1) records.  Per symbol a 32-byte record drives the branch-free state update
//        if (hi32(x) >= thi) { emit lo32(x); x >>= 32; }          thi = freq << 15  (x_max = freq << 47)
//        q = mulhi64(x, rcp) >> shift;  x += bias + q * cmpl;      cmpl = 2^16 - freq
//     == Rans64EncPut's  x = (x / freq << 16) + x % freq + start.  Streams are padded to a multiple of 32
//     records with identity records (thi = ~0, everything else 0).
struct __align__(16) Record {
    uint32_t rcp_lo, rcp_hi, thi, cmpl; // first LDS.128
    uint32_t shift, bias, escape, raw;  // second LDS.128; escape = 0 or 0x100 | nibbles
};

__global__ void __launch_bounds__(256) rans_records_kernel(TablesDev T, const int32_t *__restrict__ sym,
                                                           const int32_t *__restrict__ idx, int n_streams,
                                                           long long n_per_stream, long long n_pad,
                                                           Record *__restrict__ rec, int32_t *__restrict__ status)
{
    const long long total = (long long)n_streams * n_pad;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const long long s = g / n_pad, i = g - s * n_pad;
        Record r;
        if (i >= n_per_stream) { // identity
            r.rcp_lo = r.rcp_hi = 0; r.thi = 0xFFFFFFFFu; r.cmpl = 0; r.shift = 0; r.bias = 0; r.escape = 0; r.raw = 0;
            rec[g] = r;
            continue;
        }
        int t = idx[s * n_per_stream + i];
        if (t < 0 || t >= T.n_cdf) { // the reference has only a compiled-out assert here (UB)
            status[s] = ICM_ERR_BAD_INDEX;
            t = 0;
        }
        const int32_t *cdf = T.cdf32 + (size_t)t * T.stride;
        const int max_value = T.sizes[t] - 2;
        int v = sym[s * n_per_stream + i] - T.offsets[t];
        uint32_t raw = 0;
        if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = max_value; }
        else if (v >= max_value) { raw = (uint32_t)(2 * (v - max_value)); v = max_value; }
        const uint32_t start = (uint16_t)cdf[v];
        const uint32_t freq = (uint16_t)(cdf[v + 1] - cdf[v]);
        if (freq < 2) { // rans64.h:192-221
            r.rcp_lo = 0xFFFFFFFFu; r.rcp_hi = 0xFFFFFFFFu;
            r.shift = 0;
            r.bias = start + (1u << kPrecision) - 1;
        } else {
            const uint32_t shift = 32 - __clz(freq - 1); // ceil(log2(freq))
            const uint64_t x1 = 1ull << (shift + 31);
            const uint64_t t1 = x1 / freq;
            const uint64_t x0 = (uint64_t)(freq - 1) + ((x1 % freq) << 32);
            const uint64_t t0 = x0 / freq;
            const uint64_t rcp = t0 + (t1 << 32);
            r.rcp_lo = (uint32_t)rcp; r.rcp_hi = (uint32_t)(rcp >> 32);
            r.shift = shift - 1;
            r.bias = start;
        }
        r.thi = freq << 15;
        r.cmpl = (1u << kPrecision) - freq;
        r.escape = 0;
        if (v == max_value) {
            uint32_t nb = 0;
            while (nb < 8 && (raw >> (nb * 4)) != 0) ++nb;
            r.escape = 0x100u | nb;
        }
        r.raw = raw;
        rec[g] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// (2) serial walk, one warp per stream.  All lanes carry the state (warp-uniform control flow).
constexpr int kEncOutWords = 96; // a 32-symbol chunk emits at most 64 words (<= 52 payload bits per symbol)

struct EncState {
    uint64_t x;
    uint32_t out; // shared address of the next free word of the chunk's output buffer
};

__device__ __forceinline__ void enc_put_bits4(EncState &st, uint32_t val)
{ // Rans64EncPutBits, nbits = 4: freq = 2^12, x_max = 2^59
    const bool emit = (uint32_t)(st.x >> 32) >= (1u << 27);
    if (emit) { sts32(st.out, (uint32_t)st.x); st.out += 4; st.x >>= 32; }
    st.x = (st.x << 4) | val;
}

// records of an escaped symbol, in push order: main, count(nb), nibble_0..nibble_{nb-1}; drained back to
// front.  nb <= 8 < 15, so the count is a single nibble.  Out of line: rare on real data.
__device__ __noinline__ EncState enc_escape(EncState st, uint32_t raw, uint32_t nb)
{
    for (int j = (int)nb - 1; j >= 0; --j) enc_put_bits4(st, (raw >> (j * 4)) & 15u);
    enc_put_bits4(st, nb);
    return st;
}

// Up to 8 streams share a CTA (one warp each, nothing shared between them; ICM_ENC_WARPS overrides).  With one-warp CTAs the block
// scheduler spread a job's 32 encoders over 32 SMs for ~37 ms, and on every one of them that warp's registers kept the second
// swin_block CTA (2 x 32 768 registers fill the file) from being resident: the pipelined step lost ~4 ms to its encoders and only
// 0.5 ms to its decoders (tools/ab_coders.sh: 860-878 images/s with 1 warp per CTA, 888-906 with 8, 924-933 without encoders).
// Eight warps per CTA confine a job's encoders to 4 SMs; two warps per scheduler cost a stream ~12 % of its speed (38 -> 43 ms
// with every CTA full; a lone stream, B = 1, is as fast as before).
constexpr int kEncWarpsMax = 8;
constexpr int kEncWarpBytes = 2 * 64 * 16 + kEncOutWords * 4; // 2 432

__global__ void __launch_bounds__(32 * kEncWarpsMax, 1) rans_encode_kernel(const Record *__restrict__ rec, long long n_pad,
                                                                        uint32_t *__restrict__ words, long long cap_words,
                                                                        int32_t *__restrict__ sizes, const int32_t *__restrict__ status,
                                                                        int n_streams)
{
    extern __shared__ __align__(16) unsigned char s_enc[]; // per warp: records [2][64] uint4 (buffer, symbol * 2 + half) + output words
    // the warp index is broadcast with a shuffle so that the compiler sees warp-uniform control flow (with a plain threadIdx.x >> 5 it
    // wraps the per-symbol branches in convergence barriers: 36.6 -> 41.4 ms per stream)
    const int wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + wid;
    if (s >= n_streams) return; // whole warp
    uint4 (*s_rec)[64] = reinterpret_cast<uint4 (*)[64]>(s_enc + (size_t)wid * kEncWarpBytes);
    uint32_t *s_out = reinterpret_cast<uint32_t *>(s_enc + (size_t)wid * kEncWarpBytes + 2 * 64 * sizeof(uint4));
    const uint4 *R = reinterpret_cast<const uint4 *>(rec + (size_t)s * n_pad);
    uint32_t *top = words + (size_t)(s + 1) * cap_words; // one past the last word of this stream's scratch
    const uint32_t rec_base = smem_addr(&s_rec[0][0]), out_base = smem_addr(&s_out[0]);
    EncState st;
    st.x = kRansL;
    st.out = out_base;
    uint32_t total = 0; // words stored to global so far
    bool overflow = false;
    auto drain = [&]() {
        __syncwarp();
        const uint32_t n = (st.out - out_base) >> 2;
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t e = total + i;
            if (e < (uint32_t)cap_words) *(top - 1 - e) = s_out[i]; // descending addresses
            else overflow = true;
        }
        total += n;
        st.out = out_base;
        __syncwarp();
    };

    const long long n_chunks = n_pad / 32;
    uint4 g0 = make_uint4(0, 0, 0, 0), g1 = g0;
    if (n_chunks > 0) {
        const long long j = ((n_chunks - 1) * 32 + lane) * 2;
        g0 = __ldg(R + j); g1 = __ldg(R + j + 1);
    }
    for (long long c = n_chunks - 1; c >= 0; --c) {
        const uint32_t buf = rec_base + (uint32_t)(c & 1) * 1024u;
        sts128(buf + lane * 32, g0);
        sts128(buf + lane * 32 + 16, g1);
        if (c > 0) { // the next (earlier) chunk's loads fly while this one is coded
            const long long j = ((c - 1) * 32 + lane) * 2;
            g0 = __ldg(R + j); g1 = __ldg(R + j + 1);
        }
        drain(); // also orders the staging stores above before the reads below
        uint4 a = lds128(buf + 31 * 32), b = lds128(buf + 31 * 32 + 16);
#pragma unroll 8
        for (int k = 31; k >= 0; --k) {
            const uint32_t nk = (uint32_t)(k > 0 ? k - 1 : 0) * 32;
            const uint4 na = lds128(buf + nk), nb = lds128(buf + nk + 16); // next record, ahead of its use
            if (b.z) st = enc_escape(st, b.w, b.z & 15u);
            const bool emit = (uint32_t)(st.x >> 32) >= a.z;
            if (emit) { sts32(st.out, (uint32_t)st.x); st.out += 4; st.x >>= 32; }
            const uint64_t rcp = ((uint64_t)a.y << 32) | a.x;
            const uint64_t q = __umul64hi(st.x, rcp) >> b.x;
            st.x = st.x + b.y + q * (uint64_t)a.w;
            a = na; b = nb;
        }
    }
    // Rans64EncFlush: ptr -= 2; ptr[0] = lo; ptr[1] = hi  => hi is the "earlier" emitted word
    sts32(st.out, (uint32_t)(st.x >> 32));
    sts32(st.out + 4, (uint32_t)st.x);
    st.out += 8;
    drain();
    overflow = __any_sync(0xffffffffu, overflow);
    if (lane == 0) {
        const int32_t stat = status[s];
        sizes[s] = stat < 0 ? stat : (overflow ? ICM_ERR_CAPACITY : (int32_t)(total * 4));
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

// flag = min(flag, min(values)): folds encoder sizes / decoder statuses (negative = ICM_ERR_*) into one word that the
// host reads once per round trip instead of once per stream
__global__ void __launch_bounds__(256) min_i32_kernel(const int32_t *__restrict__ v, long long n, int32_t *__restrict__ flag)
{
    int32_t m = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = min(m, v[i]);
    for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m < 0) atomicMin(flag, m);
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
// fallback decoder (tables that do not fit the lane image): one WARP per stream, CDF rows with 32-bit entries in shared
// memory, speculative 32-entry window + ballot search.  Kept from round 1; see the lane kernel below for the fast path.
struct WDecState {
    uint32_t xl, xh;      // rANS state
    uint32_t pos;         // next stream word
    uint32_t wcur, wnxt;  // lane l: words (pos & ~31) + l and + 32 + l
    uint32_t wv;          // the next stream word, broadcast, always ready
};

struct DecCtx {
    const uint32_t *W;
    uint32_t nwords;
    int lane;
};

__device__ __forceinline__ uint32_t dec_load_block(const DecCtx &c, uint32_t block)
{
    const uint32_t j = block * 32 + c.lane;
    return j < c.nwords ? __ldg(c.W + j) : 0u; // past-the-end reads are UB in the reference; we feed zeros
}

// called after pos was incremented
__device__ __forceinline__ void dec_after_consume(WDecState &d, const DecCtx &c)
{
    if ((d.pos & 31u) == 0) { d.wcur = d.wnxt; d.wnxt = dec_load_block(c, (d.pos >> 5) + 1); }
    d.wv = __shfl_sync(0xffffffffu, d.wcur, (int)(d.pos & 31u));
}

__device__ __forceinline__ uint32_t dec_get4(WDecState &d, const DecCtx &c)
{ // Rans64DecGetBits, n_bits = 4
    const uint32_t val = d.xl & 15u;
    d.xl = (d.xl >> 4) | (d.xh << 28);
    d.xh >>= 4;
    if (d.xh == 0 && d.xl < 0x80000000u) { d.xh = d.xl; d.xl = d.wv; ++d.pos; dec_after_consume(d, c); }
    return val;
}

// everything that is not the common path: cum outside the speculative window, an escape symbol, or the
// 32-word stream window running out.  Returns the decoded value.
struct SlowArgs {
    uint32_t cum, xs_lo, xs_hi; // from the state before the symbol
    uint32_t p;                 // entries <= cum inside the speculative window
    uint32_t row_addr, size, s0, lut_addr, lut_shift;
    int offset;
};

struct SlowRet {
    WDecState d;
    int value;
};

// by value in, by value out: keeps the caller's state in registers (a reference would pin it to local memory)
__device__ __noinline__ SlowRet dec_slow(WDecState d, DecCtx c, SlowArgs a, uint32_t nxl, uint32_t nxh)
{
    uint32_t s0 = a.s0, p = a.p;
    if (p - 1u >= 31u) { // general search: per-table bucket table, then windows until one brackets cum
        s0 = lds16(a.lut_addr + ((a.cum >> a.lut_shift) << 1));
        uint32_t e, m;
        while (true) {
            e = lds32(a.row_addr + ((s0 + c.lane) << 2));
            m = __ballot_sync(0xffffffffu, e <= a.cum);
            if (m != 0xffffffffu) break;
            s0 += 31; // more than 31 symbols share this bucket
        }
        p = __popc(m); // >= 1: entry s0 is <= cum by construction of the bucket table
        const uint32_t start = __shfl_sync(0xffffffffu, e, (int)p - 1);
        const uint32_t next = __shfl_sync(0xffffffffu, e, (int)p);
        const uint32_t f = next - start;
        const uint64_t nx = (uint64_t)f * a.xs_lo + (a.cum - start) + ((uint64_t)(f * a.xs_hi) << 32);
        nxl = (uint32_t)nx; nxh = (uint32_t)(nx >> 32);
    }
    // Rans64DecAdvance tail
    d.xl = nxl; d.xh = nxh;
    if (d.xh == 0 && d.xl < 0x80000000u) { d.xh = d.xl; d.xl = d.wv; ++d.pos; dec_after_consume(d, c); }
    const int symbol = (int)(s0 + p) - 1;
    const int max_value = (int)a.size - 2;
    int value = symbol;
    if (symbol == max_value) { // escape: count nibble(s), then the payload nibbles, LSB first
        int val = (int)dec_get4(d, c);
        int nb = val;
        while (val == 15) { val = (int)dec_get4(d, c); nb += val; }
        int raw = 0;
        for (int j = 0; j < nb; ++j) { val = (int)dec_get4(d, c); raw |= val << ((j * 4) & 31); }
        value = raw >> 1;
        if (raw & 1) value = -value - 1; else value += max_value;
    }
    SlowRet r;
    r.d = d;
    r.value = value + a.offset;
    return r;
}

constexpr int kDecWarps = 16; // max streams per CTA: they share one copy of the tables in shared memory; each warp is
                              // latency-bound (one instruction every ~4 cycles), so four per scheduler still interleave

__global__ void __launch_bounds__(kDecWarps * 32) rans_decode_warp_kernel(TablesDev T, int n_streams, const uint32_t *__restrict__ words,
                                                         const int64_t *__restrict__ word_off,
                                                         const int64_t *__restrict__ nwords_arr,
                                                         uint64_t *__restrict__ state, int64_t *__restrict__ pos_arr,
                                                         const int32_t *__restrict__ idx, long long n_per_stream,
                                                         int32_t *__restrict__ out, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *s_cdf = reinterpret_cast<uint32_t *>(smem_raw);
    uint16_t *s_lut = reinterpret_cast<uint16_t *>(s_cdf + ((T.total_pad + 3) & ~3));
    uint4 *s_meta = reinterpret_cast<uint4 *>(s_lut + ((size_t)T.n_cdf << T.lut_bits)); // per table
    __shared__ __align__(16) uint4 s_sym_all[kDecWarps][64]; // per warp, per symbol: two chunks of 32
    __shared__ int32_t s_val_all[kDecWarps][32];
    // warp index broadcast from lane 0: tells the compiler it is warp-uniform (no divergence guards in the loop)
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    uint4 *s_sym = s_sym_all[wid];
    int32_t *s_val = s_val_all[wid];
    const uint32_t cdf_addr = smem_addr(s_cdf), lut_addr = smem_addr(s_lut), sym_addr = smem_addr(s_sym), val_addr = smem_addr(s_val);
    {
        const int tid = threadIdx.x, nthr = blockDim.x;
        const uint4 *g = reinterpret_cast<const uint4 *>(T.cdf_pad);
        uint4 *d = reinterpret_cast<uint4 *>(s_cdf);
        for (int i = tid; i < (T.total_pad + 3) / 4; i += nthr) d[i] = __ldg(g + i);
        const uint4 *gl = reinterpret_cast<const uint4 *>(T.lut);
        uint4 *dl = reinterpret_cast<uint4 *>(s_lut);
        for (int i = tid; i < (int)(((size_t)T.n_cdf << T.lut_bits) / 8); i += nthr) dl[i] = __ldg(gl + i);
        for (int i = tid; i < T.n_cdf; i += nthr) {
            // per table: {speculative window address, s0 + offset - 1, size - 1 - s0 (the p that means "escape"), t | s0 << 16}
            const int size = T.sizes[i], off = T.offsets[i], base = T.base[i];
            int s0 = 0;
            if (size > 32) { s0 = -off - 15; s0 = max(0, min(s0, size - 32)); }
            s_meta[i] = make_uint4(cdf_addr + (uint32_t)(base + s0) * 4u, (uint32_t)(s0 + off - 1), (uint32_t)(size - 1 - s0),
                                   (uint32_t)i | ((uint32_t)s0 << 16));
        }
    }
    __syncthreads();

    const int s = blockIdx.x * (int)(blockDim.x >> 5) + wid; // the host launches full CTAs only (no early exit:
                                                              // the compiler must see a converged warp)
    DecCtx ctx;
    ctx.W = words + word_off[s];
    ctx.nwords = (uint32_t)min((long long)nwords_arr[s], 0xFFFFFFFFLL);
    ctx.lane = lane;
    WDecState d;
    {
        const long long pos0 = pos_arr[s];
        if (pos0 < 0) { // set_stream: Rans64DecInit
            d.xl = ctx.nwords > 0 ? ctx.W[0] : 0u;
            d.xh = ctx.nwords > 1 ? ctx.W[1] : 0u;
            d.pos = 2;
        } else {
            const uint64_t x = state[s];
            d.xl = (uint32_t)x; d.xh = (uint32_t)(x >> 32);
            d.pos = (uint32_t)pos0;
        }
    }
    d.wcur = dec_load_block(ctx, d.pos >> 5);
    d.wnxt = dec_load_block(ctx, (d.pos >> 5) + 1);
    d.wv = __shfl_sync(0xffffffffu, d.wcur, (int)(d.pos & 31u));

    const int32_t *I = idx + (size_t)s * n_per_stream;
    int32_t *O = out + (size_t)s * n_per_stream;
    const uint32_t lut_shift = kPrecision - T.lut_bits;
    const long long n = n_per_stream;
    const long long n_chunks = (n + 31) / 32;
    const uint32_t lane4 = lane * 4;
    bool bad = false;
    auto stage = [&](long long c, int t) { // per-symbol metadata of chunk c into its half of s_sym
        if (t < 0 || t >= T.n_cdf) { bad = true; t = 0; }
        s_sym[(c & 1) * 32 + lane] = s_meta[t];
    };
    int ireg = (lane < n) ? __ldg(I + lane) : 0;
    s_sym[32 + lane] = s_meta[0]; // the look-ahead past the last symbol must still read a valid window address
    stage(0, ireg);
    ireg = (32 + lane < n) ? __ldg(I + 32 + lane) : 0;
    __syncwarp();
    // operands of the first symbol
    uint4 mc = lds128(sym_addr);
    uint32_t e = lds32(mc.x + lane4);
    uint32_t eprev = __shfl_up_sync(0xffffffffu, e, 1);
    uint32_t f = e - eprev;

    for (long long c = 0; c < n_chunks; ++c) {
        __syncwarp();
        if (c + 1 < n_chunks) { // stage the next chunk; its indexes were fetched one chunk ago
            stage(c + 1, ireg);
            const long long j = (c + 2) * 32 + lane;
            ireg = (j < n) ? __ldg(I + j) : 0;
        }
        __syncwarp();
        const int valid = (int)min(32LL, n - c * 32);
        const uint32_t sym_chunk = (uint32_t)(c & 1) * 512u;
#pragma unroll 2
        for (int k = 0; k < valid; ++k) {
            // ---- off the critical path: metadata of symbol k+1 (its half of s_sym was staged a chunk ago)
            const uint4 mn = lds128(sym_addr + ((sym_chunk + (uint32_t)(k + 1) * 16u) & 1023u));
            // ---- critical path
            const uint32_t cum = d.xl & 0xFFFFu;
            const uint32_t p = __popc(__ballot_sync(0xffffffffu, e <= cum)); // entries <= cum: a prefix of the lanes
            const uint32_t xs_lo = (d.xl >> 16) | (d.xh << 16), xs_hi = d.xh >> 16;
            // Rans64DecAdvance formed by every lane for the symbol that ENDS at its entry; lane p holds the real one
            const uint64_t nx = (uint64_t)f * xs_lo + (uint64_t)(cum - eprev) + ((uint64_t)(f * xs_hi) << 32);
            const uint32_t nxl = __shfl_sync(0xffffffffu, (uint32_t)nx, (int)p);
            const uint32_t nxh = __shfl_sync(0xffffffffu, (uint32_t)(nx >> 32), (int)p);
            const uint32_t en = lds32(mn.x + lane4); // speculative window of symbol k+1
            const bool rn = (nxh == 0) && (nxl < 0x80000000u);
            const uint32_t npos = d.pos + (rn ? 1u : 0u);
            int value = (int)(mc.y + p);
            const bool slow = (p - 1u >= 31u) | (p == mc.z) | (rn & ((npos & 31u) == 0));
            if (!slow) {
                d.xl = rn ? d.wv : nxl;
                d.xh = rn ? nxl : nxh;
                d.pos = npos;
                d.wv = __shfl_sync(0xffffffffu, d.wcur, (int)(npos & 31u));
            } else {
                SlowArgs a;
                a.cum = cum; a.xs_lo = xs_lo; a.xs_hi = xs_hi; a.p = p;
                const uint32_t t = mc.w & 0xFFFFu;
                a.s0 = mc.w >> 16;
                a.row_addr = mc.x - a.s0 * 4u;
                a.size = mc.z + 1u + a.s0;
                a.lut_addr = lut_addr + ((t << T.lut_bits) << 1);
                a.lut_shift = lut_shift;
                a.offset = (int)mc.y + 1 - (int)a.s0;
                const SlowRet r = dec_slow(d, ctx, a, nxl, nxh);
                d = r.d;
                value = r.value;
            }
            sts32(val_addr + (uint32_t)k * 4u, (uint32_t)value);
            // ---- operands of the next symbol
            mc = mn;
            e = en;
            eprev = __shfl_up_sync(0xffffffffu, en, 1);
            f = en - eprev;
        }
        __syncwarp();
        if (lane < valid) O[c * 32 + lane] = s_val[lane];
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        state[s] = ((uint64_t)d.xh << 32) | d.xl;
        pos_arr[s] = (int64_t)d.pos;
        if (bad) status[s] = ICM_ERR_BAD_INDEX;
    }
}


// ------------------------------------------------------------------------------------------------
// bucket decoder: one warp per stream, up to kDecMaxWarps streams per CTA share one copy of the image
constexpr int kDecMaxWarps = 16;

// warp-cooperative: load 32-word blocks until at least kRefillBelow words are ahead of position p; returns `loaded`
__device__ __forceinline__ uint32_t bucket_refill(const lane::DecConst &c, uint32_t p, uint32_t ld, int lane_id)
{
    const lane::Smem sm{};
    __syncwarp();
    while ((int)(ld - p) < lane::kRefillBelow) { lane::ring_load_block(sm, c, ld, lane_id); ld += 32; }
    __syncwarp();
    return ld;
}

// the out-of-line path of a symbol (crowded bucket / escape).  By value in, by value out: keeps the caller's state
// in registers (a reference would pin it to local memory), and one copy keeps the unrolled serial loop compact.
struct RareRet {
    lane::WarpDec d;
    uint32_t pos, loaded;
    int value;
};
__device__ __noinline__ RareRet bucket_rare(lane::DecConst c, lane::WarpDec d, uint32_t pos, uint32_t loaded, uint32_t a, uint32_t base,
                                            uint32_t maxv, uint32_t rowinfo, lane::u4 E, int lane_id)
{
    const lane::Smem sm{};
    auto refill = [&](uint32_t p, uint32_t &ld) { ld = bucket_refill(c, p, ld, lane_id); };
    RareRet r;
    r.value = lane::dec_rare(sm, c, d, pos, loaded, refill, a, base, maxv, rowinfo, E);
    r.d = d; r.pos = pos; r.loaded = loaded;
    return r;
}

__global__ void __launch_bounds__(kDecMaxWarps * 32) rans_decode_bucket_kernel(
    TablesDev T, int n_streams, const uint32_t *__restrict__ words, const int64_t *__restrict__ word_off,
    const int64_t *__restrict__ nwords_arr, uint64_t *__restrict__ state, int64_t *__restrict__ pos_arr,
    const int32_t *__restrict__ idx, long long n_per_stream, int32_t *__restrict__ out, int32_t *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t raw_addr = smem_addr(smem_raw);
    const uint32_t base = (raw_addr + lane::kAlign - 1) & ~(lane::kAlign - 1);
    {
        uint4 *d = reinterpret_cast<uint4 *>(smem_raw + (base - raw_addr));
        for (uint32_t i = threadIdx.x; i < T.image_bytes / 16; i += blockDim.x) d[i] = __ldg(T.image + i);
    }
    __syncthreads();
    // warp index broadcast from lane 0: tells the compiler it is warp-uniform (no divergence guards in the loop)
    const int lane_id = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int s = blockIdx.x * (int)(blockDim.x >> 5) + wid;
    if (s >= n_streams) return;

    const lane::Smem sm{};
    lane::DecConst c;
    c.rs = T.rs; c.M = T.M;
    c.ring = base + ((T.image_bytes + 15u) & ~15u) + (uint32_t)wid * lane::kWarpBytes;
    c.W = words + word_off[s];
    c.nwords = (uint32_t)min((long long)nwords_arr[s], 0xFFFFFFFFLL);
    const uint32_t stage = c.ring + lane::kRingBytes, outs = stage + lane::kStageBytes, meta_addr = base + T.meta_off,
                   row_addr = base + T.row_off;
    const int32_t *I = idx + (size_t)s * n_per_stream;
    int32_t *O = out + (size_t)s * n_per_stream;
    const long long n = n_per_stream, n_chunks = (n + 31) / 32;
    const int n_cdf = T.n_cdf;
    bool bad = false;

    // per-symbol table records of chunk ch (parity ch & 1); slot 32 of the other parity = its first symbol
    auto stage_chunk = [&](long long ch, int t) {
        if ((unsigned)t >= (unsigned)n_cdf) { bad = true; t = 0; } // the reference has only a compiled-out assert here (UB)
        lane::u4 r = sm.ld128(meta_addr + 16u * (uint32_t)t);
        r.x += base;
        const uint32_t par = (uint32_t)(ch & 1) * (33u * 16u);
        sm.st128(stage + par + 16u * (uint32_t)lane_id, r);
        if (lane_id == 0) sm.st128(stage + (33u * 16u - par) + 32u * 16u, r);
    };
    uint32_t pos, loaded;
    auto refill = [&](uint32_t p, uint32_t &ld) { return bucket_refill(c, p, ld, lane_id); };
    lane::WarpDec d;
    {
        const long long pos0 = pos_arr[s];
        if (pos0 < 0) { // set_stream: Rans64DecInit
            d.xl = c.nwords > 0 ? __ldg(c.W) : 0u;
            d.xh = c.nwords > 1 ? __ldg(c.W + 1) : 0u;
            pos = 2;
        } else {
            const uint64_t x = state[s];
            d.xl = (uint32_t)x; d.xh = (uint32_t)(x >> 32);
            pos = (uint32_t)pos0;
        }
    }
    loaded = pos & ~31u;
    int ireg = lane_id < n ? __ldg(I + lane_id) : 0;
    stage_chunk(0, ireg);
    ireg = 32 + lane_id < n ? __ldg(I + 32 + lane_id) : 0;
    loaded = refill(pos, loaded);
    sm.ld64(lane::ring_slot(c, pos), d.wv, d.awv);
    d.wa1 = lane::ring_slot(c, pos + 1);
    uint32_t wa1_base = d.wa1;
    lane::u4 mcur = sm.ld128(stage);
    uint32_t a = (__funnelshift_r(d.xl, d.xl, c.rs) & c.M) | mcur.x;

    for (long long ch = 0; ch < n_chunks; ++ch) {
        __syncwarp();
        if (ch + 1 < n_chunks) { // stage the next chunk; its indexes were fetched one chunk ago
            stage_chunk(ch + 1, ireg);
            const long long j = (ch + 2) * 32 + lane_id;
            ireg = j < n ? __ldg(I + j) : 0;
        }
        pos += (d.wa1 - wa1_base) >> 3;
        if ((int)(loaded - pos) < lane::kRefillBelow) loaded = refill(pos, loaded);
        d.wa1 = lane::ring_slot(c, pos + 1);
        wa1_base = d.wa1;
        __syncwarp();
        const int valid = (int)min(32LL, n - ch * 32);
        const uint32_t sbase = stage + (uint32_t)(ch & 1) * (33u * 16u);
        int k = 0;
        while (k < valid) {
            // straight run of common-path symbols; a symbol that needs the out-of-line path leaves the run, so the
            // common path is fall-through code with one not-taken forward branch per symbol
#pragma unroll 8
            for (; k < valid; ++k) {
                const lane::u4 mn = sm.ld128(sbase + 16u * (uint32_t)(k + 1)); // next symbol's table (slot 32 = next chunk's first)
                const lane::u4 E = sm.ld128_ro(a);
                int value;
                if (!lane::dec_fast(sm, c, d, a, E, (int32_t)mcur.z, mn.x, value)) break;
                sm.st32(outs + 4u * (uint32_t)k, (uint32_t)value);
                mcur = mn;
            }
            if (k < valid) { // escape or crowded bucket
                const lane::u4 mn = sm.ld128(sbase + 16u * (uint32_t)(k + 1));
                pos += (d.wa1 - wa1_base) >> 3;
                int value;
                if (!lane::dec_escape_simple(sm, c, d, pos, loaded, mcur.y, mcur.w >> 16, value)) {
                    const lane::u4 E = sm.ld128_ro(a);
                    const RareRet r = bucket_rare(c, d, pos, loaded, a, base, mcur.y, row_addr + 8u * (mcur.w & 0xFFFFu), E, lane_id);
                    d = r.d; pos = r.pos; loaded = r.loaded;
                    value = r.value;
                }
                wa1_base = d.wa1;
                a = (__funnelshift_r(d.xl, d.xl, c.rs) & c.M) | mn.x;
                sm.st32(outs + 4u * (uint32_t)k, (uint32_t)(value + (int32_t)mcur.z));
                mcur = mn;
                ++k;
            }
        }
        __syncwarp();
        if (lane_id < valid) O[ch * 32 + lane_id] = (int32_t)sm.ld32(outs + 4u * (uint32_t)lane_id);
    }
    pos += (d.wa1 - wa1_base) >> 3;
    bad = __any_sync(0xffffffffu, bad);
    if (lane_id == 0) {
        state[s] = ((uint64_t)d.xh << 32) | d.xl;
        pos_arr[s] = (int64_t)pos;
        if (bad) status[s] = ICM_ERR_BAD_INDEX;
    }
}

}  // namespace icm

// =================================================================================================
// C ABI
using namespace icm;

static int current_device() { return current_device_ordinal(); }

extern "C" int icm_tables_create(const int32_t *h_cdfs, int n_cdf, int stride, const int32_t *h_sizes,
                                 const int32_t *h_offsets, icm_tables **out)
{
    ICM_CHECK_ARG(h_cdfs && h_sizes && h_offsets && out, "icm_tables_create: null argument");
    ICM_CHECK_ARG(n_cdf > 0 && n_cdf < 65536 && stride >= 3, "icm_tables_create: bad shape n_cdf=%d stride=%d", n_cdf, stride);
    std::vector<int32_t> base(n_cdf);
    int total = 0;
    for (int t = 0; t < n_cdf; ++t) {
        const int size = h_sizes[t];
        ICM_CHECK_ARG(size >= 3 && size <= stride && size < 65536, "icm_tables_create: cdf_size[%d]=%d outside [3,%d]", t, size, stride);
        const int32_t *c = h_cdfs + (size_t)t * stride;
        ICM_CHECK_ARG(c[0] == 0 && c[size - 1] == (1 << kPrecision), "icm_tables_create: row %d is not a 16-bit CDF", t);
        for (int j = 0; j + 1 < size; ++j)
            ICM_CHECK_ARG(c[j] < c[j + 1], "icm_tables_create: row %d not strictly increasing at %d", t, j);
        base[t] = total;
        total += size + kRowPad;
    }
    // bucket decoder image: what fits under the 227 KB of dynamic shared memory of one CTA beside 8 warps' work areas
    const size_t image_budget = 226 * 1024 - lane::kAlign - 8 * lane::kWarpBytes;
    lane::Image image = lane::build_image(h_cdfs, n_cdf, stride, h_sizes, h_offsets, image_budget);
    // fallback decoder tables (rows with 32-bit entries + bucket table); its kernel also has 18 KB of static shared memory
    int lut_bits = 8;
    auto smem_need = [&](int bits) {
        return (size_t)((total + 3) & ~3) * 4 + ((size_t)n_cdf << bits) * 2 + (size_t)n_cdf * 16;
    };
    const size_t legacy_budget = (227 - 19) * 1024;
    while (lut_bits > 3 && smem_need(lut_bits) > legacy_budget) --lut_bits;
    const bool legacy_ok = smem_need(lut_bits) <= legacy_budget;
    ICM_CHECK_ARG(image.ok || legacy_ok, "icm_tables_create: tables too large for shared memory");
    const size_t lut_n = legacy_ok ? (size_t)n_cdf << lut_bits : 0;
    std::vector<uint32_t> cdf_pad(legacy_ok ? (((size_t)total + 3) & ~(size_t)3) : 0, kSentinel);
    std::vector<uint16_t> lut(lut_n);
    for (int t = 0; legacy_ok && t < n_cdf; ++t) {
        const int size = h_sizes[t];
        const int32_t *c = h_cdfs + (size_t)t * stride;
        for (int j = 0; j < size; ++j) cdf_pad[base[t] + j] = (uint32_t)c[j];
        int sidx = 0;
        for (int b = 0; b < (1 << lut_bits); ++b) {
            const int32_t lo = b << (kPrecision - lut_bits);
            while (sidx + 1 <= size - 2 && c[sidx + 1] <= lo) ++sidx;
            lut[((size_t)t << lut_bits) + b] = (uint16_t)sidx;
        }
    }
    // one device blob: cdf32 | sizes | offsets | base | cdf_pad | lut | image
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t img_bytes = image.ok ? image.bytes.size() : 0;
    const size_t o_cdf32 = 0, o_sizes = al(o_cdf32 + (size_t)n_cdf * stride * 4), o_offsets = al(o_sizes + n_cdf * 4),
                 o_base = al(o_offsets + n_cdf * 4), o_pad = al(o_base + n_cdf * 4),
                 o_lut = al(o_pad + cdf_pad.size() * 4), o_img = al(o_lut + lut_n * 2), blob_bytes = al(o_img + img_bytes + 16);
    std::vector<unsigned char> host(blob_bytes, 0);
    memcpy(&host[o_cdf32], h_cdfs, (size_t)n_cdf * stride * 4);
    memcpy(&host[o_sizes], h_sizes, n_cdf * 4);
    memcpy(&host[o_offsets], h_offsets, n_cdf * 4);
    memcpy(&host[o_base], base.data(), n_cdf * 4);
    if (!cdf_pad.empty()) memcpy(&host[o_pad], cdf_pad.data(), cdf_pad.size() * 4);
    if (lut_n) memcpy(&host[o_lut], lut.data(), lut_n * 2);
    if (img_bytes) memcpy(&host[o_img], image.bytes.data(), img_bytes);
    icm_tables *T = new (std::nothrow) icm_tables();
    ICM_CHECK_ARG(T, "icm_tables_create: out of host memory");
    if (cudaGetDevice(&T->device) != cudaSuccess) { delete T; set_error("icm_tables_create: no CUDA device"); return ICM_ERR_NO_DEVICE; }
    cudaError_t e = cudaMalloc(&T->d_blob, blob_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(T->d_blob, host.data(), blob_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("icm_tables_create: %s", cudaGetErrorString(e)); if (T->d_blob) cudaFree(T->d_blob); delete T; return ICM_ERR_CUDA; }
    char *b = (char *)T->d_blob;
    T->dev = TablesDev{n_cdf, stride, lut_bits, total,
                       (const int32_t *)(b + o_cdf32), (const int32_t *)(b + o_sizes), (const int32_t *)(b + o_offsets),
                       legacy_ok ? (const uint32_t *)(b + o_pad) : nullptr, (const int32_t *)(b + o_base),
                       legacy_ok ? (const uint16_t *)(b + o_lut) : nullptr,
                       image.ok ? (const uint4 *)(b + o_img) : nullptr, (uint32_t)img_bytes, image.meta_off, image.row_off, image.rs, image.M};
    T->smem_bytes = legacy_ok ? smem_need(lut_bits) : 0;
    T->image_smem_bytes = image.ok ? ((img_bytes + 15) & ~(size_t)15) + lane::kAlign : 0;
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
    // worst case per symbol: 16 bits + escape (count nibble + 8 payload nibbles) = 52 bits; + final state
    long long w = (n * 52 + 31) / 32 + 2;
    return (w + 31) & ~31LL;
}

struct EncLayout { size_t rec, words, offs, status, total; long long cap_words, n_pad; };
static EncLayout enc_layout(int n_streams, long long n)
{
    EncLayout L;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    L.n_pad = (n + 31) & ~31LL;
    L.cap_words = enc_cap_words(n);
    L.rec = 0;
    L.words = al(L.rec + (size_t)n_streams * L.n_pad * sizeof(Record));
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
    ICM_CHECK_ARG(t->device == current_device(), "icm_rans_encode_batch: tables were created on device %d", t->device);
    cudaStream_t st = as_stream(stream);
    // profiling aid: what does the pipeline do without the coders?  ICM_DEBUG_SKIP_CODERS=enc / dec skips one side only
    static const bool skip = getenv("ICM_DEBUG_SKIP_CODERS") != nullptr && strcmp(getenv("ICM_DEBUG_SKIP_CODERS"), "dec") != 0;
    if (skip) { ICM_CUDA(cudaMemsetAsync(d_sizes, 0, (size_t)(n_streams + 1) * 4, st)); return ICM_OK; }
    const EncLayout L = enc_layout(n_streams, n_per_stream);
    char *w = (char *)d_work;
    Record *rec = (Record *)(w + L.rec);
    uint32_t *words = (uint32_t *)(w + L.words);
    long long *offs = (long long *)(w + L.offs);
    int32_t *status = (int32_t *)(w + L.status);
    ICM_CUDA(cudaMemsetAsync(status, 0, (size_t)n_streams * 4, st));
    const long long nt = (long long)n_streams * L.n_pad;
    if (nt > 0) {
        const int grid = (int)min((nt + 255) / 256, (long long)sm_count() * 16);
        rans_records_kernel<<<grid, 256, 0, st>>>(t->dev, d_symbols, d_indexes, n_streams, n_per_stream, L.n_pad, rec, status);
        ICM_LAUNCH_CHECK();
    }
    static const int enc_warps = [] { const char *e = getenv("ICM_ENC_WARPS"); const int v = e ? atoi(e) : kEncWarpsMax; return v >= 1 && v <= kEncWarpsMax ? v : kEncWarpsMax; }();
    rans_encode_kernel<<<(n_streams + enc_warps - 1) / enc_warps, 32 * enc_warps, (size_t)enc_warps * kEncWarpBytes, st>>>(rec, L.n_pad, words, L.cap_words, d_sizes, status, n_streams);
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
    d->device = current_device();
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

// Streams (warps) per decoder CTA: 0 = automatic, else a power of two <= 16.  One stream per CTA is fastest per stream
// (a whole SM to itself); more streams per CTA share one copy of the tables in shared memory and leave more SMs to
// other kernels -- what a pipeline that overlaps the decoder with the convolutions wants.  kernel: 0 = bucket kernel
// when the tables fit, 1 = force the round-1 warp-search kernel (A/B tests).
static thread_local int g_dec_warps = 0, g_dec_kernel = 0;
extern "C" int icm_set_decoder_layout(int streams_per_cta, int kernel)
{
    ICM_CHECK_ARG(streams_per_cta >= 0 && streams_per_cta <= 16 && (streams_per_cta & (streams_per_cta - 1)) == 0,
                  "icm_set_decoder_layout: streams_per_cta %d is not 0, 1, 2, 4, 8 or 16", streams_per_cta);
    ICM_CHECK_ARG(kernel == 0 || kernel == 1, "icm_set_decoder_layout: kernel %d is not 0 or 1", kernel);
    g_dec_warps = streams_per_cta;
    g_dec_kernel = kernel;
    return ICM_OK;
}
extern "C" int icm_set_decoder_streams_per_cta(int n) { return icm_set_decoder_layout(n, g_dec_kernel); }

template <class K>
static int ensure_smem(K kernel, size_t bytes, PerDeviceSmem &configured)
{
    if (configured.needs(bytes)) {
        ICM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured.done(bytes);
    }
    return ICM_OK;
}

extern "C" int icm_rans_decoder_step(icm_rans_decoder *d, const icm_tables *t, const int32_t *d_indexes,
                                     int64_t n_per_stream, int32_t *d_out, void *stream)
{
    ICM_CHECK_ARG(d && t && d->d_words, "icm_rans_decoder_step: decoder has no stream (call set_streams first)");
    ICM_CHECK_ARG(n_per_stream >= 0 && (n_per_stream == 0 || (d_indexes && d_out)), "icm_rans_decoder_step: bad arguments");
    const int dev = current_device();
    ICM_CHECK_ARG(t->device == dev && d->device == dev, "icm_rans_decoder_step: tables (device %d) / decoder (device %d) used on device %d",
                  t->device, d->device, dev);
    if (n_per_stream == 0) return ICM_OK;
    static const bool skip = getenv("ICM_DEBUG_SKIP_CODERS") != nullptr && strcmp(getenv("ICM_DEBUG_SKIP_CODERS"), "enc") != 0;
    if (skip) { ICM_CUDA(cudaMemsetAsync(d_out, 0, (size_t)d->n_streams * n_per_stream * 4, as_stream(stream))); return ICM_OK; }
    const int S = d->n_streams;
    int warps = g_dec_warps;
    if (warps == 0) warps = S <= sm_count() / 2 ? 1 : (S <= sm_count() ? 2 : 4);
    if (t->image_smem_bytes && !(g_dec_kernel == 1 && t->smem_bytes)) {
        const size_t room = (size_t)227 * 1024 - t->image_smem_bytes;
        while (warps > 1 && (size_t)warps * lane::kWarpBytes > room) warps >>= 1;
        const size_t smem = t->image_smem_bytes + (size_t)warps * lane::kWarpBytes;
        static PerDeviceSmem conf_b;
        if (int rc = ensure_smem(rans_decode_bucket_kernel, smem, conf_b)) return rc;
        rans_decode_bucket_kernel<<<(S + warps - 1) / warps, warps * 32, smem, as_stream(stream)>>>(
            t->dev, S, d->d_words, d->d_word_off, d->d_nwords, d->d_state, d->d_pos, d_indexes, n_per_stream, d_out, d->d_status);
        ICM_LAUNCH_CHECK();
        return ICM_OK;
    }
    ICM_CHECK_ARG(t->smem_bytes, "icm_rans_decoder_step: these tables only fit the bucket decoder");
    static PerDeviceSmem conf_w;
    if (int rc = ensure_smem(rans_decode_warp_kernel, t->smem_bytes, conf_w)) return rc;
    while (S % warps) warps >>= 1; // full CTAs only
    rans_decode_warp_kernel<<<(S + warps - 1) / warps, warps * 32, t->smem_bytes, as_stream(stream)>>>(
        t->dev, S, d->d_words, d->d_word_off, d->d_nwords, d->d_state, d->d_pos, d_indexes, n_per_stream, d_out, d->d_status);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_min_i32(const int32_t *d_values, int64_t n, int32_t *d_flag, void *stream)
{
    ICM_CHECK_ARG(d_values && d_flag && n >= 0, "icm_min_i32: bad arguments");
    if (n == 0) return ICM_OK;
    min_i32_kernel<<<(unsigned)min((long long)((n + 255) / 256), 64LL), 256, 0, as_stream(stream)>>>(d_values, n, d_flag);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_rans_decoder_status_min(icm_rans_decoder *d, int32_t *d_flag, void *stream)
{
    ICM_CHECK_ARG(d && d_flag, "icm_rans_decoder_status_min: null argument");
    return icm_min_i32(d->d_status, d->n_streams, d_flag, stream);
}

extern "C" int icm_rans_decoder_status(icm_rans_decoder *d, int32_t *h_status, void *stream)
{
    ICM_CHECK_ARG(d && h_status, "icm_rans_decoder_status: null argument");
    ICM_CUDA(cudaMemcpyAsync(h_status, d->d_status, (size_t)d->n_streams * 4, cudaMemcpyDeviceToHost, as_stream(stream)));
    ICM_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return ICM_OK;
}
