// Bucket-table rANS decoder core (R4 of SURVEY.md §8a; ryg rans64.h:107-142; ans.so@0x7c40/0x7ce0) and the host-side
// builders of its tables.
//
// A stream is a strictly serial recurrence and ONE WARP carries it (all lanes execute the recurrence redundantly, so
// control flow is warp-uniform and shared-memory reads are broadcasts); the other 31 lanes exist to move data in bulk:
// per 32-symbol chunk they turn the chunk's CDF indexes into per-symbol table records, refill a ring of stream words
// and write the decoded chunk back, all coalesced.  What is left on the serial path per symbol is ~36 instructions
// with one rarely-taken branch (measured on B200: a lone warp pays ~3-5 cycles per INSTRUCTION, not per dependent
// step, so the instruction count of the loop is what sets a stream's speed):
//     LDS.128 bucket entry -> 3 compares -> 2 selects -> extract -> subtract -> IMAD.WIDE -> renorm select
// The bucket table ("image") is built on the host:
//   per CDF table a block of 2^K entries of 16 bytes, indexed by the top K bits of cum = x & 0xFFFF (K is the same
//   for all tables of a set -- the largest that fits shared memory, 7 for the 64 Gaussian tables -- so the rotate
//   amount and mask are kernel constants).  An entry lists the first three symbols that intersect the bucket as
//   w_i = start_i << 16 | freq_i (unused = ~0) and  w3 = start_3 << 16 | s0  (start_3 = 0xFFFF when the bucket has
//   <= 3 symbols; s0 = index of the first symbol).  With cumhi = cum << 16 | 0xFFFE the test "cum >= start_i" is the
//   single unsigned compare cumhi >= w_i (freq <= 0xFFFE is checked at build time), so the symbol is resolved in
//   registers from one shared-memory load -- no ballot / popc / shuffle on the chain.  Buckets with more than three
//   symbols (tails of wide tables) and escapes (symbol == max_value: count nibble(s) + payload nibbles) leave the
//   common path through ONE test: the escape symbol is never listed as a candidate, and start_3 is the start of the first
//   unlisted symbol of the bucket, so "cum >= start_3" catches both.  Escapes with a single count nibble are finished
//   inline (dec_escape_simple); the rest goes to an out-of-line function (dec_rare): a binary search over the 16-bit CDF
//   row bounded by the next bucket's s0, then the nibbles.
//   The entry address of the NEXT symbol is formed from both renormalisation outcomes before the outcome is known
//   ((rotr(v, rs) & M) | ent), so the renorm compare overlaps it; the stream-word ring stores each word together with
//   its rotated-and-masked form.
//
// Everything marked ICM_HD also compiles for the host: tests/host_sim builds the same per-symbol code for the CPU
// and checks it bit for bit against the pinned oracle without a GPU (test infrastructure only; the product library
// never runs it on the host).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#if defined(__CUDACC__)
#define ICM_HD __host__ __device__ __forceinline__
#else
#define ICM_HD inline
#endif

namespace icm {
namespace lane {

constexpr uint32_t kAlign = 4096;      // the image base is aligned to the largest entry block (2^8 * 16 B)
constexpr int kMaxBits = 8;
constexpr uint32_t kNoStart = 0xFFFFu; // start_3 of a bucket with <= 3 symbols
constexpr uint32_t kUnused = 0xFFFFFFFFu;

struct u4 { uint32_t x, y, z, w; };

// ------------------------------------------------------------------------------------------------
// host: shared-memory image of the decoder tables
struct Image {
    std::vector<unsigned char> bytes; // copied verbatim to shared memory (at a kAlign-aligned address)
    uint32_t meta_off = 0;            // u4 per table: {entry block offset, max_value, offset, t | escape-symbol start << 16}
    uint32_t row_off = 0;             // 2 x u32 per table: {row offset (bytes), nsym}
    int n_cdf = 0;
    bool ok = false;                  // false: the tables do not fit / a frequency of 65535 -> legacy kernel
    std::vector<int> bits;            // bucket bits per table (all equal to K)
    int K = 0;
    uint32_t rs = 0, M = 0;           // entry address = (rotr(v, rs) & M) | block address
};

// fraction (in cum counts) of a table left to the binary-search path with 2^k buckets
inline uint32_t unresolved_mass(const int32_t *c, int size, int k)
{
    const int nsym = size - 1;
    const uint32_t W = 65536u >> k;
    uint32_t un = 0;
    int s0 = 0;
    for (uint32_t b = 0; b < (1u << k); ++b) {
        const uint32_t lo = b * W, hi = lo + W;
        while (s0 + 1 < nsym && (uint32_t)c[s0 + 1] <= lo) ++s0;
        if (s0 + 3 < nsym && (uint32_t)c[s0 + 3] < hi) un += hi - std::max<uint32_t>((uint32_t)c[s0 + 3], lo);
    }
    return un;
}

inline Image build_image(const int32_t *cdfs, int n_cdf, int stride, const int32_t *sizes, const int32_t *offsets,
                         size_t budget_bytes)
{
    Image im;
    im.n_cdf = n_cdf;
    im.bits.assign(n_cdf, 0);
    size_t rows_bytes = 0;
    for (int t = 0; t < n_cdf; ++t) {
        const int32_t *c = cdfs + (size_t)t * stride;
        for (int j = 0; j + 1 < sizes[t]; ++j)
            if (c[j + 1] - c[j] > 0xFFFE) return im; // cumhi trick needs freq <= 0xFFFE
        if (sizes[t] - 2 > 0xFFFE) return im;
        rows_bytes += ((size_t)sizes[t] * 2 + 3) & ~(size_t)3;
    }
    const size_t fixed = (size_t)n_cdf * 16 + (size_t)n_cdf * 8 + rows_bytes;
    if (fixed + (size_t)n_cdf * 16 > budget_bytes) return im;
    // the same number of bucket bits for every table: the largest that fits (rotate amount and mask become constants)
    int K = 0;
    while (K < kMaxBits && fixed + ((size_t)n_cdf * 16 << (K + 1)) <= budget_bytes) ++K;
    for (int t = 0; t < n_cdf; ++t) im.bits[t] = K;
    im.K = K;
    // layout: entry blocks by descending size (each naturally aligned), then meta, row info, rows
    std::vector<int> order(n_cdf);
    for (int t = 0; t < n_cdf; ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return im.bits[a] > im.bits[b]; });
    std::vector<uint32_t> ent_off(n_cdf);
    uint32_t off = 0;
    for (int t : order) { ent_off[t] = off; off += 16u << im.bits[t]; }
    im.meta_off = off; off += (uint32_t)n_cdf * 16;
    im.row_off = off; off += (uint32_t)n_cdf * 8;
    std::vector<uint32_t> row_addr(n_cdf);
    for (int t = 0; t < n_cdf; ++t) { row_addr[t] = off; off += ((uint32_t)sizes[t] * 2 + 3) & ~3u; }
    im.bytes.assign((off + 15) & ~15u, 0);
    auto put32 = [&](uint32_t at, uint32_t v) { memcpy(&im.bytes[at], &v, 4); };
    for (int t = 0; t < n_cdf; ++t) {
        const int32_t *c = cdfs + (size_t)t * stride;
        const int size = sizes[t], nsym = size - 1, k = im.bits[t];
        const uint32_t W = 65536u >> k;
        int s0 = 0;
        for (uint32_t b = 0; b < (1u << k); ++b) {
            const uint32_t lo = b * W, hi = lo + W;
            while (s0 + 1 < nsym && (uint32_t)c[s0 + 1] <= lo) ++s0;
            const uint32_t e = ent_off[t] + b * 16;
            // candidates: the first three symbols that intersect the bucket, never the escape symbol (the table's last)
            int listed = 0;
            for (int i = 0; i < 3; ++i) {
                uint32_t w = kUnused;
                if (listed == i && s0 + i < nsym - 1 && (i == 0 || (uint32_t)c[s0 + i] < hi)) {
                    w = ((uint32_t)c[s0 + i] << 16) | (uint32_t)(c[s0 + i + 1] - c[s0 + i]);
                    ++listed;
                }
                put32(e + 4 * i, w);
            }
            // start of the first symbol of this bucket that is not listed: the 4th, or the escape symbol -- everything at or
            // above it takes the out-of-line path, so "cum >= start_3" is the ONLY test on the common path
            uint32_t st3 = kNoStart;
            const int nxt = s0 + listed;
            if (nxt < nsym && (listed == 0 || (uint32_t)c[nxt] < hi)) st3 = (uint32_t)c[nxt];
            put32(e + 12, (st3 << 16) | (uint32_t)s0);
        }
        // per-table record, copied per symbol by the staging lanes: {entry block, max_value, offset, t | escape start << 16}
        put32(im.meta_off + 16 * t + 0, ent_off[t]);
        put32(im.meta_off + 16 * t + 4, (uint32_t)(size - 2));
        put32(im.meta_off + 16 * t + 8, (uint32_t)offsets[t]);
        put32(im.meta_off + 16 * t + 12, (uint32_t)t | ((uint32_t)c[size - 2] << 16));
        put32(im.row_off + 8 * t + 0, row_addr[t]);
        put32(im.row_off + 8 * t + 4, (uint32_t)nsym);
        for (int j = 0; j < size; ++j) {
            const uint16_t v = (uint16_t)c[j]; // c[size-1] = 65536 wraps to 0: freq = (next - start) & 0xFFFF stays right
            memcpy(&im.bytes[row_addr[t] + 2 * j], &v, 2);
        }
    }
    im.rs = (uint32_t)(16 - K - 4) & 31u;
    im.M = ((1u << K) - 1u) << 4;
    im.ok = true;
    return im;
}

// ------------------------------------------------------------------------------------------------
// memory policies: shared memory through 32-bit shared-window addresses on the device, a byte array on the host
#if defined(__CUDA_ARCH__)
struct Smem {
    ICM_HD u4 ld128(uint32_t a) const
    {
        u4 v;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        return v;
    }
    ICM_HD u4 ld128_ro(uint32_t a) const
    { // read-only tables: not volatile, so the scheduler may move it
        u4 v;
        asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
        return v;
    }
    ICM_HD uint32_t ld32(uint32_t a) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
    ICM_HD uint32_t ld16(uint32_t a) const
    {
        uint32_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
        return v;
    }
    ICM_HD void ld64(uint32_t a, uint32_t &x, uint32_t &y) const { asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x), "=r"(y) : "r"(a)); }
    ICM_HD void st32(uint32_t a, uint32_t v) const { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
    ICM_HD void st64(uint32_t a, uint32_t x, uint32_t y) const { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
    ICM_HD void st128(uint32_t a, const u4 &v) const
    {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
};
#define ICM_LDG(p) __ldg(p)
#define ICM_ROTR(v, s) __funnelshift_r((v), (v), (s))
#define ICM_SHR64LO(lo, hi, s) __funnelshift_r((lo), (hi), (s))
#else
struct Smem {
    const unsigned char *base;
    u4 ld128(uint32_t a) const { u4 v; memcpy(&v, base + a, 16); return v; }
    u4 ld128_ro(uint32_t a) const { return ld128(a); }
    uint32_t ld32(uint32_t a) const { uint32_t v; memcpy(&v, base + a, 4); return v; }
    uint32_t ld16(uint32_t a) const { uint16_t v; memcpy(&v, base + a, 2); return v; }
    void ld64(uint32_t a, uint32_t &x, uint32_t &y) const { memcpy(&x, base + a, 4); memcpy(&y, base + a + 4, 4); }
    void st32(uint32_t a, uint32_t v) const { memcpy(const_cast<unsigned char *>(base) + a, &v, 4); }
    void st64(uint32_t a, uint32_t x, uint32_t y) const { st32(a, x); st32(a + 4, y); }
    void st128(uint32_t a, const u4 &v) const { memcpy(const_cast<unsigned char *>(base) + a, &v, 16); }
};
#define ICM_LDG(p) (*(p))
#define ICM_ROTR(v, s) ((uint32_t)(((uint64_t)(v) << 32 | (v)) >> ((s) & 31)))
#define ICM_SHR64LO(lo, hi, s) ((uint32_t)((((uint64_t)(hi) << 32) | (lo)) >> ((s) & 31)))
#endif

// ------------------------------------------------------------------------------------------------
// decoder, one stream per warp.  Shared-memory areas of a warp (addresses in the shared window):
//   ring   kRingSlots + kRingMirror slots of 8 bytes {word, (rotr(word, rs) & M)}; slot of word j = j & (kRingSlots-1);
//          the first kRingMirror slots are mirrored behind the ring so that a chunk can walk linearly
//   stage  2 x 33 per-symbol records {entry block address, max_value, offset, row info address} (slot 32 = first
//          symbol of the next chunk)
//   outs   32 decoded values
constexpr uint32_t kRingSlots = 128, kRingMirror = 40;
constexpr int kRefillBelow = 66;  // refill when fewer words than this are ahead: a chunk consumes <= 32 on the fast path
constexpr uint32_t kRingBytes = (kRingSlots + kRingMirror) * 8, kStageBytes = 2 * 33 * 16, kOutBytes = 32 * 4;
constexpr uint32_t kWarpBytes = kRingBytes + kStageBytes + kOutBytes; // 2528
constexpr int kMinAhead = 34; // words guaranteed ahead of the read position whenever the fast path runs

struct WarpDec {
    uint32_t xl, xh; // rANS state
    uint32_t wv;     // stream word at position pos (the next to be consumed)
    uint32_t awv;    // rotr(wv, rs) & M
    uint32_t wa1;    // ring address of word pos + 1 (may run into the mirror)
};

struct DecConst {
    uint32_t rs, M;       // entry address = (rotr(v, rs) & M) | block
    uint32_t ring;        // this warp's ring
    const uint32_t *W;    // the stream's words
    uint32_t nwords;
};

ICM_HD uint32_t ring_slot(const DecConst &c, uint32_t j) { return c.ring + ((j & (kRingSlots - 1)) << 3); }

// one lane's share of loading the 32-word block that starts at word `first` (a multiple of 32) into the ring
template <class SM>
ICM_HD void ring_load_block(const SM &sm, const DecConst &c, uint32_t first, int lane)
{
    const uint32_t j = first + (uint32_t)lane;
    const uint32_t w = j < c.nwords ? ICM_LDG(c.W + j) : 0u; // past the end: zeros (UB in the reference)
    const uint32_t aw = ICM_ROTR(w, c.rs) & c.M;
    const uint32_t slot = j & (kRingSlots - 1);
    sm.st64(c.ring + (slot << 3), w, aw);
    if (slot < kRingMirror) sm.st64(c.ring + ((slot + kRingSlots) << 3), w, aw);
}

// Rans64DecGetBits(4) on the ring; pos = index of d.wv
template <class SM>
ICM_HD uint32_t dec_get4(const SM &sm, const DecConst &c, WarpDec &d, uint32_t &pos)
{
    const uint32_t val = d.xl & 15u;
    d.xl = (d.xl >> 4) | (d.xh << 28);
    d.xh >>= 4;
    if (d.xh == 0 && d.xl < 0x80000000u) {
        d.xh = d.xl; d.xl = d.wv;
        ++pos;
        d.wv = sm.ld32(ring_slot(c, pos));
    }
    return val;
}

// The out-of-line path of one symbol: crowded bucket (binary search over the CDF row) and / or escape.
// `loaded` = first word index NOT yet in the ring; the caller refills when fewer than kMinAhead words are ahead, and
// the nibble loop refills by itself (through `refill`, warp-cooperative on the device) if a pathological stream
// drains the ring.  Returns the value (before the offset is added).
template <class SM, class Refill>
ICM_HD int dec_rare(const SM &sm, const DecConst &c, WarpDec &d, uint32_t &pos, uint32_t &loaded, Refill refill, uint32_t a,
                    uint32_t img_base, uint32_t maxv, uint32_t rowinfo, const u4 &E)
{
    const uint32_t cum = d.xl & 0xFFFFu, cumhi = (d.xl << 16) | 0xFFFEu;
    const uint32_t s0 = E.w & 0xFFFFu;
    const bool c1 = cumhi >= E.y, c2 = cumhi >= E.z;
    uint32_t w = c2 ? E.z : (c1 ? E.y : E.x);
    uint32_t sym = s0 + (c1 ? 1u : 0u) + (c2 ? 1u : 0u);
    if (cumhi >= E.w) { // an unlisted symbol (crowded bucket or the escape symbol): search [s0, first symbol of the next bucket]
        const uint32_t row = img_base + sm.ld32(rowinfo), nsym = sm.ld32(rowinfo + 4);
        uint32_t lo = s0, hi = nsym;
        if ((a & c.M) != c.M) { const uint32_t h = (sm.ld32(a + 28) & 0xFFFFu) + 1u; hi = h < nsym ? h : nsym; }
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sm.ld16(row + 2 * mid) <= cum) lo = mid; else hi = mid;
        }
        const uint32_t start = sm.ld16(row + 2 * lo), next = sm.ld16(row + 2 * lo + 2);
        w = (start << 16) | ((next - start) & 0xFFFFu);
        sym = lo;
    }
    // Rans64DecAdvance
    const uint32_t freq = w & 0xFFFFu, dd = cum - (w >> 16);
    const uint32_t xs_lo = ICM_SHR64LO(d.xl, d.xh, 16), xs_hi = d.xh >> 16;
    const uint64_t nx = (uint64_t)freq * xs_lo + dd;
    d.xl = (uint32_t)nx; d.xh = (uint32_t)(nx >> 32) + freq * xs_hi;
    if (d.xh == 0 && d.xl < 0x80000000u) {
        d.xh = d.xl; d.xl = d.wv;
        ++pos;
        d.wv = sm.ld32(ring_slot(c, pos));
    }
    int value = (int)sym;
    if (sym == maxv) { // escape: count nibble(s) (15 = "more"), then the payload nibbles LSB first
        int val = (int)dec_get4(sm, c, d, pos);
        int nb = val;
        while (val == 15) {
            if ((int)(loaded - pos) < 4) refill(pos, loaded);
            val = (int)dec_get4(sm, c, d, pos);
            nb += val;
        }
        int raw = 0;
        for (int j = 0; j < nb; ++j) {
            if ((int)(loaded - pos) < 4) refill(pos, loaded);
            val = (int)dec_get4(sm, c, d, pos);
            raw |= val << ((j * 4) & 31);
        }
        value = raw >> 1;
        value = (raw & 1) ? -value - 1 : value + (int)maxv;
    }
    if ((int)(loaded - pos) < kMinAhead) refill(pos, loaded);
    d.awv = sm.ld32(ring_slot(c, pos) + 4);
    d.wa1 = ring_slot(c, pos + 1);
    return value;
}

// An escape whose count fits one nibble (payload < 2^56, i.e. always for 32-bit symbols), finished without the
// out-of-line call.  Needs cum >= esc_start (else the symbol sits in a crowded bucket) and enough buffered words
// (<= 3 are consumed).  Returns false with nothing changed otherwise.
template <class SM>
ICM_HD bool dec_escape_simple(const SM &sm, const DecConst &c, WarpDec &d, uint32_t &pos, uint32_t loaded, uint32_t maxv,
                              uint32_t esc_start, int &value)
{
    const uint32_t cum = d.xl & 0xFFFFu;
    if (cum < esc_start || (int)(loaded - pos) < kMinAhead + 4) return false;
    WarpDec t = d;
    uint32_t p = pos;
    // Rans64DecAdvance over the escape symbol [esc_start, 2^16)
    const uint32_t freq = 65536u - esc_start, dd = cum - esc_start;
    const uint32_t xs_lo = ICM_SHR64LO(t.xl, t.xh, 16), xs_hi = t.xh >> 16;
    const uint64_t nx = (uint64_t)freq * xs_lo + dd;
    t.xl = (uint32_t)nx; t.xh = (uint32_t)(nx >> 32) + freq * xs_hi;
    if (t.xh == 0 && t.xl < 0x80000000u) {
        t.xh = t.xl; t.xl = t.wv;
        ++p;
        t.wv = sm.ld32(ring_slot(c, p));
    }
    const int nb = (int)dec_get4(sm, c, t, p);
    if (nb == 15) return false; // a longer count: out-of-line path
    int raw = 0;
    for (int j = 0; j < nb; ++j) raw |= (int)dec_get4(sm, c, t, p) << ((j * 4) & 31);
    value = raw >> 1;
    value = (raw & 1) ? -value - 1 : value + (int)maxv;
    d = t;
    pos = p;
    d.awv = sm.ld32(ring_slot(c, pos) + 4);
    d.wa1 = ring_slot(c, pos + 1);
    return true;
}

// The common path of one symbol.  `a` = entry address of this symbol, offset from its table record, ent_next = entry
// block of the next symbol's table.  Everything is computed before the single branch at the end (a lone warp
// issues in order: with the test in the middle, the symbol-index chain and the multiply chain could not overlap).
// Returns false, with nothing changed, when the symbol needs dec_rare.
template <class SM>
ICM_HD bool dec_fast(const SM &sm, const DecConst &c, WarpDec &d, uint32_t &a, const u4 &E, int32_t offset, uint32_t ent_next,
                     int &value)
{
    uint32_t w1, aw1;
    sm.ld64(d.wa1, w1, aw1);
    const uint32_t cum = d.xl & 0xFFFFu, cumhi = (d.xl << 16) | 0xFFFEu;
    const uint32_t xs_lo = ICM_SHR64LO(d.xl, d.xh, 16), xs_hi = d.xh >> 16;
    const bool c1 = cumhi >= E.y, c2 = cumhi >= E.z;
    const uint32_t w = c2 ? E.z : (c1 ? E.y : E.x);
    const uint32_t j = c2 ? 2u : (c1 ? 1u : 0u);
    const uint32_t sym = (E.w & 0xFFFFu) + j;
    const bool rare = cumhi >= E.w; // an unlisted symbol: crowded bucket or escape
    // Rans64DecAdvance: x = freq * (x >> 16) + cum - start, then renormalise
    const uint32_t freq = w & 0xFFFFu, dd = cum - (w >> 16);
    const uint64_t nx = (uint64_t)freq * xs_lo + dd;
    const uint32_t nxl = (uint32_t)nx, nxh = (uint32_t)(nx >> 32) + freq * xs_hi;
    const bool rn = (nxh == 0) & (nxl < 0x80000000u);
    const uint32_t a_keep = (ICM_ROTR(nxl, c.rs) & c.M) | ent_next, a_renorm = d.awv | ent_next;
    if (__builtin_expect(rare, 0)) return false;
    d.xl = rn ? d.wv : nxl;
    d.xh = rn ? nxl : nxh;
    a = rn ? a_renorm : a_keep;
    d.wv = rn ? w1 : d.wv;
    d.awv = rn ? aw1 : d.awv;
    d.wa1 = rn ? d.wa1 + 8 : d.wa1;
    value = (int)sym + offset;
    return true;
}

}  // namespace lane
}  // namespace icm
