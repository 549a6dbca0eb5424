/*
 * oracle/rans_oracle.c -- CPU restatement of the reference's native entropy-coding path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (image-compression-for-machine_b200/) never links or calls it.
 *
 * What it restates (all paths relative to /root/reference):
 *   R1  BufferedRansEncoder::encode_with_indexes  ans.cpython-38-x86_64-linux-gnu.so @0x8a10
 *   R2  BufferedRansEncoder::flush                @0x8730  + third_party/ryg_rans/rans64.h:65-103
 *   R3  RansEncoder::encode_with_indexes          @0x8d70  (= R1 + R2 on a temporary)
 *   R4  RansDecoder::set_stream/decode_stream     @0x7c40/@0x7ce0 + rans64.h:107-142
 *   R5  pmf_to_quantized_cdf                      _CXX.cpython-38-x86_64-linux-gnu.so @0x68c0
 * The C++ sources of those two modules are NOT in the reference tree (setup.py:49-79 points at
 * compressai/cpp_exts/, which is absent); they are InterDigital CompressAI 1.1.6dev0
 * (compressai/version.py:1).  The algorithm below follows the published CompressAI
 * rans_interface.cpp / ops.cpp semantics and the vendored rans64.h, and is PINNED bit-for-bit
 * against the reference's own shipped binaries (oracle/refbin.py drives them through ctypes;
 * oracle/make_golden.py records the vectors in tests/golden/rans_kat.json, and
 * tests/test_oracle_pinned.py replays them).  Call sites in the reference:
 * compressai/entropy_models/entropy_models.py:61,228,277; compressai/models/stf.py:698,727-729,
 * 751-752,766.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION 16
#define BYPASS_BITS 4
#define BYPASS_MAX 15
#define RANS_L (1ull << 31) /* rans64.h:59 */

/* ------------------------------------------------------------------------------------------ */
/* R5: pmf -> quantised cdf.  float32 arithmetic for the rounding step, as in the binary.      */
int orc_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf)
{
    if (n < 1 || precision < 1 || precision > 31) return -1;
    const int len = n + 1;
    cdf[0] = 0;
    for (int i = 0; i < n; ++i)
        cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
    uint32_t total = 0; /* 32-bit accumulation (std::accumulate with an int seed) */
    for (int i = 0; i < len; ++i) total += cdf[i];
    if (total == 0) return -2;
    for (int i = 0; i < len; ++i)
        cdf[i] = (uint32_t)((((uint64_t)1 << precision) * (uint64_t)cdf[i]) / total);
    for (int i = 1; i < len; ++i) cdf[i] += cdf[i - 1];
    cdf[len - 1] = 1u << precision;
    /* zero-frequency repair: steal one count from the least-frequent symbol that can spare it */
    for (int i = 0; i < len - 1; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        uint32_t best = ~0u;
        int donor = -1;
        for (int j = 0; j < len - 1; ++j) {
            uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < best) { best = f; donor = j; }
        }
        if (donor < 0) return -3;
        if (donor < i) { for (int j = donor + 1; j <= i; ++j) cdf[j]--; }
        else           { for (int j = i + 1; j <= donor; ++j) cdf[j]++; }
    }
    return len;
}

/* ------------------------------------------------------------------------------------------ */
/* R1: symbol -> list of coder records {start, range, bypass}.                                 */
typedef struct { uint16_t start, range; uint8_t bypass; } rec_t;

typedef struct { rec_t *v; size_t n, cap; } recvec_t;

static int push(recvec_t *r, uint16_t start, uint16_t range, uint8_t bypass)
{
    if (r->n == r->cap) {
        size_t nc = r->cap ? r->cap * 2 : 1024;
        rec_t *nv = (rec_t *)realloc(r->v, nc * sizeof(rec_t));
        if (!nv) return -1;
        r->v = nv; r->cap = nc;
    }
    r->v[r->n].start = start; r->v[r->n].range = range; r->v[r->n].bypass = bypass;
    r->n++;
    return 0;
}

static int build_records(recvec_t *r, const int32_t *sym, const int32_t *idx, int64_t n,
                         const int32_t *cdfs, int n_cdf, int cdf_stride,
                         const int32_t *sizes, const int32_t *offsets)
{
    for (int64_t i = 0; i < n; ++i) {
        const int32_t t = idx[i];
        if (t < 0 || t >= n_cdf) return -2; /* the reference only asserts (compiled out): UB there */
        const int32_t *cdf = cdfs + (size_t)t * cdf_stride;
        const int32_t max_value = sizes[t] - 2;
        if (max_value < 0 || max_value + 1 >= cdf_stride) return -3;
        int32_t v = sym[i] - offsets[t];
        uint32_t raw = 0;
        if (v < 0) { raw = (uint32_t)(-2 * v - 1); v = max_value; }
        else if (v >= max_value) { raw = (uint32_t)(2 * (v - max_value)); v = max_value; }
        if (push(r, (uint16_t)cdf[v], (uint16_t)(cdf[v + 1] - cdf[v]), 0)) return -1;
        if (v == max_value) {
            int32_t nb = 0;
            while ((raw >> (nb * BYPASS_BITS)) != 0) ++nb;
            int32_t val = nb;
            while (val >= BYPASS_MAX) { if (push(r, BYPASS_MAX, BYPASS_MAX + 1, 1)) return -1; val -= BYPASS_MAX; }
            if (push(r, (uint16_t)val, (uint16_t)(val + 1), 1)) return -1;
            for (int32_t j = 0; j < nb; ++j) {
                const uint32_t nib = (raw >> (j * BYPASS_BITS)) & BYPASS_MAX;
                if (push(r, (uint16_t)nib, (uint16_t)(nib + 1), 1)) return -1;
            }
        }
    }
    return 0;
}

/* R2: drain records back-to-front through the 64-bit rANS state (rans64.h:77-103). */
static int64_t flush_records(const recvec_t *r, uint8_t *out, int64_t cap)
{
    /* the reference allocates exactly r->n words and under-runs for <3 records; we allocate +2 */
    const size_t nwords = r->n + 2;
    uint32_t *buf = (uint32_t *)malloc(nwords * sizeof(uint32_t));
    if (!buf) return -1;
    uint32_t *end = buf + nwords, *ptr = end;
    uint64_t x = RANS_L;
    for (size_t k = r->n; k-- > 0;) {
        const rec_t s = r->v[k];
        if (!s.bypass) {
            const uint64_t x_max = ((RANS_L >> PRECISION) << 32) * (uint64_t)s.range;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = ((x / s.range) << PRECISION) + (x % s.range) + s.start;
        } else {
            const uint64_t x_max = ((RANS_L >> 16) << 32) * (uint64_t)(1u << (16 - BYPASS_BITS));
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = (x << BYPASS_BITS) | s.start;
        }
    }
    ptr -= 2;
    ptr[0] = (uint32_t)x;
    ptr[1] = (uint32_t)(x >> 32);
    const int64_t nbytes = (int64_t)(end - ptr) * 4;
    if (nbytes > cap) { free(buf); return -4; }
    memcpy(out, ptr, (size_t)nbytes);
    free(buf);
    return nbytes;
}

/* Buffered encoder object: several encode calls, one flush (stf.py:698,727-729). */
void *orc_encoder_new(void) { return calloc(1, sizeof(recvec_t)); }
void orc_encoder_free(void *e) { if (e) { free(((recvec_t *)e)->v); free(e); } }

int orc_encoder_push(void *e, const int32_t *sym, const int32_t *idx, int64_t n,
                     const int32_t *cdfs, int n_cdf, int cdf_stride,
                     const int32_t *sizes, const int32_t *offsets)
{
    return build_records((recvec_t *)e, sym, idx, n, cdfs, n_cdf, cdf_stride, sizes, offsets);
}

int64_t orc_encoder_flush(void *e, uint8_t *out, int64_t cap)
{
    recvec_t *r = (recvec_t *)e;
    int64_t nb = flush_records(r, out, cap);
    r->n = 0; /* reusable after flush, like the reference object */
    return nb;
}

int64_t orc_encoder_pending(void *e) { return (int64_t)((recvec_t *)e)->n; }

/* R3: one-shot encode. Returns number of bytes written at out[0..], or <0 on error. */
int64_t orc_rans_encode(const int32_t *sym, const int32_t *idx, int64_t n,
                        const int32_t *cdfs, int n_cdf, int cdf_stride,
                        const int32_t *sizes, const int32_t *offsets,
                        uint8_t *out, int64_t cap)
{
    recvec_t r = {0, 0, 0};
    int rc = build_records(&r, sym, idx, n, cdfs, n_cdf, cdf_stride, sizes, offsets);
    int64_t nb = rc ? rc : flush_records(&r, out, cap);
    free(r.v);
    return nb;
}

/* ------------------------------------------------------------------------------------------ */
/* R4: decoder with persistent {state, word position} (set_stream + N x decode_stream).        */
typedef struct { uint64_t x; const uint32_t *words; int64_t nwords, pos; } dec_t;

static inline uint32_t next_word(dec_t *d)
{
    /* reading past the end is UB in the reference; we return 0 and keep going */
    return d->pos < d->nwords ? d->words[d->pos++] : (d->pos++, 0u);
}

static inline uint32_t get_bits(dec_t *d)
{
    const uint32_t val = (uint32_t)(d->x & BYPASS_MAX);
    d->x >>= BYPASS_BITS;
    if (d->x < RANS_L) d->x = (d->x << 32) | next_word(d);
    return val;
}

/* state_io[0] = rANS state, state_io[1] = next word index.  If state_io[1] < 0 the stream is
 * (re)initialised from its first two words (set_stream). */
int orc_rans_decode(const uint8_t *stream, int64_t nbytes, const int32_t *idx, int64_t n,
                    const int32_t *cdfs, int n_cdf, int cdf_stride,
                    const int32_t *sizes, const int32_t *offsets,
                    int32_t *out, int64_t *state_io)
{
    dec_t d;
    d.words = (const uint32_t *)stream;
    d.nwords = nbytes / 4;
    if (state_io[1] < 0) {
        if (d.nwords < 2) return -5;
        d.x = (uint64_t)d.words[0] | ((uint64_t)d.words[1] << 32);
        d.pos = 2;
    } else {
        d.x = (uint64_t)state_io[0];
        d.pos = state_io[1];
    }
    for (int64_t i = 0; i < n; ++i) {
        const int32_t t = idx[i];
        if (t < 0 || t >= n_cdf) return -2;
        const int32_t *cdf = cdfs + (size_t)t * cdf_stride;
        const int32_t size = sizes[t];
        const int32_t max_value = size - 2;
        const uint32_t cum = (uint32_t)(d.x & 0xFFFFu);
        int32_t j = 0;
        while (j < size && !((uint32_t)cdf[j] > cum)) ++j; /* the reference's linear find_if */
        const int32_t s = j - 1;
        const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
        d.x = (uint64_t)freq * (d.x >> PRECISION) + cum - start;
        if (d.x < RANS_L) d.x = (d.x << 32) | next_word(&d);
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)get_bits(&d);
            int32_t nb = val;
            while (val == BYPASS_MAX) { val = (int32_t)get_bits(&d); nb += val; }
            int32_t raw = 0;
            for (int32_t k = 0; k < nb; ++k) { val = (int32_t)get_bits(&d); raw |= val << (k * BYPASS_BITS); }
            value = raw >> 1;
            if (raw & 1) value = -value - 1; else value += max_value;
        }
        out[i] = value + offsets[t];
    }
    state_io[0] = (int64_t)d.x;
    state_io[1] = d.pos;
    return 0;
}
