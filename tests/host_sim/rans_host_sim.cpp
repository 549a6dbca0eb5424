// TEST INFRASTRUCTURE.  Builds the per-symbol decoder code of csrc/rans_lane.cuh for the HOST so that the bucket-table
// image, the in-register bucket resolution, the binary-search path, the escape path and the stream-word ring can be
// checked bit for bit against the pinned oracle without a GPU (tests/test_rans_host_sim.py).  The product library never
// runs this on the CPU; the same inline functions are what rans_decode_bucket_kernel (csrc/rans.cu) executes.
#include "rans_lane.cuh"

#include <cstdio>

using namespace icm::lane;

extern "C" {

// returns the image size in bytes (0 = tables do not fit), fills bits[n_cdf] with the bucket bits
int sim_image_info(const int32_t *cdfs, int n_cdf, int stride, const int32_t *sizes, const int32_t *offsets, long long budget, int *bits)
{
    Image im = build_image(cdfs, n_cdf, stride, sizes, offsets, (size_t)budget);
    if (!im.ok) return 0;
    for (int t = 0; t < n_cdf; ++t) bits[t] = im.bits[t];
    return (int)im.bytes.size();
}

// decode `n` symbols in `n_steps` calls (step boundaries in steps[]), state carried across like the kernel does.
// The loop below mirrors rans_decode_bucket_kernel (csrc/rans.cu) with the 32 lanes of the warp emulated one after the
// other; dec_fast / dec_escape_simple / dec_rare / ring_load_block are the very functions the kernel runs.
// rare_count: symbols that took the out-of-line path (crowded buckets, long escapes).
int sim_decode(const int32_t *cdfs, int n_cdf, int stride, const int32_t *sizes, const int32_t *offsets, long long budget,
               const uint32_t *words, int nwords, const int32_t *idx, long long n_total, const long long *steps, int n_steps,
               int32_t *out, long long *rare_count)
{
    Image im = build_image(cdfs, n_cdf, stride, sizes, offsets, (size_t)budget);
    if (!im.ok) return -1;
    const uint32_t base = 3 * kAlign; // the image at a kAlign-aligned "shared address", this warp's areas behind it
    std::vector<unsigned char> smem(base + im.bytes.size() + 16 + kWarpBytes);
    memcpy(&smem[base], im.bytes.data(), im.bytes.size());
    Smem sm{smem.data()};
    DecConst c;
    c.rs = im.rs; c.M = im.M;
    c.ring = base + (((uint32_t)im.bytes.size() + 15u) & ~15u);
    c.W = words; c.nwords = (uint32_t)nwords;
    const uint32_t stage = c.ring + kRingBytes, outs = stage + kStageBytes, meta_addr = base + im.meta_off;
    auto refill_to = [&](uint32_t p, uint32_t ld) {
        while ((int)(ld - p) < kRefillBelow) { for (int lane = 0; lane < 32; ++lane) ring_load_block(sm, c, ld, lane); ld += 32; }
        return ld;
    };
    auto refill = [&](uint32_t p, uint32_t &ld) { ld = refill_to(p, ld); };
    uint64_t state = 0;
    long long pos_saved = -1, k0 = 0;
    *rare_count = 0;
    for (int st = 0; st < n_steps; ++st) {
        const long long n = steps[st] - k0;
        const int32_t *I = idx + k0;
        int32_t *O = out + k0;
        const long long n_chunks = (n + 31) / 32;
        auto stage_chunk = [&](long long ch) {
            for (int lane = 0; lane < 32; ++lane) {
                const long long j = ch * 32 + lane;
                const int t = j < n ? I[j] : 0;
                u4 r = sm.ld128(meta_addr + 16u * (uint32_t)t);
                r.x += base;
                const uint32_t par = (uint32_t)(ch & 1) * (33u * 16u);
                sm.st128(stage + par + 16u * (uint32_t)lane, r);
                if (lane == 0) sm.st128(stage + (33u * 16u - par) + 32u * 16u, r);
            }
        };
        WarpDec d;
        uint32_t pos, loaded;
        if (pos_saved < 0) { d.xl = nwords > 0 ? words[0] : 0; d.xh = nwords > 1 ? words[1] : 0; pos = 2; }
        else { d.xl = (uint32_t)state; d.xh = (uint32_t)(state >> 32); pos = (uint32_t)pos_saved; }
        loaded = pos & ~31u;
        stage_chunk(0);
        loaded = refill_to(pos, loaded);
        sm.ld64(ring_slot(c, pos), d.wv, d.awv);
        d.wa1 = ring_slot(c, pos + 1);
        uint32_t wa1_base = d.wa1;
        u4 mcur = sm.ld128(stage);
        uint32_t a = (ICM_ROTR(d.xl, c.rs) & c.M) | mcur.x;
        for (long long ch = 0; ch < n_chunks; ++ch) {
            if (ch + 1 < n_chunks) stage_chunk(ch + 1);
            pos += (d.wa1 - wa1_base) >> 3;
            if ((int)(loaded - pos) < kRefillBelow) loaded = refill_to(pos, loaded);
            d.wa1 = ring_slot(c, pos + 1);
            wa1_base = d.wa1;
            const int valid = (int)std::min<long long>(32, n - ch * 32);
            const uint32_t sbase = stage + (uint32_t)(ch & 1) * (33u * 16u);
            for (int k = 0; k < valid; ++k) {
                const u4 mn = sm.ld128(sbase + 16u * (uint32_t)(k + 1));
                const u4 E = sm.ld128_ro(a);
                int value;
                if (!dec_fast(sm, c, d, a, E, (int32_t)mcur.z, mn.x, value)) {
                    pos += (d.wa1 - wa1_base) >> 3;
                    if (!dec_escape_simple(sm, c, d, pos, loaded, mcur.y, mcur.w >> 16, value)) {
                        ++*rare_count;
                        value = dec_rare(sm, c, d, pos, loaded, refill, a, base, mcur.y, base + im.row_off + 8u * (mcur.w & 0xFFFFu), E);
                    }
                    value += (int32_t)mcur.z;
                    wa1_base = d.wa1;
                    a = (ICM_ROTR(d.xl, c.rs) & c.M) | mn.x;
                }
                sm.st32(outs + 4u * (uint32_t)k, (uint32_t)value);
                mcur = mn;
            }
            for (int lane = 0; lane < valid; ++lane) O[ch * 32 + lane] = (int32_t)sm.ld32(outs + 4u * (uint32_t)lane);
        }
        pos += (d.wa1 - wa1_base) >> 3;
        state = ((uint64_t)d.xh << 32) | d.xl;
        pos_saved = pos;
        k0 = steps[st];
    }
    return (int)pos_saved;
}

}  // extern "C"
