// Implicit-GEMM convolution / linear layer on the Blackwell tensor cores (rows T3-T11 of SURVEY.md §8a).
//
//   D[pixel, cout] = act( sum_{tap, cin} A[pixel + tap, cin] * W[cout, tap, cin] + bias ) (+ residual)
//
// One CTA computes a 128-pixel x BN-channel output tile:
//   warp 0     TMA producer: per k-step one 4-D box of the channels-last activation
//              (64 channels x TW x TH pixels, shifted by the filter tap; the zero padding of the convolution
//              is TMA's out-of-bounds zero fill, stride-2 convolutions use the tensor map's element strides)
//              and one 2-D box of the packed weights, both landing 128B-swizzled in shared memory;
//   warp 1     allocates TMEM and issues tcgen05.mma (M=128, N=BN, K=16, bf16 x bf16 -> fp32 in TMEM);
//   warps 2-17 epilogue (four warps per TMEM lane quarter, taking 16-column chunks in turn): tcgen05.ld the
//              accumulator rows, bias + activation (+ residual) in registers, vectorised stores (optionally
//              with the PixelShuffle permutation folded into the address).
// Shared memory and TMEM are sized so that two (small-K layers: up to four) CTAs share an SM: the epilogue of
// one tile then overlaps the loads and MMAs of another, which is what the memory-bound Swin linears need.
// The K loop order (tap-major, then 64-channel chunks) is fixed and there is no split-K, so every output
// element is reduced in the same order whatever the batch size or tile shape: the encoder and decoder
// sides of the codec see bit-identical means/scales (SURVEY.md §7 "Encoder/decoder determinism").
#include "umma.cuh"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

namespace icm {

struct ConvParams {
    int Ho, Wo;            // output spatial size
    int TW_log2;           // tile = TW x TH pixels, TW * TH == 128
    int tiles_w, tiles_h;
    int k_chunks;          // ceil(Cin / 64)
    int last_k16;          // K = 16 steps that hold real channels in the last 64-channel chunk of a tap (1..4)
    int Cin_pad;           // channels per tap in the packed weight
    int KH, KW, stride, pad;
    int BN, Cout, stages, tmem_cols; // tmem_cols = two accumulators
    int n_tiles, total_tiles;
    unsigned long long fd_n, fd_w, fd_h; // ceil(2^40 / d) for d = n_tiles, tiles_w, tiles_h (0 when d == 1)
    int act, out_dtype, pixel_shuffle;
    long long out_pitch, res_pitch;
    int wide_st, wide_ld;    // output / second-operand rows and base are 32-byte aligned: 256-bit stores / loads
    int res_dtype, res_mode; // residual operand: fp32 / bf16; added after act (0), before act (1), multiplied (2)
    const float *bias;
    const void *residual;
    void *out;
    int *sched;              // {next tile to hand out beyond the first wave, CTAs finished}: this stream's tile counter
    int dyn_first;           // the first tile of a CTA also comes from the counter
    // grouped launch (icm_conv2d_grouped): G convolutions of one geometry, tile id = g * tiles_per_group + tile in group
    int groups, tiles_per_group;
    unsigned long long fd_g;
    int w_group_rows;            // packed-weight rows between groups
    int bias_group_stride;       // floats between the groups' biases in global memory
    long long out_group_stride;  // output elements between groups
    int has_tail;                // the last 64-channel chunk of a tap is read at tail_ch[g]
    int a_img[ICM_MAX_CONV_GROUPS];   // first image of group g in the input tensor
    int tail_ch[ICM_MAX_CONV_GROUPS];
};

// The activation switch sits OUTSIDE the 16-element loop (one uniform branch per chunk): with it inside, the
// unrolled epilogue carried every activation's code and a branch chain per element (~32 instructions per
// output; the stage-0 GELU linear was issue-bound at a fifth of the HBM roofline).
__device__ __forceinline__ void apply_act16(float (&v)[16], int act)
{
    switch (act) {
    case ICM_ACT_GELU:
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
        break;
    case ICM_ACT_HALF_TANH:
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.5f * tanhf(v[j]);
        break;
    case ICM_ACT_SIGMOID:
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = mufu_rcp(1.0f + mufu_ex2(-1.4426950408889634f * v[j]));
        break;
    case ICM_ACT_RSQRT:
#pragma unroll
        for (int j = 0; j < 16; ++j) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[j])); v[j] = y; }
        break;
    case ICM_ACT_SQRT:
#pragma unroll
        for (int j = 0; j < 16; ++j) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(v[j])); v[j] = y; }
        break;
    default:
        break;
    }
}

constexpr int EPI_WARPS = 16;                      // four per TMEM lane quarter
constexpr int TQ = 4;                              // depth of the tile-id queue between the scheduler thread and its consumers
constexpr int CONV_THREADS = (2 + EPI_WARPS) * 32; // TMA warp + MMA warp + epilogue warps

// Persistent with a DYNAMIC tile scheduler: every tile of a CTA, the first included, comes from an atomic counter.  The TMA thread is the scheduler: it fetches the tile id one tile ahead and hands
// it to the MMA thread and the epilogue warps through a small mbarrier-guarded queue in shared memory.  A CTA that
// starts late -- its SM was held by an rANS coder CTA of another stream, which cannot share an SM with this kernel's
// ~200 KB of shared memory -- therefore finds no work left instead of delaying the launch by the tiles a static
// round-robin would have reserved for it, so the grid never has to leave SMs unused "in case".  The last CTA to
// finish resets the counter for the next launch on the stream.  Which CTA computes a tile does not affect its result
// (no cross-tile reduction), so outputs stay bit-identical.  The TMA pipeline runs across tile boundaries, and the
// accumulator is double-buffered in TMEM, so the epilogue of tile i overlaps the loads and MMAs of tile i+1.
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const ConvParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    // layout: [stages][A 16 KB][B BN*128 B] | barriers | tmem slot | bias
    const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)p.BN * BK * 2;
    const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(tiles + (size_t)p.stages * stage_bytes);
    uint64_t *empty_bar = full_bar + MAX_STAGES;
    uint64_t *acc_full = empty_bar + MAX_STAGES; // [2] MMA -> epilogue
    uint64_t *acc_empty = acc_full + 2;          // [2] epilogue -> MMA
    uint64_t *tq_full = acc_empty + 2;           // [TQ] scheduler (TMA thread) -> MMA thread + epilogue warps: s_tile[slot] is set
    uint64_t *tq_empty = tq_full + TQ;           // [TQ] consumers -> scheduler
    int *s_tile = reinterpret_cast<int *>(tq_empty + TQ); // [TQ] tile ids, -1 = no more tiles
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(s_tile + TQ);
    float *s_bias = reinterpret_cast<float *>(tmem_slot + 4); // [n_tiles * BN], zeros without a bias

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TW = 1 << p.TW_log2, TH = BM >> p.TW_log2;
    const int k_iters = p.KH * p.KW * p.k_chunks;
    const uint32_t acc_stride = (uint32_t)p.tmem_cols >> 1; // TMEM columns between the two accumulators

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_WARPS); }
        for (int a = 0; a < TQ; ++a) { mbar_init(&tq_full[a], 1); mbar_init(&tq_empty[a], EPI_WARPS + 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const int per_group = p.n_tiles * p.BN;
        for (int i = threadIdx.x; i < p.groups * per_group; i += blockDim.x) {
            const int g = i / per_group, j = i - g * per_group;
            s_bias[i] = (p.bias && j < p.Cout) ? p.bias[(long long)g * p.bias_group_stride + j] : 0.f;
        }
    }
    if (warp == 1) { // TMEM allocation is warp-collective
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // tile id -> (n tile fastest, then w, h, image)
    // (three runtime integer divisions cost ~100 instructions per tile: 10 % of the epilogue of a K = 48 linear)
    auto fdiv = [](uint32_t n, unsigned long long m) -> uint32_t { return m ? (uint32_t)(((unsigned long long)n * m) >> 40) : n; };
    auto tile_coords = [&](int tile, int &n0, int &w0, int &h0, int &b, int &g) {
        uint32_t t = (uint32_t)tile, q = fdiv(t, p.fd_g);
        if (p.groups > 1) { g = (int)q; t -= q * (uint32_t)p.tiles_per_group; } else g = 0;
        q = fdiv(t, p.fd_n);
        const int nt = (int)(t - q * (uint32_t)p.n_tiles); t = q;
        q = fdiv(t, p.fd_w);
        const int tw_i = (int)(t - q * (uint32_t)p.tiles_w); t = q;
        q = fdiv(t, p.fd_h);
        const int th_i = (int)(t - q * (uint32_t)p.tiles_h);
        n0 = nt * p.BN; w0 = tw_i * TW; h0 = th_i * TH; b = (int)q;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            int stage = 0;
            uint32_t phase = 0;
            // first tile: blockIdx.x, or (dyn_first) drawn from the counter like every later one, so that a CTA that starts late
            // -- its SM was held by a coder CTA of another stream -- owns no tile at all and leaves at once
            int tile = p.dyn_first ? atomicAdd(p.sched, 1) : (int)blockIdx.x;
            int slot = 0;
            uint32_t qphase = 0;
            while (true) {
                mbar_wait(&tq_empty[slot], qphase ^ 1);
                s_tile[slot] = tile < p.total_tiles ? tile : -1;
                mbar_arrive(&tq_full[slot]); // release semantics: the id is visible to whoever observes the phase
                if (++slot == TQ) { slot = 0; qphase ^= 1; }
                if (tile >= p.total_tiles) break;
                const int next = (p.dyn_first ? 0 : (int)gridDim.x) + atomicAdd(p.sched, 1); // in flight while this tile's loads are issued
                int n0, w0, h0, b, g;
                tile_coords(tile, n0, w0, h0, b, g);
                tile = next;
                const int img = b + p.a_img[g], wrow = n0 + g * p.w_group_rows;
                const int tail_c = p.has_tail ? p.tail_ch[g] : (p.k_chunks - 1) * BK;
                // (tap, chunk) -> (dy, dx, channel, weight column) advance by counters: this one thread issues every load of the
                // CTA, and a k-step used to start with two runtime integer divisions (it / k_chunks, tap / KW: ~25 dependent
                // instructions each at ~5 cycles apiece in a lone thread).  For N <= 176 the four MMAs of a k-step take only
                // 160-350 cycles (tools/umma_issue_bench.cu: 40 / 48 / 64 / 96 cycles per MMA at N = 32 / 64 / 128 / 192) and
                // the TMA ingest about as long, so this prologue was on the critical path: conv family 34.5 -> 33.7 ms per step.
                const int wx0 = w0 * p.stride - p.pad, hy0 = h0 * p.stride - p.pad, last_chunk = p.k_chunks - 1;
                int chunk = 0, dx = 0, dy = 0, wcol = 0, ch = 0;
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char *sa = tiles + (size_t)stage * stage_bytes;
                    unsigned char *sb = sa + a_bytes;
                    mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
                    tma_load_4d(&map_a, &full_bar[stage], sa, chunk == last_chunk ? tail_c : ch, wx0 + dx, hy0 + dy, img);
                    tma_load_2d(&map_w, &full_bar[stage], sb, wcol + ch, wrow);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    ch += BK;
                    if (++chunk == p.k_chunks) { // next tap
                        chunk = 0; ch = 0; wcol += p.Cin_pad;
                        if (++dx == p.KW) { dx = 0; ++dy; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        // ONE thread runs the whole issue loop (waits included), and the shared-memory descriptors advance by addition:
        // a clock64 timeline showed ~130 cycles of single-thread overhead per tcgen05.mma and ~150 per mbarrier wait with a
        // per-iteration elect / reconvergence and the descriptor rebuilt from the address each k-step -- more than the
        // MMAs themselves take for N <= 176, which is why those layers sat at 30-70 % of the tensor peak.
        if (elect_one()) {
            const uint64_t desc0 = make_smem_desc(smem_u32(tiles));
            const uint64_t step = (uint64_t)(stage_bytes >> 4), boff = (uint64_t)(a_bytes >> 4); // in the 16-byte address field
            uint64_t da = desc0;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int slot = 0;
            uint32_t qphase = 0;
            while (true) {
                mbar_wait(&tq_full[slot], qphase);
                const int tile = *reinterpret_cast<volatile int *>(&s_tile[slot]);
                mbar_arrive(&tq_empty[slot]); // only whether there is a tile matters here
                if (++slot == TQ) { slot = 0; qphase ^= 1; }
                if (tile < 0) break;
                mbar_wait(&acc_empty[acc], acc_phase ^ 1); // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)acc * acc_stride;
                int chunk = 0;
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t db = da + boff;
                    // the channels beyond Cin in a tap's last chunk are zero in both operands (TMA fill, weight padding):
                    // their K = 16 steps are skipped, which is exact (Cin = 224: 14 instead of 16 MMAs per tap)
                    const int n_k = (chunk == p.k_chunks - 1) ? p.last_k16 : BK / UMMA_K;
                    if (++chunk == p.k_chunks) chunk = 0;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 bytes (16 bf16) along K inside the swizzle row: +2 in the 16-byte address field
                        if (k < n_k) umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);                  // frees the smem slot when these MMAs retire
                    if (it == k_iters - 1) umma_commit(&acc_full[acc]); // accumulator complete
                    if (++stage == p.stages) { stage = 0; phase ^= 1; da = desc0; } else da += step;
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..17)
        const int q = warp & 3;           // TMEM lane quarter this warp may read
        const int grp = (warp - 2) >> 2;  // the four warps of a quarter take column chunks grp, grp+4, ...
        const int r = q * 32 + lane;
        const int th = r >> p.TW_log2, tw = r & (TW - 1);
        const int Cq = p.pixel_shuffle ? p.Cout / (p.pixel_shuffle * p.pixel_shuffle) : 0;
        const int n_chunks = p.BN / 16;
        int acc = 0;
        uint32_t acc_phase = 0;
        // one 16-column chunk: bias, second operand, activation, store
        auto finish = [&](uint32_t (&accv)[16], int n, bool valid, long long pix, int b, int oh, int ow, int g) {
            if (!valid || n >= p.Cout) return;
            float v[16];
            {
                const float4 *bp = reinterpret_cast<const float4 *>(s_bias + g * (p.n_tiles * p.BN) + n);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 t = bp[j];
                    v[4 * j] = __uint_as_float(accv[4 * j]) + t.x; v[4 * j + 1] = __uint_as_float(accv[4 * j + 1]) + t.y;
                    v[4 * j + 2] = __uint_as_float(accv[4 * j + 2]) + t.z; v[4 * j + 3] = __uint_as_float(accv[4 * j + 3]) + t.w;
                }
            }
            // second operand of the epilogue: fp32 (residual stream) or bf16 (activations); loaded where it is applied so
            // that no 16-register copy of it is live across the activation
            auto combine = [&](int mode) {
                float rv[16];
                if (p.res_dtype == ICM_OUT_F32) {
                    const float *rp = reinterpret_cast<const float *>(p.residual) + pix * p.res_pitch + n;
                    if (p.wide_ld) {
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(rv[8 * j]), "=f"(rv[8 * j + 1]), "=f"(rv[8 * j + 2]),
                                         "=f"(rv[8 * j + 3]), "=f"(rv[8 * j + 4]), "=f"(rv[8 * j + 5]), "=f"(rv[8 * j + 6]), "=f"(rv[8 * j + 7]) : "l"(rp + 8 * j));
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) { const float4 t = __ldg(reinterpret_cast<const float4 *>(rp) + j); rv[4 * j] = t.x; rv[4 * j + 1] = t.y; rv[4 * j + 2] = t.z; rv[4 * j + 3] = t.w; }
                    }
                } else {
                    const __nv_bfloat16 *rp = reinterpret_cast<const __nv_bfloat16 *>(p.residual) + pix * p.res_pitch + n;
                    uint32_t u[8];
                    if (p.wide_ld) {
                        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]),
                                     "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "l"(rp));
                    } else {
                        const uint4 t0 = __ldg(reinterpret_cast<const uint4 *>(rp)), t1 = __ldg(reinterpret_cast<const uint4 *>(rp) + 1);
                        u[0] = t0.x; u[1] = t0.y; u[2] = t0.z; u[3] = t0.w; u[4] = t1.x; u[5] = t1.y; u[6] = t1.z; u[7] = t1.w;
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u[k]));
                        rv[2 * k] = f.x; rv[2 * k + 1] = f.y;
                    }
                }
                if (mode == 2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] *= rv[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += rv[j];
                }
            };
            if (p.residual && p.res_mode == 1) combine(1);
            apply_act16(v, p.act);
            long long off;
            if (p.pixel_shuffle) {
                const int rr = p.pixel_shuffle;
                const int quad = n / Cq, c = n - quad * Cq;
                const int i = quad / rr, jj = quad - i * rr;
                off = (((long long)b * p.Ho * rr + (long long)oh * rr + i) * ((long long)p.Wo * rr) + (long long)ow * rr + jj) * p.out_pitch + c;
            } else {
                off = pix * p.out_pitch + n;
            }
            if (p.residual && p.res_mode != 1) combine(p.res_mode);
            off += (long long)g * p.out_group_stride;
            // 32-byte stores (st.global.v8, sm_100): one full sector per lane and instruction instead of two half-sector writes
            if (p.out_dtype == ICM_OUT_F32) {
                float *op = reinterpret_cast<float *>(p.out) + off;
                if (p.wide_st) {
#pragma unroll
                    for (int j = 0; j < 2; ++j)
                        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op + 8 * j), "f"(v[8 * j]), "f"(v[8 * j + 1]),
                                     "f"(v[8 * j + 2]), "f"(v[8 * j + 3]), "f"(v[8 * j + 4]), "f"(v[8 * j + 5]), "f"(v[8 * j + 6]), "f"(v[8 * j + 7]) : "memory");
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) reinterpret_cast<float4 *>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            } else {
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    pk[j] = *reinterpret_cast<const uint32_t *>(&h2);
                }
                __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(p.out) + off;
                if (p.wide_st) {
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(op), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                                 "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                } else {
                    reinterpret_cast<uint4 *>(op)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    reinterpret_cast<uint4 *>(op)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        };
        int slot = 0;
        uint32_t qphase = 0;
        while (true) {
            mbar_wait(&tq_full[slot], qphase);
            const int tile = *reinterpret_cast<volatile int *>(&s_tile[slot]);
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[slot]);
            if (++slot == TQ) { slot = 0; qphase ^= 1; }
            if (tile < 0) break;
            int n0, w0, h0, b, g;
            tile_coords(tile, n0, w0, h0, b, g);
            const int oh = h0 + th, ow = w0 + tw;
            const bool valid = (oh < p.Ho) && (ow < p.Wo);
            const long long pix = ((long long)b * p.Ho + oh) * p.Wo + ow;
            mbar_wait(&acc_full[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(q * 32) << 16);
            // Software pipeline over this warp's chunks: the TMEM load of chunk i+1 is in flight while chunk i is
            // finished (ncu: 35 % of the epilogue's stall samples sat on the exposed tcgen05.ld / bias latency), and the
            // accumulator goes back to the MMA warp as soon as the last load has landed, before its chunk is finished.
            uint32_t ra[16], rb[16];
            int c16 = grp;
            bool released = false;
            auto release = [&]() {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);
                released = true;
            };
            if (c16 < n_chunks) tmem_ld16(tmem_d + (uint32_t)(c16 * 16), ra);
            while (c16 < n_chunks) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c16 + 4 < n_chunks) tmem_ld16(tmem_d + (uint32_t)((c16 + 4) * 16), rb); else release();
                finish(ra, n0 + c16 * 16, valid, pix, b, oh, ow, g);
                c16 += 4;
                if (c16 >= n_chunks) break;
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c16 + 4 < n_chunks) tmem_ld16(tmem_d + (uint32_t)((c16 + 4) * 16), ra); else release();
                finish(rb, n0 + c16 * 16, valid, pix, b, oh, ow, g);
                c16 += 4;
            }
            if (!released) release(); // a warp with no chunk of this tile
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
    if (threadIdx.x == 0) { // every CTA has drawn its last tile id before it counts itself as finished
        if (atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) {
            p.sched[0] = 0;
            p.sched[1] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------ weight packing
// OIHW fp32 -> bf16 [Cout_pad][KH*KW][Cin_pad]; with pixel_shuffle = r the output channels are permuted so
// that GEMM column (i*r + j) * (Cout / r^2) + c holds conv channel c*r^2 + i*r + j (nn.PixelShuffle order).
__global__ void pack_weight_kernel(const float *__restrict__ w, int Cout, int Cin, int KH, int KW, int Cin_pad, int Cout_pad,
                                   int ps, __nv_bfloat16 *__restrict__ out)
{
    const long long total = (long long)Cout_pad * KH * KW * Cin_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin_pad);
        long long t = i / Cin_pad;
        const int tap = (int)(t % (KH * KW));
        const int n = (int)(t / (KH * KW));
        float v = 0.f;
        if (n < Cout && ci < Cin) {
            int co = n;
            if (ps > 1) {
                const int Cq = Cout / (ps * ps);
                const int quad = n / Cq, c = n - quad * Cq;
                co = c * ps * ps + quad;
            }
            v = w[((long long)co * Cin + ci) * KH * KW + tap];
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------------------------ host side
// Tile counter of the dynamic scheduler: {next, finished}, one pair per (device, stream).  Launches on one stream run
// in order and the last CTA of a launch zeroes the pair, so a single pair per stream is enough; launches on different
// streams never share one.  Entries live for the life of the process (streams are few and long-lived here).
int *tile_counter(cudaStream_t st)
{
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, int *> slots;
    const std::pair<int, cudaStream_t> key(current_device_ordinal(), st);
    std::lock_guard<std::mutex> lock(mu);
    auto it = slots.find(key);
    if (it != slots.end()) return it->second;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
        set_error("tile_counter: first use of a stream inside a graph capture (run the path once before capturing)");
        return nullptr;
    }
    int *d = nullptr;
    if (cudaMalloc(&d, 256) != cudaSuccess || cudaMemset(d, 0, 256) != cudaSuccess) { // cudaMemset: synchronous, done before any launch uses it
        set_error("tile_counter: cudaMalloc failed");
        return nullptr;
    }
    slots[key] = d;
    return d;
}

static int pick_tile_w_log2(int Ho, int Wo)
{
    long long best = -1;
    int best_l = 7;
    for (int l = 7; l >= 3; --l) {
        const int TW = 1 << l, TH = BM >> l;
        const long long covered = (long long)((Wo + TW - 1) / TW) * TW * ((Ho + TH - 1) / TH) * TH;
        if (best < 0 || covered < best) { best = covered; best_l = l; }
    }
    return best_l;
}


}  // namespace icm

using namespace icm;

// Cap on the number of persistent CTAs (= SMs) icm_conv2d may occupy, 0 = all.  With the dynamic tile scheduler a
// launch no longer has to leave room for the rANS coder CTAs of other streams; a pipeline may still cap its
// throughput-bound launches so that a few SMs stay free for latency-bound ones.
extern "C" int icm_set_conv_sm_limit(int n_sms)
{
    ICM_CHECK_ARG(n_sms >= 0, "icm_set_conv_sm_limit: negative limit");
    set_persistent_grid_limit(n_sms);
    return ICM_OK;
}

static int launch_conv(const icm_conv_args *a, const icm_conv_groups *grp, void *stream)
{
    ICM_CHECK_ARG(a && a->in && a->weight && a->out, "icm_conv2d: null argument");
    const int G = grp ? grp->groups : 1;
    ICM_CHECK_ARG(G >= 1 && G <= ICM_MAX_CONV_GROUPS, "icm_conv2d_grouped: groups=%d outside 1..%d", G, ICM_MAX_CONV_GROUPS);
    ICM_CHECK_ARG(G == 1 || a->residual == nullptr, "icm_conv2d_grouped: a residual operand needs groups == 1");
    ICM_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0, "icm_conv2d: empty input");
    ICM_CHECK_ARG(a->Cin > 0 && a->Cin <= a->in_pitch && a->in_pitch % 8 == 0, "icm_conv2d: Cin=%d in_pitch=%d (pitch must be a multiple of 8 and >= Cin)", a->Cin, a->in_pitch);
    ICM_CHECK_ARG(a->Cout > 0 && a->Cout % 16 == 0, "icm_conv2d: Cout=%d must be a multiple of 16", a->Cout);
    ICM_CHECK_ARG(a->KH >= 1 && a->KW >= 1 && a->stride >= 1 && a->stride <= 2 && a->pad >= 0, "icm_conv2d: bad filter geometry");
    ICM_CHECK_ARG(((uintptr_t)a->in & 15) == 0 && ((uintptr_t)a->weight & 15) == 0 && ((uintptr_t)a->out & 15) == 0, "icm_conv2d: pointers must be 16-byte aligned");
    const int ps = a->pixel_shuffle;
    ICM_CHECK_ARG(ps == 0 || (ps >= 2 && a->Cout % (ps * ps) == 0 && (a->Cout / (ps * ps)) % 16 == 0), "icm_conv2d: pixel_shuffle=%d incompatible with Cout=%d", ps, a->Cout);
    ICM_CHECK_ARG(ps == 0 || a->residual == nullptr, "icm_conv2d: residual with pixel_shuffle is unsupported");
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("icm_conv2d: cuTensorMapEncodeTiled unavailable (no CUDA driver)"); return ICM_ERR_NO_DEVICE; }

    ConvParams p{};
    p.Ho = (a->H + 2 * a->pad - a->KH) / a->stride + 1;
    p.Wo = (a->W + 2 * a->pad - a->KW) / a->stride + 1;
    ICM_CHECK_ARG(p.Ho > 0 && p.Wo > 0, "icm_conv2d: empty output");
    p.TW_log2 = pick_tile_w_log2(p.Ho, p.Wo);
    const int TW = 1 << p.TW_log2, TH = BM >> p.TW_log2;
    p.tiles_w = (p.Wo + TW - 1) / TW;
    p.tiles_h = (p.Ho + TH - 1) / TH;
    const int Cin = a->Cin; // channels actually read; the packed weight pads each tap to a multiple of 64
    p.k_chunks = (Cin + BK - 1) / BK;
    p.Cin_pad = p.k_chunks * BK;
    p.last_k16 = (Cin - (p.k_chunks - 1) * BK + UMMA_K - 1) / UMMA_K;
    p.KH = a->KH; p.KW = a->KW; p.stride = a->stride; p.pad = a->pad;
    p.Cout = a->Cout;
    const int n_tiles = (a->Cout + 255) / 256;
    p.BN = ((a->Cout + n_tiles - 1) / n_tiles + 15) & ~15;
    // small problems: more, narrower N tiles so that more SMs take part
    const long long m_tiles = (long long)a->B * p.tiles_w * p.tiles_h;
    while (p.BN > 64 && p.BN % 32 == 0 && G * m_tiles * ((a->Cout + p.BN - 1) / p.BN) * 2 <= sm_count()) p.BN /= 2;
    p.tmem_cols = 64;
    while (p.tmem_cols < 2 * p.BN) p.tmem_cols *= 2; // two accumulators, power-of-two allocation, <= 512
    const int k_iters = p.KH * p.KW * p.k_chunks;
    const uint32_t a_bytes = BM * BK * 2, b_bytes = (uint32_t)p.BN * BK * 2;
    const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
    // one persistent CTA per SM: the pipeline is as deep as shared memory allows and runs across tiles
    p.stages = (int)((200 * 1024) / stage_bytes);
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    if (p.stages < 2) p.stages = 2;
    p.act = a->act; p.out_dtype = a->out_dtype; p.pixel_shuffle = ps;
    p.out_pitch = a->out_pitch; p.res_pitch = a->res_pitch;
    p.bias = a->bias; p.residual = a->residual; p.out = a->out;
    p.res_dtype = a->res_dtype; p.res_mode = a->res_mode;
    p.groups = G;
    int in_images = a->B;
    if (grp) {
        ICM_CHECK_ARG(grp->weight_group_rows >= ((a->Cout + 15) & ~15) || G == 1, "icm_conv2d_grouped: weight_group_rows smaller than the padded Cout");
        ICM_CHECK_ARG(grp->in_images >= a->B, "icm_conv2d_grouped: in_images < B");
        p.w_group_rows = (int)grp->weight_group_rows;
        p.bias_group_stride = (int)grp->bias_group_stride;
        p.out_group_stride = grp->out_group_stride;
        in_images = grp->in_images;
        for (int g = 0; g < G; ++g) {
            ICM_CHECK_ARG(grp->in_image_offset[g] >= 0 && grp->in_image_offset[g] + a->B <= grp->in_images, "icm_conv2d_grouped: group %d reads images outside the input", g);
            p.a_img[g] = grp->in_image_offset[g];
            p.tail_ch[g] = grp->tail_channel[g];
            if (grp->tail_channel[g] >= 0) p.has_tail = 1;
        }
        if (p.has_tail) {
            const int tail = Cin - (p.k_chunks - 1) * BK; // channels of the tail chunk
            for (int g = 0; g < G; ++g)
                ICM_CHECK_ARG(grp->tail_channel[g] >= 0 && grp->tail_channel[g] % 8 == 0 && grp->tail_channel[g] + tail <= a->in_pitch,
                              "icm_conv2d_grouped: tail_channel[%d]=%d (need every group set, 16-byte aligned, inside the row)", g, grp->tail_channel[g]);
        }
    }
    {
        const long long es = a->out_dtype == ICM_OUT_F32 ? 4 : 2;
        const int cq = ps ? a->Cout / (ps * ps) : 16;
        const long long rs = a->res_dtype == ICM_OUT_F32 ? 4 : 2;
        p.wide_ld = a->residual && ((uintptr_t)a->residual % 32 == 0) && ((long long)a->res_pitch * rs % 32 == 0);
        p.wide_st = ((uintptr_t)a->out % 32 == 0) && ((long long)a->out_pitch * es % 32 == 0) && (cq * es % 32 == 0);
    }
    ICM_CHECK_ARG(a->res_mode >= 0 && a->res_mode <= 2 && (a->res_dtype == ICM_OUT_F32 || a->res_dtype == ICM_OUT_BF16), "icm_conv2d: bad residual mode/dtype");
    ICM_CHECK_ARG(a->out_pitch % (a->out_dtype == ICM_OUT_F32 ? 4 : 8) == 0, "icm_conv2d: out_pitch=%d breaks 16-byte store alignment", a->out_pitch);
    ICM_CHECK_ARG(!a->residual || (a->res_pitch % (a->res_dtype == ICM_OUT_F32 ? 4 : 8) == 0 && ((uintptr_t)a->residual & 15) == 0), "icm_conv2d: residual must be 16-byte aligned");
    ICM_CHECK_ARG(!a->bias || ((uintptr_t)a->bias & 15) == 0, "icm_conv2d: bias must be 16-byte aligned");

    CUtensorMap map_a, map_w;
    {
        // a tail chunk is read at a channel of its own (up to the row pitch); otherwise channels beyond Cin are zero-filled
        cuuint64_t dims[4] = {(cuuint64_t)(p.has_tail ? a->in_pitch : Cin), (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)in_images};
        cuuint64_t strides[3] = {(cuuint64_t)a->in_pitch * 2, (cuuint64_t)a->W * a->in_pitch * 2, (cuuint64_t)a->H * a->W * a->in_pitch * 2};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)((TW - 1) * a->stride + 1), (cuuint32_t)((TH - 1) * a->stride + 1), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)a->stride, (cuuint32_t)a->stride, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(a->in), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("icm_conv2d: cuTensorMapEncodeTiled(A) failed (%d)", (int)r); return ICM_ERR_CUDA; }
    }
    {
        const cuuint64_t ktot = (cuuint64_t)p.KH * p.KW * p.Cin_pad;
        const int cout_pad = (a->Cout + 15) & ~15;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)(G > 1 ? (long long)(G - 1) * p.w_group_rows + cout_pad : cout_pad)};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)p.BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(a->weight), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("icm_conv2d: cuTensorMapEncodeTiled(W) failed (%d)", (int)r); return ICM_ERR_CUDA; }
    }
    p.n_tiles = (a->Cout + p.BN - 1) / p.BN;
    const size_t bias_bytes = (size_t)G * p.n_tiles * p.BN * 4;
    ICM_CHECK_ARG(bias_bytes <= 16 * 1024, "icm_conv2d: Cout=%d too wide for the bias staging area", a->Cout);
    const size_t smem_bytes = (size_t)p.stages * stage_bytes + 1024 /*alignment slack*/ + (2 * MAX_STAGES + 4 + 2 * TQ) * 8 + TQ * 4 + 16 + bias_bytes;
    static PerDeviceSmem configured;
    if (configured.needs(smem_bytes)) {
        ICM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured.done(227 * 1024);
    }
    ICM_CHECK_ARG(G * m_tiles * p.n_tiles < (1 << 24) && p.tiles_w < 65536 && p.tiles_h < 65536, "icm_conv2d: too many tiles");
    p.tiles_per_group = (int)(m_tiles * p.n_tiles);
    p.total_tiles = G * p.tiles_per_group;
    auto magic = [](int d) -> unsigned long long { return d <= 1 ? 0ull : ((1ull << 40) + (unsigned long long)d - 1) / (unsigned long long)d; };
    p.fd_n = magic(p.n_tiles); p.fd_w = magic(p.tiles_w); p.fd_h = magic(p.tiles_h);
    p.fd_g = G > 1 ? magic(p.tiles_per_group) : 0;
    p.sched = tile_counter(as_stream(stream));
    if (!p.sched) return ICM_ERR_CUDA;
    static const bool dyn_first = getenv("ICM_CONV_STATIC_FIRST") == nullptr; // A/B switch
    p.dyn_first = dyn_first ? 1 : 0;
    const int max_ctas = persistent_grid_limit();
    const int grid = p.total_tiles < max_ctas ? p.total_tiles : max_ctas;
    conv_igemm_kernel<<<grid, CONV_THREADS, smem_bytes, as_stream(stream)>>>(map_a, map_w, p);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_conv2d(const icm_conv_args *a, void *stream) { return launch_conv(a, nullptr, stream); }

extern "C" int icm_conv2d_grouped(const icm_conv_args *a, const icm_conv_groups *g, void *stream)
{
    ICM_CHECK_ARG(g, "icm_conv2d_grouped: null group description");
    return launch_conv(a, g, stream);
}

extern "C" int icm_pack_conv_weight(const float *d_w_oihw, int Cout, int Cin, int KH, int KW, int Cin_pad, int Cout_pad,
                                    int pixel_shuffle, void *d_out_bf16, void *stream)
{
    ICM_CHECK_ARG(d_w_oihw && d_out_bf16, "icm_pack_conv_weight: null argument");
    ICM_CHECK_ARG(Cin_pad >= Cin && Cin_pad % 64 == 0 && Cout_pad >= Cout, "icm_pack_conv_weight: bad padding");
    const long long total = (long long)Cout_pad * KH * KW * Cin_pad;
    const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 8);
    pack_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_w_oihw, Cout, Cin, KH, KW, Cin_pad, Cout_pad, pixel_shuffle,
                                                           reinterpret_cast<__nv_bfloat16 *>(d_out_bf16));
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}
