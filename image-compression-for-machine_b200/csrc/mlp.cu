// Fused Swin MLP on the Blackwell tensor cores (rows T3/T5 of SURVEY.md §8a: stf.py:34-40 inside :194-198):
//
//   x[row, :] += fc2( GELU( fc1( h[row, :] ) ) )          h = LayerNorm2(x) in bf16 (icm_layernorm), x the fp32 stream
//
// The 4C-wide hidden activation never leaves the SM.  Per 128-row tile the hidden dimension is walked in chunks of
// 64 columns: GEMM1 (128 x 64 x C) into one of two TMEM accumulators, an epilogue that adds the bias, applies GELU and
// writes the bf16 chunk straight into shared memory in the 128-byte-swizzled K-major layout the next MMA reads, and
// GEMM2 (128 x C x 64) accumulating the output tile in a third TMEM region.  The MMA warp issues GEMM1 of chunk j+1
// before GEMM2 of chunk j, so the tensor core works while the 16 epilogue warps run GELU.
//   warp 0      TMA: the h tile (double-buffered across tiles), then per chunk a W1 block [64 x C] and a W2 block
//               [C x 64] into two 2-slot rings (the weights come from L2 every tile: 4C*C*4 bytes)
//   warp 1      tcgen05.mma issuer
//   warps 2-17  epilogue 1 (per chunk) and epilogue 2 (bias + fp32 residual, 256-bit loads / stores)
// Separate fc1 / fc2 launches wrote and re-read the hidden tensor (2 x 2.4 GB per stage-0 block at 64 images).
// The K order of both products is that of conv.cu (ascending 64-channel chunks, four K = 16 MMAs each), so the result
// is bit-identical to the unfused path: the encoder and decoder sides may mix them freely.
#include "umma.cuh"

namespace icm {

constexpr int MLP_EPI_WARPS = 16;
constexpr int MLP_THREADS = (2 + MLP_EPI_WARPS) * 32;
constexpr int HC = 64; // hidden columns per chunk == one swizzle row of the second product's A operand
constexpr int MAX_RING = 8;

struct MlpParams {
    long long M;
    int C, k1_chunks, n_hc, tiles; // k1_chunks = ceil(C / 64), n_hc = 4C / 64
    int tmem_cols, bias_in_smem; // C = 192 has no shared memory left for the biases: read through L1 instead
    int ring, a1_bufs;           // depth of the W1 / W2 rings; h-tile buffers (2, or 1 when shared memory is short)
    const float *b1, *b2;
    float *x;
};

struct alignas(16) MlpBars {
    uint64_t a1_full[2], a1_empty[2], w1_full[MAX_RING], w1_empty[MAX_RING], w2_full[MAX_RING], w2_empty[MAX_RING];
    uint64_t d1_full[2], d1_empty[2], a2_full[2], a2_empty[2], d2_full, d2_empty;
    uint32_t tmem_slot, pad;
};

__global__ void __launch_bounds__(MLP_THREADS, 1)
swin_mlp_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_w1,
                const __grid_constant__ CUtensorMap map_w2, const MlpParams p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const uint32_t a1_bytes = (uint32_t)p.k1_chunks * (BM * 128);    // h tile: k1_chunks swizzled [128 x 64] blocks
    const uint32_t w1_bytes = (uint32_t)p.k1_chunks * (HC * 128);    // W1 block: [64 hidden rows x C]
    const uint32_t w2_bytes = (((uint32_t)p.C * 128) + 1023) & ~1023u; // W2 block: [C rows x 64 hidden]
    const uint32_t a2_bytes = BM * 128;                              // GELU chunk [128 x 64]
    unsigned char *a1 = base, *w1 = a1 + p.a1_bufs * a1_bytes, *w2 = w1 + p.ring * w1_bytes, *a2 = w2 + p.ring * w2_bytes;
    MlpBars *bar = reinterpret_cast<MlpBars *>(a2 + 2 * a2_bytes);
    float *s_b1 = reinterpret_cast<float *>(bar + 1), *s_b2 = s_b1 + 4 * p.C;

    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar->a1_full[i], 1); mbar_init(&bar->a1_empty[i], 1);
            mbar_init(&bar->d1_full[i], 1); mbar_init(&bar->d1_empty[i], MLP_EPI_WARPS);
            mbar_init(&bar->a2_full[i], MLP_EPI_WARPS); mbar_init(&bar->a2_empty[i], 1);
        }
        for (int i = 0; i < MAX_RING; ++i) {
            mbar_init(&bar->w1_full[i], 1); mbar_init(&bar->w1_empty[i], 1);
            mbar_init(&bar->w2_full[i], 1); mbar_init(&bar->w2_empty[i], 1);
        }
        mbar_init(&bar->d2_full, 1); mbar_init(&bar->d2_empty, MLP_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.bias_in_smem) {
        for (int i = threadIdx.x; i < 4 * p.C; i += blockDim.x) s_b1[i] = p.b1[i];
        for (int i = threadIdx.x; i < p.C; i += blockDim.x) s_b2[i] = p.b2[i];
    }
    const float *b1p = p.bias_in_smem ? s_b1 : p.b1, *b2p = p.bias_in_smem ? s_b2 : p.b2;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar->tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = bar->tmem_slot;
    const uint32_t d2_col = 2 * HC; // TMEM columns: D1[0] at 0, D1[1] at 64, D2 at 128

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
            auto load_h = [&](int tile, int slot, uint32_t parity) {
                mbar_wait(&bar->a1_empty[slot], parity ^ 1);
                mbar_expect_tx(&bar->a1_full[slot], a1_bytes);
                for (int kc = 0; kc < p.k1_chunks; ++kc)
                    tma_load_2d(&map_h, &bar->a1_full[slot], a1 + slot * a1_bytes + kc * (BM * 128), kc * 64, tile * BM);
            };
            int it = 0;                      // tiles done by this CTA
            int ws = 0;                      // weight ring slot and its phase
            uint32_t wpar = 0;
            const int ab = p.a1_bufs;        // h tile t lives in buffer t % ab, phase (t / ab) & 1
            if ((int)blockIdx.x < p.tiles) load_h(blockIdx.x, 0, 0);
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
                const int next = tile + gridDim.x;
                if (ab == 2 && next < p.tiles) load_h(next, (it + 1) & 1, (uint32_t)(((it + 1) >> 1) & 1)); // prefetch the next h tile
                for (int j = 0; j < p.n_hc; ++j) {
                    const int s = ws;
                    const uint32_t par = wpar;
                    if (++ws == p.ring) { ws = 0; wpar ^= 1; }
                    mbar_wait(&bar->w1_empty[s], par ^ 1);
                    mbar_expect_tx(&bar->w1_full[s], w1_bytes);
                    for (int kc = 0; kc < p.k1_chunks; ++kc)
                        tma_load_2d(&map_w1, &bar->w1_full[s], w1 + s * w1_bytes + kc * (HC * 128), kc * 64, j * HC);
                    mbar_wait(&bar->w2_empty[s], par ^ 1);
                    mbar_expect_tx(&bar->w2_full[s], (uint32_t)p.C * 128);
                    tma_load_2d(&map_w2, &bar->w2_full[s], w2 + s * w2_bytes, j * HC, 0);
                }
                // single h buffer: the next tile's h can only be requested once every GEMM1 of this tile has read it
                if (ab == 1 && next < p.tiles) load_h(next, 0, (uint32_t)((it + 1) & 1));
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // (one thread runs the whole role, waits included: no per-call elect / reconvergence)
        if (elect_one()) {
        const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HC >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.C >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        int it = 0;
        uint32_t c1 = 0, c2 = 0; // chunks issued to GEMM1 / GEMM2 over the whole kernel (D1 / A2 buffer = count & 1)
        int r1 = 0, r2 = 0;      // W1 / W2 ring slots and phases
        uint32_t rp1 = 0, rp2 = 0;
        auto gemm1 = [&](int slot_a, bool last) {
            const int s = c1 & 1;
            const uint32_t par = (c1 >> 1) & 1;
            mbar_wait(&bar->w1_full[r1], rp1);
            mbar_wait(&bar->d1_empty[s], par ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const uint32_t d = tmem_base + (uint32_t)s * HC;
                for (int kc = 0; kc < p.k1_chunks; ++kc) {
                    const uint64_t da = make_smem_desc(smem_u32(a1 + slot_a * a1_bytes + kc * (BM * 128)));
                    const uint64_t db = make_smem_desc(smem_u32(w1 + r1 * w1_bytes + kc * (HC * 128)));
                    const int n_k = (kc == p.k1_chunks - 1) ? (p.C - kc * 64 + UMMA_K - 1) / UMMA_K : BK / UMMA_K; // skip all-zero K steps
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)
                        if (k < n_k) umma_bf16(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc1, (kc | k) != 0);
                }
                umma_commit(&bar->w1_empty[r1]);
                umma_commit(&bar->d1_full[s]);
                if (last) umma_commit(&bar->a1_empty[slot_a]); // every GEMM1 of this tile has read the h tile
            }
            ++c1;
            if (++r1 == p.ring) { r1 = 0; rp1 ^= 1; }
        };
        auto gemm2 = [&](bool first, bool last) {
            const int s = c2 & 1;
            const uint32_t par = (c2 >> 1) & 1;
            mbar_wait(&bar->a2_full[s], par);
            mbar_wait(&bar->w2_full[r2], rp2);
            if (first) mbar_wait(&bar->d2_empty, (uint32_t)((it & 1) ^ 1)); // epilogue 2 of the previous tile has drained D2
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                const uint64_t da = make_smem_desc(smem_u32(a2 + s * a2_bytes)), db = make_smem_desc(smem_u32(w2 + r2 * w2_bytes));
#pragma unroll
                for (int k = 0; k < HC / UMMA_K; ++k) umma_bf16(tmem_base + d2_col, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc2, (!first) || k != 0);
                umma_commit(&bar->a2_empty[s]);
                umma_commit(&bar->w2_empty[r2]);
                if (last) umma_commit(&bar->d2_full);
            }
            ++c2;
            if (++r2 == p.ring) { r2 = 0; rp2 ^= 1; }
        };
        // One continuous chunk stream across tiles: GEMM1 of a chunk is always issued before GEMM2 of the previous chunk,
        // also when the two belong to different tiles -- otherwise the first GEMM1 of a tile waited for the last GELU chunk
        // of the tile before, and the 16 epilogue warps idled for a full MMA round trip once per tile.
        int n_my = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) ++n_my;
        const long long total = (long long)n_my * p.n_hc;
        int j1 = 0, t1 = 0; // chunk / tile ordinal of the next GEMM1
        int j2 = 0;         // chunk of the next GEMM2 (its tile ordinal is `it`, used for the D2 hand-shake)
        for (long long g = 0; g <= total; ++g) {
            if (g < total) {
                const int slot_a = p.a1_bufs == 2 ? (t1 & 1) : 0;
                if (j1 == 0) mbar_wait(&bar->a1_full[slot_a], (uint32_t)(p.a1_bufs == 2 ? ((t1 >> 1) & 1) : (t1 & 1)));
                gemm1(slot_a, j1 == p.n_hc - 1);
                if (++j1 == p.n_hc) { j1 = 0; ++t1; }
            }
            if (g > 0) {
                gemm2(j2 == 0, j2 == p.n_hc - 1);
                if (++j2 == p.n_hc) { j2 = 0; ++it; }
            }
        }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps
        const int q = warp & 3;          // TMEM lane quarter
        const int grp = (warp - 2) >> 2; // 16-column group of a 64-column chunk
        const int r = q * 32 + lane;     // row of the tile
        int it = 0, prev_tile = -1;
        uint32_t ce = 0; // chunks finished
        // epilogue 2: output tile + bias + residual, in place.  It runs one chunk late -- after the first GELU chunk of the NEXT
        // tile -- so that the last GEMM2 of its own tile completes behind useful work instead of in front of a stalled epilogue.
        auto epilogue2 = [&](int tile2, int it2) {
            mbar_wait(&bar->d2_full, (uint32_t)(it2 & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long grow = (long long)tile2 * BM + r;
            const int n_c16 = p.C / 16;
            uint32_t acc[16];
            bool have = false;
            if (grp < n_c16) { tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + d2_col + (uint32_t)(grp * 16), acc); have = true; }
            // C <= 192: at most three chunks per warp (grp, grp + 4, grp + 8); loads of later chunks are issued one ahead
            for (int c16 = grp; c16 < n_c16; c16 += 4) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(acc[k]);
                if (c16 + 4 < n_c16) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + d2_col + (uint32_t)((c16 + 4) * 16), acc);
                else {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar->d2_empty);
                }
                if (grow < p.M) {
                    float *xp = p.x + grow * p.C + c16 * 16;
                    const float4 *bp = reinterpret_cast<const float4 *>(b2p + c16 * 16);
                    float rv[16];
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=f"(rv[8 * k]), "=f"(rv[8 * k + 1]), "=f"(rv[8 * k + 2]),
                                     "=f"(rv[8 * k + 3]), "=f"(rv[8 * k + 4]), "=f"(rv[8 * k + 5]), "=f"(rv[8 * k + 6]), "=f"(rv[8 * k + 7]) : "l"(xp + 8 * k));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 t = bp[k];
                        v[4 * k] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) v[k] += rv[k]; // same order as conv.cu: (acc + bias) + residual
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(xp + 8 * k), "f"(v[8 * k]), "f"(v[8 * k + 1]),
                                     "f"(v[8 * k + 2]), "f"(v[8 * k + 3]), "f"(v[8 * k + 4]), "f"(v[8 * k + 5]), "f"(v[8 * k + 6]), "f"(v[8 * k + 7]) : "memory");
                }
            }
            if (!have) { // a warp without a chunk of the output tile still owes its arrival
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->d2_empty);
            }
        };
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++it) {
            {   // epilogue 2 reads this thread's residual row once per tile: ask L2 for it now (ncu: a quarter of the stall
                // samples sat on those loads with their full HBM latency exposed)
                const long long grow = (long long)tile * BM + r;
                if (grow < p.M)
                    for (int c16 = grp; c16 < p.C / 16; c16 += 4)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + grow * p.C + c16 * 16) : "memory");
            }
            for (int j = 0; j < p.n_hc; ++j, ++ce) {
                const int s = ce & 1;
                const uint32_t par = (ce >> 1) & 1;
                float4 bq[4]; // the chunk's bias does not depend on the accumulator: fetch it before waiting
                {
                    const float4 *bp = reinterpret_cast<const float4 *>(b1p + j * HC + grp * 16);
#pragma unroll
                    for (int k = 0; k < 4; ++k) bq[k] = bp[k];
                }
                mbar_wait(&bar->d1_full[s], par);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t acc[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * HC + grp * 16), acc);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->d1_empty[s]);
                float v[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 t = bq[k];
                    v[4 * k] = gelu_erf(__uint_as_float(acc[4 * k]) + t.x); v[4 * k + 1] = gelu_erf(__uint_as_float(acc[4 * k + 1]) + t.y);
                    v[4 * k + 2] = gelu_erf(__uint_as_float(acc[4 * k + 2]) + t.z); v[4 * k + 3] = gelu_erf(__uint_as_float(acc[4 * k + 3]) + t.w);
                }
                uint32_t pk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
                    pk[k] = *reinterpret_cast<const uint32_t *>(&h2);
                }
                mbar_wait(&bar->a2_empty[s], par ^ 1); // GEMM2 of the chunk that used this buffer has read it
                unsigned char *row = a2 + s * a2_bytes + r * 128;
                const int j0 = grp * 2; // 16-byte piece index inside the 128-byte row, XOR-swizzled by the row
                *reinterpret_cast<uint4 *>(row + ((j0 ^ (r & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *reinterpret_cast<uint4 *>(row + (((j0 + 1) ^ (r & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar->a2_full[s]);
                if (j == 0 && prev_tile >= 0) epilogue2(prev_tile, it - 1);
            }
            prev_tile = tile;
        }
        if (prev_tile >= 0) epilogue2(prev_tile, it - 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

}  // namespace icm

using namespace icm;

extern "C" int icm_swin_mlp(const void *d_h_bf16, const void *d_w1_packed, const float *d_b1, const void *d_w2_packed, const float *d_b2,
                            float *d_x, int64_t rows, int C, void *stream)
{
    ICM_CHECK_ARG(d_h_bf16 && d_w1_packed && d_w2_packed && d_x && d_b1 && d_b2, "icm_swin_mlp: null argument");
    ICM_CHECK_ARG(((uintptr_t)d_b1 & 15) == 0 && ((uintptr_t)d_b2 & 15) == 0, "icm_swin_mlp: biases must be 16-byte aligned");
    ICM_CHECK_ARG(rows > 0, "icm_swin_mlp: no rows");
    if (!(C == 48 || C == 96 || C == 192)) { set_error("icm_swin_mlp: C=%d is not built (48, 96, 192)", C); return ICM_ERR_UNSUPPORTED; }
    ICM_CHECK_ARG(((uintptr_t)d_h_bf16 & 15) == 0 && ((uintptr_t)d_x & 31) == 0 && ((uintptr_t)d_w1_packed & 15) == 0 && ((uintptr_t)d_w2_packed & 15) == 0,
                  "icm_swin_mlp: misaligned pointer");
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("icm_swin_mlp: cuTensorMapEncodeTiled unavailable (no CUDA driver)"); return ICM_ERR_NO_DEVICE; }
    MlpParams p{};
    p.M = rows; p.C = C;
    p.k1_chunks = (C + 63) / 64;
    p.n_hc = 4 * C / HC;
    p.tiles = (int)((rows + BM - 1) / BM);
    p.tmem_cols = 256;
    while (p.tmem_cols < 2 * HC + C) p.tmem_cols *= 2;
    p.b1 = d_b1; p.b2 = d_b2; p.x = d_x;
    const int k1pad = p.k1_chunks * 64;
    CUtensorMap map_h, map_w1, map_w2;
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)BM}, estr[2] = {1, 1};
        CUresult r = enc(&map_h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_h_bf16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("icm_swin_mlp: cuTensorMapEncodeTiled(h) failed (%d)", (int)r); return ICM_ERR_CUDA; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)k1pad, (cuuint64_t)(4 * C)};
        cuuint64_t strides[1] = {(cuuint64_t)k1pad * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)HC}, estr[2] = {1, 1};
        CUresult r = enc(&map_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_w1_packed), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("icm_swin_mlp: cuTensorMapEncodeTiled(W1) failed (%d)", (int)r); return ICM_ERR_CUDA; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)(4 * C), (cuuint64_t)C};
        cuuint64_t strides[1] = {(cuuint64_t)(4 * C) * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)C}, estr[2] = {1, 1};
        CUresult r = enc(&map_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_w2_packed), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("icm_swin_mlp: cuTensorMapEncodeTiled(W2) failed (%d)", (int)r); return ICM_ERR_CUDA; }
    }
    const size_t a1_bytes = (size_t)p.k1_chunks * (BM * 128), w1_bytes = (size_t)p.k1_chunks * (HC * 128);
    const size_t w2_bytes = (((size_t)C * 128) + 1023) & ~(size_t)1023, a2_bytes = BM * 128;
    // The weight blocks of a chunk are used once and re-fetched from L2 for every tile; with a 2-slot ring that ~1.4 us
    // TMA round trip sat on the critical path of every second chunk (2 700 cycles per chunk measured against a 1 000-cycle
    // GELU bound).  The rings are as deep as shared memory allows; at C = 192 the h tile gives up its second buffer.
    const size_t budget = 227 * 1024, fixed = 1024 + 2 * a2_bytes + sizeof(MlpBars) + 16;
    p.a1_bufs = 2;
    p.ring = (int)((budget - fixed - 2 * a1_bytes) / (w1_bytes + w2_bytes));
    if (p.ring < 3) { p.a1_bufs = 1; p.ring = (int)((budget - fixed - a1_bytes) / (w1_bytes + w2_bytes)); }
    if (p.ring > MAX_RING) p.ring = MAX_RING;
    ICM_CHECK_ARG(p.ring >= 2, "icm_swin_mlp: shared memory budget exceeded");
    size_t smem_bytes = fixed + p.a1_bufs * a1_bytes + (size_t)p.ring * (w1_bytes + w2_bytes);
    p.bias_in_smem = smem_bytes + (size_t)5 * C * 4 <= budget;
    if (p.bias_in_smem) smem_bytes += (size_t)5 * C * 4;
    ICM_CHECK_ARG(smem_bytes <= 227 * 1024, "icm_swin_mlp: shared memory budget exceeded (%zu bytes)", smem_bytes);
    static PerDeviceSmem configured;
    if (configured.needs(smem_bytes)) {
        ICM_CUDA(cudaFuncSetAttribute(swin_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured.done(227 * 1024);
    }
    const int max_ctas = persistent_grid_limit();
    const int grid = p.tiles < max_ctas ? p.tiles : max_ctas;
    swin_mlp_kernel<<<grid, MLP_THREADS, smem_bytes, as_stream(stream)>>>(map_h, map_w1, map_w2, p);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}
