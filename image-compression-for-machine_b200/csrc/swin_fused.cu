// Fused Swin transformer block for the narrow STF stages (C = 48, 96): T3 + T4 + T5 of SURVEY.md §8a in ONE pass over
// the residual stream (stf.py:149-199):
//
//     x += proj( softmax( q k^T / 4 + B[rel] + mask ) v ),   q, k, v = qkv( LayerNorm1(x) )        (attention half)
//     x += fc2( GELU( fc1( LayerNorm2(x) ) ) )                                                      (MLP half)
//
// At C = 48 / 96 these stages are HBM-bound: six launches per block (LayerNorm, qkv, attention, proj, LayerNorm, MLP)
// move the fp32 residual stream and the bf16 intermediates ~14 bytes-per-element-times through memory, 3.75 ms per
// stage-0 block at 64 images, while the arithmetic is 0.36 TFLOP.  Here ONE WARP owns ONE 4x4 WINDOW (16 tokens = the M
// of mma.sync.m16n8k16) and keeps everything in registers from the load of x to its store:
//   * x rows are loaded in the A-fragment layout (lane (g, tq) holds channels 4tq..4tq+3 of every 16-channel group of
//     rows g and g+8: one 16-byte load per group, a quad covers 64 contiguous bytes), LayerNorm is a quad reduction;
//   * every GEMM-shaped piece is mma.sync bf16 with fp32 accumulation; the accumulator fragment of one product is
//     the A fragment of the next (q -> S, P -> PV, attention output -> proj, GELU(fc1) -> fc2), K's accumulators are
//     S's B fragments as they are, V's go through movmatrix;
//   * the contraction index inside a 16-group is permuted (fragment slots (2tq, 2tq+1, 8+2tq, 9+2tq) <-> channels
//     4tq..4tq+3) consistently on both operands, and output features are assigned to accumulator columns by the same
//     permutation, so every weight fragment is ONE 8-byte shared-memory load and the final accumulators line up with
//     the x registers for the residual add;
//   * weights live in shared memory as bf16 rows, in the order the fragments read them, with a stride of 8 or 24
//     (mod 32) words: conflict-free LDS.64.
// The cyclic shift, the window partition / reverse and the SW-MSA region mask are index arithmetic, as in
// window_attention_kernel (transforms.cu).  A window reads and writes only its own 16 tokens, so x is updated in place.
// Results are deterministic and batch-invariant (fixed reduction order per window), which is all the codec needs:
// encoder and decoder run the same kernel.
#include <stdlib.h>

#include "mma_sync.cuh"
#include "umma.cuh"

namespace icm {

namespace sf {

constexpr int WIN = 4, HD = 16;

// shared-memory row stride (in bf16 elements) for a [N][K] weight: words per row == 8 or 24 (mod 32)
__host__ __device__ constexpr int wstride(int K)
{
    int words = K / 2;
    while ((words % 32) != 8 && (words % 32) != 24) words += 4; // K is a multiple of 16: words is a multiple of 8
    return words * 2;
}

template <int C, bool ATTN, bool MLP>
struct Layout {
    static constexpr int S_C = wstride(C), S_H = wstride(4 * C);
    static constexpr size_t qkv_w = 0;
    static constexpr size_t proj_w = qkv_w + (ATTN ? (size_t)3 * C * S_C * 2 : 0);
    static constexpr size_t fc1_w = proj_w + (ATTN ? (size_t)C * S_C * 2 : 0);
    static constexpr size_t fc2_w = fc1_w + (MLP ? (size_t)4 * C * S_C * 2 : 0);
    static constexpr size_t f32_base = fc2_w + (MLP ? (size_t)C * S_H * 2 : 0);
    // fp32 parameters: ln1 g/b [2C], qkv bias [3C], proj bias [C], rel-pos table [49 * heads], ln2 g/b [2C], fc1 bias [4C], fc2 bias [C]
    static constexpr int heads = C / HD;
    static constexpr int o_ln1 = 0, o_bqkv = 2 * C, o_bproj = 5 * C, o_rel = 6 * C, o_ln2 = 6 * C + ((49 * heads + 3) & ~3),
                         o_b1 = o_ln2 + 2 * C, o_b2 = o_b1 + 4 * C, n_f32 = o_b2 + C;
    static constexpr size_t bytes = f32_base + (size_t)n_f32 * 4;
};

struct Params {
    float *x;  // [B*H*W, C] fp32 residual stream, updated in place
    int B, H, W, shift;
    const __nv_bfloat16 *w_qkv, *w_proj, *w_fc1, *w_fc2; // [N][ld] bf16 rows (icm_pack_conv_weight layout for 1x1: ld = Cin padded to 64)
    int ld_c, ld_h;                                      // row pitch of the K = C and K = 4C weights
    const float *ln1_g, *ln1_b, *b_qkv, *b_proj, *rel_table, *ln2_g, *ln2_b, *b_fc1, *b_fc2;
    int *sched; // {next window to hand out, CTAs finished}: the stream's scheduling counter pair (csrc/conv.cu tile_counter)
    int dynamic; // 0: static grid-stride walk (A/B switch ICM_SWIN_STATIC)
};

// feature held by accumulator column (2tq + e) of n-tile nt.  The weight row for column g of n-tile nt is feature
// 16 (nt >> 1) + 4 (g >> 1) + 2 (nt & 1) + (g & 1); rows are stored in shared memory in (n-tile, g) order (prow below), so
// that the eight rows a B-fragment load touches are consecutive: with a stride of 8 or 24 (mod 32) words the four rows
// of a half-warp then cover all 32 banks exactly once (rows 0,1,4,5 of the natural order would collide two by two).
__device__ __forceinline__ int feat(int nt, int tq, int e) { return (nt >> 1) * 16 + 4 * tq + 2 * (nt & 1) + e; }
__device__ __forceinline__ int wrow(int nt, int g) { return nt * 8 + g; }
__device__ __forceinline__ int prow(int n) { return ((n >> 4) * 2 + ((n >> 1) & 1)) * 8 + ((n >> 2) & 3) * 2 + (n & 1); } // feature -> stored row

__device__ __forceinline__ uint2 lds64(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint2 *>(p); }

// LayerNorm of the warp's 16 x C tile held as xr[kt] = channels kt*16 + 4tq..+3 of rows g (a) and g+8 (b); writes the bf16
// A fragments of every 16-channel group
template <int C>
__device__ __forceinline__ void layernorm_frag(const float4 (&xa)[C / 16], const float4 (&xb)[C / 16], const float *gb, int tq,
                                               uint32_t (&af)[C / 16][4])
{
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int kt = 0; kt < C / 16; ++kt) {
        sa += (xa[kt].x + xa[kt].y) + (xa[kt].z + xa[kt].w);
        sb += (xb[kt].x + xb[kt].y) + (xb[kt].z + xb[kt].w);
    }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    const float ma = sa * (1.0f / C), mb = sb * (1.0f / C);
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int kt = 0; kt < C / 16; ++kt) {
        float d;
        d = xa[kt].x - ma; qa += d * d; d = xa[kt].y - ma; qa += d * d; d = xa[kt].z - ma; qa += d * d; d = xa[kt].w - ma; qa += d * d;
        d = xb[kt].x - mb; qb += d * d; d = xb[kt].y - mb; qb += d * d; d = xb[kt].z - mb; qb += d * d; d = xb[kt].w - mb; qb += d * d;
    }
    qa += __shfl_xor_sync(0xffffffffu, qa, 1); qa += __shfl_xor_sync(0xffffffffu, qa, 2);
    qb += __shfl_xor_sync(0xffffffffu, qb, 1); qb += __shfl_xor_sync(0xffffffffu, qb, 2);
    const float ra = rsqrtf(qa * (1.0f / C) + 1e-5f), rb = rsqrtf(qb * (1.0f / C) + 1e-5f);
#pragma unroll
    for (int kt = 0; kt < C / 16; ++kt) {
        const float4 gm = *reinterpret_cast<const float4 *>(gb + kt * 16 + 4 * tq), bt = *reinterpret_cast<const float4 *>(gb + C + kt * 16 + 4 * tq);
        // slots (2tq, 2tq+1) = channels (4tq, 4tq+1); slots (8+2tq, 9+2tq) = channels (4tq+2, 4tq+3)
        af[kt][0] = pack_bf16((xa[kt].x - ma) * ra * gm.x + bt.x, (xa[kt].y - ma) * ra * gm.y + bt.y);
        af[kt][1] = pack_bf16((xb[kt].x - mb) * rb * gm.x + bt.x, (xb[kt].y - mb) * rb * gm.y + bt.y);
        af[kt][2] = pack_bf16((xa[kt].z - ma) * ra * gm.z + bt.z, (xa[kt].w - ma) * ra * gm.w + bt.w);
        af[kt][3] = pack_bf16((xb[kt].z - mb) * rb * gm.z + bt.z, (xb[kt].w - mb) * rb * gm.w + bt.w);
    }
}

// acc[w] (one n-tile pair = 16 output features starting at n-tile 2*j) = bias + A_w[16 x K] * W[rows of the pair][K]^T for the
// warp's NW windows: every weight fragment is loaded from shared memory once and used NW times
template <int KT, int NW>
__device__ __forceinline__ void gemm_pair(float (&acc)[NW][2][4], const uint32_t (&af)[NW][KT][4], const __nv_bfloat16 *w, int stride, int j,
                                          const float *bias, int g, int tq)
{
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int nt = 2 * j + p;
        const float2 b = *reinterpret_cast<const float2 *>(bias + feat(nt, tq, 0));
#pragma unroll
        for (int w_ = 0; w_ < NW; ++w_) { acc[w_][p][0] = b.x; acc[w_][p][1] = b.y; acc[w_][p][2] = b.x; acc[w_][p][3] = b.y; }
        const __nv_bfloat16 *wr = w + (size_t)wrow(nt, g) * stride + 4 * tq;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            const uint2 bf = lds64(wr + kt * 16);
#pragma unroll
            for (int w_ = 0; w_ < NW; ++w_) mma_bf16_16816(acc[w_][p], af[w_][kt], bf.x, bf.y);
        }
    }
}

// GELU for the register-resident MLP: 0.5 v (1 + tanh(sqrt(2/pi) (v + 0.044715 v^3))) with tanh.approx -- 6 instructions and
// ONE MUFU op per element (the erf-polynomial form of the conv epilogues takes 10 and two).  |error| against erf-GELU
// <= 5e-4 absolute, an eighth of the bf16 rounding of the value it is stored as; the kernel is instruction-issue-bound
// (ncu: 3 100 instructions per window, 1 150 of them GELU) and both sides of the codec run this same code.
__device__ __forceinline__ float gelu_tanh(float v)
{
    const float v2 = v * v;
    const float inner = v * fmaf(v2, 0.0356774081f, 0.7978845608f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
    const float hv = 0.5f * v;
    return fmaf(hv, t, hv);
}

// WARPS warps per CTA (MINB CTAs per SM), NW windows per warp at a time.  Measured on B200, 64 images (tools/swin_shape_ab.py; every
// shape gives the same bits):
//   C = 48: 8 warps x 2 CTAs, NW = 1: 1.52 ms per block; 10 / 12 warps x 2 CTAs (96 / 80 registers, spills): 1.54 / 1.62;
//           NW = 2 (every weight fragment feeds two windows' MMAs: half the shared-memory traffic, 225 registers, one CTA):
//           1.89 with 8 warps, 1.72 with 12 -- the kernel is bound by its dependent-issue chains, not by the LSU pipe;
//   C = 96: one window takes 170-190 registers, one CTA per SM: 8 / 12 / 14 / 16 warps: 1.46 / 1.31-1.34 / 1.67 / 1.62 ms
//           (attention + MLP passes; 12 warps cap the kernel at 168 registers with 30-70 bytes of spills, 14 and 16 at 128).
template <int C, bool ATTN, bool MLP, int WARPS, int NW, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) swin_block_kernel(const Params p)
{
    using L = Layout<C, ATTN, MLP>;
    constexpr int KT = C / 16, NT = C / 8, heads = C / HD;
    extern __shared__ __align__(16) unsigned char smem[];
    __nv_bfloat16 *s_qkv = reinterpret_cast<__nv_bfloat16 *>(smem + L::qkv_w), *s_proj = reinterpret_cast<__nv_bfloat16 *>(smem + L::proj_w);
    __nv_bfloat16 *s_fc1 = reinterpret_cast<__nv_bfloat16 *>(smem + L::fc1_w), *s_fc2 = reinterpret_cast<__nv_bfloat16 *>(smem + L::fc2_w);
    float *s_f = reinterpret_cast<float *>(smem + L::f32_base);
    {   // stage the weights (8-byte pieces: both pitches are multiples of 4 elements) and the fp32 parameters
        auto stage_w = [&](__nv_bfloat16 *dst, const __nv_bfloat16 *src, int N, int K, int ld, int stride) {
            const int per_row = K / 4;
            for (int i = threadIdx.x; i < N * per_row; i += blockDim.x) {
                const int n = i / per_row, c = (i - n * per_row) * 4;
                *reinterpret_cast<uint2 *>(dst + (size_t)prow(n) * stride + c) = __ldg(reinterpret_cast<const uint2 *>(src + (size_t)n * ld + c));
            }
        };
        auto stage_f = [&](int off, const float *src, int n) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) s_f[off + i] = src ? src[i] : 0.f;
        };
        if (ATTN) {
            stage_w(s_qkv, p.w_qkv, 3 * C, C, p.ld_c, L::S_C);
            stage_w(s_proj, p.w_proj, C, C, p.ld_c, L::S_C);
            stage_f(L::o_ln1, p.ln1_g, C); stage_f(L::o_ln1 + C, p.ln1_b, C);
            stage_f(L::o_bqkv, p.b_qkv, 3 * C); stage_f(L::o_bproj, p.b_proj, C);
            stage_f(L::o_rel, p.rel_table, 49 * heads);
        }
        if (MLP) {
            stage_w(s_fc1, p.w_fc1, 4 * C, C, p.ld_c, L::S_C);
            stage_w(s_fc2, p.w_fc2, C, 4 * C, p.ld_h, L::S_H);
            stage_f(L::o_ln2, p.ln2_g, C); stage_f(L::o_ln2 + C, p.ln2_b, C);
            stage_f(L::o_b1, p.b_fc1, 4 * C); stage_f(L::o_b2, p.b_fc2, C);
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int H = p.H, W = p.W, shift = p.shift;
    const int nWw = W / WIN, nWh = H / WIN;
    const long long n_win = (long long)p.B * nWh * nWw;
    // relative-position-bias offsets of this lane's 2 x 4 score elements (rows g / g+8, keys 2tq, 2tq+1, 8+2tq, 9+2tq):
    // they depend on the lane only, not on the window or the head (stf.py:70-80)
    int rel[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = g + 8 * r, j = (c >> 1) * 8 + 2 * tq + (c & 1);
            rel[r][c] = L::o_rel + (((i >> 2) - (j >> 2) + WIN - 1) * (2 * WIN - 1) + ((i & 3) - (j & 3) + WIN - 1)) * heads;
        }
    // Windows are handed out dynamically, NW per warp at a time, from a counter in global memory (the next index is fetched while
    // the current windows are processed).  A static grid-stride walk reserves a fixed share of the windows for every CTA, and in
    // the serving pipeline a CTA whose SM is held by an rANS decoder CTA of another stream (~190 KB of shared memory, 5 ms per
    // step) starts only when another CTA of this launch has finished -- the launch then takes twice as long.
    int nxt = ((int)blockIdx.x * WARPS + warp) * NW;
    if (p.dynamic) {
        if (lane == 0) nxt = atomicAdd(p.sched, NW);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
    }
    while (nxt < n_win) {
        const long long win0 = nxt;
        if (!p.dynamic) nxt += (int)gridDim.x * WARPS * NW;
        else if (lane == 0) nxt = atomicAdd(p.sched, NW); // in flight while these windows are processed
        float4 *pa[NW], *pb[NW];
        float4 xa[NW][KT], xb[NW][KT];
        uint32_t masked[NW];
        bool live[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            // an odd tail: the surplus slot recomputes window win0 (loaded before anything is stored) and stores nothing
            live[w] = win0 + w < n_win;
            long long t = live[w] ? win0 + w : win0;
            const int ww = (int)(t % nWw); t /= nWw;
            const int wh = (int)(t % nWh);
            const int b = (int)(t / nWh);
            // the two window tokens whose rows this lane holds (g and g + 8), in the shifted and the original grid
            long long tokA, tokB;
            int labA = 0, labB = 0;
            auto locate = [&](int tok, long long &token, int &label) {
                const int hs = wh * WIN + (tok >> 2), ws = ww * WIN + (tok & 3);
                int h = hs + shift, w2 = ws + shift;
                if (h >= H) h -= H;
                if (w2 >= W) w2 -= W;
                token = ((long long)b * H + h) * W + w2;
                if (shift > 0) label = 3 * (hs < H - WIN ? 0 : (hs < H - shift ? 1 : 2)) + (ws < W - WIN ? 0 : (ws < W - shift ? 1 : 2));
            };
            locate(g, tokA, labA);
            locate(g + 8, tokB, labB);
            // SW-MSA mask of this lane's score elements, one bit each: query and key in different regions (stf.py:316-334);
            // a property of the window, not of the head
            masked[w] = 0;
            if (ATTN && shift > 0) {
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int j = (c >> 1) * 8 + 2 * tq + (c & 1);
                        const int hj = wh * WIN + (j >> 2), wj = ww * WIN + (j & 3);
                        const int lj = 3 * (hj < H - WIN ? 0 : (hj < H - shift ? 1 : 2)) + (wj < W - WIN ? 0 : (wj < W - shift ? 1 : 2));
                        if (lj != (r ? labB : labA)) masked[w] |= 1u << (r * 4 + c);
                    }
            }
            pa[w] = reinterpret_cast<float4 *>(p.x + tokA * C) + tq;
            pb[w] = reinterpret_cast<float4 *>(p.x + tokB * C) + tq;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) { xa[w][kt] = pa[w][4 * kt]; xb[w][kt] = pb[w][4 * kt]; }
        }

        if (ATTN) {
            uint32_t af[NW][KT][4];
            float po[NW][NT][4]; // proj accumulators
#pragma unroll
            for (int w = 0; w < NW; ++w) layernorm_frag<C>(xa[w], xb[w], s_f + L::o_ln1, tq, af[w]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 bb = *reinterpret_cast<const float2 *>(s_f + L::o_bproj + feat(nt, tq, 0));
#pragma unroll
                for (int w = 0; w < NW; ++w) { po[w][nt][0] = bb.x; po[w][nt][1] = bb.y; po[w][nt][2] = bb.x; po[w][nt][3] = bb.y; }
            }
#pragma unroll 1
            for (int head = 0; head < heads; ++head) {
                float q[NW][2][4], k[NW][2][4], v[NW][2][4];
                gemm_pair<KT, NW>(q, af, s_qkv, L::S_C, head, s_f + L::o_bqkv, g, tq);
                gemm_pair<KT, NW>(k, af, s_qkv, L::S_C, heads + head, s_f + L::o_bqkv, g, tq);
                gemm_pair<KT, NW>(v, af, s_qkv, L::S_C, 2 * heads + head, s_f + L::o_bqkv, g, tq);
                float relb[2][4]; // B[rel] of this head: the same for every window
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) relb[r][c] = s_f[rel[r][c] + head];
                uint32_t oa[NW][4];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    // S = q k^T: q's accumulators are the A fragment, k's the B fragments (keys 0..7 from rows g, keys 8..15 from rows g+8)
                    uint32_t qa[4];
                    qa[0] = pack_bf16(q[w][0][0], q[w][0][1]); qa[1] = pack_bf16(q[w][0][2], q[w][0][3]);
                    qa[2] = pack_bf16(q[w][1][0], q[w][1][1]); qa[3] = pack_bf16(q[w][1][2], q[w][1][3]);
                    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_bf16_16816(s0, qa, pack_bf16(k[w][0][0], k[w][0][1]), pack_bf16(k[w][1][0], k[w][1][1]));
                    mma_bf16_16816(s1, qa, pack_bf16(k[w][0][2], k[w][0][3]), pack_bf16(k[w][1][2], k[w][1][3]));
                    float sc[2][4]; // [row g / g+8][key slot: 2tq, 2tq+1, 8+2tq, 9+2tq]
                    sc[0][0] = s0[0]; sc[0][1] = s0[1]; sc[0][2] = s1[0]; sc[0][3] = s1[1];
                    sc[1][0] = s0[2]; sc[1][1] = s0[3]; sc[1][2] = s1[2]; sc[1][3] = s1[3];
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        float mx = -1e30f;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            // q k^T * head_dim ** -0.5 (stf.py:63,100) + B[rel] (+ mask: -100, stf.py:334)
                            float a = fmaf(sc[r][c], 0.25f, relb[r][c]);
                            if (masked[w] & (1u << (r * 4 + c))) a += -100.0f;
                            sc[r][c] = a;
                            mx = fmaxf(mx, a);
                        }
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                        float den = 0.f;
#pragma unroll
                        for (int c = 0; c < 4; ++c) { sc[r][c] = __expf(sc[r][c] - mx); den += sc[r][c]; }
                        den += __shfl_xor_sync(0xffffffffu, den, 1);
                        den += __shfl_xor_sync(0xffffffffu, den, 2);
                        const float inv = 1.0f / den;
#pragma unroll
                        for (int c = 0; c < 4; ++c) sc[r][c] *= inv;
                    }
                    uint32_t pf[4]; // P as the A fragment of O = P V
                    pf[0] = pack_bf16(sc[0][0], sc[0][1]); pf[1] = pack_bf16(sc[1][0], sc[1][1]);
                    pf[2] = pack_bf16(sc[0][2], sc[0][3]); pf[3] = pack_bf16(sc[1][2], sc[1][3]);
                    // V's accumulators hold V[token g / g+8][d]; the B fragment wants V[token 2tq..][d = g]: transposed 8x8 blocks
                    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
                    mma_bf16_16816(o0, pf, movmatrix_trans(pack_bf16(v[w][0][0], v[w][0][1])), movmatrix_trans(pack_bf16(v[w][0][2], v[w][0][3])));
                    mma_bf16_16816(o1, pf, movmatrix_trans(pack_bf16(v[w][1][0], v[w][1][1])), movmatrix_trans(pack_bf16(v[w][1][2], v[w][1][3])));
                    // the head's output is k-group `head` of proj's input
                    oa[w][0] = pack_bf16(o0[0], o0[1]); oa[w][1] = pack_bf16(o0[2], o0[3]);
                    oa[w][2] = pack_bf16(o1[0], o1[1]); oa[w][3] = pack_bf16(o1[2], o1[3]);
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint2 bf = lds64(s_proj + (size_t)wrow(nt, g) * L::S_C + head * 16 + 4 * tq);
#pragma unroll
                    for (int w = 0; w < NW; ++w) mma_bf16_16816(po[w][nt], oa[w], bf.x, bf.y);
                }
            }
            // residual: accumulator column (2tq + e) of n-tile 2kt + h2 is channel kt*16 + 4tq + 2*h2 + e = the x registers
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    xa[w][kt].x += po[w][2 * kt][0]; xa[w][kt].y += po[w][2 * kt][1]; xa[w][kt].z += po[w][2 * kt + 1][0]; xa[w][kt].w += po[w][2 * kt + 1][1];
                    xb[w][kt].x += po[w][2 * kt][2]; xb[w][kt].y += po[w][2 * kt][3]; xb[w][kt].z += po[w][2 * kt + 1][2]; xb[w][kt].w += po[w][2 * kt + 1][3];
                }
        }
        if (MLP) {
            uint32_t af[NW][KT][4];
            float mo[NW][NT][4]; // fc2 accumulators
#pragma unroll
            for (int w = 0; w < NW; ++w) layernorm_frag<C>(xa[w], xb[w], s_f + L::o_ln2, tq, af[w]);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 bb = *reinterpret_cast<const float2 *>(s_f + L::o_b2 + feat(nt, tq, 0));
#pragma unroll
                for (int w = 0; w < NW; ++w) { mo[w][nt][0] = bb.x; mo[w][nt][1] = bb.y; mo[w][nt][2] = bb.x; mo[w][nt][3] = bb.y; }
            }
#pragma unroll(NW == 1 ? 2 : 1)
            for (int j = 0; j < 4 * C / 16; ++j) { // 16 hidden features at a time
                float h[NW][2][4];
                gemm_pair<KT, NW>(h, af, s_fc1, L::S_C, j, s_f + L::o_b1, g, tq);
                uint32_t ha[NW][4];
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    ha[w][0] = pack_bf16(gelu_tanh(h[w][0][0]), gelu_tanh(h[w][0][1])); ha[w][1] = pack_bf16(gelu_tanh(h[w][0][2]), gelu_tanh(h[w][0][3]));
                    ha[w][2] = pack_bf16(gelu_tanh(h[w][1][0]), gelu_tanh(h[w][1][1])); ha[w][3] = pack_bf16(gelu_tanh(h[w][1][2]), gelu_tanh(h[w][1][3]));
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const uint2 bf = lds64(s_fc2 + (size_t)wrow(nt, g) * L::S_H + j * 16 + 4 * tq);
#pragma unroll
                    for (int w = 0; w < NW; ++w) mma_bf16_16816(mo[w][nt], ha[w], bf.x, bf.y);
                }
            }
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    xa[w][kt].x += mo[w][2 * kt][0]; xa[w][kt].y += mo[w][2 * kt][1]; xa[w][kt].z += mo[w][2 * kt + 1][0]; xa[w][kt].w += mo[w][2 * kt + 1][1];
                    xb[w][kt].x += mo[w][2 * kt][2]; xb[w][kt].y += mo[w][2 * kt][3]; xb[w][kt].z += mo[w][2 * kt + 1][2]; xb[w][kt].w += mo[w][2 * kt + 1][3];
                }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (live[w]) {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) { pa[w][4 * kt] = xa[w][kt]; pb[w][4 * kt] = xb[w][kt]; }
            }
        if (p.dynamic) nxt = __shfl_sync(0xffffffffu, nxt, 0);
    }
    __syncthreads(); // every warp of this CTA has drawn its last index
    if (p.dynamic && threadIdx.x == 0 && atomicAdd(p.sched + 1, 1) == (int)gridDim.x - 1) { p.sched[0] = 0; p.sched[1] = 0; } // last CTA re-arms the pair
}

template <int C, bool ATTN, bool MLP, int WARPS, int NW, int MINB>
static int launch(const Params &p, cudaStream_t st)
{
    using L = Layout<C, ATTN, MLP>;
    static_assert(L::bytes <= 227 * 1024, "weights do not fit shared memory");
    auto kernel = swin_block_kernel<C, ATTN, MLP, WARPS, NW, MINB>;
    static PerDeviceSmem configured;
    if (L::bytes > 48 * 1024 && configured.needs(L::bytes)) {
        ICM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
        configured.done(L::bytes);
    }
    const long long n_win = (long long)p.B * (p.H / WIN) * (p.W / WIN);
    static int per_sm = 0; // resident CTAs per SM (a property of the kernel: asked once)
    if (per_sm == 0 && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS * 32, L::bytes) != cudaSuccess || per_sm < 1)) per_sm = 1;
    const long long want = (n_win + WARPS * NW - 1) / (WARPS * NW);
    const long long cap = (long long)persistent_grid_limit() * per_sm;
    Params q = p;
    q.sched = tile_counter(st);
    if (!q.sched) return ICM_ERR_CUDA;
    static const bool dyn = getenv("ICM_SWIN_STATIC") == nullptr;
    q.dynamic = dyn ? 1 : 0;
    kernel<<<(unsigned)(want < cap ? want : cap), WARPS * 32, L::bytes, st>>>(q);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

// Shape of the launch per instantiation (the fastest of the table above).  ICM_SWIN_SHAPE=<warps><windows><CTAs per SM> selects one of the
// other compiled shapes for A/B measurements (1012: C = 48 with 10 warps; 811: C = 96 with 8 warps, the round-2 shape).
template <int C, bool ATTN, bool MLP>
static int launch_shaped(const Params &p, cudaStream_t st)
{
    static const int shape = getenv("ICM_SWIN_SHAPE") ? atoi(getenv("ICM_SWIN_SHAPE")) : 0;
    if constexpr (C == 48) {
        if (shape == 1012) return launch<C, ATTN, MLP, 10, 1, 2>(p, st);
        return launch<C, ATTN, MLP, 8, 1, 2>(p, st);
    } else {
        if (shape == 811) return launch<C, ATTN, MLP, 8, 1, 1>(p, st);
        return launch<C, ATTN, MLP, 12, 1, 1>(p, st);
    }
}

}  // namespace sf
}  // namespace icm

using namespace icm;

extern "C" int icm_swin_block(float *d_x, int B, int H, int W, int C, int heads, int window, int shift, int parts,
                              const void *d_w_qkv, const float *d_b_qkv, const void *d_w_proj, const float *d_b_proj,
                              const float *d_rel_table, const float *d_ln1_g, const float *d_ln1_b,
                              const void *d_w_fc1, const float *d_b_fc1, const void *d_w_fc2, const float *d_b_fc2,
                              const float *d_ln2_g, const float *d_ln2_b, void *stream)
{
    ICM_CHECK_ARG(d_x && B > 0 && H > 0 && W > 0, "icm_swin_block: bad arguments");
    ICM_CHECK_ARG(parts >= 1 && parts <= 3, "icm_swin_block: parts must be 1 (attention half), 2 (MLP half) or 3 (both)");
    if (window != sf::WIN || C != heads * sf::HD || (C != 48 && C != 96)) {
        set_error("icm_swin_block: built for window 4, head_dim 16, C in {48, 96} (got window %d, C %d, heads %d)", window, C, heads);
        return ICM_ERR_UNSUPPORTED;
    }
    if (H % sf::WIN || W % sf::WIN) { set_error("icm_swin_block: H=%d W=%d must be multiples of the window", H, W); return ICM_ERR_UNSUPPORTED; }
    ICM_CHECK_ARG(shift >= 0 && shift < sf::WIN, "icm_swin_block: bad shift");
    const bool attn = parts & 1, mlp = parts & 2;
    ICM_CHECK_ARG(!attn || (d_w_qkv && d_w_proj && d_rel_table && d_ln1_g && d_ln1_b), "icm_swin_block: attention parameters missing");
    ICM_CHECK_ARG(!mlp || (d_w_fc1 && d_w_fc2 && d_ln2_g && d_ln2_b), "icm_swin_block: MLP parameters missing");
    sf::Params p{};
    p.x = d_x; p.B = B; p.H = H; p.W = W; p.shift = shift;
    p.w_qkv = (const __nv_bfloat16 *)d_w_qkv; p.w_proj = (const __nv_bfloat16 *)d_w_proj;
    p.w_fc1 = (const __nv_bfloat16 *)d_w_fc1; p.w_fc2 = (const __nv_bfloat16 *)d_w_fc2;
    p.ld_c = (C + 63) / 64 * 64; p.ld_h = (4 * C + 63) / 64 * 64;
    p.ln1_g = d_ln1_g; p.ln1_b = d_ln1_b; p.b_qkv = d_b_qkv; p.b_proj = d_b_proj; p.rel_table = d_rel_table;
    p.ln2_g = d_ln2_g; p.ln2_b = d_ln2_b; p.b_fc1 = d_b_fc1; p.b_fc2 = d_b_fc2;
    cudaStream_t st = as_stream(stream);
    if (C == 48) {
        if (attn && mlp) return sf::launch_shaped<48, true, true>(p, st);
        return attn ? sf::launch_shaped<48, true, false>(p, st) : sf::launch_shaped<48, false, true>(p, st);
    }
    // C = 96: the four weight matrices together (221 KB + padding) exceed one CTA's shared memory: two passes
    if (attn) { if (int rc = sf::launch_shaped<96, true, false>(p, st)) return rc; }
    if (mlp) { if (int rc = sf::launch_shaped<96, false, true>(p, st)) return rc; }
    return ICM_OK;
}
