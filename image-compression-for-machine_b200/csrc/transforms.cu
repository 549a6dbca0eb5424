// Memory-bound pieces of the STF transforms (rows T1, T2, T4, T6, T10 of SURVEY.md §8a): LayerNorm (with the
// PatchMerging gather folded in), shifted-window attention, PatchEmbed and the 48->3 output convolution.
// The GEMM-shaped work (qkv / proj / MLP / merge / split linears and all 3x3 / 5x5 convolutions) is in
// conv.cu on the tensor cores; these kernels keep its operands in channels-last bf16 so that no
// permute/contiguous/roll/window_partition copy of the reference (stf.py:42-53,97,167-191) exists at all.
#include "common.cuh"
#include "mma_sync.cuh"

namespace icm {

// ------------------------------------------------------------------------------------------------
// LayerNorm over C (eps 1e-5).  LPR lanes share one row, every lane holds NQ float4 (LPR * NQ * 4 >= C), so a warp
// normalises 32 / LPR rows at once with 16-byte loads: at C = 48 that is 8 rows and 48 bytes in flight per lane
// (the one-row-per-warp version had 8 and ran at a quarter of the HBM roofline).
// gather: row (b,h2,w2) is the concatenation of the four tokens (2h2+dh, 2w2+dw) in the order (0,0),(1,0),(0,1),(1,1)
// (stf.py:225-229), zero beyond H/W; C/4 is a multiple of 4, so a float4 never straddles two source tokens.
template <typename OutT, int LPR, int NQ>
__global__ void __launch_bounds__(256) layernorm_kernel(const float *__restrict__ in, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, OutT *__restrict__ out,
                                                        long long rows, int C, int gather, int H, int W)
{
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool row_ok = row < rows;
    const int Cs = gather ? C / 4 : C; // channels per source token
    int H2 = 0, W2 = 0, b = 0, h2 = 0, w2 = 0;
    if (gather) {
        H2 = (H + 1) / 2; W2 = (W + 1) / 2;
        long long t = row_ok ? row : 0;
        w2 = (int)(t % W2); t /= W2;
        h2 = (int)(t % H2); b = (int)(t / H2);
    }
    float4 v[NQ];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int c = 4 * (sub + LPR * i);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && c < C) {
            if (gather) {
                const int k = c / Cs, cc = c - k * Cs;
                const int hh = 2 * h2 + (k & 1), ww = 2 * w2 + (k >> 1);
                if (hh < H && ww < W) x = __ldg(reinterpret_cast<const float4 *>(in + (((long long)b * H + hh) * W + ww) * Cs + cc));
            } else {
                x = __ldg(reinterpret_cast<const float4 *>(in + row * C + c));
            }
        }
        v[i] = x;
        sum += (x.x + x.y) + (x.z + x.w);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        if (4 * (sub + LPR * i) < C) {
            const float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
            sq += (a * a + bb * bb) + (cc * cc + d * d);
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + 1e-5f);
    if (!row_ok) return;
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
        const int c = 4 * (sub + LPR * i);
        if (c < C) {
            const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c)), be = __ldg(reinterpret_cast<const float4 *>(beta + c));
            float4 y;
            y.x = (v[i].x - mean) * rstd * g.x + be.x;
            y.y = (v[i].y - mean) * rstd * g.y + be.y;
            y.z = (v[i].z - mean) * rstd * g.z + be.z;
            y.w = (v[i].w - mean) * rstd * g.w + be.w;
            if constexpr (sizeof(OutT) == 2) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(y.x, y.y), hi = __floats2bfloat162_rn(y.z, y.w);
                uint2 u;
                u.x = *reinterpret_cast<const uint32_t *>(&lo);
                u.y = *reinterpret_cast<const uint32_t *>(&hi);
                *reinterpret_cast<uint2 *>(out + row * C + c) = u;
            } else {
                *reinterpret_cast<float4 *>(out + row * C + c) = y;
            }
        }
    }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float *__restrict__ in, long long rows, int C, long long in_pitch,
                                                        __nv_bfloat16 *__restrict__ out, long long out_pitch)
{
    const long long total = rows * (C / 4);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (C / 4);
        const int c = (int)(i - r * (C / 4)) * 4;
        const float4 f = *reinterpret_cast<const float4 *>(in + r * in_pitch + c);
        const __nv_bfloat162 a = __floats2bfloat162_rn(f.x, f.y), b2 = __floats2bfloat162_rn(f.z, f.w);
        uint2 u;
        u.x = *reinterpret_cast<const uint32_t *>(&a);
        u.y = *reinterpret_cast<const uint32_t *>(&b2);
        *reinterpret_cast<uint2 *>(out + r * out_pitch + c) = u;
    }
}

// ------------------------------------------------------------------------------------------------
// Window attention, window 4x4 (16 tokens), head_dim 16 (every STF stage: 48/3 = 96/6 = 192/12 = 384/24).
// One warp = one (window, head): S = Q K^T and O = P V are each two mma.sync.m16n8k16 (bf16 in, fp32 accumulate),
// with the operand fragments loaded straight from the channels-last qkv rows (the S accumulator fragment IS the
// A fragment of the second product; V is transposed in registers with movmatrix).  The scalar version spent
// ~1100 instructions per warp on 512 shared-memory loads + 512 FMAs and was issue-bound at a quarter of the
// HBM roofline; this one is ~100 instructions and streams qkv once.  The cyclic shift, the window
// partition/reverse and the SW-MSA region mask are index arithmetic (stf.py:42-53,166-191,316-334).
constexpr int WIN = 4, NTOK = 16, HD = 16;
constexpr int ATT_WARPS = 8, ATT_JOBS_PER_WARP = 4; // per CTA: 32 (window, head) jobs share one copy of the bias table

__global__ void __launch_bounds__(ATT_WARPS * 32) window_attention_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ out,
                                                               const float *__restrict__ bias_table, int B, int H, int W, int C,
                                                               int heads, int shift)
{
    extern __shared__ float s_bias[]; // [49][heads]
    for (int i = threadIdx.x; i < 49 * heads; i += blockDim.x) s_bias[i] = bias_table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int nWw = W / WIN, nWh = H / WIN;
    const long long jobs = (long long)B * nWh * nWw * heads;
    const int g = lane >> 2, tq = lane & 3;
    for (int it = 0; it < ATT_JOBS_PER_WARP; ++it) {
    const long long job = ((long long)blockIdx.x * ATT_WARPS + warp) * ATT_JOBS_PER_WARP + it;
    if (job >= jobs) break; // warp-uniform
    const int head = (int)(job % heads);
    long long t = job / heads;
    const int ww = (int)(t % nWw); t /= nWw;
    const int wh = (int)(t % nWh);
    const int b = (int)(t / nWh);

    // the two window tokens whose rows this lane touches (g and g + 8), in the shifted and the original grid
    long long tokA, tokB;
    int labA = 0, labB = 0;
    auto locate = [&](int tok, long long &token, int &label) {
        const int hs = wh * WIN + (tok >> 2), ws = ww * WIN + (tok & 3);
        int h = hs + shift, w = ws + shift;
        if (h >= H) h -= H;
        if (w >= W) w -= W;
        token = ((long long)b * H + h) * W + w;
        if (shift > 0) label = 3 * (hs < H - WIN ? 0 : (hs < H - shift ? 1 : 2)) + (ws < W - WIN ? 0 : (ws < W - shift ? 1 : 2));
    };
    locate(g, tokA, labA);
    locate(g + 8, tokB, labB);
    // The contraction index d of S = Q K^T and the output column d of O = P V may be permuted freely as long as Q, K, V
    // and O agree: fragment slots (2tq, 2tq+1, 8+2tq, 9+2tq) hold head channels 4tq .. 4tq+3, so every operand is ONE
    // 8-byte load per row and lane (4 lanes = one full 32-byte sector) and the output one 8-byte store.
    const uint2 *rowA = reinterpret_cast<const uint2 *>(qkv + tokA * 3 * C + head * HD) + tq;
    const uint2 *rowB = reinterpret_cast<const uint2 *>(qkv + tokB * 3 * C + head * HD) + tq;
    const int cw = C / 4; // 8-byte words between q, k and v of a token
    const uint2 qA = __ldg(rowA), qB = __ldg(rowB), kA = __ldg(rowA + cw), kB = __ldg(rowB + cw), vA = __ldg(rowA + 2 * cw), vB = __ldg(rowB + 2 * cw);
    uint32_t qa[4];
    qa[0] = qA.x; qa[1] = qB.x; qa[2] = qA.y; qa[3] = qB.y;
    const uint32_t kA0 = kA.x, kA1 = kA.y, kB0 = kB.x, kB1 = kB.y, vA0 = vA.x, vA1 = vA.y, vB0 = vB.x, vB1 = vB.y;

    // S[i][j] for i in {g, g+8}, j in {2tq, 2tq+1, 8+2tq, 9+2tq}
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16_16816(s0, qa, kA0, kA1); // keys 0..7
    mma_bf16_16816(s1, qa, kB0, kB1); // keys 8..15
    float sc[2][4]; // [row g / g+8][j slot]
    sc[0][0] = s0[0]; sc[0][1] = s0[1]; sc[0][2] = s1[0]; sc[0][3] = s1[1];
    sc[1][0] = s0[2]; sc[1][1] = s0[3]; sc[1][2] = s1[2]; sc[1][3] = s1[3];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = g + 8 * r, ih = i >> 2, iw = i & 3;
        const int li = r ? labB : labA;
        float mx = -1e30f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = (c >> 1) * 8 + 2 * tq + (c & 1), jh = j >> 2, jw = j & 3;
            float a = sc[r][c] * 0.25f; // head_dim ** -0.5 (a power of two: same as scaling q first)
            a += s_bias[((ih - jh + WIN - 1) * (2 * WIN - 1) + (iw - jw + WIN - 1)) * heads + head];
            if (shift > 0) {
                const int hj = wh * WIN + jh, wj = ww * WIN + jw;
                const int lj = 3 * (hj < H - WIN ? 0 : (hj < H - shift ? 1 : 2)) + (wj < W - WIN ? 0 : (wj < W - shift ? 1 : 2));
                if (lj != li) a += -100.0f;
            }
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
    // P as the A fragment of O = P V:  a0 = P[g][2tq..], a1 = P[g+8][2tq..], a2 = P[g][8+2tq..], a3 = P[g+8][8+2tq..]
    uint32_t pa[4];
    pa[0] = pack_bf16(sc[0][0], sc[0][1]); pa[1] = pack_bf16(sc[1][0], sc[1][1]);
    pa[2] = pack_bf16(sc[0][2], sc[0][3]); pa[3] = pack_bf16(sc[1][2], sc[1][3]);
    // B fragment for output columns d = dt*8 + g: (V[2tq][d], V[2tq+1][d]) and (V[8+2tq][d], V[9+2tq][d]) = transposed 8x8 blocks
    const uint32_t b00 = movmatrix_trans(vA0), b01 = movmatrix_trans(vB0); // d tile 0: keys 0..7, keys 8..15
    const uint32_t b10 = movmatrix_trans(vA1), b11 = movmatrix_trans(vB1); // d tile 1
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16_16816(o0, pa, b00, b01);
    mma_bf16_16816(o1, pa, b10, b11);
    // fragment columns (2tq, 2tq+1) of d tile 0 / 1 are head channels (4tq, 4tq+1) / (4tq+2, 4tq+3)
    *(reinterpret_cast<uint2 *>(out + tokA * C + head * HD) + tq) = make_uint2(pack_bf16(o0[0], o0[1]), pack_bf16(o1[0], o1[1]));
    *(reinterpret_cast<uint2 *>(out + tokB * C + head * HD) + tq) = make_uint2(pack_bf16(o0[2], o0[3]), pack_bf16(o1[2], o1[3]));
    }
}

// ------------------------------------------------------------------------------------------------
// PatchEmbed: Conv2d(3 -> C, k=2, s=2) + LayerNorm(C); C <= 64.  One thread per output token.
__global__ void __launch_bounds__(128) patch_embed_kernel(const float *__restrict__ img, const float *__restrict__ w,
                                                          const float *__restrict__ bias, const float *__restrict__ gamma,
                                                          const float *__restrict__ beta, float *__restrict__ tokens,
                                                          int B, int H, int W, int C)
{
    __shared__ __align__(16) float s_w[64 * 12];
    __shared__ float s_b[64], s_g[64], s_be[64];
    extern __shared__ float s_o[]; // [blockDim.x][C + 1]
    for (int i = threadIdx.x; i < C * 12; i += blockDim.x) s_w[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_b[i] = bias[i]; s_g[i] = gamma[i]; s_be[i] = beta[i]; }
    __syncthreads();
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    const long long total = (long long)B * H2 * W2;
    const long long idx0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = idx0 < total;
    const long long idx = live ? idx0 : total - 1;
    const int w2 = (int)(idx % W2);
    const int h2 = (int)((idx / W2) % H2);
    const int b = (int)(idx / ((long long)W2 * H2));
    float x[12]; // [ci][kh][kw], zero padded on the bottom/right (stf.py:368-372)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
            for (int kw = 0; kw < 2; ++kw) {
                const int hh = 2 * h2 + kh, ww = 2 * w2 + kw;
                x[ci * 4 + kh * 2 + kw] = (hh < H && ww < W) ? img[(((long long)b * 3 + ci) * H + hh) * W + ww] : 0.f;
            }
    float y[64];
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 64; ++c) {
        float a = 0.f;
        if (c < C) {
            a = s_b[c];
            // the 12 weights of a channel as three broadcast LDS.128 (one LDS.32 per multiply-add made the kernel LSU-bound:
            // 0.91 ms for 64 images at 1.6 TB/s); same accumulation order as before
            const float4 *w4 = reinterpret_cast<const float4 *>(s_w + c * 12);
            const float4 wa = w4[0], wb = w4[1], wc = w4[2];
            a += wa.x * x[0]; a += wa.y * x[1]; a += wa.z * x[2]; a += wa.w * x[3];
            a += wb.x * x[4]; a += wb.y * x[5]; a += wb.z * x[6]; a += wb.w * x[7];
            a += wc.x * x[8]; a += wc.y * x[9]; a += wc.z * x[10]; a += wc.w * x[11];
            sum += a;
        }
        y[c] = a;
    }
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < 64; ++c) if (c < C) { const float d = y[c] - mean; sq += d * d; }
    const float rstd = rsqrtf(sq / (float)C + 1e-5f);
    // each thread owns one token (C contiguous floats): go through shared memory so that the CTA's 128 x C block,
    // which is contiguous in the token tensor, is written with coalesced stores
    float *mine = s_o + threadIdx.x * (C + 1);
    if (live) {
#pragma unroll
        for (int c = 0; c < 64; ++c) if (c < C) mine[c] = (y[c] - mean) * rstd * s_g[c] + s_be[c];
    }
    __syncthreads();
    const long long first = (long long)blockIdx.x * blockDim.x;
    const int n_tok = (int)min((long long)blockDim.x, total - first);
    float *dst = tokens + first * C;
    // (row, column) walk without a division per element: i advances by blockDim.x = q * C + r
    const int q = blockDim.x / C, r = blockDim.x - q * C;
    int row = threadIdx.x / C, col = threadIdx.x - row * C;
    for (int i = threadIdx.x; i < n_tok * C; i += blockDim.x) {
        dst[i] = s_o[row * (C + 1) + col];
        row += q; col += r;
        if (col >= C) { col -= C; ++row; }
    }
}

// ------------------------------------------------------------------------------------------------
// Output convolution C -> 3, 3x3, pad 1, bf16 channels-last in, fp32 NCHW image out (stf.py:466,784).
// 16x16 pixel tile per CTA with an 18x18 halo tile in shared memory.
constexpr int FT = 16;

__global__ void __launch_bounds__(256) final_conv_kernel(const __nv_bfloat16 *__restrict__ in, const float *__restrict__ w,
                                                         const float *__restrict__ bias, float *__restrict__ img, int B, int H,
                                                         int W, int C, int clamp01)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    const int CP = C / 2 + 1; // 32-bit words per pixel (+1 pad: odd stride, conflict-free)
    uint32_t *s_in = reinterpret_cast<uint32_t *>(fsm);                       // [(FT+2)^2][CP]
    float *s_w = reinterpret_cast<float *>(s_in + (FT + 2) * (FT + 2) * CP);   // [3][9][C]
    const int b = blockIdx.z, h0 = blockIdx.y * FT, w0 = blockIdx.x * FT;
    for (int i = threadIdx.x; i < 3 * 9 * C; i += blockDim.x) {
        // torch layout [co][ci][kh][kw] -> [co][tap][ci]
        const int ci = i % C, tap = (i / C) % 9, co = i / (9 * C);
        s_w[i] = w[(co * C + ci) * 9 + tap];
    }
    const int words = C / 2;
    for (int i = threadIdx.x; i < (FT + 2) * (FT + 2) * words; i += blockDim.x) {
        const int cw = i % words, pix = i / words;
        const int hh = h0 + pix / (FT + 2) - 1, ww = w0 + pix % (FT + 2) - 1;
        uint32_t v = 0;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
            v = reinterpret_cast<const uint32_t *>(in + (((long long)b * H + hh) * W + ww) * C)[cw];
        s_in[pix * CP + cw] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x % FT, ty = threadIdx.x / FT;
    const int oh = h0 + ty, ow = w0 + tx;
    float acc0 = bias[0], acc1 = bias[1], acc2 = bias[2];
    for (int tap = 0; tap < 9; ++tap) {
        const int pix = (ty + tap / 3) * (FT + 2) + tx + tap % 3;
        const uint32_t *src = s_in + pix * CP;
        const float *w0p = s_w + (0 * 9 + tap) * C, *w1p = s_w + (1 * 9 + tap) * C, *w2p = s_w + (2 * 9 + tap) * C;
        for (int cw = 0; cw < words; ++cw) {
            const uint32_t u = src[cw];
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u));
            acc0 += f.x * w0p[2 * cw] + f.y * w0p[2 * cw + 1];
            acc1 += f.x * w1p[2 * cw] + f.y * w1p[2 * cw + 1];
            acc2 += f.x * w2p[2 * cw] + f.y * w2p[2 * cw + 1];
        }
    }
    if (oh < H && ow < W) {
        if (clamp01) { acc0 = fminf(fmaxf(acc0, 0.f), 1.f); acc1 = fminf(fmaxf(acc1, 0.f), 1.f); acc2 = fminf(fmaxf(acc2, 0.f), 1.f); }
        const long long plane = (long long)H * W;
        float *dst = img + (long long)b * 3 * plane + (long long)oh * W + ow;
        dst[0] = acc0; dst[plane] = acc1; dst[2 * plane] = acc2;
    }
}

// ------------------------------------------------------------------------------------------------
// Output convolution 48 -> 3 on the tensor cores: per warp 32 pixels of one image row as two m16n8k16 tiles
// (M = 16 pixels, N = 8 output channels of which 3 are real, K = 16 input channels), 9 taps x 3 k-steps each.
// The CTA (8 warps = 8 rows x 32 pixels) stages its 10 x 34 halo tile in shared memory with a 112-byte pixel
// pitch (conflict-free ldmatrix rows); the 27 weight fragments live in registers; results leave through a small
// shared-memory transpose as 128-byte rows of the three NCHW planes.  The CUDA-core version above (kept for
// other channel counts) was FMA/LDS-bound at 5.5 ms per 64 images; N = 16 on tcgen05 would be bound by the
// 9-fold re-read of the A tile from L2 instead.
constexpr int FC_C = 48, FC_TW = 32, FC_TH = 8, FC_VT = 4, FC_PITCH = 112; // bytes per halo pixel (96 + 16 pad)

__global__ void __launch_bounds__(256) final_conv48_mma_kernel(const __nv_bfloat16 *__restrict__ in, const float *__restrict__ w,
                                                               const float *__restrict__ bias, float *__restrict__ img, int B, int H,
                                                               int W, int clamp01)
{
    __shared__ __align__(16) unsigned char s_tile[(FC_TH + 2) * (FC_TW + 2) * FC_PITCH];
    __shared__ float s_out[FC_TH][3][FC_TW];
    __shared__ uint2 s_wf[27][32];
    const int b = blockIdx.z, w0 = blockIdx.x * FC_TW;
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    // weight fragments, built once per CTA in lane order: B[k = channel][n = co] = w[co][channel][tap]; lane (g, t) holds
    // k = 2t, 2t+1 (b0) and 2t+8, 2t+9 (b1) of n = g
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
        const int l = i & 31, frag = i >> 5, tap = frag / 3, ks = frag - tap * 3;
        const int fg = l >> 2, ft = l & 3;
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        if (fg < 3) {
            const int c0 = ks * 16 + 2 * ft;
            f[0] = w[(fg * FC_C + c0) * 9 + tap];     f[1] = w[(fg * FC_C + c0 + 1) * 9 + tap];
            f[2] = w[(fg * FC_C + c0 + 8) * 9 + tap]; f[3] = w[(fg * FC_C + c0 + 9) * 9 + tap];
        }
        s_wf[frag][l] = make_uint2(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]));
    }
    const int g = lane >> 2, t = lane & 3;
    for (int vt = 0; vt < FC_VT; ++vt) {
    const int h0 = (blockIdx.y * FC_VT + vt) * FC_TH;
    if (h0 >= H) break;
    if (vt) __syncthreads(); // the previous tile has been consumed
    // halo tile: 6 x 16-byte pieces per pixel
    for (int i = threadIdx.x; i < (FC_TH + 2) * (FC_TW + 2) * 6; i += blockDim.x) {
        const int piece = i % 6, pix = i / 6;
        const int hh = h0 + pix / (FC_TW + 2) - 1, ww = w0 + pix % (FC_TW + 2) - 1;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
            v = __ldg(reinterpret_cast<const uint4 *>(in + (((long long)b * H + hh) * W + ww) * FC_C) + piece);
        *reinterpret_cast<uint4 *>(s_tile + pix * FC_PITCH + piece * 16) = v;
    }
    __syncthreads();
    const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(s_tile);
    // ldmatrix.x4 row address of this lane: pixel row (lane % 16), channel half (lane / 16)
    const int lrow = lane & 15, lhalf = lane >> 4;
    float acc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[mt][j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3, dx = tap % 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const uint32_t row_addr = tile_addr + (uint32_t)(((warp + dy) * (FC_TW + 2) + mt * 16 + lrow + dx) * FC_PITCH + lhalf * 16);
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                uint32_t a[4];
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(row_addr + ks * 32));
                const uint2 wb = s_wf[tap * 3 + ks][lane];
                mma_bf16_16816(acc[mt], a, wb.x, wb.y);
            }
        }
    }
    // acc[mt][0..1]: pixel mt*16 + g, co 2t, 2t+1;  acc[mt][2..3]: pixel mt*16 + g + 8
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        if (t == 0) {
            s_out[warp][0][mt * 16 + g] = acc[mt][0]; s_out[warp][1][mt * 16 + g] = acc[mt][1];
            s_out[warp][0][mt * 16 + g + 8] = acc[mt][2]; s_out[warp][1][mt * 16 + g + 8] = acc[mt][3];
        } else if (t == 1) {
            s_out[warp][2][mt * 16 + g] = acc[mt][0];
            s_out[warp][2][mt * 16 + g + 8] = acc[mt][2];
        }
    }
    __syncwarp();
    const int oh = h0 + warp, ow = w0 + lane;
    if (oh < H && ow < W) {
        const long long plane = (long long)H * W;
        float *dst = img + (long long)b * 3 * plane + (long long)oh * W + ow;
#pragma unroll
        for (int co = 0; co < 3; ++co) {
            float v = s_out[warp][co][lane] + bias[co];
            if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
            dst[co * plane] = v;
        }
    }
    }
}

}  // namespace icm

using namespace icm;

extern "C" int icm_layernorm(const float *d_in, const float *d_gamma, const float *d_beta, void *d_out, int out_dtype,
                             int64_t rows, int C, int gather, int B, int H, int W, void *stream)
{
    ICM_CHECK_ARG(d_in && d_gamma && d_beta && d_out, "icm_layernorm: null argument");
    ICM_CHECK_ARG(C > 0 && C <= 768, "icm_layernorm: C=%d outside (0,768]", C);
    ICM_CHECK_ARG(!gather || (C % 4 == 0 && rows == (int64_t)B * ((H + 1) / 2) * ((W + 1) / 2)), "icm_layernorm: gather shape mismatch");
    ICM_CHECK_ARG(rows > 0, "icm_layernorm: no rows");
    ICM_CHECK_ARG(C % 4 == 0 && (!gather || C % 16 == 0), "icm_layernorm: C=%d must be a multiple of 4 (16 with gather)", C);
    cudaStream_t st = as_stream(stream);
#define ICM_LN_LAUNCH(LPR, NQ)                                                                                             \
    do {                                                                                                                   \
        const unsigned grid = (unsigned)((rows + 8 * (32 / LPR) - 1) / (8 * (32 / LPR)));                                  \
        if (out_dtype == ICM_OUT_BF16)                                                                                     \
            layernorm_kernel<__nv_bfloat16, LPR, NQ><<<grid, 256, 0, st>>>(d_in, d_gamma, d_beta, (__nv_bfloat16 *)d_out, rows, C, gather, H, W); \
        else                                                                                                               \
            layernorm_kernel<float, LPR, NQ><<<grid, 256, 0, st>>>(d_in, d_gamma, d_beta, (float *)d_out, rows, C, gather, H, W);  \
    } while (0)
    if (C <= 48) ICM_LN_LAUNCH(4, 3);
    else if (C <= 96) ICM_LN_LAUNCH(8, 3);
    else if (C <= 192) ICM_LN_LAUNCH(16, 3);
    else if (C <= 384) ICM_LN_LAUNCH(32, 3);
    else ICM_LN_LAUNCH(32, 6);
#undef ICM_LN_LAUNCH
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_cast_bf16(const float *d_in, int64_t rows, int C, int64_t in_pitch, void *d_out, int64_t out_pitch, void *stream)
{
    ICM_CHECK_ARG(d_in && d_out && rows > 0 && C > 0 && C % 4 == 0 && in_pitch % 4 == 0 && out_pitch % 4 == 0, "icm_cast_bf16: bad arguments");
    const long long total = rows * (C / 4);
    const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 16);
    cast_bf16_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_in, rows, C, in_pitch, (__nv_bfloat16 *)d_out, out_pitch);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_window_attention(const void *d_qkv, void *d_out, const float *d_bias_table, int B, int H, int W,
                                    int C, int heads, int window, int shift, void *stream)
{
    ICM_CHECK_ARG(d_qkv && d_out && d_bias_table, "icm_window_attention: null argument");
    if (window != WIN || C != heads * HD) { set_error("icm_window_attention: only window 4 / head_dim 16 is built (got window %d, head_dim %d)", window, heads ? C / heads : 0); return ICM_ERR_UNSUPPORTED; }
    if (H % WIN || W % WIN) { set_error("icm_window_attention: H=%d W=%d must be multiples of the window (pad the image to a multiple of 64 as the reference's eval does)", H, W); return ICM_ERR_UNSUPPORTED; }
    ICM_CHECK_ARG(shift >= 0 && shift < WIN, "icm_window_attention: bad shift");
    const long long jobs = (long long)B * (H / WIN) * (W / WIN) * heads;
    const int per_cta = ATT_WARPS * ATT_JOBS_PER_WARP;
    const unsigned grid = (unsigned)((jobs + per_cta - 1) / per_cta);
    window_attention_kernel<<<grid, ATT_WARPS * 32, 49 * heads * sizeof(float), as_stream(stream)>>>(
        (const __nv_bfloat16 *)d_qkv, (__nv_bfloat16 *)d_out, d_bias_table, B, H, W, C, heads, shift);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_patch_embed(const float *d_img, const float *d_w, const float *d_b, const float *d_gamma,
                               const float *d_beta, float *d_tokens, int B, int H, int W, int C, void *stream)
{
    ICM_CHECK_ARG(d_img && d_w && d_b && d_gamma && d_beta && d_tokens, "icm_patch_embed: null argument");
    ICM_CHECK_ARG(C > 0 && C <= 64 && B > 0 && H > 0 && W > 0, "icm_patch_embed: bad shape");
    const long long total = (long long)B * ((H + 1) / 2) * ((W + 1) / 2);
    patch_embed_kernel<<<(unsigned)((total + 127) / 128), 128, (size_t)128 * (C + 1) * sizeof(float), as_stream(stream)>>>(d_img, d_w, d_b, d_gamma, d_beta, d_tokens, B, H, W, C);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_final_conv(const void *d_in_bf16, const float *d_w, const float *d_b, float *d_img, int B, int H,
                              int W, int C, int clamp01, void *stream)
{
    ICM_CHECK_ARG(d_in_bf16 && d_w && d_b && d_img, "icm_final_conv: null argument");
    ICM_CHECK_ARG(C > 0 && C % 2 == 0 && C <= 96, "icm_final_conv: C=%d unsupported", C);
    if (C == FC_C && (((uintptr_t)d_in_bf16) & 15) == 0) { // the reference's end_conv[2]: tensor-core kernel
        dim3 g((W + FC_TW - 1) / FC_TW, (H + FC_TH * FC_VT - 1) / (FC_TH * FC_VT), B);
        final_conv48_mma_kernel<<<g, 256, 0, as_stream(stream)>>>((const __nv_bfloat16 *)d_in_bf16, d_w, d_b, d_img, B, H, W, clamp01);
        ICM_LAUNCH_CHECK();
        return ICM_OK;
    }
    const size_t smem = (size_t)(FT + 2) * (FT + 2) * (C / 2 + 1) * 4 + (size_t)27 * C * 4;
    dim3 grid((W + FT - 1) / FT, (H + FT - 1) / FT, B);
    static PerDeviceSmem configured;
    if (smem > 48 * 1024 && configured.needs(smem)) {
        ICM_CHECK_ARG(smem <= 227 * 1024, "icm_final_conv: C=%d needs %zu bytes of shared memory", C, smem);
        ICM_CUDA(cudaFuncSetAttribute(final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.done(smem);
    }
    final_conv_kernel<<<grid, 256, smem, as_stream(stream)>>>((const __nv_bfloat16 *)d_in_bf16, d_w, d_b, d_img, B, H, W, C, clamp01);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}
