// Pieces of the WACNN ("cnn" / "cnn2") transforms that are not plain convolutions (row T11 of SURVEY.md §8a):
//   - image <-> channels-last conversions around the 5x5 stride-2 convolutions / transposed convolutions
//     (compressai/models/cnn.py:31-52, compressai/models/utils.py:114-132),
//   - the elementwise parts of GDN (x^2 feeding a 1x1 GEMM whose epilogue does x * rsqrt(.), gdn.py:62-75)
//     and of the gated attention block  out = x + a(x) * sigmoid(b(x))  (layers.py:83-89),
//   - window attention with window 8 / head_dim 24 and window 4 / head_dim 40, shifted, with relative-position
//     bias and region mask and no padding (win_attention.py:90-207),
//   - packing of ConvTranspose2d(k5, s2, p2, op1) weights into four 3x3 phase filters so that the transposed
//     convolution runs on the same tcgen05 implicit-GEMM kernel with the PixelShuffle(2) store.
#include "common.cuh"
#include "mma_sync.cuh"

#include <stdlib.h>

namespace icm {

// NCHW fp32 image -> NHWC bf16 with `pitch` channels (channels >= C are zero)
__global__ void __launch_bounds__(256) image_to_nhwc_kernel(const float *__restrict__ img, __nv_bfloat16 *__restrict__ out, int B, int C,
                                                            long long P, int pitch)
{
    const long long total = (long long)B * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / P, p = i - b * P;
        __nv_bfloat16 *dst = out + i * pitch;
        for (int c = 0; c < pitch; ++c) dst[c] = __float2bfloat16_rn(c < C ? img[(b * C + c) * P + p] : 0.f);
    }
}

// NHWC (bf16, `pitch` channels) -> NCHW fp32 image with C channels, optional clamp to [0,1]
__global__ void __launch_bounds__(256) nhwc_to_image_kernel(const __nv_bfloat16 *__restrict__ in, float *__restrict__ img, int B, int C,
                                                            long long P, int pitch, int clamp01)
{
    const long long total = (long long)B * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / P, p = i - b * P;
        for (int c = 0; c < C; ++c) {
            float v = __bfloat162float(in[i * pitch + c]);
            if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
            img[(b * C + c) * P + p] = v;
        }
    }
}

// mode 0: out = x * x    mode 1: out = a * s + x    mode 2: the same, written as fp32
// (bf16 in, fp32 math); rows x C with pitches
__global__ void __launch_bounds__(256) eltwise_kernel(int mode, const __nv_bfloat16 *__restrict__ a, long long pa,
                                                      const __nv_bfloat16 *__restrict__ s, long long ps,
                                                      const __nv_bfloat16 *__restrict__ x, long long px,
                                                      void *__restrict__ out_v, long long po, long long rows, int C)
{
    const int c8 = C / 8;
    const long long total = rows * c8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c8;
        const int c = (int)(i - r * c8) * 8;
        const uint4 vx = *reinterpret_cast<const uint4 *>(x + r * px + c);
        const __nv_bfloat162 *hx = reinterpret_cast<const __nv_bfloat162 *>(&vx);
        uint4 vo;
        __nv_bfloat162 *ho = reinterpret_cast<__nv_bfloat162 *>(&vo);
        if (mode == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(hx[k]); ho[k] = __floats2bfloat162_rn(f.x * f.x, f.y * f.y); }
        } else {
            const uint4 va = *reinterpret_cast<const uint4 *>(a + r * pa + c);
            const uint4 vs = *reinterpret_cast<const uint4 *>(s + r * ps + c);
            const __nv_bfloat162 *ha = reinterpret_cast<const __nv_bfloat162 *>(&va);
            const __nv_bfloat162 *hs = reinterpret_cast<const __nv_bfloat162 *>(&vs);
            float f[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fa = __bfloat1622float2(ha[k]), fs = __bfloat1622float2(hs[k]), fx = __bfloat1622float2(hx[k]);
                f[2 * k] = fa.x * fs.x + fx.x; f[2 * k + 1] = fa.y * fs.y + fx.y;
                ho[k] = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
            }
            if (mode == 2) {
                float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out_v) + r * po + c);
                o[0] = make_float4(f[0], f[1], f[2], f[3]);
                o[1] = make_float4(f[4], f[5], f[6], f[7]);
                continue;
            }
        }
        *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(out_v) + r * po + c) = vo;
    }
}

// ------------------------------------------------------------------------------------------------
// Window attention, one warp per (window, head); K and V of the window/head in shared memory (fp32); every
// lane owns WIN*WIN/32 query tokens (or one, for 16-token windows) and runs an online softmax over the keys.
template <int WIN, int HD>
__global__ void __launch_bounds__(128) win_attention_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ out,
                                                            const float *__restrict__ bias_table, int B, int H, int W, int C,
                                                            int heads, int shift)
{
    constexpr int N = WIN * WIN;
    constexpr int NB = (2 * WIN - 1) * (2 * WIN - 1);
    extern __shared__ float wsm[];
    float *s_bias = wsm;                       // [NB][heads]
    float *s_kv = wsm + NB * heads;            // [4 warps][2][N][HD + 1]
    for (int i = threadIdx.x; i < NB * heads; i += blockDim.x) s_bias[i] = bias_table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    float *s_k = s_kv + (size_t)warp * 2 * N * (HD + 1), *s_v = s_k + N * (HD + 1);
    const int nWw = W / WIN, nWh = H / WIN;
    const long long jobs = (long long)B * nWh * nWw * heads;
    const long long job = (long long)blockIdx.x * 4 + warp;
    if (job >= jobs) return;
    const int head = (int)(job % heads);
    long long t = job / heads;
    const int ww = (int)(t % nWw); t /= nWw;
    const int wh = (int)(t % nWh);
    const int b = (int)(t / nWh);
    const float scale = rsqrtf((float)HD);
    auto token_of = [&](int tok) -> long long { // window token -> position in the (unshifted) feature map
        int h = wh * WIN + tok / WIN + shift, w = ww * WIN + tok % WIN + shift;
        if (h >= H) h -= H;
        if (w >= W) w -= W;
        return ((long long)b * H + h) * W + w;
    };
    auto label_of = [&](int tok) -> int {
        const int hs = wh * WIN + tok / WIN, ws = ww * WIN + tok % WIN;
        return 3 * (hs < H - WIN ? 0 : (hs < H - shift ? 1 : 2)) + (ws < W - WIN ? 0 : (ws < W - shift ? 1 : 2));
    };
    // stage K and V
    for (int tok = lane; tok < N; tok += 32) {
        const __nv_bfloat16 *row = qkv + token_of(tok) * 3 * C + head * HD;
#pragma unroll
        for (int d8 = 0; d8 < HD / 8; ++d8) {
            const uint4 kk = *reinterpret_cast<const uint4 *>(row + C + d8 * 8);
            const uint4 vv = *reinterpret_cast<const uint4 *>(row + 2 * C + d8 * 8);
            const __nv_bfloat162 *hk = reinterpret_cast<const __nv_bfloat162 *>(&kk), *hv = reinterpret_cast<const __nv_bfloat162 *>(&vv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 fk = __bfloat1622float2(hk[k]), fv = __bfloat1622float2(hv[k]);
                s_k[tok * (HD + 1) + d8 * 8 + 2 * k] = fk.x; s_k[tok * (HD + 1) + d8 * 8 + 2 * k + 1] = fk.y;
                s_v[tok * (HD + 1) + d8 * 8 + 2 * k] = fv.x; s_v[tok * (HD + 1) + d8 * 8 + 2 * k + 1] = fv.y;
            }
        }
    }
    __syncwarp();
    for (int tok = lane; tok < N; tok += 32) {
        const long long token = token_of(tok);
        const __nv_bfloat16 *row = qkv + token * 3 * C + head * HD;
        float q[HD], acc[HD];
#pragma unroll
        for (int d8 = 0; d8 < HD / 8; ++d8) {
            const uint4 qq = *reinterpret_cast<const uint4 *>(row + d8 * 8);
            const __nv_bfloat162 *hq = reinterpret_cast<const __nv_bfloat162 *>(&qq);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(hq[k]); q[d8 * 8 + 2 * k] = f.x * scale; q[d8 * 8 + 2 * k + 1] = f.y * scale; }
        }
#pragma unroll
        for (int d = 0; d < HD; ++d) acc[d] = 0.f;
        const int ih = tok / WIN, iw = tok % WIN;
        const int label = shift > 0 ? label_of(tok) : 0;
        float m = -1e30f, l = 0.f;
        for (int j = 0; j < N; ++j) {
            float sc = 0.f;
#pragma unroll
            for (int d = 0; d < HD; ++d) sc += q[d] * s_k[j * (HD + 1) + d];
            const int jh = j / WIN, jw = j % WIN;
            sc += s_bias[((ih - jh + WIN - 1) * (2 * WIN - 1) + (iw - jw + WIN - 1)) * heads + head];
            if (shift > 0 && label_of(j) != label) sc += -100.0f;
            const float mn = fmaxf(m, sc);
            const float corr = __expf(m - mn), pj = __expf(sc - mn);
            l = l * corr + pj;
#pragma unroll
            for (int d = 0; d < HD; ++d) acc[d] = acc[d] * corr + pj * s_v[j * (HD + 1) + d];
            m = mn;
        }
        const float inv = 1.0f / l;
        __nv_bfloat16 *dst = out + token * C + head * HD;
#pragma unroll
        for (int d8 = 0; d8 < HD / 8; ++d8) {
            uint4 o;
            __nv_bfloat162 *ho = reinterpret_cast<__nv_bfloat162 *>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) ho[k] = __floats2bfloat162_rn(acc[d8 * 8 + 2 * k] * inv, acc[d8 * 8 + 2 * k + 1] * inv);
            *reinterpret_cast<uint4 *>(dst + d8 * 8) = o;
        }
    }
}

// ConvTranspose2d(Cin -> Cout, k5, s2, p2, output_padding 1), torch weight [Cin][Cout][5][5]:
//   out[2m+py, 2n+px] = sum over taps (dy,dx) in {-1,0,1}^2 of in[m+dy, n+dx] * w[ci][co][py + 2(1-dy)][px + 2(1-dx)]
// (kernel index 5 does not exist: those phase/tap pairs are zero).  Packed as a 3x3 conv with 4*Cq output
// channels, phase-major (quad = py*2+px, channel quad*Cq + co, Cq = Cout rounded up to 16), for the
// PixelShuffle(2) store of conv_igemm_kernel.
__global__ void pack_deconv_weight_kernel(const float *__restrict__ w, int Cin, int Cout, int Cin_pad, int Cq,
                                          __nv_bfloat16 *__restrict__ out)
{
    const long long total = (long long)4 * Cq * 9 * Cin_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin_pad);
        long long t = i / Cin_pad;
        const int tap = (int)(t % 9);
        const int n = (int)(t / 9);
        const int quad = n / Cq, co = n - quad * Cq;
        const int py = quad >> 1, px = quad & 1;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int ky = py + 2 * (1 - dy), kx = px + 2 * (1 - dx);
        float v = 0.f;
        if (ci < Cin && co < Cout && ky < 5 && kx < 5) v = w[(((long long)ci * Cout + co) * 5 + ky) * 5 + kx];
        out[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace icm

using namespace icm;

extern "C" int icm_image_to_nhwc(const float *d_img, void *d_out_bf16, int B, int C, int H, int W, int pitch, void *stream)
{
    ICM_CHECK_ARG(d_img && d_out_bf16 && B > 0 && C > 0 && C <= pitch && H > 0 && W > 0, "icm_image_to_nhwc: bad arguments");
    const long long total = (long long)B * H * W;
    const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 16);
    image_to_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_img, (__nv_bfloat16 *)d_out_bf16, B, C, (long long)H * W, pitch);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_nhwc_to_image(const void *d_in_bf16, float *d_img, int B, int C, int H, int W, int pitch, int clamp01, void *stream)
{
    ICM_CHECK_ARG(d_in_bf16 && d_img && B > 0 && C > 0 && C <= pitch && H > 0 && W > 0, "icm_nhwc_to_image: bad arguments");
    const long long total = (long long)B * H * W;
    const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 16);
    nhwc_to_image_kernel<<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16 *)d_in_bf16, d_img, B, C, (long long)H * W, pitch, clamp01);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_eltwise_bf16(int mode, const void *d_a, int64_t pitch_a, const void *d_s, int64_t pitch_s, const void *d_x,
                                int64_t pitch_x, void *d_out, int64_t pitch_out, int64_t rows, int C, void *stream)
{
    ICM_CHECK_ARG((mode == 0 || ((mode == 1 || mode == 2) && d_a && d_s)) && d_x && d_out, "icm_eltwise_bf16: bad mode or null argument");
    ICM_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && pitch_x % 8 == 0 && pitch_out % 8 == 0 && pitch_a % 8 == 0 && pitch_s % 8 == 0,
                  "icm_eltwise_bf16: C and pitches must be multiples of 8");
    const long long total = rows * (C / 8);
    const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 16);
    eltwise_kernel<<<grid, 256, 0, as_stream(stream)>>>(mode, (const __nv_bfloat16 *)d_a, pitch_a, (const __nv_bfloat16 *)d_s, pitch_s,
                                                        (const __nv_bfloat16 *)d_x, pitch_x, d_out, pitch_out, rows, C);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

// Tensor-core version (mma.sync m16n8k16): one warp per (window, head).  K and V fragments of the whole window are
// loaded once, straight from the channels-last qkv rows; the query rows are walked 16 at a time: S = Q K^T (head_dim
// padded to a multiple of 16 with zero fragments), bias + shift mask + softmax on the accumulator fragment, which then
// serves as the A operand of O = P V (V transposed in registers with movmatrix).  The scalar kernel above spends
// 2 * N * HD FMAs and as many shared-memory loads per query row (12 ms per 64 images of 768x512 in WACNN).
template <int WIN, int HD>
__global__ void __launch_bounds__(128) win_attention_mma_kernel(const __nv_bfloat16 *__restrict__ qkv, __nv_bfloat16 *__restrict__ out,
                                                                const float *__restrict__ bias_table, int B, int H, int W, int C,
                                                                int heads, int shift)
{
    constexpr int NT = WIN * WIN;        // tokens per window
    constexpr int MT = NT / 16;          // 16-row query tiles
    constexpr int NKT = NT / 8;          // 8-key tiles
    constexpr int KS = (HD + 15) / 16;   // K = 16 steps of Q K^T
    constexpr int DT = HD / 8;           // 8-column tiles of the output
    constexpr int PK = NT / 16;          // K = 16 steps of P V
    constexpr int NB = (2 * WIN - 1) * (2 * WIN - 1);
    static_assert(NT % 16 == 0 && HD % 8 == 0, "window / head_dim not tileable");
    extern __shared__ float s_bias[]; // [NB][heads]
    for (int i = threadIdx.x; i < NB * heads; i += blockDim.x) s_bias[i] = bias_table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int g = lane >> 2, t = lane & 3;
    const int nWw = W / WIN, nWh = H / WIN;
    const long long jobs = (long long)B * nWh * nWw * heads;
    const float scale = rsqrtf((float)HD);
    for (int rep = 0; rep < 2; ++rep) { // two (window, head) jobs per warp share one copy of the bias table
    const long long job = ((long long)blockIdx.x * 4 + warp) * 2 + rep;
    if (job >= jobs) break; // warp-uniform
    const int head = (int)(job % heads);
    long long tt = job / heads;
    const int ww = (int)(tt % nWw); tt /= nWw;
    const int wh = (int)(tt % nWh);
    const int b = (int)(tt / nWh);
    auto token_of = [&](int tok) -> long long { // window token -> position in the (unshifted) feature map
        int h = wh * WIN + tok / WIN + shift, w = ww * WIN + tok % WIN + shift;
        if (h >= H) h -= H;
        if (w >= W) w -= W;
        return ((long long)b * H + h) * W + w;
    };
    auto label_of = [&](int tok) -> int {
        const int hs = wh * WIN + tok / WIN, ws = ww * WIN + tok % WIN;
        return 3 * (hs < H - WIN ? 0 : (hs < H - shift ? 1 : 2)) + (ws < W - WIN ? 0 : (ws < W - shift ? 1 : 2));
    };
    // K as the B operand of S (k = channel pair, n = key), V in its natural fragment layout, then transposed
    uint32_t kf[NKT][KS][2], vf[NKT][DT];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
        const uint32_t *row = reinterpret_cast<const uint32_t *>(qkv + token_of(nt * 8 + g) * 3 * C + head * HD);
        const uint32_t *rk = row + C / 2, *rv = row + C;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            kf[nt][ks][0] = (ks * 16 < HD) ? __ldg(rk + ks * 8 + t) : 0u;          // channels ks*16 + 2t, +1
            kf[nt][ks][1] = (ks * 16 + 8 < HD) ? __ldg(rk + ks * 8 + 4 + t) : 0u;  // channels ks*16 + 8 + 2t, +1
        }
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) vf[nt][dt] = __ldg(rv + dt * 4 + t);       // V[key][dt*8 + 2t, +1]
    }
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) vf[nt][dt] = movmatrix_trans(vf[nt][dt]);  // -> (V[2t][dt*8+g], V[2t+1][dt*8+g]) of key tile nt
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
        const int tokA = mt * 16 + g, tokB = tokA + 8;
        const long long tA = token_of(tokA), tB = token_of(tokB);
        const uint32_t *qA = reinterpret_cast<const uint32_t *>(qkv + tA * 3 * C + head * HD);
        const uint32_t *qB = reinterpret_cast<const uint32_t *>(qkv + tB * 3 * C + head * HD);
        float sacc[NKT][4];
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            uint32_t qa[4];
            qa[0] = (ks * 16 < HD) ? __ldg(qA + ks * 8 + t) : 0u;
            qa[1] = (ks * 16 < HD) ? __ldg(qB + ks * 8 + t) : 0u;
            qa[2] = (ks * 16 + 8 < HD) ? __ldg(qA + ks * 8 + 4 + t) : 0u;
            qa[3] = (ks * 16 + 8 < HD) ? __ldg(qB + ks * 8 + 4 + t) : 0u;
#pragma unroll
            for (int nt = 0; nt < NKT; ++nt) mma_bf16_16816(sacc[nt], qa, kf[nt][ks][0], kf[nt][ks][1]);
        }
        // bias + mask + softmax over the NT keys of rows tokA (slots 0, 1) and tokB (slots 2, 3)
        const int labA = shift > 0 ? label_of(tokA) : 0, labB = shift > 0 ? label_of(tokB) : 0;
        float mxA = -1e30f, mxB = -1e30f;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int j = nt * 8 + 2 * t + c, jh = j / WIN, jw = j % WIN;
                const float bA = s_bias[((tokA / WIN - jh + WIN - 1) * (2 * WIN - 1) + (tokA % WIN - jw + WIN - 1)) * heads + head];
                const float bB = s_bias[((tokB / WIN - jh + WIN - 1) * (2 * WIN - 1) + (tokB % WIN - jw + WIN - 1)) * heads + head];
                float vA = sacc[nt][c] * scale + bA, vB = sacc[nt][2 + c] * scale + bB;
                if (shift > 0) {
                    const int lj = label_of(j);
                    if (lj != labA) vA += -100.0f;
                    if (lj != labB) vB += -100.0f;
                }
                sacc[nt][c] = vA; sacc[nt][2 + c] = vB;
                mxA = fmaxf(mxA, vA); mxB = fmaxf(mxB, vB);
            }
        mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 1)); mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 2));
        mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 1)); mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 2));
        float dA = 0.f, dB = 0.f;
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                sacc[nt][c] = __expf(sacc[nt][c] - mxA); dA += sacc[nt][c];
                sacc[nt][2 + c] = __expf(sacc[nt][2 + c] - mxB); dB += sacc[nt][2 + c];
            }
        dA += __shfl_xor_sync(0xffffffffu, dA, 1); dA += __shfl_xor_sync(0xffffffffu, dA, 2);
        dB += __shfl_xor_sync(0xffffffffu, dB, 1); dB += __shfl_xor_sync(0xffffffffu, dB, 2);
        const float iA = 1.0f / dA, iB = 1.0f / dB;
        // O = P V with the (unnormalised) probabilities as A fragments
        float oacc[DT][4];
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) { oacc[dt][0] = oacc[dt][1] = oacc[dt][2] = oacc[dt][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < PK; ++kk) {
            uint32_t pa[4];
            pa[0] = pack_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);         pa[1] = pack_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
            pa[2] = pack_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]); pa[3] = pack_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
            for (int dt = 0; dt < DT; ++dt) mma_bf16_16816(oacc[dt], pa, vf[2 * kk][dt], vf[2 * kk + 1][dt]);
        }
        uint32_t *oA = reinterpret_cast<uint32_t *>(out + tA * C + head * HD), *oB = reinterpret_cast<uint32_t *>(out + tB * C + head * HD);
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
            oA[dt * 4 + t] = pack_bf16(oacc[dt][0] * iA, oacc[dt][1] * iA);
            oB[dt * 4 + t] = pack_bf16(oacc[dt][2] * iB, oacc[dt][3] * iB);
        }
    }
    }
}

template <int WIN, int HD>
static int launch_win_attention(const void *qkv, void *out, const float *bias, int B, int H, int W, int C, int heads, int shift, void *stream)
{
    constexpr int N = WIN * WIN, NB = (2 * WIN - 1) * (2 * WIN - 1);
    const long long jobs = (long long)B * (H / WIN) * (W / WIN) * heads;
    static const bool scalar = getenv("ICM_WACNN_SCALAR_ATTENTION") != nullptr; // the CUDA-core kernel, kept for comparison
    if (!scalar && NB * heads * sizeof(float) <= 48 * 1024) {
        win_attention_mma_kernel<WIN, HD><<<(unsigned)((jobs + 7) / 8), 128, (size_t)NB * heads * sizeof(float), as_stream(stream)>>>(
            (const __nv_bfloat16 *)qkv, (__nv_bfloat16 *)out, bias, B, H, W, C, heads, shift);
        ICM_LAUNCH_CHECK();
        return ICM_OK;
    }
    const size_t smem = ((size_t)NB * heads + (size_t)4 * 2 * N * (HD + 1)) * sizeof(float);
    static PerDeviceSmem configured;
    if (smem > 48 * 1024 && configured.needs(smem)) {
        ICM_CUDA(cudaFuncSetAttribute(win_attention_kernel<WIN, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.done(smem);
    }
    win_attention_kernel<WIN, HD><<<(unsigned)((jobs + 3) / 4), 128, smem, as_stream(stream)>>>(
        (const __nv_bfloat16 *)qkv, (__nv_bfloat16 *)out, bias, B, H, W, C, heads, shift);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_window_attention_wacnn(const void *d_qkv, void *d_out, const float *d_bias_table, int B, int H, int W, int C,
                                          int heads, int window, int shift, void *stream)
{
    ICM_CHECK_ARG(d_qkv && d_out && d_bias_table && heads > 0, "icm_window_attention_wacnn: null argument");
    ICM_CHECK_ARG(shift >= 0 && shift < window, "icm_window_attention_wacnn: bad shift");
    if (H % window || W % window) { set_error("icm_window_attention_wacnn: H=%d W=%d must be multiples of the window %d (the reference does not pad either)", H, W, window); return ICM_ERR_INVALID_ARG; }
    const int hd = C / heads;
    if (window == 8 && hd == 24 && C == heads * 24) return launch_win_attention<8, 24>(d_qkv, d_out, d_bias_table, B, H, W, C, heads, shift, stream);
    if (window == 4 && hd == 40 && C == heads * 40) return launch_win_attention<4, 40>(d_qkv, d_out, d_bias_table, B, H, W, C, heads, shift, stream);
    set_error("icm_window_attention_wacnn: built for (window 8, head_dim 24) and (window 4, head_dim 40); got window %d head_dim %d", window, hd);
    return ICM_ERR_UNSUPPORTED;
}

extern "C" int icm_pack_deconv_weight(const float *d_w_iohw, int Cin, int Cout, int Cin_pad, int Cq, void *d_out_bf16, void *stream)
{
    ICM_CHECK_ARG(d_w_iohw && d_out_bf16 && Cin > 0 && Cout > 0 && Cin_pad >= Cin && Cin_pad % 64 == 0 && Cq >= Cout && Cq % 16 == 0,
                  "icm_pack_deconv_weight: bad arguments");
    const long long total = (long long)4 * Cq * 9 * Cin_pad;
    const int grid = (int)min((total + 255) / 256, (long long)sm_count() * 8);
    pack_deconv_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_w_iohw, Cin, Cout, Cin_pad, Cq, (__nv_bfloat16 *)d_out_bf16);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}
