// Training-step kernels (BASELINE.json configs[4]): the memory-bound pieces around the transforms'
// autograd graph.
//
//   gc_train_fwd / gc_train_bwd   GaussianConditional.forward in training mode (entropy_models.py:645-659 with
//                                 quantize "noise" :126-135, _likelihood :626-643, the two LowerBounds of bound_ops.py:21-62)
//                                 fused with the straight-through y_hat of stf.py:622 (ops.py:20-34): one pass forward,
//                                 one pass backward, instead of ~20 elementwise autograd nodes per slice.
//   sumsq / clip_coef             torch.nn.utils.clip_grad_norm_ (train.py:208-209) over ONE flat gradient buffer, the
//                                 coefficient stays on the device (no host synchronisation in the step).
//   adam_step                     torch.optim.Adam (train.py:161-168; defaults betas (0.9, 0.999), eps 1e-8, no weight
//                                 decay) over the flat parameter / gradient / moment buffers, gradient scale (clip
//                                 coefficient x 1/world) read from the device.
// All are HBM-bound: the forward moves 16 B in + 8 B out per latent element, the backward 24 B in + 12 B out, Adam 16 B in
// + 12 B out per parameter (28 B x 99.86 M parameters = 2.8 GB per step).
#include "common.cuh"

namespace icm {

constexpr float kRsqrt2 = 0.70710678118654752440f;
constexpr float kRsqrt2Pi = 0.39894228040143267794f;

struct Rows {  // element (r, i) of a [rows, n] tensor whose rows are `stride` elements apart
    float *p;
    long long stride;
};

__device__ __forceinline__ float std_cdf(float x) { return 0.5f * erfcf(-kRsqrt2 * x); }
__device__ __forceinline__ float std_pdf(float x) { return kRsqrt2Pi * __expf(-0.5f * x * x); }

__global__ void gc_train_fwd_kernel(Rows y, Rows noise, Rows mu, Rows scale, long long n, float scale_bound, float lik_bound,
                                    Rows lik, Rows y_hat)
{
    const long long r = blockIdx.y;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 yv = *reinterpret_cast<const float4 *>(y.p + r * y.stride + i);
        const float4 nv = *reinterpret_cast<const float4 *>(noise.p + r * noise.stride + i);
        const float4 mv = *reinterpret_cast<const float4 *>(mu.p + r * mu.stride + i);
        const float4 sv = *reinterpret_cast<const float4 *>(scale.p + r * scale.stride + i);
        const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, na[4] = {nv.x, nv.y, nv.z, nv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w}, sa[4] = {sv.x, sv.y, sv.z, sv.w};
        float l[4], h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float v = fabsf((ya[k] + na[k]) - ma[k]);
            const float s = fmaxf(sa[k], scale_bound);
            const float raw = std_cdf((0.5f - v) / s) - std_cdf((-0.5f - v) / s);
            l[k] = fmaxf(raw, lik_bound);
            h[k] = rintf(ya[k] - ma[k]) + ma[k];  // torch.round: half to even
        }
        *reinterpret_cast<float4 *>(lik.p + r * lik.stride + i) = make_float4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<float4 *>(y_hat.p + r * y_hat.stride + i) = make_float4(h[0], h[1], h[2], h[3]);
    }
}

// grad inputs: g_lik (gradient of the bounded likelihood), g_hat (gradient of the straight-through y_hat).
// d lik / d v = (pdf(b) - pdf(a)) / s,  d lik / d s = (pdf(b) b - pdf(a) a) / s  with a = (0.5 - v) / s, b = (-0.5 - v) / s;
// LowerBound passes a gradient where its input is >= the bound or the gradient is negative (bound_ops.py:33-36).
__global__ void gc_train_bwd_kernel(Rows y, Rows noise, Rows mu, Rows scale, Rows g_lik, Rows g_hat, long long n, float scale_bound,
                                    float lik_bound, Rows g_y, Rows g_mu, Rows g_scale)
{
    const long long r = blockIdx.y;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        const float4 yv = *reinterpret_cast<const float4 *>(y.p + r * y.stride + i);
        const float4 nv = *reinterpret_cast<const float4 *>(noise.p + r * noise.stride + i);
        const float4 mv = *reinterpret_cast<const float4 *>(mu.p + r * mu.stride + i);
        const float4 sv = *reinterpret_cast<const float4 *>(scale.p + r * scale.stride + i);
        const float4 glv = *reinterpret_cast<const float4 *>(g_lik.p + r * g_lik.stride + i);
        const float4 ghv = *reinterpret_cast<const float4 *>(g_hat.p + r * g_hat.stride + i);
        const float ya[4] = {yv.x, yv.y, yv.z, yv.w}, na[4] = {nv.x, nv.y, nv.z, nv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w}, sa[4] = {sv.x, sv.y, sv.z, sv.w};
        const float gl[4] = {glv.x, glv.y, glv.z, glv.w}, gh[4] = {ghv.x, ghv.y, ghv.z, ghv.w};
        float oy[4], om[4], os[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float d = (ya[k] + na[k]) - ma[k];
            const float v = fabsf(d), sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
            const float s = fmaxf(sa[k], scale_bound), inv = 1.0f / s;
            const float a = (0.5f - v) * inv, b = (-0.5f - v) * inv;
            const float raw = std_cdf(a) - std_cdf(b);
            const float g = (raw >= lik_bound || gl[k] < 0.f) ? gl[k] : 0.f;
            const float pa = std_pdf(a), pb = std_pdf(b);
            const float dv = g * (pb - pa) * inv * sgn;
            const float ds = g * (pb * b - pa * a) * inv;
            oy[k] = dv + gh[k];
            om[k] = -dv;  // the straight-through y_hat = round(y - mu) + mu has no gradient towards mu (-g + g)
            os[k] = (sa[k] >= scale_bound || ds < 0.f) ? ds : 0.f;
        }
        *reinterpret_cast<float4 *>(g_y.p + r * g_y.stride + i) = make_float4(oy[0], oy[1], oy[2], oy[3]);
        *reinterpret_cast<float4 *>(g_mu.p + r * g_mu.stride + i) = make_float4(om[0], om[1], om[2], om[3]);
        *reinterpret_cast<float4 *>(g_scale.p + r * g_scale.stride + i) = make_float4(os[0], os[1], os[2], os[3]);
    }
}

__global__ void sumsq_kernel(const float *__restrict__ x, long long n, float *out)
{
    float acc = 0.f;
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4 *>(x)[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = x[(n4 << 2) + threadIdx.x]; acc += v * v; }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) atomicAdd(out, acc);
    }
}

// coef = pre * min(1, max_norm / (pre * sqrt(sumsq) + 1e-6)): the gradients in the buffer are `1 / pre` times the ones the
// norm is defined on (sums over `world` ranks, pre = 1 / world); max_norm <= 0 switches clipping off (train.py:208).
__global__ void clip_coef_kernel(const float *sumsq, float max_norm, float pre, float *coef, float *norm_out)
{
    const float norm = pre * sqrtf(*sumsq);
    float c = 1.f;
    if (max_norm > 0.f) c = fminf(1.f, max_norm / (norm + 1e-6f));
    *coef = pre * c;
    if (norm_out) *norm_out = norm;
}

// Step counter and bias corrections kept on the device ({t, 1 - b1^t, 1 / sqrt(1 - b2^t)}), so that a training step captured in a
// CUDA graph advances them on every replay.
__global__ void adam_tick_kernel(float *state, float b1, float b2)
{
    const float t = state[0] + 1.f;
    state[0] = t;
    state[1] = (float)(1.0 - pow((double)b1, (double)t));
    state[2] = (float)(1.0 / sqrt(1.0 - pow((double)b2, (double)t)));
}

__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, long long n,
                            float lr, float b1, float b2, float eps, float bc1, float rsqrt_bc2, const float *gscale_dev, float gscale,
                            const float *state)
{
    if (state) { bc1 = state[1]; rsqrt_bc2 = state[2]; }
    const float gs = gscale_dev ? gscale * *gscale_dev : gscale;
    const float step = lr / bc1;
    const long long n4 = n >> 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pv = reinterpret_cast<float4 *>(p)[i], mv = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gv = reinterpret_cast<const float4 *>(g)[i];
        float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
        const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            ma[k] = b1 * ma[k] + (1.f - b1) * gk;
            va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
            pa[k] -= step * (ma[k] / (sqrtf(va[k]) * rsqrt_bc2 + eps));
        }
        reinterpret_cast<float4 *>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
        reinterpret_cast<float4 *>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
        reinterpret_cast<float4 *>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float gk = g[i] * gs;
        const float mk = b1 * m[i] + (1.f - b1) * gk, vk = b2 * v[i] + (1.f - b2) * gk * gk;
        m[i] = mk; v[i] = vk;
        p[i] -= step * (mk / (sqrtf(vk) * rsqrt_bc2 + eps));
    }
}

static bool rows_ok(const icm_rows &t, int64_t n) { return t.ptr && ((uintptr_t)t.ptr & 15) == 0 && t.stride % 4 == 0 && t.stride >= 0 && n % 4 == 0; }
static Rows as_rows(const icm_rows &t) { return Rows{(float *)t.ptr, (long long)t.stride}; }

static dim3 rows_grid(int64_t rows, int64_t n)
{
    long long bx = (n / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8 / (rows < 1 ? 1 : rows) + 1;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    return dim3((unsigned)bx, (unsigned)rows, 1);
}

}  // namespace icm

using namespace icm;

extern "C" int icm_gc_train_forward(icm_rows y, icm_rows noise, icm_rows mu, icm_rows scale, int64_t rows, int64_t n, float scale_bound,
                                    float likelihood_bound, icm_rows likelihood, icm_rows y_hat, void *stream)
{
    ICM_CHECK_ARG(rows > 0 && rows <= 65535 && n > 0, "icm_gc_train_forward: rows=%lld n=%lld", (long long)rows, (long long)n);
    ICM_CHECK_ARG(rows_ok(y, n) && rows_ok(noise, n) && rows_ok(mu, n) && rows_ok(scale, n) && rows_ok(likelihood, n) && rows_ok(y_hat, n),
                  "icm_gc_train_forward: every tensor needs a 16-byte aligned base, row strides and n multiples of 4");
    gc_train_fwd_kernel<<<rows_grid(rows, n), 256, 0, as_stream(stream)>>>(as_rows(y), as_rows(noise), as_rows(mu), as_rows(scale), n, scale_bound,
                                                                            likelihood_bound, as_rows(likelihood), as_rows(y_hat));
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_gc_train_backward(icm_rows y, icm_rows noise, icm_rows mu, icm_rows scale, icm_rows g_likelihood, icm_rows g_y_hat,
                                     int64_t rows, int64_t n, float scale_bound, float likelihood_bound, icm_rows g_y, icm_rows g_mu,
                                     icm_rows g_scale, void *stream)
{
    ICM_CHECK_ARG(rows > 0 && rows <= 65535 && n > 0, "icm_gc_train_backward: rows=%lld n=%lld", (long long)rows, (long long)n);
    ICM_CHECK_ARG(rows_ok(y, n) && rows_ok(noise, n) && rows_ok(mu, n) && rows_ok(scale, n) && rows_ok(g_likelihood, n) && rows_ok(g_y_hat, n) &&
                      rows_ok(g_y, n) && rows_ok(g_mu, n) && rows_ok(g_scale, n),
                  "icm_gc_train_backward: every tensor needs a 16-byte aligned base, row strides and n multiples of 4");
    gc_train_bwd_kernel<<<rows_grid(rows, n), 256, 0, as_stream(stream)>>>(as_rows(y), as_rows(noise), as_rows(mu), as_rows(scale), as_rows(g_likelihood),
                                                                            as_rows(g_y_hat), n, scale_bound, likelihood_bound, as_rows(g_y), as_rows(g_mu),
                                                                            as_rows(g_scale));
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_grad_sumsq(const float *d_x, int64_t n, float *d_sumsq, void *stream)
{
    ICM_CHECK_ARG(d_x && d_sumsq && n > 0 && ((uintptr_t)d_x & 15) == 0, "icm_grad_sumsq: null / misaligned argument");
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_x, n, d_sumsq);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_clip_coef(const float *d_sumsq, float max_norm, float pre_scale, float *d_coef, float *d_norm, void *stream)
{
    ICM_CHECK_ARG(d_sumsq && d_coef && pre_scale > 0.f, "icm_clip_coef: null argument");
    clip_coef_kernel<<<1, 1, 0, as_stream(stream)>>>(d_sumsq, max_norm, pre_scale, d_coef, d_norm);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_adam_step(float *d_param, const float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, int step, float *d_step_state, const float *d_grad_scale, float grad_scale, void *stream)
{
    ICM_CHECK_ARG(d_param && d_grad && d_exp_avg && d_exp_avg_sq && n > 0 && (step >= 1 || d_step_state), "icm_adam_step: null argument or step < 1");
    ICM_CHECK_ARG((((uintptr_t)d_param | (uintptr_t)d_grad | (uintptr_t)d_exp_avg | (uintptr_t)d_exp_avg_sq) & 15) == 0, "icm_adam_step: buffers must be 16-byte aligned");
    double bc1 = 1.0, bc2 = 1.0;
    if (d_step_state) {
        adam_tick_kernel<<<1, 1, 0, as_stream(stream)>>>(d_step_state, beta1, beta2);
        ICM_LAUNCH_CHECK();
    } else {
        bc1 = 1.0 - pow((double)beta1, step);
        bc2 = 1.0 - pow((double)beta2, step);
    }
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, lr, beta1, beta2, eps, (float)bc1,
                                                                   (float)(1.0 / sqrt(bc2)), d_grad_scale, grad_scale, d_step_state);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

// ------------------------------------------------------------------------------------------------ LayerNorm, training
// nn.LayerNorm forward + backward for the Swin blocks of the training step (stf.py:155,197,232,256,379; 55 calls per step).
// torch's own three kernels per call (forward, input gradient, gamma/beta gradient) took 15 ms of an 86 ms step on the
// [262 144, 48] ... [1 024, 768] token tensors of a 16 x 256 x 256 batch; these move each tensor once.
//   forward : y = (x - mean) * rstd * gamma + beta, mean / rstd kept per row; y in fp32 or (under autocast) bf16
//   backward: dx = rstd * (g gamma - mean_c(g gamma) - xhat mean_c(g gamma xhat)); dgamma = sum_rows g xhat; dbeta = sum_rows g
// LPR lanes share a row (C / 4 float4 chunks dealt round-robin), so a warp works on 32 / LPR rows at once; the per-channel
// gamma / beta gradients are accumulated in registers over the rows a lane sees, folded per CTA in shared memory and added to
// global memory with one atomicAdd per channel and CTA (summation order across CTAs is not fixed: fine for training).
namespace icm {

constexpr int LN_NQ = 6;       // float4 chunks per lane: C <= 4 * 32 * 6 = 768
constexpr int LN_THREADS = 256;

__device__ __forceinline__ float group_sum(float v, int lpr)
{
    for (int o = lpr >> 1; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float4 load4(const void *p, long long elem, int is_bf16)
{
    if (!is_bf16) return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p) + elem);
    const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __nv_bfloat16 *>(p) + elem);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

__global__ void __launch_bounds__(LN_THREADS) ln_train_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                                                                 const float *__restrict__ beta, void *__restrict__ y, int y_bf16,
                                                                 float *__restrict__ mean, float *__restrict__ rstd, long long rows, int C,
                                                                 int lpr)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & (lpr - 1), grp = lane / lpr, gpw = 32 / lpr;
    const int C4 = C >> 2;
    const long long stride = (long long)gridDim.x * (LN_THREADS / 32) * gpw;
    // every lane of a warp must take part in the shuffles: loop on the warp's first row, guard the row itself
    for (long long base = ((long long)blockIdx.x * (LN_THREADS / 32) + warp) * gpw; base < rows; base += stride) {
        const long long row = base + grp;
        const bool live = row < rows;
        float4 v[LN_NQ];
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < LN_NQ; ++q) {
            const int c4 = sub + q * lpr;
            v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && c4 < C4) v[q] = *reinterpret_cast<const float4 *>(x + row * C + 4 * c4);
            s += (v[q].x + v[q].y) + (v[q].z + v[q].w);
        }
        const float mu = group_sum(s, lpr) / (float)C;
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < LN_NQ; ++q) {
            if (sub + q * lpr < C4) {
                float d;
                d = v[q].x - mu; ss += d * d; d = v[q].y - mu; ss += d * d; d = v[q].z - mu; ss += d * d; d = v[q].w - mu; ss += d * d;
            }
        }
        const float rs = rsqrtf(group_sum(ss, lpr) / (float)C + 1e-5f);
        if (!live) continue;
        if (sub == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
        for (int q = 0; q < LN_NQ; ++q) {
            const int c4 = sub + q * lpr;
            if (c4 >= C4) continue;
            const float4 g = *reinterpret_cast<const float4 *>(gamma + 4 * c4), b = *reinterpret_cast<const float4 *>(beta + 4 * c4);
            const float4 o = make_float4((v[q].x - mu) * rs * g.x + b.x, (v[q].y - mu) * rs * g.y + b.y, (v[q].z - mu) * rs * g.z + b.z,
                                         (v[q].w - mu) * rs * g.w + b.w);
            if (y_bf16) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
                *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(y) + row * C + 4 * c4) =
                    make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
            } else {
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(y) + row * C + 4 * c4) = o;
            }
        }
    }
}

__global__ void __launch_bounds__(LN_THREADS) ln_train_bwd_kernel(const float *__restrict__ x, const void *__restrict__ g, int g_bf16,
                                                                 const float *__restrict__ gamma, const float *__restrict__ mean,
                                                                 const float *__restrict__ rstd, float *__restrict__ dx,
                                                                 float *__restrict__ dgamma, float *__restrict__ dbeta, long long rows, int C, int lpr)
{
    extern __shared__ float s_acc[]; // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += LN_THREADS) s_acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & (lpr - 1), grp = lane / lpr, gpw = 32 / lpr;
    const int C4 = C >> 2;
    const long long stride = (long long)gridDim.x * (LN_THREADS / 32) * gpw;
    float4 ag[LN_NQ], ab[LN_NQ], gm[LN_NQ];
#pragma unroll
    for (int q = 0; q < LN_NQ; ++q) {
        ag[q] = ab[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c4 = sub + q * lpr;
        gm[q] = c4 < C4 ? *reinterpret_cast<const float4 *>(gamma + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long base = ((long long)blockIdx.x * (LN_THREADS / 32) + warp) * gpw; base < rows; base += stride) {
        const long long row = base + grp;
        const bool live = row < rows;
        const float mu = live ? mean[row] : 0.f, rs = live ? rstd[row] : 0.f;
        float4 xh[LN_NQ], gg[LN_NQ];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < LN_NQ; ++q) {
            const int c4 = sub + q * lpr;
            xh[q] = gg[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && c4 < C4) {
                const float4 xv = *reinterpret_cast<const float4 *>(x + row * C + 4 * c4);
                const float4 gv = load4(g, row * C + 4 * c4, g_bf16);
                xh[q] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                ag[q].x += gv.x * xh[q].x; ag[q].y += gv.y * xh[q].y; ag[q].z += gv.z * xh[q].z; ag[q].w += gv.w * xh[q].w;
                ab[q].x += gv.x; ab[q].y += gv.y; ab[q].z += gv.z; ab[q].w += gv.w;
                gg[q] = make_float4(gv.x * gm[q].x, gv.y * gm[q].y, gv.z * gm[q].z, gv.w * gm[q].w);
                s1 += (gg[q].x + gg[q].y) + (gg[q].z + gg[q].w);
                s2 += (gg[q].x * xh[q].x + gg[q].y * xh[q].y) + (gg[q].z * xh[q].z + gg[q].w * xh[q].w);
            }
        }
        s1 = group_sum(s1, lpr) / (float)C;
        s2 = group_sum(s2, lpr) / (float)C;
        if (!live) continue;
#pragma unroll
        for (int q = 0; q < LN_NQ; ++q) {
            const int c4 = sub + q * lpr;
            if (c4 >= C4) continue;
            *reinterpret_cast<float4 *>(dx + row * C + 4 * c4) =
                make_float4(rs * (gg[q].x - s1 - xh[q].x * s2), rs * (gg[q].y - s1 - xh[q].y * s2), rs * (gg[q].z - s1 - xh[q].z * s2),
                            rs * (gg[q].w - s1 - xh[q].w * s2));
        }
    }
#pragma unroll
    for (int q = 0; q < LN_NQ; ++q) {
        const int c4 = sub + q * lpr;
        if (c4 >= C4) continue;
        atomicAdd(&s_acc[4 * c4], ag[q].x); atomicAdd(&s_acc[4 * c4 + 1], ag[q].y); atomicAdd(&s_acc[4 * c4 + 2], ag[q].z); atomicAdd(&s_acc[4 * c4 + 3], ag[q].w);
        atomicAdd(&s_acc[C + 4 * c4], ab[q].x); atomicAdd(&s_acc[C + 4 * c4 + 1], ab[q].y); atomicAdd(&s_acc[C + 4 * c4 + 2], ab[q].z); atomicAdd(&s_acc[C + 4 * c4 + 3], ab[q].w);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += LN_THREADS) { atomicAdd(dgamma + i, s_acc[i]); atomicAdd(dbeta + i, s_acc[C + i]); }
}

static int ln_lanes_per_row(int C)
{
    int lpr = 4;
    while (lpr < 32 && lpr < C / 4) lpr *= 2;
    return lpr;
}

static unsigned ln_grid(int64_t rows, int lpr, int per_sm)
{
    const long long rows_per_cta = (LN_THREADS / 32) * (32 / lpr);
    long long need = (rows + rows_per_cta - 1) / rows_per_cta, cap = (long long)sm_count() * per_sm;
    if (need > cap) need = cap;
    return (unsigned)(need < 1 ? 1 : need);
}

}  // namespace icm

extern "C" int icm_layernorm_train_forward(const float *d_x, const float *d_gamma, const float *d_beta, void *d_y, int y_dtype,
                                           float *d_mean, float *d_rstd, int64_t rows, int C, void *stream)
{
    ICM_CHECK_ARG(d_x && d_gamma && d_beta && d_y && d_mean && d_rstd, "icm_layernorm_train_forward: null argument");
    ICM_CHECK_ARG(rows > 0 && C >= 4 && C % 4 == 0 && C <= 4 * 32 * LN_NQ, "icm_layernorm_train_forward: rows=%lld C=%d (C must be a multiple of 4, <= 768)", (long long)rows, C);
    ICM_CHECK_ARG((((uintptr_t)d_x | (uintptr_t)d_gamma | (uintptr_t)d_beta) & 15) == 0 && ((uintptr_t)d_y & 7) == 0, "icm_layernorm_train_forward: misaligned pointer");
    const int lpr = ln_lanes_per_row(C);
    ln_train_fwd_kernel<<<ln_grid(rows, lpr, 16), LN_THREADS, 0, as_stream(stream)>>>(d_x, d_gamma, d_beta, d_y, y_dtype == ICM_OUT_BF16, d_mean, d_rstd, rows, C, lpr);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}

extern "C" int icm_layernorm_train_backward(const float *d_x, const void *d_grad_y, int grad_dtype, const float *d_gamma, const float *d_mean,
                                            const float *d_rstd, float *d_grad_x, float *d_grad_gamma, float *d_grad_beta, int64_t rows, int C,
                                            void *stream)
{
    ICM_CHECK_ARG(d_x && d_grad_y && d_gamma && d_mean && d_rstd && d_grad_x && d_grad_gamma && d_grad_beta, "icm_layernorm_train_backward: null argument");
    ICM_CHECK_ARG(rows > 0 && C >= 4 && C % 4 == 0 && C <= 4 * 32 * LN_NQ, "icm_layernorm_train_backward: rows=%lld C=%d", (long long)rows, C);
    ICM_CHECK_ARG((((uintptr_t)d_x | (uintptr_t)d_gamma | (uintptr_t)d_grad_x) & 15) == 0 && ((uintptr_t)d_grad_y & 7) == 0, "icm_layernorm_train_backward: misaligned pointer");
    ICM_CUDA(cudaMemsetAsync(d_grad_gamma, 0, (size_t)C * 4, as_stream(stream)));
    ICM_CUDA(cudaMemsetAsync(d_grad_beta, 0, (size_t)C * 4, as_stream(stream)));
    const int lpr = ln_lanes_per_row(C);
    ln_train_bwd_kernel<<<ln_grid(rows, lpr, 4), LN_THREADS, (size_t)2 * C * 4, as_stream(stream)>>>(d_x, d_grad_y, grad_dtype == ICM_OUT_BF16, d_gamma, d_mean, d_rstd,
                                                                                                        d_grad_x, d_grad_gamma, d_grad_beta, rows, C, lpr);
    ICM_LAUNCH_CHECK();
    return ICM_OK;
}
