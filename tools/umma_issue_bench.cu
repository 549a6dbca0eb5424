// Two microbenchmarks behind DESIGN.md 4.1 / 4.1b (what bounds the N <= 176 layers of conv_igemm_kernel and swin_mlp_kernel):
//  (1) the single-thread tcgen05.mma issue loop: cycles per MMA as a function of N and of how many TMEM accumulators the
//      consecutive MMAs rotate over, plus the cost of an already-satisfied mbarrier wait and of the tcgen05 fence;
//  (2) L2 -> shared memory TMA bandwidth per SM: every CTA streams 16 KB boxes (one k-step's A operand) of an L2-resident
//      buffer through an 8-slot ring, 1 to 148 CTAs at a time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I image-compression-for-machine_b200/csrc \
//        tools/umma_issue_bench.cu -o tools/_bin/umma_issue_bench && tools/_bin/umma_issue_bench
// Standalone (no torch): a run takes a second.  Operands are zeros; only timing is observed.
#include "umma.cuh"

#include <vector>

using namespace icm;

struct Result { long long issue, done, wait_sat, fence; };

__global__ void __launch_bounds__(64) issue_bench(int N, int n_acc, int n_mma, int rotate_smem, Result *out)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar_sat;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    constexpr int kStages = 4, kStageBytes = 16384 + 32768; // A 128 x 64 bf16, B up to 256 x 64 bf16
    for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(tiles)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar_sat, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && elect_one()) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t acc_stride = (uint32_t)((N + 31) & ~31);
        const uint64_t desc0 = make_smem_desc(smem_u32(tiles));
        const uint64_t boff = 16384 >> 4, sstep = kStageBytes >> 4;
        uint32_t phase = 0;
        // warm-up
        for (int i = 0; i < 16; ++i) umma_bf16(tmem_base, desc0 + (uint64_t)((i & 3) * 2), desc0 + boff + (uint64_t)((i & 3) * 2), idesc, i != 0);
        umma_commit(&bar);
        mbar_wait(&bar, phase); phase ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t0 = clock64();
        int acc = 0, stage = 0;
        for (int i = 0; i < n_mma; i += 4) {
            const uint64_t da = desc0 + (rotate_smem ? (uint64_t)stage * sstep : 0);
            const uint32_t d = tmem_base + (uint32_t)acc * acc_stride;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // n_acc > 1: consecutive MMAs go to different accumulators (k-steps of n_acc tiles interleaved)
                const uint32_t dk = n_acc > 1 ? tmem_base + (uint32_t)((acc + k) & (n_acc - 1)) * acc_stride : d; // n_acc = 2, 4
                umma_bf16(dk, da + (uint64_t)(k * 2), da + boff + (uint64_t)(k * 2), idesc, 1u);
            }
            if (++stage == kStages) stage = 0;
            if (n_acc > 1 && ++acc == n_acc) acc = 0;
        }
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, phase); phase ^= 1;
        const long long t2 = clock64();
        // an already-satisfied wait: bar_sat completed phase 0 once
        mbar_arrive(&bar_sat);
        mbar_wait(&bar_sat, 0);
        const long long t3 = clock64();
        for (int i = 0; i < 256; ++i) mbar_wait(&bar_sat, 0);
        const long long t4 = clock64();
        for (int i = 0; i < 256; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long t5 = clock64();
        if (blockIdx.x == 0) *out = Result{t1 - t0, t2 - t0, (t4 - t3) / 256, (t5 - t4) / 256};
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// (2) one thread per CTA keeps 8 TMA loads of 16 KB in flight; nothing consumes them
__global__ void __launch_bounds__(32) tma_bw(const __grid_constant__ CUtensorMap map, int n_loads, int rows_total, long long *cycles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (elect_one()) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
        int row = (int)(((long long)blockIdx.x * 128 * 37) % rows_total);
        const int stride = (int)gridDim.x * 128;
        const long long t0 = clock64();
        for (int i = 0; i < n_loads + 8; ++i) {
            const int s = i & 7;
            if (i >= 8) mbar_wait(&full[s], (uint32_t)(((i >> 3) - 1) & 1)); // the previous load into this slot has landed
            if (i < n_loads) {
                mbar_expect_tx(&full[s], 16384);
                tma_load_2d(&map, &full[s], tiles + s * 16384, 0, row);
                row += stride;
                if (row >= rows_total) row -= rows_total;
            }
        }
        cycles[blockIdx.x] = clock64() - t0;
    }
}

static int tma_bandwidth()
{
    const int rows_total = 1 << 16; // pixels; at the widest pitch 2^16 x 512 B = 32 MB: L2-resident after the first pass
    void *buf;
    if (cudaMalloc(&buf, (size_t)rows_total * 512) != cudaSuccess) return 1;
    cudaMemset(buf, 0, (size_t)rows_total * 512);
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    long long *d_cyc;
    cudaMalloc(&d_cyc, 1024 * sizeof(long long));
    const size_t smem = 8 * 16384 + 1024;
    cudaFuncSetAttribute(tma_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("\nL2 -> shared memory, TMA 2-D boxes of 16 KB (64 channels x 128 pixels, SWIZZLE_128B), 8 in flight per CTA, L2-resident buffer (second pass);\n"
           "pitch = bytes between pixels (channels-last activation with C = pitch / 2 channels), c0 = first channel of the box\n");
    printf("%5s %5s %5s | %12s %12s | %12s\n", "pitch", "c0", "CTAs", "B/clk/SM avg", "B/clk/SM min", "B/clk chip");
    const int n_loads = 2048;
    for (int pitch : {128, 256, 352, 384, 448, 512})
        for (int c0 : {0, 64}) {
            if (c0 * 2 + 128 > pitch) continue;
            CUtensorMap map;
            cuuint64_t dims[2] = {(cuuint64_t)(pitch / 2), (cuuint64_t)rows_total}, strides[1] = {(cuuint64_t)pitch};
            cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
            if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (char *)buf + c0 * 2, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("tensor map failed (pitch %d)\n", pitch);
                continue;
            }
            for (int grid : {1, 148}) {
                std::vector<long long> h(grid);
                for (int pass = 0; pass < 2; ++pass) {
                    tma_bw<<<grid, 32, smem>>>(map, n_loads, rows_total, d_cyc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                }
                cudaMemcpy(h.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                double sum = 0, mn = 1e30;
                for (long long c : h) { const double bw = (double)n_loads * 16384 / (double)c; sum += bw; mn = bw < mn ? bw : mn; }
                printf("%5d %5d %5d | %12.1f %12.1f | %12.0f\n", pitch, c0, grid, sum / grid, mn, sum);
            }
        }
    return 0;
}

int main()
{
    Result *d_out;
    cudaMalloc(&d_out, sizeof(Result));
    const size_t smem = 4 * (16384 + 32768) + 1024;
    cudaFuncSetAttribute(issue_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int n_mma = 4096;
    printf("cycles per tcgen05.mma (M = 128, K = 16, bf16), %d MMAs from one thread; issue = until the last one is issued, done = until the commit arrives\n", n_mma);
    printf("%5s %5s %6s %6s | %8s %8s | %10s\n", "grid", "N", "n_acc", "rotA", "issue", "done", "ideal(N/2)");
    for (int grid : {1, 148})
        for (int N : {32, 64, 128, 192, 224, 256})
            for (int n_acc : {1, 2, 4})
                for (int rot : {0, 1}) {
                    if (((N + 31) & ~31) * n_acc > 512) continue;
                    if (grid == 148 && rot == 0) continue;
                    issue_bench<<<grid, 64, smem>>>(N, n_acc, n_mma, rot, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
                    Result r;
                    cudaMemcpy(&r, d_out, sizeof(r), cudaMemcpyDeviceToHost);
                    printf("%5d %5d %6d %6d | %8.1f %8.1f | %10.1f   satisfied mbarrier wait %lld, tcgen05.fence %lld cycles\n", grid, N, n_acc, rot,
                           (double)r.issue / n_mma, (double)r.done / n_mma, N / 2.0, r.wait_sat, r.fence);
                }
    return tma_bandwidth();
}
