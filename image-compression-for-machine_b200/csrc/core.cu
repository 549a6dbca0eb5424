// Library plumbing: error string, launch counter, device query, and the host-side
// pmf -> quantised-CDF routine (R5: compressai._CXX.pmf_to_quantized_cdf, _CXX.so@0x68c0,
// called from entropy_models.py:60-63,172-180).
#include "common.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <vector>

namespace icm {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device_ordinal()
{
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDeviceOrdinals ? dev : 0;
}

int sm_count()
{
    static int cached[kMaxDeviceOrdinals] = {};
    const int dev = current_device_ordinal();
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

static thread_local int g_persistent_sm_limit = 0;
int persistent_grid_limit()
{
    const int n = sm_count();
    return (g_persistent_sm_limit > 0 && g_persistent_sm_limit < n) ? g_persistent_sm_limit : n;
}
void set_persistent_grid_limit(int n) { g_persistent_sm_limit = n; }


}  // namespace icm

extern "C" const char *icm_last_error(void) { return icm::g_error; }
extern "C" int icm_abi_version(void) { return 1; }
extern "C" int64_t icm_launch_count(void) { return icm::g_launches.load(); }
extern "C" int64_t icm_note_graph_launches(int64_t n)
{
    if (n > 0) icm::g_launches.fetch_add(n, std::memory_order_relaxed);
    return icm::g_launches.load();
}

// Frequencies are rounded in float32 (the reference rounds `float p * (1 << precision)`), rescaled to sum
// to 2^precision by integer division, and zero-width bins are repaired by taking one count from the
// least-frequent bin that can spare it.  The donor scan keeps the reference's "first smallest freq > 1"
// tie-break, which fixes the result bit for bit.
extern "C" int icm_pmf_to_quantized_cdf(const float *h_pmf, int n, int precision, uint32_t *h_out)
{
    ICM_CHECK_ARG(h_pmf && h_out, "icm_pmf_to_quantized_cdf: null argument");
    ICM_CHECK_ARG(n >= 1 && precision >= 1 && precision <= 31, "icm_pmf_to_quantized_cdf: bad n=%d precision=%d", n, precision);
    const int bins = n;
    std::vector<uint32_t> freq(bins);
    uint32_t total = 0;
    const float scale = (float)(1u << precision);
    for (int i = 0; i < bins; ++i) {
        freq[i] = (uint32_t)std::round(h_pmf[i] * scale);
        total += freq[i];
    }
    ICM_CHECK_ARG(total != 0, "icm_pmf_to_quantized_cdf: pmf sums to zero");
    h_out[0] = 0;
    uint32_t run = 0;
    for (int i = 0; i < bins; ++i) {
        run += (uint32_t)((((uint64_t)1 << precision) * (uint64_t)freq[i]) / total);
        h_out[i + 1] = run;
    }
    h_out[bins] = 1u << precision;
    for (int i = 0; i < bins; ++i) {
        if (h_out[i] != h_out[i + 1]) continue;
        uint32_t best = ~0u;
        int donor = -1;
        for (int j = 0; j < bins; ++j) {
            const uint32_t f = h_out[j + 1] - h_out[j];
            if (f > 1 && f < best) { best = f; donor = j; }
        }
        ICM_CHECK_ARG(donor >= 0, "icm_pmf_to_quantized_cdf: no bin can donate a count (too many symbols for the precision)");
        if (donor < i) for (int j = donor + 1; j <= i; ++j) --h_out[j];
        else           for (int j = i + 1; j <= donor; ++j) ++h_out[j];
    }
    return n + 1;
}
