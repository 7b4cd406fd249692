// analysis.cu -- annotation-analysis path: downconvert + Welch PSD (C-ABI entry points).
#include "engine_internal.h"
#include "analysis_kernels.cuh"
using namespace sa;
struct sa_engine : public sa::Engine {};
extern "C" {
int32_t sa_lowpass_taps(int32_t, double*) { return set_error(SA_ERR_UNSUPPORTED, "not built yet"); }
int32_t sa_downconvert(sa_engine*, const void*, uint64_t, int32_t, int32_t, uint64_t, uint64_t, double, int32_t, int32_t,
                       double*, double*, uint64_t*) { return set_error(SA_ERR_UNSUPPORTED, "not built yet"); }
int32_t sa_psd_welch(sa_engine*, const double*, const double*, uint64_t, double, uint32_t, uint64_t, int32_t, double*, double*) {
    return set_error(SA_ERR_UNSUPPORTED, "not built yet"); }
int32_t sa_downconvert_psd_batch(sa_engine*, const void*, uint64_t, int32_t, int32_t, double, const sa_annotation*, uint32_t,
                                 uint32_t, uint64_t, int32_t, double*, const uint64_t*, double*) {
    return set_error(SA_ERR_UNSUPPORTED, "not built yet"); }
int32_t sa_downconvert_psd_batch_device(sa_engine*, const void*, uint64_t, int32_t, int32_t, double, const sa_annotation*, uint32_t,
                                        uint32_t, uint64_t, int32_t, double*, const uint64_t*, double*, void*) {
    return set_error(SA_ERR_UNSUPPORTED, "not built yet"); }
}
