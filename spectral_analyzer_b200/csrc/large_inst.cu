// large_inst.cu -- instantiations of the four-step kernels: -DSA_LARGE_PREC=1|2 -DSA_LARGE_N=<nfft>.
#include "large_fft_kernels.cuh"

namespace sa {
namespace {
#if SA_LARGE_N == 16384
constexpr int LN1 = 128, LN2 = 128;
#elif SA_LARGE_N == 32768
constexpr int LN1 = 128, LN2 = 256;
#else
constexpr int LN1 = 256, LN2 = 256;
#endif
struct Registrar {
    Registrar() {
#if SA_LARGE_PREC == 1
        register_large_kernel(make_large_info<float, LN1, LN2, DK_CF32, false>(1));
        register_large_kernel(make_large_info<float, LN1, LN2, DK_CF32, true>(1));
        register_large_kernel(make_large_info<float, LN1, LN2, DK_CI16, false>(1));
        register_large_kernel(make_large_info<float, LN1, LN2, DK_CI16, true>(1));
        register_large_kernel(make_large_info<float, LN1, LN2, DK_C8, false>(1));
        register_large_kernel(make_large_info<float, LN1, LN2, DK_C8, true>(1));
#else
        register_large_kernel(make_large_info<double, LN1, LN2, DK_CF32, true>(2));
        register_large_kernel(make_large_info<double, LN1, LN2, DK_CI16, true>(2));
        register_large_kernel(make_large_info<double, LN1, LN2, DK_C8, true>(2));
        register_large_kernel(make_large_info<double, LN1, LN2, DK_CF64, true>(2));
#endif
    }
};
static Registrar registrar_instance;
}  // namespace
}  // namespace sa
