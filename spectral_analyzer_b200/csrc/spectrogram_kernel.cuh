// spectrogram_kernel.cuh -- decode -> frame -> window -> FFT -> |X| -> dB (-> RGBA) in one kernel.
//
// Replaces, for a whole batch of frames, the loop of S/controllers/MainController.java:980-999
// around S/services/SpectralService.java:33-85 (and, for SA_OUT_RGBA8, the per-pixel colour
// mapping of MainController.java:1273-1285 with getColorForMagnitude :926-957).
// Persistent CTAs walk frame blocks with a grid stride; every frame is read straight from global
// memory with coalesced per-point loads (thread t takes points t + TPF*q), transformed in
// registers/shared memory (fft_core.cuh) and written as one fft-shifted row.
#pragma once
#include "decode.cuh"

namespace sa {

enum { OUT_F32_DB = 0, OUT_F64_DB = 1, OUT_RGBA8 = 2 };
enum { DBM_MAG_1E10 = 0, DBM_POWER = 1 };

struct SpecArgs {
    LoadParams  lp;
    long long   n_samples;     // IQ pairs readable from lp.base
    long long   start_sample;  // MainController.java:984 currentSampleOffset
    long long   hop;
    long long   n_frames;
    const void* window;        // T[N] or nullptr
    const void* twiddle;       // cpx<T>[Geo::TW_ELEMS]
    void*       out;
    int         out_kind;
    int         db_mode;
    double      eof_fill;      // -150.0, MainController.java:996-997
    // renderSpectrogram parameters (RGBA8 only)
    float       conv;          // 10*log10(fs/N) + 20*log10(N), :1273-1274
    float       min_db;
    float       inv_range;     // 1/(max_db - min_db), :929
    int         cmap;
};

__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrt_approx(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// SpectralService.java:80-81: 20*log10(abs + 1e-10)
__device__ __forceinline__ float to_db(float re, float im, int mode) {
    const float p = __fmaf_rn(re, re, im * im);
    if (mode == DBM_MAG_1E10) return 6.02059991327962f * lg2_approx(sqrt_approx(p) + 1e-10f);
    return 3.01029995663981f * lg2_approx(p + 1e-20f);
}
__device__ __forceinline__ double to_db(double re, double im, int mode) {
    if (mode == DBM_MAG_1E10) return 20.0 * log10(sqrt(re * re + im * im) + 1e-10);
    return 10.0 * log10(re * re + im * im + 1e-20);
}

// getColorForMagnitude (MainController.java:926-957) on float components, channels packed
// R | G<<8 | B<<16 | A<<24 with floor(c*255 + 0.5)
__device__ __forceinline__ uint32_t colormap_rgba(float db, const SpecArgs& a) {
    float n = (db - a.conv - a.min_db) * a.inv_range;
    n = fminf(fmaxf(n, 0.0f), 1.0f);           // NaN -> 0 via fmaxf
    float r, g, b;
    if (a.cmap == 1) {                          // Heatmap :944-953
        if (n < 0.2f)      { r = 0.f; g = 0.f; b = 0.f; }
        else if (n < 0.5f) { const float u = (n - 0.2f) / 0.3f; r = u; g = 0.f; b = 1.0f + (0.0f - 1.0f) * u; }
        else               { const float u = (n - 0.5f) / 0.5f; r = 1.f; g = u; b = 0.f; }
    } else { r = g = b = n; }                   // Grayscale :939-942
    const uint32_t R = (uint32_t)floorf(__fmaf_rn(r, 255.0f, 0.5f));
    const uint32_t G = (uint32_t)floorf(__fmaf_rn(g, 255.0f, 0.5f));
    const uint32_t B = (uint32_t)floorf(__fmaf_rn(b, 255.0f, 0.5f));
    return R | (G << 8) | (B << 16) | 0xFF000000u;
}

template <typename T, int N, int DK, bool WIN>
__global__ void __launch_bounds__(Geo<T, N>::CTA, Geo<T, N>::MINB)
spectrogram_kernel(const SpecArgs a) {
    using G = Geo<T, N>;
    constexpr int P = G::P, TPF = G::TPF, FPC = G::FPC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int fl = threadIdx.x / TPF;            // frame slot in this CTA
    const int t  = threadIdx.x % TPF;            // thread within the frame
    cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw) + (size_t)fl * G::SM_ELEMS;
    const cpx<T>* tw = reinterpret_cast<const cpx<T>*>(a.twiddle);

    T win[WIN ? P : 1];
    if constexpr (WIN) {
        const T* w = reinterpret_cast<const T*>(a.window);
#pragma unroll
        for (int q = 0; q < P; q++) win[q] = __ldg(&w[t + TPF * q]);
    }

    const long long n_blocks = (a.n_frames + FPC - 1) / FPC;
    for (long long fb = blockIdx.x; fb < n_blocks; fb += gridDim.x) {
        const long long frame = fb * FPC + fl;
        const long long s0 = a.start_sample + frame * a.hop;            // MainController.java:984
        const bool in_grid = frame < a.n_frames;
        const bool readable = in_grid && (s0 + N <= a.n_samples);       // :987
        cpx<T> v[P];
        if (readable) {
            if (a.lp.swap) {
#pragma unroll
                for (int q = 0; q < P; q++) v[q] = Loader<T, DK>::template load<true>(a.lp, s0 + t + TPF * q);
            } else {
#pragma unroll
                for (int q = 0; q < P; q++) v[q] = Loader<T, DK>::template load<false>(a.lp, s0 + t + TPF * q);
            }
            if constexpr (WIN) {
#pragma unroll
                for (int q = 0; q < P; q++) { v[q].x *= win[q]; v[q].y *= win[q]; }
            }
        } else {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = mk2<T>((T)0, (T)0);
        }

        fft_frame<T, N>(v, t, sm, tw);

        if (!in_grid) continue;
        // out[(k + N/2) % N], SpectralService.java:76-82 ; k = t + TPF*q
        const size_t row = (size_t)frame * N;
        const int k0 = (t + N / 2) & (N - 1);
        if (a.out_kind == OUT_F32_DB) {
            float* o = reinterpret_cast<float*>(a.out) + row;
            if (readable) {
#pragma unroll
                for (int q = 0; q < P; q++) o[(k0 + TPF * q) & (N - 1)] = (float)to_db(v[q].x, v[q].y, a.db_mode);
            } else {
#pragma unroll
                for (int q = 0; q < P; q++) o[(k0 + TPF * q) & (N - 1)] = (float)a.eof_fill;
            }
        } else if (a.out_kind == OUT_F64_DB) {
            double* o = reinterpret_cast<double*>(a.out) + row;
#pragma unroll
            for (int q = 0; q < P; q++)
                o[(k0 + TPF * q) & (N - 1)] = readable ? (double)to_db(v[q].x, v[q].y, a.db_mode) : a.eof_fill;
        } else {
            uint32_t* o = reinterpret_cast<uint32_t*>(a.out) + row;
#pragma unroll
            for (int q = 0; q < P; q++) {
                const float db = readable ? (float)to_db(v[q].x, v[q].y, a.db_mode) : (float)a.eof_fill;
                o[(k0 + TPF * q) & (N - 1)] = colormap_rgba(db, a);
            }
        }
    }
}

// ---- registry ----
struct SpecKernelInfo {
    const void* fn;
    int prec;       // 1 f32, 2 f64 (SA_PREC_*)
    int n;
    int dk;
    int win;
    int cta;
    int fpc;
    int minb;
    size_t smem;
    int p;          // points per thread
    int np;         // passes
    int radix[4];
};

void register_spec_kernel(const SpecKernelInfo& k);   // engine.cu

template <typename T, int N, int DK, bool WIN>
SpecKernelInfo make_spec_info(int prec) {
    using G = Geo<T, N>;
    using PL = Plan<T, N>;
    SpecKernelInfo k;
    k.fn = (const void*)&spectrogram_kernel<T, N, DK, WIN>;
    k.prec = prec; k.n = N; k.dk = DK; k.win = WIN ? 1 : 0;
    k.cta = G::CTA; k.fpc = G::FPC; k.minb = G::MINB; k.smem = G::SMEM_BYTES;
    k.p = G::P; k.np = PL::NP;
    for (int i = 0; i < 4; i++) k.radix[i] = PL::radix(i);
    return k;
}

}  // namespace sa
