// analysis_kernels.cuh -- annotation analysis: NCO mix + FIR decimate, and Welch PSD.
//
// Replaces S/services/ExtractDownConvertService.java:54-117 (decode loop :74-100 + the JDSP
// Resampler calls :106,:111-112) and the JDSP PowerSpectralDensity.calculatePsdWelch call of
// S/controllers/AnalysisDialogController.java:308-312.  JDSP is not vendored in the reference, so
// the arithmetic follows this repository's documented spec (DESIGN.md "downconvert / Welch"):
//   y[n]  = x[n] * exp(-2 pi i f n)                      n = 0 at the first extracted sample
//   conv  : z[m] = sum_{k=0}^{8D} h[k] y[mD - k]         Hamming-windowed sinc, causal, zero history
//   fast  : z[m] = (1/D) sum_{k<D} y[mD + k]             moving average, then decimate
//   M     = count / D
//   Welch : Hann, hop nfft/4, two-sided, mean |FFT|^2 / (fs sum w^2), 10 log10, fft-shifted
#pragma once
#include "decode.cuh"

namespace sa {

struct DcAnn {
    long long start_sample;          // first extracted sample of the capture
    long long count;
    unsigned long long phase_step;   // frac(freq_off) * 2^64 : NCO phase is an exact 64-bit accumulator
    long long out_off;               // offset (in doubles) of this annotation's re block in the output
    long long m_out;                 // count / down
    int down;
    int fast;
    int nb;                          // outputs per tile of the staged kernel
    int taps_off;                    // offset (floats) of this D's tap block in the tap table
};

struct DcArgs {
    LoadParams lp;
    long long n_samples;             // IQ pairs readable from lp.base
    const DcAnn* anns;
    const float* taps;               // per D: h[0..8D] natural order, then ht[r*8 + p] = h[D*p + r]
    double* out;                     // planar: re[M] then im[M] per annotation
};

constexpr int kDcThreads = 256;
constexpr int kDcStage = 8192;       // staged samples per tile (shared memory budget)
constexpr int kDcMaxDown = 512;      // larger decimations take the warp-per-output kernel

// x * exp(-2 pi i phase), phase from the top 32 bits of the 64-bit accumulator
__device__ __forceinline__ float2 nco_mix(float2 x, unsigned long long phase) {
    const float ang = (float)(int)(unsigned)(phase >> 32) * (6.283185307179586f / 4294967296.0f);
    float s, c;
    __sincosf(ang, &s, &c);
    return make_float2(__fmaf_rn(x.x, c, x.y * s), __fmaf_rn(x.y, c, -x.x * s));
}

template <int DK>
__device__ __forceinline__ float2 load_mixed(const DcArgs& a, const DcAnn& an, long long n) {
    // n relative to the annotation start; zero history before it (causal filter)
    if (n < 0) return make_float2(0.f, 0.f);
    const long long g = an.start_sample + n;
    cpx<float> x = a.lp.swap ? Loader<float, DK>::template load<true>(a.lp, g)
                             : Loader<float, DK>::template load<false>(a.lp, g);
    return nco_mix(make_float2(x.x, x.y), an.phase_step * (unsigned long long)n);
}

// e^{-2 pi i phase / 2^64} as (cos, -sin)
__device__ __forceinline__ float2 nco_phasor(unsigned long long phase) {
    const float ang = (float)(int)(unsigned)(phase >> 32) * (6.283185307179586f / 4294967296.0f);
    float s, c;
    __sincosf(ang, &s, &c);
    return make_float2(c, -s);
}
__device__ __forceinline__ float2 cmul(float2 x, float2 p) {
    return make_float2(__fmaf_rn(x.x, p.x, -x.y * p.y), __fmaf_rn(x.x, p.y, x.y * p.x));
}

// Stages the decoded + mixed samples n = nlo .. nlo + n_stage - 1 of one tile: stage[i + (i / D) * pad].
// Samples are fetched as 16-byte groups (2 cf32 / 4 ci16 / 8 cu8 pairs) aligned in GLOBAL memory, kDcUnroll
// groups in flight per thread; the NCO phasor of a group's first sample comes from the exact 64-bit phase and
// advances by one complex multiply per sample inside the group.
constexpr int kDcUnroll = 4;

template <int DK>
__device__ __forceinline__ void dc_stage_tile(const DcArgs& a, const DcAnn& an, float2* __restrict__ stage,
                                              const long long nlo, const int n_stage, const int D, const int pad) {
    constexpr int bps = DK == DK_CF32 ? 8 : DK == DK_CI16 ? 4 : DK == DK_C8 ? 2 : 16;
    constexpr int V = 16 / bps;                                  // samples per 16-byte group
    const char* base = reinterpret_cast<const char*>(a.lp.base);
    const long long g_lo = an.start_sample + nlo;                // global sample of staged index 0 (may be < 0)
    // staged index of the first group is -shift: groups start on 16-byte boundaries of the capture
    const int shift = (int)((((unsigned long long)(uintptr_t)base / bps) + (unsigned long long)(g_lo + (1LL << 40) * V)) % V);
    const int n_groups = (n_stage + shift + V - 1) / V;
    const float2 wstep = nco_phasor(an.phase_step);
    for (int k0 = threadIdx.x; k0 < n_groups; k0 += kDcThreads * kDcUnroll) {
        uint4 raw[kDcUnroll];
        bool vec[kDcUnroll];
#pragma unroll
        for (int u = 0; u < kDcUnroll; u++) {
            const int k = k0 + u * kDcThreads;
            const long long n0 = nlo + (long long)k * V - shift;  // annotation-relative sample of the group
            const long long g0 = an.start_sample + n0;
            vec[u] = (k < n_groups) && (n0 >= 0) && (g0 + V <= a.n_samples);
            raw[u] = make_uint4(0u, 0u, 0u, 0u);
            if (vec[u]) raw[u] = __ldg(reinterpret_cast<const uint4*>(base + g0 * bps));
        }
#pragma unroll
        for (int u = 0; u < kDcUnroll; u++) {
            const int k = k0 + u * kDcThreads;
            if (k >= n_groups) break;
            const int i0 = k * V - shift;
            const long long n0 = nlo + i0;
            int q = (i0 >= 0 ? i0 : 0) / D, rem = (i0 >= 0 ? i0 : 0) - q * D;
            float2 ph = nco_phasor(an.phase_step * (unsigned long long)n0);
            const uint32_t w[4] = { raw[u].x, raw[u].y, raw[u].z, raw[u].w };
#pragma unroll
            for (int j = 0; j < V; j++) {
                const int i = i0 + j;
                float2 x;
                if (vec[u]) {
                    cpx<float> d;
                    if (a.lp.swap) {
                        if constexpr (DK == DK_CF32) d = Loader<float, DK>::template decode<true>(a.lp, make_uint2(w[2 * j], w[2 * j + 1]));
                        else if constexpr (DK == DK_CI16) d = Loader<float, DK>::template decode<true>(a.lp, w[j]);
                        else if constexpr (DK == DK_C8) d = Loader<float, DK>::template decode<true>(a.lp, (uint16_t)(w[j / 2] >> (16 * (j & 1))));
                        else d = Loader<float, DK>::template decode<true>(a.lp, raw[u]);
                    } else {
                        if constexpr (DK == DK_CF32) d = Loader<float, DK>::template decode<false>(a.lp, make_uint2(w[2 * j], w[2 * j + 1]));
                        else if constexpr (DK == DK_CI16) d = Loader<float, DK>::template decode<false>(a.lp, w[j]);
                        else if constexpr (DK == DK_C8) d = Loader<float, DK>::template decode<false>(a.lp, (uint16_t)(w[j / 2] >> (16 * (j & 1))));
                        else d = Loader<float, DK>::template decode<false>(a.lp, raw[u]);
                    }
                    x = cmul(make_float2(d.x, d.y), ph);
                } else {
                    // group touching the zero history or the end of the capture: per-sample path
                    x = (i >= 0 && i < n_stage) ? load_mixed<DK>(a, an, n0 + j) : make_float2(0.f, 0.f);
                }
                if (i >= 0 && i < n_stage) stage[i + q * pad] = x;
                if (i >= 0) { if (++rem == D) { rem = 0; q++; } }
                if (j + 1 < V) ph = cmul(ph, wstep);
            }
        }
    }
}

// Staged kernel (down <= kDcMaxDown): one tile = an.nb consecutive outputs of one annotation.
// Samples are decoded and mixed ONCE into shared memory; thread b then forms the 8 polyphase
// partial sums C_p[b] = sum_r h[Dp + r] y[bD - r] of input block b, and
// z[m] = sum_p C_p[m - p] + h[8D] y[(m-8)D].
template <int DK>
__global__ void __launch_bounds__(kDcThreads)
downconvert_kernel(const DcArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DcAnn an = a.anns[blockIdx.y];
    const int D = an.down;
    const long long m0 = (long long)blockIdx.x * an.nb;
    if (an.nb == 0 || m0 >= an.m_out) return;      // nb == 0: handled by downconvert_wide_kernel
    const int nbt = (int)min((long long)an.nb, an.m_out - m0);
    const int pad = (D & 1) ? 0 : 1;                 // odd stride between blocks: conflict-free LDS.64
    float2* stage = reinterpret_cast<float2*>(smem_raw);
    const int halo = an.fast ? 0 : 7;
    const int nblk = nbt + halo;
    const int n_stage = an.fast ? nbt * D : nblk * D + 1;
    const int stage_phys = n_stage + n_stage / D + 2;
    float4* csm = reinterpret_cast<float4*>(stage + ((stage_phys + 1) & ~1));   // C_p[b]: 8 float2 per block
    const float* h = a.taps + an.taps_off;
    const float* ht = h + 8 * D + 1;
    const long long nlo = an.fast ? m0 * D : (m0 - 8) * D;

    dc_stage_tile<DK>(a, an, stage, nlo, n_stage, D, pad);
    __syncthreads();
    double* out_re = a.out + an.out_off + m0;
    double* out_im = out_re + an.m_out;
    if (an.fast) {
        for (int j = threadIdx.x; j < nbt; j += kDcThreads) {
            const float2* s = stage + j * (D + pad);
            float sr = 0.f, si = 0.f;
            for (int k = 0; k < D; k++) { sr += s[k].x; si += s[k].y; }
            const float inv = 1.0f / (float)D;
            out_re[j] = (double)(sr * inv);
            out_im[j] = (double)(si * inv);
        }
        return;
    }
    for (int b = threadIdx.x; b < nblk; b += kDcThreads) {
        float2 acc[8];
#pragma unroll
        for (int p = 0; p < 8; p++) acc[p] = make_float2(0.f, 0.f);
        // block b (local) holds staged indices (b+1)*D - r, r = 0..D-1
        const float2* s_hi = stage + (b + 1) * (D + pad);       // r = 0 sits here (next block's slot 0)
        const float2* s_lo = stage + b * (D + pad) + D;         // r > 0: b*(D+pad) + D - r
        const float4* t4 = reinterpret_cast<const float4*>(ht);
        for (int r = 0; r < D; r++) {
            const float2 s = (r == 0) ? s_hi[0] : s_lo[-r];
            const float4 ta = t4[2 * r], tb = t4[2 * r + 1];
            acc[0].x = __fmaf_rn(ta.x, s.x, acc[0].x); acc[0].y = __fmaf_rn(ta.x, s.y, acc[0].y);
            acc[1].x = __fmaf_rn(ta.y, s.x, acc[1].x); acc[1].y = __fmaf_rn(ta.y, s.y, acc[1].y);
            acc[2].x = __fmaf_rn(ta.z, s.x, acc[2].x); acc[2].y = __fmaf_rn(ta.z, s.y, acc[2].y);
            acc[3].x = __fmaf_rn(ta.w, s.x, acc[3].x); acc[3].y = __fmaf_rn(ta.w, s.y, acc[3].y);
            acc[4].x = __fmaf_rn(tb.x, s.x, acc[4].x); acc[4].y = __fmaf_rn(tb.x, s.y, acc[4].y);
            acc[5].x = __fmaf_rn(tb.y, s.x, acc[5].x); acc[5].y = __fmaf_rn(tb.y, s.y, acc[5].y);
            acc[6].x = __fmaf_rn(tb.z, s.x, acc[6].x); acc[6].y = __fmaf_rn(tb.z, s.y, acc[6].y);
            acc[7].x = __fmaf_rn(tb.w, s.x, acc[7].x); acc[7].y = __fmaf_rn(tb.w, s.y, acc[7].y);
        }
        float4* c = csm + 4 * b;
        c[0] = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
        c[1] = make_float4(acc[2].x, acc[2].y, acc[3].x, acc[3].y);
        c[2] = make_float4(acc[4].x, acc[4].y, acc[5].x, acc[5].y);
        c[3] = make_float4(acc[6].x, acc[6].y, acc[7].x, acc[7].y);
    }
    __syncthreads();
    const float2* c2 = reinterpret_cast<const float2*>(csm);
    const float h_last = h[8 * D];
    for (int j = threadIdx.x; j < nbt; j += kDcThreads) {
        // output m = m0 + j uses local blocks (j + 7 - p), p = 0..7, and staged sample j*D
        const float2 s = stage[j * (D + pad)];
        float zr = h_last * s.x, zi = h_last * s.y;
#pragma unroll
        for (int p = 0; p < 8; p++) { const float2 c = c2[(j + 7 - p) * 8 + p]; zr += c.x; zi += c.y; }
        out_re[j] = (double)zr;
        out_im[j] = (double)zi;
    }
}

// Fallback for very large decimations: one warp per output, lanes stride over the taps.
template <int DK>
__global__ void __launch_bounds__(256)
downconvert_wide_kernel(const DcArgs a) {
    const DcAnn an = a.anns[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (an.nb != 0 || m >= an.m_out) return;
    const int D = an.down;
    const float* h = a.taps + an.taps_off;
    float zr = 0.f, zi = 0.f;
    if (an.fast) {
        for (int k = lane; k < D; k += 32) { const float2 y = load_mixed<DK>(a, an, m * D + k); zr += y.x; zi += y.y; }
        zr /= (float)D; zi /= (float)D;
    } else {
        for (int k = lane; k <= 8 * D; k += 32) {
            const float2 y = load_mixed<DK>(a, an, m * D - k);
            zr = __fmaf_rn(h[k], y.x, zr); zi = __fmaf_rn(h[k], y.y, zi);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { zr += __shfl_xor_sync(0xffffffffu, zr, o); zi += __shfl_xor_sync(0xffffffffu, zi, o); }
    if (lane == 0) {
        a.out[an.out_off + m] = (double)zr;
        a.out[an.out_off + an.m_out + m] = (double)zi;
    }
}

// ---------------- Welch ----------------
struct WelchSig {
    const double* re;      // planar FP64 input (the downconverter's output rows)
    const double* im;
    long long n;           // samples
    double scale;          // 1 / (nseg * fs * sum w^2)
    long long nseg;
};

struct WelchArgs {
    const WelchSig* sigs;
    long long hop;
    const float* window;
    const void* twiddle;
    float* partial;        // [sig][slot][N], slot = split * FPC + frame slot
    int nsplit;
    double* out_db;        // [sig][N], fft-shifted
};

template <int N>
__global__ void __launch_bounds__(Geo<float, N>::CTA, Geo<float, N>::MINB)
welch_accum_kernel(const WelchArgs a) {
    using G = Geo<float, N>;
    constexpr int P = G::P, TPF = G::TPF, FPC = G::FPC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int fl = threadIdx.x / TPF, t = threadIdx.x % TPF;
    float2* sm = reinterpret_cast<float2*>(smem_raw) + (size_t)fl * G::SM_ELEMS;
    const float2* tw = reinterpret_cast<const float2*>(a.twiddle);
    if constexpr (G::TW_SMEM) {      // one-warp-per-frame plans read their twiddles from shared memory
        float2* tsm = reinterpret_cast<float2*>(smem_raw + G::SMEM_BYTES);
        for (int i = threadIdx.x; i < (int)(G::TW_BYTES / sizeof(float2)); i += G::CTA) tsm[i] = __ldg(&tw[i]);
        __syncthreads();
        tw = tsm;
    }
    const TwSeed<float> seed = load_tw_seed<float, N>(reinterpret_cast<const float2*>(a.twiddle), t);
    const WelchSig sg = a.sigs[blockIdx.y];
    float acc[P];
#pragma unroll
    for (int q = 0; q < P; q++) acc[q] = 0.f;
    const long long stride = (long long)a.nsplit * FPC;
    const long long iters = (sg.nseg + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long seg = it * stride + (long long)blockIdx.x * FPC + fl;
        const bool valid = seg < sg.nseg;
        float2 v[P];
        if (valid) {
            const double* re = sg.re + seg * a.hop + t;
            const double* im = sg.im + seg * a.hop + t;
#pragma unroll
            for (int q = 0; q < P; q++) {
                const float w = __ldg(&a.window[t + TPF * q]);
                v[q] = make_float2((float)__ldg(&re[TPF * q]) * w, (float)__ldg(&im[TPF * q]) * w);
            }
        } else {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = make_float2(0.f, 0.f);
        }
        fft_frame<float, N, false>(v, t, sm, tw, nullptr, seed);
        if (valid) {
#pragma unroll
            for (int q = 0; q < P; q++) acc[q] += __fmaf_rn(v[q].x, v[q].x, v[q].y * v[q].y);
        }
    }
    float* part = a.partial + (((size_t)blockIdx.y * a.nsplit + blockIdx.x) * FPC + fl) * N;
#pragma unroll
    for (int q = 0; q < P; q++) part[t + TPF * q] = acc[q];
}

// sums the partial spectra in a fixed order (deterministic), scales to density, dB, fft-shift
__global__ void welch_finalize_kernel(const WelchArgs a, const int n, const int slots) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const WelchSig sg = a.sigs[blockIdx.y];
    const float* part = a.partial + (size_t)blockIdx.y * slots * n + k;
    double s = 0.0;
    for (int i = 0; i < slots; i++) s += (double)part[(size_t)i * n];
    const double v = sg.nseg > 0 ? 10.0 * log10(s * sg.scale + 1e-30) : __longlong_as_double(0x7ff8000000000000LL);
    a.out_db[(size_t)blockIdx.y * n + ((k + n / 2) & (n - 1))] = v;
}

}  // namespace sa
