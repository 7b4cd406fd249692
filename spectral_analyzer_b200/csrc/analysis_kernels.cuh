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
    unsigned qmagic;                 // ceil(2^32 / down) (0 for down == 1): staged-index / down by umulhi
};

struct DcArgs {
    LoadParams lp;
    long long n_samples;             // IQ pairs readable from lp.base
    const DcAnn* anns;
    const float* taps;               // per D: h[0..8D] natural order, then ht[r*8 + p] = h[D*p + r]
    double* out;                     // planar: re[M] then im[M] per annotation
    int ann_base;                    // first annotation of this launch (launches are grouped by down)
};

// Taps of ONE decimation factor as a kernel parameter: every lane of a warp uses the same 8 taps per
// step, so they are read through the constant bank (uniform datapath) instead of the LSU.
constexpr int kDcParamMaxDown = 256;
struct DcTapParams {
    float h_last;                    // h[8D]
    int   down;
    int   pad_[2];
    float ht[8 * kDcParamMaxDown];   // ht[r*8 + p] = h[D*p + r]
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
// One sample per thread and step (coalesced), kDcUnroll independent loads in flight per thread before the
// first is consumed; the NCO phase of every sample comes from the exact 64-bit accumulator.  Samples before
// the annotation start (n < 0) are the zero history of the causal filter.
constexpr int kDcUnrollMax = 16;      // loads in flight per thread (16-byte cf64 pairs: 8, register budget); C3: 8 -> 16 = +4 %

__device__ __forceinline__ void sts64(uint32_t addr, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

template <int DK, bool SWAP, bool INTERIOR>
__device__ __forceinline__ void dc_stage_tile_impl(const DcArgs& a, const DcAnn& an, float2* __restrict__ stage,
                                                   const long long nlo, const int n_stage, const int pad) {
    using LD = Loader<float, DK>;
    using raw_t = typename LD::raw_t;
    constexpr int kDcUnroll = sizeof(raw_t) <= 8 ? kDcUnrollMax : kDcUnrollMax / 2;
    const raw_t* src = reinterpret_cast<const raw_t*>(a.lp.base) + (an.start_sample + nlo) + threadIdx.x;
    // the phasor advances kDcThreads samples per step: one complex multiply by wstep (packed: P = (c, -s),
    // Q = i P, P' = P.x W + P.y (i W)), re-seeded from the exact 64-bit phase at every batch of kDcUnroll samples
    const float2 wstep = nco_phasor(an.phase_step * (unsigned long long)kDcThreads);
    const pk2 W = pack2(wstep.x, wstep.y), iW = pack2(-wstep.y, wstep.x);
    const unsigned qmagic = pad ? an.qmagic : 0u;                // i / D == umulhi(i, ceil(2^32 / D)) for i*D < 2^32
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    for (int i = threadIdx.x; i < n_stage; i += kDcThreads * kDcUnroll, src += kDcThreads * kDcUnroll) {
        raw_t raw[kDcUnroll];
#pragma unroll
        for (int u = 0; u < kDcUnroll; u++) {
            const int ii = i + u * kDcThreads;
            raw[u] = raw_t();
            if (ii < n_stage && (INTERIOR || nlo + ii >= 0)) raw[u] = __ldg(src + u * kDcThreads);
        }
        const float2 ph0 = nco_phasor(an.phase_step * (unsigned long long)(nlo + i));
        pk2 P = pack2(ph0.x, ph0.y);
#pragma unroll
        for (int u = 0; u < kDcUnroll; u++) {
            const int ii = i + u * kDcThreads;
            const cpx<float> d = LD::template decode<SWAP>(a.lp, raw[u]);
            float px, py;
            unpack2(P, px, py);
            // y = d * P = d.x (P.x, P.y) + d.y (-P.y, P.x)
            pk2 Y = fma2(pack2(d.y, d.y), pack2(-py, px), mul2(pack2(d.x, d.x), P));
            if (!INTERIOR && nlo + ii < 0) Y = pack2(0.f, 0.f);
            float yx, yy;
            unpack2(Y, yx, yy);
            if (ii < n_stage) sts64(stage_s + 8u * (unsigned)(ii + (int)__umulhi((unsigned)ii, qmagic)), yx, yy);
            if (u + 1 < kDcUnroll) P = fma2(pack2(py, py), iW, mul2(pack2(px, px), W));
        }
    }
}

template <int DK>
__device__ __forceinline__ void dc_stage_tile(const DcArgs& a, const DcAnn& an, float2* __restrict__ stage,
                                              const long long nlo, const int n_stage, const int pad) {
    if (nlo >= 0) {
        if (a.lp.swap) dc_stage_tile_impl<DK, true, true>(a, an, stage, nlo, n_stage, pad);
        else           dc_stage_tile_impl<DK, false, true>(a, an, stage, nlo, n_stage, pad);
    } else {
        if (a.lp.swap) dc_stage_tile_impl<DK, true, false>(a, an, stage, nlo, n_stage, pad);
        else           dc_stage_tile_impl<DK, false, false>(a, an, stage, nlo, n_stage, pad);
    }
}

// Staged kernel (down <= kDcMaxDown): one tile = an.nb consecutive outputs of one annotation.
// Samples are decoded and mixed ONCE into shared memory; thread b then forms the 8 polyphase
// partial sums C_p[b] = sum_r h[Dp + r] y[bD - r] of input block b, and
// z[m] = sum_p C_p[m - p] + h[8D] y[(m-8)D].
template <int DK, bool PTAPS>
__global__ void __launch_bounds__(kDcThreads, 4)
downconvert_kernel(const DcArgs a, const __grid_constant__ DcTapParams tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DcAnn an = a.anns[a.ann_base + blockIdx.y];
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

    dc_stage_tile<DK>(a, an, stage, nlo, n_stage, pad);
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
    float2* csm_t = reinterpret_cast<float2*>(csm);               // C_p[b] at csm_t[p * nblk + b]
    for (int b = threadIdx.x; b < nblk; b += kDcThreads) {
        // accumulators packed over p: ax[i] = (Re C_2i, Re C_2i+1), ay[i] likewise (FFMA2: the tap pairs
        // arrive adjacent from the 128-bit loads, the sample is broadcast into both lanes)
        pk2 ax[4], ay[4];
#pragma unroll
        for (int p = 0; p < 4; p++) { ax[p] = pack2(0.f, 0.f); ay[p] = pack2(0.f, 0.f); }
        // block b (local) holds staged indices (b+1)*D - r, r = 0..D-1
        const float2* s_hi = stage + (b + 1) * (D + pad);       // r = 0 sits here (next block's slot 0)
        const float2* s_lo = stage + b * (D + pad) + D;         // r > 0: b*(D+pad) + D - r
        const float4* t4 = reinterpret_cast<const float4*>(ht);
        for (int r = 0; r < D; r++) {
            const float2 s = (r == 0) ? s_hi[0] : s_lo[-r];
            float4 ta, tb;
            if constexpr (PTAPS) {
                ta = make_float4(tp.ht[8 * r], tp.ht[8 * r + 1], tp.ht[8 * r + 2], tp.ht[8 * r + 3]);
                tb = make_float4(tp.ht[8 * r + 4], tp.ht[8 * r + 5], tp.ht[8 * r + 6], tp.ht[8 * r + 7]);
            } else {
                ta = __ldg(&t4[2 * r]); tb = __ldg(&t4[2 * r + 1]);
            }
            const pk2 sx = pack2(s.x, s.x), sy = pack2(s.y, s.y);
            const pk2 t01 = pack2(ta.x, ta.y), t23 = pack2(ta.z, ta.w), t45 = pack2(tb.x, tb.y), t67 = pack2(tb.z, tb.w);
            ax[0] = fma2(t01, sx, ax[0]); ay[0] = fma2(t01, sy, ay[0]);
            ax[1] = fma2(t23, sx, ax[1]); ay[1] = fma2(t23, sy, ay[1]);
            ax[2] = fma2(t45, sx, ax[2]); ay[2] = fma2(t45, sy, ay[2]);
            ax[3] = fma2(t67, sx, ax[3]); ay[3] = fma2(t67, sy, ay[3]);
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
            float x0, x1, y0, y1;
            unpack2(ax[p], x0, x1); unpack2(ay[p], y0, y1);
            csm_t[(2 * p) * nblk + b] = make_float2(x0, y0);
            csm_t[(2 * p + 1) * nblk + b] = make_float2(x1, y1);
        }
    }
    __syncthreads();
    const float h_last = h[8 * D];
    for (int j = threadIdx.x; j < nbt; j += kDcThreads) {
        // output m = m0 + j uses local blocks (j + 7 - p), p = 0..7, and staged sample j*D
        const float2 s = stage[j * (D + pad)];
        float zr = h_last * s.x, zi = h_last * s.y;
#pragma unroll
        for (int p = 0; p < 8; p++) { const float2 c = csm_t[p * nblk + (j + 7 - p)]; zr += c.x; zi += c.y; }
        out_re[j] = (double)zr;
        out_im[j] = (double)zi;
    }
}

// Fallback for very large decimations: one warp per output, lanes stride over the taps.
template <int DK>
__global__ void __launch_bounds__(256)
downconvert_wide_kernel(const DcArgs a) {
    const DcAnn an = a.anns[a.ann_base + blockIdx.y];
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
