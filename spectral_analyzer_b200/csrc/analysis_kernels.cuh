// analysis_kernels.cuh -- annotation analysis: NCO mix + FIR decimate, and Welch PSD.
//
// Replaces S/services/ExtractDownConvertService.java:54-117 (decode loop :74-100 + the JDSP
// Resampler calls :106,:111-112) and the JDSP PowerSpectralDensity.calculatePsdWelch call of
// S/controllers/AnalysisDialogController.java:308-312.  JDSP is not vendored in the reference, so
// every choice it makes is a parameter of the engine's analysis profile (sa_analysis_config,
// include/sa_engine.h); the defaults are this repository's documented spec (DESIGN.md "downconvert / Welch"):
//   y[n]  = x[n] * exp(-2 pi i f n)                      n = 0 at the first extracted sample, y = 0 outside [0, count)
//   conv  : z[m] = sum_{k<L} h[k] y[mD + off - k]        default h: Hamming-windowed sinc, L = 8D+1; off = 0 (causal),
//                                                        (L-1)/2 (same) or L-1 (valid)
//   fast  : z[m] = (1/D) sum_{k<D} y[mD + k]             moving average, then decimate
//   M     = floor(count / D) | ceil(count / D) | (count - L)/D + 1 (valid)
//   Welch : window, hop, optional per-segment mean removal, two-sided, mean |FFT|^2 scaled 1/(fs sum w^2) (density) or
//           1/(sum w)^2 (spectrum), 10 log10, fft-shifted; FP32 or FP64 transforms; any nfft (powers of two through
//           the Stockham kernels, everything else through the direct DFT kernel: the reference's short-signal branch,
//           AnalysisDialogController.java:304-307, passes nfft = signal length)
#pragma once
#include "decode.cuh"
#include "spectrogram_mid_kernel.cuh"

namespace sa {

struct DcAnn {
    long long start_sample;          // first extracted sample of the capture
    long long count;
    unsigned long long phase_step;   // frac(freq_off) * 2^64 : NCO phase is an exact 64-bit accumulator
    long long out_off;               // offset (in doubles) of this annotation's re block in the FP64 output
    long long scr_off;               // offset (in float2) of this annotation in the FP32 scratch rows
    long long m_out;                 // outputs
    long long in_off;                // delay shift `off` above
    int down;
    int fast;
    int nb;                          // outputs per tile of the staged kernel (0: warp-per-output kernel)
    int taps_off;                    // offset (floats) of this annotation's tap block in the tap table
    int n_taps;                      // L (warp-per-output kernel)
    unsigned qmagic;                 // ceil(2^32 / down) (0 for down == 1): staged-index / down by umulhi
};

struct DcArgs {
    LoadParams lp;
    long long n_samples;             // IQ pairs readable from lp.base
    const DcAnn* anns;
    const float* taps;               // per block: h[0..8D] natural order (zero padded), then ht[r*8 + p] = h[D*p + r]
    double* out;                     // planar: re[M] then im[M] per annotation (the Java double[2][M]); may be NULL
    float2* scratch;                 // interleaved FP32 rows for the Welch kernel (stays in L2); may be NULL
    int ann_base;                    // first annotation of this launch (launches are grouped by down)
    int tiles_per_cta;
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
// Shape of the pipelined downconverter per input type (measured on C3, 500 x 2^20 samples, D 16, and over D = 2..32 with
// tools/dc_matrix.py, same box):
//   3 CTAs/SM, 80 registers, tap loop as written                      cf32 1.661 ms
//   + next tap row's loads issued under this row's FMAs (SWP)          cf32 1.712 ms  -- needs > 85 registers: spills
//   SWP at 2 CTAs/SM (127 registers, no spills)                        cf32 1.618 ms (1.5-4 % faster at every D);
//                                                                      ci16 / cu8 2-6 % SLOWER at every D
// cf32 holds 34 registers of raw samples in flight per thread, the integer types 17: cf32 takes the 2-CTA software-pipelined
// shape, the others stay at 3 CTAs.  -DSA_DC_SWP=0|1 / -DSA_DC_PIPE_CTAS=n force one shape for all types.
#ifndef SA_DC_PIPE_CTAS
#define SA_DC_PIPE_CTAS 0       // 0: per type
#endif
#ifndef SA_DC_SWP
#define SA_DC_SWP -1            // -1: per type
#endif
template <int DK> __host__ __device__ constexpr bool dc_swp() { return SA_DC_SWP < 0 ? DK == DK_CF32 : SA_DC_SWP != 0; }
template <int DK> __host__ __device__ constexpr int dc_pipe_ctas() { return SA_DC_PIPE_CTAS > 0 ? SA_DC_PIPE_CTAS : (DK == DK_CF32 ? 2 : 3); }
constexpr int kDcStage = 8192;       // staged samples per tile (shared memory budget)
constexpr int kDcMaxDown = 512;      // larger decimations take the warp-per-output kernel
constexpr int kDcPipeLoads = 17;     // pipelined variant: the whole tile (<= 17 x 256 samples) is one register batch
constexpr int kDcPipeTiles = 8;      // consecutive tiles of one annotation per CTA (pipelined variant)
__host__ __device__ constexpr int dc_tap_smem_bytes(int down) { return (8 * down * (int)sizeof(float) + 15) & ~15; }   // pipelined variant

// The reference's cf64 stride bug, reproduced only with strict_reference (ExtractDownConvertService.java:60-67,79-81):
// IQ pair i is read as the two doubles at byte offset 8 i (re = d[i], im = d[i+1]).
constexpr int DK_CF64_S8 = 4;
template <typename T> struct Loader<T, DK_CF64_S8> : Loader<T, DK_CF64> {};

template <int DK> struct DcRaw {
    using LD = Loader<float, DK>;
    using raw_t = typename LD::raw_t;
    static __device__ __forceinline__ raw_t ld(const void* base, long long i) { return __ldg(reinterpret_cast<const raw_t*>(base) + i); }
};
template <> struct DcRaw<DK_CF64_S8> {
    using LD = Loader<float, DK_CF64>;
    using raw_t = uint4;
    static __device__ __forceinline__ raw_t ld(const void* base, long long i) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(base) + i), b = __ldg(reinterpret_cast<const uint2*>(base) + i + 1);
        return make_uint4(a.x, a.y, b.x, b.y);
    }
};

// x * exp(-2 pi i phase), phase from the top 32 bits of the 64-bit accumulator
__device__ __forceinline__ float2 nco_mix(float2 x, unsigned long long phase) {
    const float ang = (float)(int)(unsigned)(phase >> 32) * (6.283185307179586f / 4294967296.0f);
    float s, c;
    __sincosf(ang, &s, &c);
    return make_float2(__fmaf_rn(x.x, c, x.y * s), __fmaf_rn(x.y, c, -x.x * s));
}

template <int DK>
__device__ __forceinline__ float2 load_mixed(const DcArgs& a, const DcAnn& an, long long n) {
    // n relative to the annotation start; zero outside [0, count)
    if (n < 0 || n >= an.count) return make_float2(0.f, 0.f);
    using R = DcRaw<DK>;
    const typename R::raw_t raw = R::ld(a.lp.base, an.start_sample + n);
    const cpx<float> x = a.lp.swap ? R::LD::template decode<true>(a.lp, raw) : R::LD::template decode<false>(a.lp, raw);
    return nco_mix(make_float2(x.x, x.y), an.phase_step * (unsigned long long)n);
}

// e^{-2 pi i phase / 2^64} as (cos, -sin)
__device__ __forceinline__ float2 nco_phasor(unsigned long long phase) {
    const float ang = (float)(int)(unsigned)(phase >> 32) * (6.283185307179586f / 4294967296.0f);
    float s, c;
    __sincosf(ang, &s, &c);
    return make_float2(c, -s);
}

// the same evaluated in FP64 from the full 64-bit phase and rounded once: for phasors that are multiplied up in a
// recurrence, where the 4e-7 absolute error of the fast sincos (and of a 24-bit angle) would grow with the step count
__device__ __forceinline__ float2 nco_phasor_acc(unsigned long long phase) {
    double s, c;
    sincospi((double)(long long)phase * (1.0 / 9223372036854775808.0), &s, &c);
    return make_float2((float)c, (float)-s);
}

__device__ __forceinline__ float2 lds64(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}

// One register batch of a tile: U coalesced loads per thread (sample i0 + tid + 256 u of the tile), all issued before
// the first is consumed.  i0 starts at -qoff (below), so that a half-warp's 16 samples never straddle a pad element
// when D is a multiple of 16 (conflict-free staging stores); the lanes left of sample 0 idle.  Samples outside the annotation ([lo, hi) in tile coordinates) are not read.
// INTERIOR (tile-uniform): every staged sample of the tile lies inside the annotation: one range test (tile end) per
// sample instead of two, and no zeroing select in the staging step.
template <int DK, int U, bool INTERIOR>
__device__ __forceinline__ void dc_load_batch_impl(const DcArgs& a, const DcAnn& an, typename DcRaw<DK>::raw_t (&raw)[U],
                                                   const long long nlo, const int i0, const int lo, const int hi) {
    using R = DcRaw<DK>;
    const long long s = an.start_sample + nlo + i0 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int ii = i0 + (int)threadIdx.x + u * kDcThreads;
        raw[u] = typename R::raw_t();      // (skipping this for interior tiles measured SLOWER: C3 1.66 -> 1.74 ms on one box)
        if (INTERIOR ? ((unsigned)ii < (unsigned)hi) : (ii >= lo && ii < hi)) raw[u] = R::ld(a.lp.base, s + u * kDcThreads);
    }
}
template <int DK, int U>
__device__ __forceinline__ void dc_load_batch(const DcArgs& a, const DcAnn& an, typename DcRaw<DK>::raw_t (&raw)[U],
                                              const long long nlo, const int i0, const int n_stage, const int lo, const int hi) {
    if (lo == 0 && hi == n_stage) dc_load_batch_impl<DK, U, true>(a, an, raw, nlo, i0, lo, hi);
    else dc_load_batch_impl<DK, U, false>(a, an, raw, nlo, i0, lo, hi);
}

// Decodes and mixes a register batch into the staged tile: stage[i + ((i + qoff) / D) * pad].  qoff = D - 1 (FIR tiles)
// puts the pad element of an even D AFTER the samples i = 0 mod D, so the D samples y[(b+1)D - r], r = 0..D-1, that
// input block b multiplies are contiguous (one address stream, no first-row special case in the tap loop) while
// y[jD] stays at j (D + pad); qoff = 0 (box-car tiles) keeps y[jD .. jD + D) contiguous.  The NCO phasor of the batch
// is seeded from the exact 64-bit phase and advances 256 samples per step by one packed complex multiply.
template <int DK, bool SWAP, int U, bool INTERIOR>
__device__ __forceinline__ void dc_stage_batch_impl(const DcArgs& a, const DcAnn& an, const typename DcRaw<DK>::raw_t (&raw)[U],
                                                    float2* __restrict__ stage, const long long nlo, const int i0,
                                                    const int n_stage, const int lo, const int hi, const int pad, const int qoff) {
    using LD = typename DcRaw<DK>::LD;
    const float2 wstep = nco_phasor(an.phase_step * (unsigned long long)kDcThreads);
    const pk2 W = pack2(wstep.x, wstep.y), iW = pack2(-wstep.y, wstep.x);
    const unsigned qmagic = pad ? an.qmagic : 0u;                // i / D == umulhi(i, ceil(2^32 / D)) for i*D < 2^32
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const float2 ph0 = nco_phasor(an.phase_step * (unsigned long long)(nlo + i0 + (long long)threadIdx.x));
    pk2 P = pack2(ph0.x, ph0.y);
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int ii = i0 + (int)threadIdx.x + u * kDcThreads;
        const cpx<float> d = LD::template decode<SWAP>(a.lp, raw[u]);
        float px, py;
        unpack2(P, px, py);
        // y = d * P = d.x (P.x, P.y) + d.y (-P.y, P.x)
        pk2 Y = fma2(pack2(d.y, d.y), pack2(-py, px), mul2(pack2(d.x, d.x), P));
        if (!INTERIOR && (ii < lo || ii >= hi)) Y = pack2(0.f, 0.f);             // zero history / zero tail of the filter
        float yx, yy;
        unpack2(Y, yx, yy);
        if ((unsigned)ii < (unsigned)n_stage) sts64(stage_s + 8u * (unsigned)(ii + (int)__umulhi((unsigned)(ii + qoff), qmagic)), yx, yy);
        if (u + 1 < U) P = fma2(pack2(py, py), iW, mul2(pack2(px, px), W));
    }
}
template <int DK, bool SWAP, int U>
__device__ __forceinline__ void dc_stage_batch(const DcArgs& a, const DcAnn& an, const typename DcRaw<DK>::raw_t (&raw)[U],
                                               float2* __restrict__ stage, const long long nlo, const int i0,
                                               const int n_stage, const int lo, const int hi, const int pad, const int qoff) {
    if (lo == 0 && hi == n_stage)
        dc_stage_batch_impl<DK, SWAP, U, true>(a, an, raw, stage, nlo, i0, n_stage, lo, hi, pad, qoff);
    else
        dc_stage_batch_impl<DK, SWAP, U, false>(a, an, raw, stage, nlo, i0, n_stage, lo, hi, pad, qoff);
}

constexpr int kDcUnrollMax = 16;      // loads in flight per thread of the non-pipelined variant (16-byte pairs: 8)

// Staged kernel (down <= kDcMaxDown): one tile = an.nb consecutive outputs of one annotation.
// Samples are decoded and mixed ONCE into shared memory; thread b then forms the 8 polyphase
// partial sums C_p[b] = sum_r h[Dp + r] y[bD - r] of input block b, and
// z[m] = sum_p C_p[m - p] + h[8D] y[(m-8)D]   (tile-local indices; `in_off` shifts the whole tile).
// PIPE: a CTA walks kDcPipeTiles consecutive tiles of its annotation; the raw samples of tile i+1 are loaded
// into registers right after tile i has been staged and travel under tile i's FIR, combine and stores.
// (Ablation, removed: a compile-time decimation factor with the loop over r unrolled.  FFMA2 takes its scalar operand
// from a vector register only, so the taps still travel LDCU -> MOV -> register pair: 19 instructions per tap row
// against 13 for the run-time loop, whose LDC.64 lands the tap pairs in vector registers directly.)
template <int DK, bool PTAPS, bool PIPE>
__global__ void __launch_bounds__(kDcThreads, PIPE ? dc_pipe_ctas<DK>() : 4)
downconvert_kernel(const DcArgs a, const __grid_constant__ DcTapParams tp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using raw_t = typename DcRaw<DK>::raw_t;
    constexpr int U = PIPE ? kDcPipeLoads : (sizeof(raw_t) <= 8 ? kDcUnrollMax : kDcUnrollMax / 2);
    const DcAnn an = a.anns[a.ann_base + blockIdx.y];
    const int D = an.down;
    if (an.nb == 0) return;                           // handled by downconvert_wide_kernel
    const long long n_tiles = (an.m_out + an.nb - 1) / an.nb;
    long long tile = (long long)blockIdx.x * a.tiles_per_cta;
    const long long tile_end = min(n_tiles, tile + a.tiles_per_cta);
    if (tile >= tile_end) return;
    const int pad = (D & 1) ? 0 : 1;                  // odd stride between blocks: conflict-free LDS.64
    const int qoff = an.fast ? 0 : D - 1;             // where the pad element sits (dc_stage_batch)
    const int halo = an.fast ? 0 : 7;
    const float* h = a.taps + an.taps_off;
    const float* ht = h + 8 * D + 1;
    const float h_last = PTAPS ? tp.h_last : h[8 * D];
    // PIPE: the transposed tap rows sit in front of the staged tile (two broadcast LDS.128 per tap row instead of four
    // LDC.64 from the parameter bank); they become visible with the barrier that follows the first staging step
    const int tap_bytes = PIPE ? dc_tap_smem_bytes(D) : 0;
    const uint32_t taps_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
    if constexpr (PIPE) {
        float* tsm = reinterpret_cast<float*>(smem_raw);
        for (int i = threadIdx.x; i < 8 * D; i += kDcThreads) tsm[i] = __ldg(&ht[i]);
    }
    float2* stage = reinterpret_cast<float2*>(smem_raw + tap_bytes);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);

    // geometry of a tile (the last one of an annotation may be short)
    auto geom = [&](long long tl, int& nbt, int& nblk, int& n_stage, long long& nlo, int& lo, int& hi) {
        const long long m0 = tl * an.nb;
        nbt = (int)min((long long)an.nb, an.m_out - m0);
        nblk = nbt + halo;
        n_stage = an.fast ? nbt * D : nblk * D + 1;
        nlo = an.fast ? m0 * D : (m0 - 8) * D + an.in_off;
        lo = (int)max(0LL, min((long long)n_stage, -nlo));
        hi = (int)max(0LL, min((long long)n_stage, an.count - nlo));
    };
    int nbt, nblk, n_stage, lo, hi;
    long long nlo;
    geom(tile, nbt, nblk, n_stage, nlo, lo, hi);
    const int nst0 = n_stage, phys0 = n_stage + n_stage / D + 2;
    raw_t raw[U];
    if constexpr (PIPE) dc_load_batch<DK, U>(a, an, raw, nlo, -qoff, n_stage, lo, hi);
    for (; tile < tile_end; tile++) {
        const long long m0 = tile * an.nb;
        // ---- stage: decode + mix once into shared memory
        if constexpr (PIPE) {
            if (a.lp.swap) dc_stage_batch<DK, true, U>(a, an, raw, stage, nlo, -qoff, n_stage, lo, hi, pad, qoff);
            else           dc_stage_batch<DK, false, U>(a, an, raw, stage, nlo, -qoff, n_stage, lo, hi, pad, qoff);
        } else {
            for (int i0 = -qoff; i0 < n_stage; i0 += kDcThreads * U) {
                dc_load_batch<DK, U>(a, an, raw, nlo, i0, n_stage, lo, hi);
                if (a.lp.swap) dc_stage_batch<DK, true, U>(a, an, raw, stage, nlo, i0, n_stage, lo, hi, pad, qoff);
                else           dc_stage_batch<DK, false, U>(a, an, raw, stage, nlo, i0, n_stage, lo, hi, pad, qoff);
            }
        }
        __syncthreads();
        const int c_nbt = nbt, c_nblk = nblk, c_nstage = n_stage;
        if constexpr (PIPE) {           // next tile's samples fly under this tile's arithmetic
            if (tile + 1 < tile_end) {
                geom(tile + 1, nbt, nblk, n_stage, nlo, lo, hi);
                dc_load_batch<DK, U>(a, an, raw, nlo, -qoff, n_stage, lo, hi);
            }
        }
        const int stage_phys = (c_nstage == nst0) ? phys0 : c_nstage + c_nstage / D + 2;      // one division per CTA, not per tile
        float2* csm_t = stage + ((stage_phys + 1) & ~1);               // C_p[b] at csm_t[p * nblk + b]
        double* out_re = a.out ? a.out + an.out_off + m0 : nullptr;
        double* out_im = a.out ? out_re + an.m_out : nullptr;
        float2* scr = a.scratch ? a.scratch + an.scr_off + m0 : nullptr;
        if (an.fast) {
            for (int j = threadIdx.x; j < c_nbt; j += kDcThreads) {
                const float2* s = stage + j * (D + pad);
                float sr = 0.f, si = 0.f;
                for (int k = 0; k < D; k++) { sr += s[k].x; si += s[k].y; }
                const float inv = 1.0f / (float)D;
                if (out_re) { out_re[j] = (double)(sr * inv); out_im[j] = (double)(si * inv); }
                if (scr) scr[j] = make_float2(sr * inv, si * inv);
            }
        } else {
            for (int b = threadIdx.x; b < c_nblk; b += kDcThreads) {
                // accumulators packed over p: ax[i] = (Re C_2i, Re C_2i+1), ay[i] likewise (FFMA2: the tap pairs
                // arrive adjacent from the 128-bit loads, the sample is broadcast into both lanes)
                pk2 ax[4], ay[4];
#pragma unroll
                for (int p = 0; p < 4; p++) { ax[p] = pack2(0.f, 0.f); ay[p] = pack2(0.f, 0.f); }
                // block b (local) multiplies y[(b+1)D - r], r = 0..D-1: staged at (b+1)(D+pad) - r
                const uint32_t s_top = stage_s + 8u * (unsigned)((b + 1) * (D + pad));
                const float4* t4 = reinterpret_cast<const float4*>(ht);
                // (unrolling this loop by 4 to amortise its uniform-datapath control, ~4 of 52 instructions per input sample,
                // measured slower: C3 1.66 -> 1.74 ms)
                float2 s_n;
                float4 ta_n, tb_n;
                if constexpr (PIPE && dc_swp<DK>()) { s_n = lds64(s_top); ta_n = lds128(taps_s); tb_n = lds128(taps_s + 16u); }
                for (int r = 0; r < D; r++) {
                    float2 s;
                    float4 ta, tb;
                    if constexpr (PIPE && dc_swp<DK>()) {
                        // software pipeline: row r + 1 is in flight under row r's arithmetic (the read past the last row
                        // lands in the block below / the first staged sample: valid shared memory, value unused)
                        s = s_n; ta = ta_n; tb = tb_n;
                        s_n = lds64(s_top - 8u * (unsigned)(r + 1));
                        ta_n = lds128(taps_s + 32u * (unsigned)(r + 1)); tb_n = lds128(taps_s + 32u * (unsigned)(r + 1) + 16u);
                    } else {
                        s = lds64(s_top - 8u * (unsigned)r);
                        if constexpr (PIPE) {
                            ta = lds128(taps_s + 32u * (unsigned)r); tb = lds128(taps_s + 32u * (unsigned)r + 16u);
                        } else if constexpr (PTAPS) {
                            ta = make_float4(tp.ht[8 * r], tp.ht[8 * r + 1], tp.ht[8 * r + 2], tp.ht[8 * r + 3]);
                            tb = make_float4(tp.ht[8 * r + 4], tp.ht[8 * r + 5], tp.ht[8 * r + 6], tp.ht[8 * r + 7]);
                        } else {
                            ta = __ldg(&t4[2 * r]); tb = __ldg(&t4[2 * r + 1]);
                        }
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
                    csm_t[(2 * p) * c_nblk + b] = make_float2(x0, y0);
                    csm_t[(2 * p + 1) * c_nblk + b] = make_float2(x1, y1);
                }
            }
            __syncthreads();
            for (int j = threadIdx.x; j < c_nbt; j += kDcThreads) {
                // output m = m0 + j uses local blocks (j + 7 - p), p = 0..7, and staged sample j*D
                const float2 s = stage[j * (D + pad)];
                float zr = h_last * s.x, zi = h_last * s.y;
#pragma unroll
                for (int p = 0; p < 8; p++) { const float2 c = csm_t[p * c_nblk + (j + 7 - p)]; zr += c.x; zi += c.y; }
                if (out_re) { out_re[j] = (double)zr; out_im[j] = (double)zi; }
                if (scr) scr[j] = make_float2(zr, zi);
            }
        }
        __syncthreads();                 // stage / C_p are rewritten by the next tile
    }
}

// Fallback for very large decimations or tap counts above 8D+1: one warp per output, lanes stride over the taps.
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
        for (int k = lane; k < an.n_taps; k += 32) {
            const float2 y = load_mixed<DK>(a, an, m * D + an.in_off - k);
            zr = __fmaf_rn(h[k], y.x, zr); zi = __fmaf_rn(h[k], y.y, zi);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { zr += __shfl_xor_sync(0xffffffffu, zr, o); zi += __shfl_xor_sync(0xffffffffu, zi, o); }
    if (lane == 0) {
        if (a.out) {
            a.out[an.out_off + m] = (double)zr;
            a.out[an.out_off + an.m_out + m] = (double)zi;
        }
        if (a.scratch) a.scratch[an.scr_off + m] = make_float2(zr, zi);
    }
}

// ---------------- row-per-thread downconverter (even decimation 4..32, rows on the 16-byte grid) ----------------
// The staged kernel above spends two thirds of its instructions moving samples (register batch -> decode -> NCO
// recurrence -> padded shared-memory store -> LDS.64 per tap row).  Here the RAW tile (NT rows of D samples) goes
// global -> shared memory with 16-byte cp.async (no registers in flight, no per-sample address arithmetic: a thread's
// chunks differ by immediates), and thread b owns ROW b = D consecutive raw samples, which it reads back as 128-bit loads
// (rows an odd number of chunks apart, or packed and XOR-swizzled: the 8 lanes of a quarter-warp hit 8 different 16-byte
// bank groups).  Row b is mixed with the NCO and filtered into the 8 polyphase partial sums
//     S_p[b] = sum_i h[(p+1)D - i - sg] y[b][i],      z[m] = sum_p S_p[row(m) - p] + lone tap,
// which go through shared memory once (DESIGN.md K5 has the history of the two variants below and their ncu records):
//   MODE 0 (ablation record): NCO split e^{-i th (n_b + i)} = P_b T[i] with T[] and the taps G[i][p] as shared-memory tables
//           read by lane-uniform LDS.128, FFMA2 arithmetic, P_b applied to the 8 sums -- half the staged kernel's
//           instructions and the same time: bound by the shared-memory pipe;
//   MODE 1 (shipped): taps as immediates of the kernel-parameter constant bank, NCO as a per-row recurrence.
// cp.async needs 16-byte aligned sources, i.e. rows that start on a chunk boundary of the recording.  sg (the residue of
// start_sample + in_off) picks one of two equivalent decompositions of z[m] = sum_k h[k] y[q - k], q = mD + in_off:
//   sg = 0: rows START at q - (p+1)D: taps k = 1..8D from rows j..j+7, lone tap h[0] on the first sample of row j+8
//   sg = 1: rows END   at q - pD    : taps k = 0..8D-1 from rows j+1..j+8, lone tap h[8D] on the last sample of row j
// (j = output index within the tile, NT - 8 outputs per NT rows).  Samples outside the annotation are zeroed in the
// tiles that touch its ends; chunks outside the recording are not read (cp.async src-size).
// Input types: cf32 (2 samples per 16-byte chunk), ci16 (4), cu8 / ci8 (8), either byte order; the integer types are decoded
// in the tap loop (the same exact one-FMA decodes as everywhere else).  A row must be whole chunks (D a multiple of the
// samples per chunk) and start on a chunk boundary of the recording: (start_sample + in_off) mod SPC must be 0 (sg = 0)
// or SPC - 1 (sg = 1); every other alignment takes the staged kernel above.
template <int DK> struct DcRowsSpc { static constexpr int value = DK == DK_CF32 ? 2 : (DK == DK_CI16 ? 4 : 8); };
// SWZ: rows are packed (CPR chunks apart) and chunk c of row r sits at position c ^ swz(r) instead of rows being padded to
// an odd stride: the same conflict-free 128-bit row reads (8 consecutive rows hit 8 different 16-byte bank groups) in
// 8/9 of the shared memory, which is one more CTA per SM for the 128-row double-buffered shape.
#ifndef SA_DC_ROWS_ALIAS
#define SA_DC_ROWS_ALIAS 1
#endif
template <int CPR> __host__ __device__ constexpr int dc_rows_swz(int row) {
    return CPR >= 8 ? (row & 7) : (CPR == 4 ? ((row >> 1) & 3) : (CPR == 2 ? ((row >> 2) & 1) : 0));
}
template <int D, int NT = 256, int NBUF = 1, int SPC = 2, bool SWZ = false> struct DcRowsGeo {
    static_assert(D >= 2 && D <= 32, "row kernel: decimation 2..32");
    static_assert(D % SPC == 0 && (NBUF == 1 || NBUF == 2), "row kernel geometry: rows are whole 16-byte chunks");
    static_assert(!SWZ || ((D / SPC) & (D / SPC - 1)) == 0, "swizzled rows: a power-of-two number of chunks per row");
    static constexpr int CPR = D / SPC;                  // 16-byte chunks per row
    static constexpr int RS = SWZ ? CPR : (CPR | 1);     // row stride in chunks (odd unless swizzled)
    static constexpr int RAW_BYTES = NT * RS * 16;       // one raw tile: NT rows
    static constexpr int CSM_BYTES = 9 * NT * 8;
    static constexpr int G_BYTES = D * 32;               // G[i][p]
    static constexpr int T_BYTES = D * 16;               // (T.x, T.y, -T.y, T.x)
    // ALIAS: with two raw buffers the partial sums are written over the raw tile that has just been consumed (one more
    // barrier per tile, no shared memory of their own)
    static constexpr bool ALIAS = SA_DC_ROWS_ALIAS && NBUF == 2 && SWZ && RAW_BYTES >= CSM_BYTES;
    static constexpr int SMEM = 128 + NBUF * RAW_BYTES + (ALIAS ? 0 : CSM_BYTES) + G_BYTES + T_BYTES;
    static constexpr int BY_SMEM = (227 * 1024) / (SMEM + 1024);
    static constexpr int BY_WARPS = 2048 / NT;
    static constexpr int BY_REGS = 65536 / (NT * 64);    // aim: 64 registers per thread at most when that raises occupancy
    static constexpr int MINB = BY_SMEM < BY_WARPS ? (BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS) : (BY_WARPS < BY_REGS ? BY_WARPS : BY_REGS);
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_partial(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

template <int D, bool INTERIOR>
__device__ __forceinline__ void dc_rows_fir(const uint32_t row_s, const uint32_t g_s, const uint32_t t_s, const long long n_row,
                                            const long long count, pk2 (&ax)[4], pk2 (&ay)[4], float2& y_first, float2& y_last) {
#pragma unroll
    for (int c = 0; c < D / 2; c++) {
        float4 x = lds128(row_s + 16u * (unsigned)c);
        if constexpr (!INTERIOR) {
            const long long n = n_row + 2 * c;
            if (n < 0 || n >= count) { x.x = 0.f; x.y = 0.f; }
            if (n + 1 < 0 || n + 1 >= count) { x.z = 0.f; x.w = 0.f; }
        }
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const int i = 2 * c + hh;
            const float xr = hh ? x.z : x.x, xi = hh ? x.w : x.y;
            const float4 tt = lds128(t_s + 16u * (unsigned)i);
            const pk2 y = fma2(pack2(xi, xi), pack2(tt.z, tt.w), mul2(pack2(xr, xr), pack2(tt.x, tt.y)));
            float yr, yi;
            unpack2(y, yr, yi);
            if (i == 0) y_first = make_float2(yr, yi);
            if (i == D - 1) y_last = make_float2(yr, yi);
            const float4 ga = lds128(g_s + 32u * (unsigned)i), gb = lds128(g_s + 32u * (unsigned)i + 16u);
            const pk2 sx = pack2(yr, yr), sy = pack2(yi, yi);
            const pk2 g01 = pack2(ga.x, ga.y), g23 = pack2(ga.z, ga.w), g45 = pack2(gb.x, gb.y), g67 = pack2(gb.z, gb.w);
            ax[0] = fma2(g01, sx, ax[0]); ay[0] = fma2(g01, sy, ay[0]);
            ax[1] = fma2(g23, sx, ax[1]); ay[1] = fma2(g23, sy, ay[1]);
            ax[2] = fma2(g45, sx, ax[2]); ay[2] = fma2(g45, sy, ay[2]);
            ax[3] = fma2(g67, sx, ax[3]); ay[3] = fma2(g67, sy, ay[3]);
        }
    }
}

// Variant with NO table loads in the tap loop (MODE 1).  MODE 0 above is bound by the shared-memory pipe (ncu: MIO
// throttle + short scoreboard, LSU wavefronts 62 %: a lane-uniform LDS.128 still costs two wavefronts, and a sample
// needs three of them).  Here the taps are immediates of the kernel-parameter constant bank (`FFMA R, R, c[0][imm], R`:
// scalar FFMA instead of FFMA2, because FFMA2 cannot take a constant operand) and the NCO is a per-row recurrence
// P_{i+1} = P_i W seeded with the row's exact phase (15 steps), so the only shared-memory reads are the raw row itself.
template <int D> struct DcRowsTaps {
    float g[2][D][8];                 // g[sg][i][p] = h[(p+1)D - i - sg]
    float lone[2];                    // h[0], h[8D]
    float pad_[2];
};

// the SPC samples of one 16-byte chunk
template <int DK, bool SWAP> __device__ __forceinline__ void dc_rows_decode(const LoadParams& lp, const uint4 w, float2 (&x)[DcRowsSpc<DK>::value]) {
    const uint32_t ww[4] = { w.x, w.y, w.z, w.w };
    if constexpr (DK == DK_CF32) {
#pragma unroll
        for (int j = 0; j < 2; j++) { const cpx<float> v = Loader<float, DK_CF32>::template decode<SWAP>(lp, make_uint2(ww[2 * j], ww[2 * j + 1])); x[j] = make_float2(v.x, v.y); }
    } else if constexpr (DK == DK_CI16) {
#pragma unroll
        for (int j = 0; j < 4; j++) { const cpx<float> v = Loader<float, DK_CI16>::template decode<SWAP>(lp, ww[j]); x[j] = make_float2(v.x, v.y); }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) { const cpx<float> v = Loader<float, DK_C8>::template decode<SWAP>(lp, (uint16_t)(ww[j >> 1] >> (16 * (j & 1)))); x[j] = make_float2(v.x, v.y); }
    }
}

template <int DK, bool SWAP, int D, bool INTERIOR, int SG>
__device__ __forceinline__ void dc_rows_fir_c(const DcRowsTaps<D>& tp, const LoadParams& lp, const uint32_t row_s, const uint32_t swz16, const long long n_row,
                                              const long long count, float2 P, const float2 W, float (&sr)[8], float (&si)[8], float2& y_lone) {
    constexpr int SPC = DcRowsSpc<DK>::value;
#pragma unroll
    for (int c = 0; c < D / SPC; c++) {
        uint4 raw;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(row_s + ((16u * (unsigned)c) ^ swz16)) : "memory");
        float2 x[SPC];
        dc_rows_decode<DK, SWAP>(lp, raw, x);
#pragma unroll
        for (int hh = 0; hh < SPC; hh++) {
            const int i = SPC * c + hh;
            if constexpr (!INTERIOR) {
                const long long n = n_row + i;
                if (n < 0 || n >= count) x[hh] = make_float2(0.f, 0.f);
            }
            const float xr = x[hh].x, xi = x[hh].y;
            const float yr = __fmaf_rn(xr, P.x, -xi * P.y), yi = __fmaf_rn(xr, P.y, xi * P.x);
            if (i == (SG ? D - 1 : 0)) y_lone = make_float2(yr, yi);
            if (i + 1 < D) P = make_float2(__fmaf_rn(P.x, W.x, -P.y * W.y), __fmaf_rn(P.x, W.y, P.y * W.x));
#pragma unroll
            for (int p = 0; p < 8; p++) {
                sr[p] = __fmaf_rn(tp.g[SG][i][p], yr, sr[p]);
                si[p] = __fmaf_rn(tp.g[SG][i][p], yi, si[p]);
            }
        }
    }
}

template <int DK, bool SWAP, int D>
__device__ __forceinline__ void dc_rows_fir_pick(const DcRowsTaps<D>& tp, const LoadParams& lp, const int sg, const bool interior, const uint32_t row_s,
                                                 const uint32_t swz16, const long long n_row, const long long count, const float2 P, const float2 W,
                                                 float (&sr)[8], float (&si)[8], float2& yl) {
    if (sg) {
        if (interior) dc_rows_fir_c<DK, SWAP, D, true, 1>(tp, lp, row_s, swz16, n_row, count, P, W, sr, si, yl);
        else          dc_rows_fir_c<DK, SWAP, D, false, 1>(tp, lp, row_s, swz16, n_row, count, P, W, sr, si, yl);
    } else {
        if (interior) dc_rows_fir_c<DK, SWAP, D, true, 0>(tp, lp, row_s, swz16, n_row, count, P, W, sr, si, yl);
        else          dc_rows_fir_c<DK, SWAP, D, false, 0>(tp, lp, row_s, swz16, n_row, count, P, W, sr, si, yl);
    }
}

template <int DK, int D, int MODE, int NT, int NBUF, bool SWZ>
__global__ void __launch_bounds__(NT, DcRowsGeo<D, NT, NBUF, DcRowsSpc<DK>::value, SWZ>::MINB)
downconvert_rows_kernel(const DcArgs a, const __grid_constant__ DcRowsTaps<D> tp) {
    constexpr int SPC = DcRowsSpc<DK>::value, BPS = 16 / SPC;
    static_assert(MODE == 1 || (DK == DK_CF32 && !SWZ), "the table variant is cf32 little-endian, padded rows only");
    using G = DcRowsGeo<D, NT, NBUF, SPC, SWZ>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DcAnn an = a.anns[a.ann_base + blockIdx.y];
    constexpr int NB = NT - 8;
    const long long n_tiles = (an.m_out + NB - 1) / NB;
    long long tile = (long long)blockIdx.x * a.tiles_per_cta;
    const long long tile_end = min(n_tiles, tile + a.tiles_per_cta);
    if (tile >= tile_end) return;
    const int t = threadIdx.x;
    const uint32_t smem_s = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 127u) & ~127u;
    const uint32_t raw_s = smem_s, csm_own = raw_s + NBUF * G::RAW_BYTES, g_s = csm_own + (G::ALIAS ? 0 : G::CSM_BYTES), t_s = g_s + G::G_BYTES;
    const int sg = ((an.start_sample + an.in_off) & (SPC - 1)) ? 1 : 0;       // eligible annotations: residue 0 or SPC - 1
    float h_lone;
    float2 W = make_float2(1.f, 0.f);
    if constexpr (MODE == 0) {
        // tables of this annotation: taps of the chosen decomposition, NCO phasors of the D positions of a row
        const float* h = a.taps + an.taps_off;
        for (int idx = t; idx < 8 * D; idx += NT) {
            const int i = idx >> 3, p = idx & 7;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(g_s + 4u * (unsigned)idx), "f"(__ldg(&h[(p + 1) * D - i - sg])) : "memory");
        }
        if (t < D) {
            const float2 T = nco_phasor(an.phase_step * (unsigned long long)t);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(t_s + 16u * (unsigned)t), "f"(T.x), "f"(T.y), "f"(-T.y), "f"(T.x) : "memory");
        }
        h_lone = __ldg(&h[sg ? 8 * D : 0]);
    } else {
        h_lone = tp.lone[sg];
        W = nco_phasor_acc(an.phase_step);
    }
    const int lone_row = sg ? 0 : 8;                          // row (relative to j) that holds the lone sample

    // this thread's chunks of a tile: chunk g = t + NT k, row g / CPR, position g % CPR
    // (NT a multiple of CPR: the chunks of a thread are kDstStep apart; otherwise -- 3, 5, 6, 7 ... chunks per row -- the row and
    // position of every chunk are computed, divisions by a compile-time constant)
    constexpr bool kRegular = NT % G::CPR == 0;
    auto chunk_dst = [&](int g) { const int row = g / G::CPR, c = g % G::CPR; return 16u * (unsigned)(row * G::RS + (c ^ (SWZ ? dc_rows_swz<G::CPR>(row) : 0))); };
    const uint32_t dst_t = raw_s + chunk_dst(t);
    const uint32_t swz16 = SWZ ? 16u * (unsigned)dc_rows_swz<G::CPR>(t) : 0u;
    constexpr uint32_t kDstStep = (NT / G::CPR) * G::RS * 16;
    auto tile_n0 = [&](long long tl) { return (tl * NB - 8 - sg) * D + an.in_off + sg; };       // sample of row 0 (annotation-relative)
    auto issue = [&](long long tl, uint32_t dst0) {
        const long long s0 = an.start_sample + tile_n0(tl);                                     // a chunk boundary by construction
        const char* src = reinterpret_cast<const char*>(a.lp.base) + BPS * s0 + 16 * (long long)t;
        if (s0 >= 0 && s0 + (long long)NT * D <= a.n_samples) {
#pragma unroll
            for (int k = 0; k < G::CPR; k++)
                cp_async16(kRegular ? dst0 + kDstStep * k : dst0 - chunk_dst(t) + chunk_dst(t + NT * k), src + (size_t)NT * 16 * k);
        } else {
#pragma unroll
            for (int k = 0; k < G::CPR; k++) {
                const long long s = s0 + SPC * ((long long)t + NT * k);
                const int bytes = s < 0 ? 0 : (int)max(0LL, min((long long)SPC, a.n_samples - s)) * BPS;
                cp_async16_partial(kRegular ? dst0 + kDstStep * k : dst0 - chunk_dst(t) + chunk_dst(t + NT * k),
                                   bytes ? (const void*)(src + (size_t)NT * 16 * k) : a.lp.base, bytes);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(tile, dst_t);
    uint32_t buf = 0;                                       // byte offset of the current raw tile (NBUF == 2)
    for (; tile < tile_end; tile++) {
        const uint32_t row_s = raw_s + buf + (unsigned)(t * G::RS * 16);
        const long long m0 = tile * NB;
        const int nbt = (int)min((long long)NB, an.m_out - m0);
        const long long n0 = tile_n0(tile);
        const long long n_row = n0 + (long long)t * D;
        const bool interior = n0 >= 0 && n0 + (long long)NT * D <= an.count;
        const float2 P = nco_phasor(an.phase_step * (unsigned long long)n_row);     // phasor of the row's first sample
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                    // the tile has landed; C_p of the previous tile has been consumed
        const uint32_t csm_s = G::ALIAS ? raw_s + buf : csm_own;   // ALIAS: over the tile this iteration consumes
        if constexpr (NBUF == 2) {                          // the next tile flies under this tile's tap loop
            buf ^= (uint32_t)G::RAW_BYTES;
            if (tile + 1 < tile_end) issue(tile + 1, dst_t + buf);
        }
        if constexpr (MODE == 0) {
            pk2 ax[4], ay[4];
#pragma unroll
            for (int p = 0; p < 4; p++) { ax[p] = pack2(0.f, 0.f); ay[p] = pack2(0.f, 0.f); }
            float2 y_first = make_float2(0.f, 0.f), y_last = y_first;
            if (interior) dc_rows_fir<D, true>(row_s, g_s, t_s, n_row, an.count, ax, ay, y_first, y_last);
            else          dc_rows_fir<D, false>(row_s, g_s, t_s, n_row, an.count, ax, ay, y_first, y_last);
            // C_p[b] = P_b S_p[b]
            const pk2 Px = pack2(P.x, P.x), Py = pack2(P.y, P.y), nPy = pack2(-P.y, -P.y);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const pk2 re = fma2(ay[k], nPy, mul2(ax[k], Px)), im = fma2(ay[k], Px, mul2(ax[k], Py));
                float r0, r1, i0, i1;
                unpack2(re, r0, r1); unpack2(im, i0, i1);
                sts64(csm_s + 8u * (unsigned)((2 * k) * NT + t), r0, i0);
                sts64(csm_s + 8u * (unsigned)((2 * k + 1) * NT + t), r1, i1);
            }
            const float2 yl = sg ? y_last : y_first;
            const float lr = h_lone * yl.x, li = h_lone * yl.y;
            sts64(csm_s + 8u * (unsigned)(8 * NT + t), lr * P.x - li * P.y, lr * P.y + li * P.x);
        } else {
            float sr[8], si[8];
#pragma unroll
            for (int p = 0; p < 8; p++) { sr[p] = 0.f; si[p] = 0.f; }
            float2 yl = make_float2(0.f, 0.f);
            if constexpr (DK == DK_C8) {
                dc_rows_fir_pick<DK, false, D>(tp, a.lp, sg, interior, row_s, swz16, n_row, an.count, P, W, sr, si, yl);
            } else {
                if (a.lp.swap) dc_rows_fir_pick<DK, true, D>(tp, a.lp, sg, interior, row_s, swz16, n_row, an.count, P, W, sr, si, yl);
                else           dc_rows_fir_pick<DK, false, D>(tp, a.lp, sg, interior, row_s, swz16, n_row, an.count, P, W, sr, si, yl);
            }
            if constexpr (G::ALIAS) __syncthreads();        // every row of the tile has been read
#pragma unroll
            for (int p = 0; p < 8; p++) sts64(csm_s + 8u * (unsigned)(p * NT + t), sr[p], si[p]);
            sts64(csm_s + 8u * (unsigned)(8 * NT + t), h_lone * yl.x, h_lone * yl.y);
        }
        __syncthreads();                                    // C_p complete; the raw tile is free
        if constexpr (NBUF == 1) { if (tile + 1 < tile_end) issue(tile + 1, dst_t); }   // next tile flies under the combine step and the stores
        if (t < nbt) {
            float2 z = lds64(csm_s + 8u * (unsigned)(8 * NT + t + lone_row));
#pragma unroll
            for (int p = 0; p < 8; p++) { const float2 c = lds64(csm_s + 8u * (unsigned)(p * NT + t + sg + 7 - p)); z.x += c.x; z.y += c.y; }
            if (a.out) { double* o = a.out + an.out_off + m0 + t; o[0] = (double)z.x; o[an.m_out] = (double)z.y; }
            if (a.scratch) a.scratch[an.scr_off + m0 + t] = z;
        }
    }
}

// Box-car ("fast") mode of the same idea: z[m] = (1/D) sum_{k<D} y[mD + k] is the mean of ONE row, so there are no partial
// sums, no halo and no taps -- thread b mixes and adds its own row and writes output m0 + b.  The decimation factor is a
// run-time value here (any multiple of the samples per 16-byte chunk whose double-buffered 128-row tile fits the shared
// memory budget); rows are padded to an odd number of chunks.  Needs start_sample on a chunk boundary (in_off is 0).
constexpr int kDcFastRows = 128;
__host__ __device__ constexpr int dc_fast_smem_bytes(int cpr) { return 128 + 2 * kDcFastRows * (cpr | 1) * 16; }
constexpr int kDcFastSmemMax = 72 * 1024;          // three CTAs per SM

template <int DK, bool SWAP, bool INTERIOR>
__device__ __forceinline__ float2 dc_rows_boxcar(const LoadParams& lp, const uint32_t row_s, const int cpr, const long long n_row,
                                                 const long long count, float2 P, const float2 W) {
    constexpr int SPC = DcRowsSpc<DK>::value;
    float sr = 0.f, si = 0.f;
    for (int c = 0; c < cpr; c++) {
        uint4 raw;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "r"(row_s + 16u * (unsigned)c) : "memory");
        float2 x[SPC];
        dc_rows_decode<DK, SWAP>(lp, raw, x);
#pragma unroll
        for (int hh = 0; hh < SPC; hh++) {
            if constexpr (!INTERIOR) {
                const long long n = n_row + SPC * c + hh;
                if (n < 0 || n >= count) x[hh] = make_float2(0.f, 0.f);
            }
            sr += __fmaf_rn(x[hh].x, P.x, -x[hh].y * P.y);
            si += __fmaf_rn(x[hh].x, P.y, x[hh].y * P.x);
            P = make_float2(__fmaf_rn(P.x, W.x, -P.y * W.y), __fmaf_rn(P.x, W.y, P.y * W.x));
        }
    }
    return make_float2(sr, si);
}

template <int DK>
__global__ void __launch_bounds__(kDcFastRows, 3)
downconvert_rows_fast_kernel(const DcArgs a) {
    constexpr int SPC = DcRowsSpc<DK>::value, BPS = 16 / SPC, NT = kDcFastRows;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DcAnn an = a.anns[a.ann_base + blockIdx.y];
    const long long n_tiles = (an.m_out + NT - 1) / NT;
    long long tile = (long long)blockIdx.x * a.tiles_per_cta;
    const long long tile_end = min(n_tiles, tile + a.tiles_per_cta);
    if (tile >= tile_end) return;
    const int t = threadIdx.x, D = an.down, cpr = D / SPC, rs = cpr | 1;
    const uint32_t raw_s = ((uint32_t)__cvta_generic_to_shared(smem_raw) + 127u) & ~127u;
    const uint32_t raw_bytes = (uint32_t)(NT * rs * 16);
    const unsigned magic = cpr > 1 ? (unsigned)(0xFFFFFFFFu / (unsigned)cpr + 1u) : 0u;       // g / cpr = umulhi(g, magic), g < NT cpr
    const float2 W = nco_phasor_acc(an.phase_step);
    const float inv = 1.0f / (float)D;
    auto issue = [&](long long tl, uint32_t dst_base) {
        const long long s0 = an.start_sample + tl * NT * D;                                   // a chunk boundary (host-checked)
        const char* src = reinterpret_cast<const char*>(a.lp.base) + BPS * s0 + 16 * (long long)t;
        const bool inside = s0 >= 0 && s0 + (long long)NT * D <= a.n_samples;
        for (int k = 0; k < cpr; k++) {
            const int g = t + NT * k;
            const int row = cpr > 1 ? (int)__umulhi((unsigned)g, magic) : g;
            const uint32_t dst = dst_base + 16u * (unsigned)(row * rs + (g - row * cpr));
            if (inside) {
                cp_async16(dst, src + (size_t)NT * 16 * k);
            } else {
                const long long s = s0 + (long long)SPC * g;
                const int bytes = s < 0 ? 0 : (int)max(0LL, min((long long)SPC, a.n_samples - s)) * BPS;
                cp_async16_partial(dst, bytes ? (const void*)(src + (size_t)NT * 16 * k) : a.lp.base, bytes);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(tile, raw_s);
    uint32_t buf = 0;
    for (; tile < tile_end; tile++) {
        const long long m0 = tile * NT;
        const int nbt = (int)min((long long)NT, an.m_out - m0);
        const uint32_t row_s = raw_s + buf + (unsigned)(t * rs * 16);
        const long long n_row = (m0 + t) * D;
        const float2 P = nco_phasor(an.phase_step * (unsigned long long)n_row);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                    // the tile has landed; the other buffer has been read by every thread
        buf ^= raw_bytes;
        if (tile + 1 < tile_end) issue(tile + 1, raw_s + buf);
        const bool interior = (m0 + NT) * D <= an.count;
        float2 z;
        if constexpr (DK == DK_C8) {
            z = interior ? dc_rows_boxcar<DK, false, true>(a.lp, row_s, cpr, n_row, an.count, P, W)
                         : dc_rows_boxcar<DK, false, false>(a.lp, row_s, cpr, n_row, an.count, P, W);
        } else if (a.lp.swap) {
            z = interior ? dc_rows_boxcar<DK, true, true>(a.lp, row_s, cpr, n_row, an.count, P, W)
                         : dc_rows_boxcar<DK, true, false>(a.lp, row_s, cpr, n_row, an.count, P, W);
        } else {
            z = interior ? dc_rows_boxcar<DK, false, true>(a.lp, row_s, cpr, n_row, an.count, P, W)
                         : dc_rows_boxcar<DK, false, false>(a.lp, row_s, cpr, n_row, an.count, P, W);
        }
        if (t < nbt) {
            z.x *= inv; z.y *= inv;
            if (a.out) { double* o = a.out + an.out_off + m0 + t; o[0] = (double)z.x; o[an.m_out] = (double)z.y; }
            if (a.scratch) a.scratch[an.scr_off + m0 + t] = z;
        }
    }
}

// ---------------- Welch ----------------
struct WelchSig {
    const double* re;      // planar FP64 input (the downconverter's output rows), or
    const double* im;
    const float2* f32;     // interleaved FP32 rows (the downconverter's scratch); used when not NULL
    long long n;           // samples
    double scale;          // 1 / (nseg * fs * sum w^2)  |  1 / (nseg * (sum w)^2)
    long long nseg;
};

struct WelchArgs {
    const WelchSig* sigs;
    long long hop;
    const void* window;    // T[N]
    const void* twiddle;
    const void* aux;       // welch_accum_mid_kernel: roots W_N^j (twiddle = its pass-1 pair table)
    void* partial;         // T [sig][slot][N], slot = split * FPC + frame slot
    int nsplit;
    int detrend;           // 1: subtract the segment's mean before the window
    int window_id;         // direct kernel: window evaluated in the kernel
    double* out_db;        // [sig][N], fft-shifted
    int* ticket;           // welch_accum_mid_kernel: task counter (zeroed before the launch); a task = (signal, split)
    int n_tasks;           // signals x nsplit
};

template <typename T> __device__ __forceinline__ cpx<T> welch_load(const WelchSig& sg, long long i) {
    if (sg.f32) { const float2 v = __ldg(&sg.f32[i]); return mk2<T>((T)v.x, (T)v.y); }
    return mk2<T>((T)__ldg(&sg.re[i]), (T)__ldg(&sg.im[i]));
}

// mean over the TPF threads x P values of one frame (detrend = constant)
template <typename T, int TPF, int P, int CTA>
__device__ __forceinline__ cpx<T> frame_mean(const cpx<T> (&v)[P], cpx<T>* red) {
    T sx = 0, sy = 0;
#pragma unroll
    for (int q = 0; q < P; q++) { sx += v[q].x; sy += v[q].y; }
    constexpr int W = TPF < 32 ? TPF : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
    if constexpr (TPF > 32) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        constexpr int WPF = TPF / 32;                 // warps per frame
        __syncthreads();
        if (lane == 0) red[warp] = mk2<T>(sx, sy);
        __syncthreads();
        const int w0 = (warp / WPF) * WPF;
        sx = 0; sy = 0;
#pragma unroll
        for (int k = 0; k < WPF; k++) { sx += red[w0 + k].x; sy += red[w0 + k].y; }
    }
    const T inv = (T)1 / (T)(TPF * P);
    return mk2<T>(sx * inv, sy * inv);
}

template <typename T, int N>
__global__ void __launch_bounds__(Geo<T, N>::CTA, Geo<T, N>::MINB)
welch_accum_kernel(const WelchArgs a) {
    using G = Geo<T, N>;
    constexpr int P = G::P, TPF = G::TPF, FPC = G::FPC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ cpx<T> red[32];
    const int fl = threadIdx.x / TPF, t = threadIdx.x % TPF;
    cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw) + (size_t)fl * G::SM_ELEMS;
    const cpx<T>* tw = reinterpret_cast<const cpx<T>*>(a.twiddle);
    if constexpr (G::TW_SMEM) {      // one-warp-per-frame plans read their twiddles from shared memory
        cpx<T>* tsm = reinterpret_cast<cpx<T>*>(smem_raw + G::SMEM_BYTES);
        for (int i = threadIdx.x; i < (int)(G::TW_BYTES / sizeof(cpx<T>)); i += G::CTA) tsm[i] = __ldg(&tw[i]);
        __syncthreads();
        tw = tsm;
    }
    const TwSeed<T> seed = load_tw_seed<T, N>(reinterpret_cast<const cpx<T>*>(a.twiddle), t);
    const T* win = reinterpret_cast<const T*>(a.window);
    const WelchSig sg = a.sigs[blockIdx.y];
    T acc[P];
#pragma unroll
    for (int q = 0; q < P; q++) acc[q] = (T)0;
    const long long stride = (long long)a.nsplit * FPC;
    const long long iters = (sg.nseg + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long seg = it * stride + (long long)blockIdx.x * FPC + fl;
        const bool valid = seg < sg.nseg;
        cpx<T> v[P];
        if (valid) {
            const long long s0 = seg * a.hop + t;
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = welch_load<T>(sg, s0 + TPF * q);
        } else {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = mk2<T>((T)0, (T)0);
        }
        if (a.detrend) {                 // launch-uniform
            const cpx<T> m = frame_mean<T, TPF, P, G::CTA>(v, red);
#pragma unroll
            for (int q = 0; q < P; q++) { v[q].x -= m.x; v[q].y -= m.y; }
        }
#pragma unroll
        for (int q = 0; q < P; q++) { const T w = __ldg(&win[t + TPF * q]); v[q].x *= w; v[q].y *= w; }
        fft_frame<T, N, false>(v, t, sm, tw, nullptr, seed);
        if (valid) {
#pragma unroll
            for (int q = 0; q < P; q++) acc[q] += fma_t(v[q].x, v[q].x, v[q].y * v[q].y);
        }
    }
    T* part = reinterpret_cast<T*>(a.partial) + (((size_t)blockIdx.y * a.nsplit + blockIdx.x) * FPC + fl) * N;
#pragma unroll
    for (int q = 0; q < P; q++) part[t + TPF * q] = acc[q];
}

// FP32 segments of 2048 .. 16384 points on the small-radix-first plan of spectrogram_mid_kernel.cuh (R0 x 32 x 32, pass-1
// twiddles from a tiny shared-memory table, pass-2 twiddles by the register recurrence, window folded into the first
// butterfly stage) instead of the general plan's two table-driven passes: thread t owns S = 32 / R0 CONSECUTIVE samples of
// each of the R0 slices of a segment -- 128-bit loads from the downconverter's FP32 rows when they are 16-byte aligned.
// CTA_: threads per CTA, a multiple of the TPF threads of one segment (the spectrogram kernel's geometry, or one segment
// per CTA so that the CTAs are small enough to share an SM with the downconverter's, see run_batch_device).
#ifndef SA_WELCH_L2_PREFETCH
#define SA_WELCH_L2_PREFETCH 0      // L2 prefetch of a slot's next segment: measured no faster (C3 1.226 against 1.212 ms)
#endif
template <int N, int CTA_>
__global__ void __launch_bounds__(CTA_, 512 / CTA_)
welch_accum_mid_kernel(const WelchArgs a) {
    using G = MidGeo<N>;
    constexpr int P = 32, R0 = G::R0, S = G::S, TPF = G::TPF, FPC = CTA_ / TPF;
    static_assert(CTA_ % TPF == 0 && FPC >= 1 && FPC <= G::FPC, "Welch CTA: whole segments, at most the spectrogram kernel's");
    constexpr size_t EX_BYTES = (size_t)FPC * G::SM_ELEMS * sizeof(float2);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float2 red[32];
    const int fl = threadIdx.x / TPF, t = threadIdx.x % TPF;
    float2* sm = reinterpret_cast<float2*>(smem_raw) + (size_t)fl * G::SM_ELEMS;
    TwPair<float>* t1 = reinterpret_cast<TwPair<float>*>(smem_raw + EX_BYTES);
    float* wsm = reinterpret_cast<float*>(smem_raw + EX_BYTES + G::T1_BYTES);
    mid_setup_tables<N, true>(a.twiddle, a.window, t1, wsm);
    __syncthreads();
    const float* win = wsm + t * G::WROW;
    TwSeed<float> seed;
    {
        const float2* root = reinterpret_cast<const float2*>(a.aux);
        seed.om = __ldg(&root[t]);
        seed.oh = __ldg(&root[(16 * t) & (N - 1)]);
        seed.q_lo = seed.om; seed.q_hi = seed.om;
    }
    const TwPair<float>* t1_row = t1 + (t % R0);
#ifndef SA_WELCH_REC1
#define SA_WELCH_REC1 1
#endif
    TwSeed<float> seed1;                                     // pass-1 twiddle recurrence: om = W^(t mod R0), oh = om^16
    {
        const TwPair<float> m0 = ldg_tw<true>(t1_row), m1 = ldg_tw<true>(t1_row + R0);
        seed1.om = m1.lo; seed1.oh = m0.hi; seed1.q_lo = seed1.om; seed1.q_hi = seed1.om;
    }
    // Persistent CTAs: the tables above are set up once per CTA, then the CTA draws (signal, split) tasks from a ticket
    // counter until none is left (C3: 1000 tasks on 148 CTAs; with one CTA per task the table setup, the CTA launch and
    // the tail of every CTA were ~4 % of the step).  A task's partial spectrum does not depend on which CTA computes it.
    __shared__ int s_task;
  for (;;) {
    __syncthreads();                                         // the previous task is done with the exchange buffers and s_task
    if (threadIdx.x == 0) s_task = atomicAdd(a.ticket, 1);
    __syncthreads();
    const int task = s_task;
    if (task >= a.n_tasks) break;
    const int task_sig = task / a.nsplit, task_split = task - task_sig * a.nsplit;
    const WelchSig sg = a.sigs[task_sig];
    // 128-bit loads: FP32 rows whose every slice start is 16-byte aligned (S t is even for every plan; hop and row start must be)
    const bool vec = sg.f32 && ((reinterpret_cast<uintptr_t>(sg.f32) & 15) == 0) && ((a.hop & 1) == 0);
    float acc[P];
#pragma unroll
    for (int q = 0; q < P; q++) acc[q] = 0.f;
    const long long stride = (long long)a.nsplit * FPC;
    const long long iters = (sg.nseg + stride - 1) / stride;
    for (long long it = 0; it < iters; it++) {
        const long long seg = it * stride + (long long)task_split * FPC + fl;
        const bool valid = seg < sg.nseg;
        // a slot whose segment does not exist skips the transform: the slots of a CTA synchronise on their own named
        // barriers (one slot per CTA: the test is CTA-uniform), only the mean removal uses CTA-wide barriers.  29
        // segments over 4 CTAs x 2 slots: 15 instead of 16 CTA steps per signal.
        if (!valid && !a.detrend) continue;
#if SA_WELCH_L2_PREFETCH
        // the slot's NEXT segment (stride segments ahead: no sample in common with this one) is pulled into L2 while this
        // one is transformed: N * 8 bytes = N / 16 lines of 128 bytes, N / 16 / TPF = 2 prefetches per thread
        if (sg.f32 && seg + stride < sg.nseg) {
            const char* nx = reinterpret_cast<const char*>(sg.f32 + (seg + stride) * a.hop);
#pragma unroll
            for (int i = 0; i < N / 16 / TPF; i++) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 128 * (t + TPF * i)));
        }
#endif
        float2 v[P];
        if (valid) {
            const long long s0 = seg * a.hop + (long long)S * t;
            if (vec) {
#pragma unroll
                for (int m = 0; m < R0; m++) {
                    const float4* p = reinterpret_cast<const float4*>(sg.f32 + s0 + (long long)m * (N / R0));
#pragma unroll
                    for (int i = 0; i < S; i += 2) {
                        const float4 x = __ldg(p + i / 2);
                        v[i + S * m] = make_float2(x.x, x.y);
                        v[i + 1 + S * m] = make_float2(x.z, x.w);
                    }
                }
            } else {
#pragma unroll
                for (int m = 0; m < R0; m++)
#pragma unroll
                    for (int i = 0; i < S; i++) v[i + S * m] = welch_load<float>(sg, s0 + (long long)m * (N / R0) + i);
            }
        } else {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = make_float2(0.f, 0.f);
        }
        if (a.detrend) {                 // launch-uniform
            const float2 m = frame_mean<float, TPF, P, CTA_>(v, red);
#pragma unroll
            for (int q = 0; q < P; q++) { v[q].x -= m.x; v[q].y -= m.y; }
        }
        mid_fft<N, true, SA_WELCH_REC1 != 0>(v, t, fl, sm, win, t1_row, seed, &seed1);
        if (valid) {
#pragma unroll
            for (int q = 0; q < P; q++) acc[q] += __fmaf_rn(v[q].x, v[q].x, v[q].y * v[q].y);
        }
    }
    // one partial spectrum per CTA: the segment slots of the CTA are added up here in a fixed order (through the exchange
    // buffers, which are free now), so the finalize step reads nsplit rows per signal instead of nsplit x FPC
    if constexpr (FPC > 1) {
        __syncthreads();                                     // every slot has read back its last exchange
        float* mine = reinterpret_cast<float*>(sm);
        if (fl > 0) {
#pragma unroll
            for (int q = 0; q < P; q++) mine[t + TPF * q] = acc[q];
        }
        __syncthreads();
        if (fl == 0) {
#pragma unroll
            for (int f = 1; f < FPC; f++) {
                const float* other = reinterpret_cast<const float*>(reinterpret_cast<const float2*>(smem_raw) + (size_t)f * G::SM_ELEMS);
#pragma unroll
                for (int q = 0; q < P; q++) acc[q] += other[t + TPF * q];
            }
        }
    }
    if (fl == 0) {
        float* part = reinterpret_cast<float*>(a.partial) + (size_t)task * N;
#pragma unroll
        for (int q = 0; q < P; q++) part[t + TPF * q] = acc[q];
    }
  }
}

// sums the partial spectra in a fixed order (deterministic), scales, dB, fft-shift
template <typename T>
__global__ void welch_finalize_kernel(const WelchArgs a, const int n, const int slots) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const WelchSig sg = a.sigs[blockIdx.y];
    const T* part = reinterpret_cast<const T*>(a.partial) + (size_t)blockIdx.y * slots * n + k;
    double s = 0.0;
    for (int i = 0; i < slots; i++) s += (double)part[(size_t)i * n];
    const double v = sg.nseg > 0 ? 10.0 * log10(s * sg.scale + 1e-30) : __longlong_as_double(0x7ff8000000000000LL);
    a.out_db[(size_t)blockIdx.y * n + ((k + n / 2) & (n - 1))] = v;
}

// ---- any transform length: direct DFT (the reference's short-signal branch passes nfft = signal length) ----
// Grid (bins / 256, signals).  The CTA builds W_nfft^j (FP64 sincospi, rounded once to T) in shared memory; the segment
// goes through shared memory in tiles (mean removed, window evaluated in FP64); thread k accumulates bin k with the
// exact twiddle index (k n) mod nfft kept by an integer recurrence.  O(nfft^2) per segment: meant for nfft < 8192.
constexpr int kDirectTile = 1024;

__device__ __forceinline__ double window_value(int id, int i, int n) {
    const double x = 2.0 * (double)i / (double)n;            // in units of pi
    switch (id) {
        case SA_WIN_HANN:     return 0.5 - 0.5 * cospi(x);
        case SA_WIN_HAMMING:  return 0.54 - 0.46 * cospi(x);
        case SA_WIN_BLACKMAN: return 0.42 - 0.5 * cospi(x) + 0.08 * cospi(2 * x);
        case SA_WIN_BLACKMAN_HARRIS: return 0.35875 - 0.48829 * cospi(x) + 0.14128 * cospi(2 * x) - 0.01168 * cospi(3 * x);
        default: return 1.0;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
psd_direct_kernel(const WelchArgs a, const int nfft) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double red_x[8], red_y[8];
    cpx<T>* tw = reinterpret_cast<cpx<T>*>(smem_raw);
    cpx<T>* xs = tw + nfft;
    const WelchSig sg = a.sigs[blockIdx.y];
    const int k = blockIdx.x * 256 + threadIdx.x;
    for (int j = threadIdx.x; j < nfft; j += 256) {
        double s, c;
        sincospi(2.0 * (double)j / (double)nfft, &s, &c);
        tw[j] = mk2<T>((T)c, (T)(-s));
    }
    T acc = (T)0;
    for (long long seg = 0; seg < sg.nseg; seg++) {
        const long long s0 = seg * a.hop;
        double mx = 0.0, my = 0.0;
        if (a.detrend) {
            double sx = 0.0, sy = 0.0;
            for (int i = threadIdx.x; i < nfft; i += 256) { const cpx<double> v = welch_load<double>(sg, s0 + i); sx += v.x; sy += v.y; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); }
            __syncthreads();
            if ((threadIdx.x & 31) == 0) { red_x[threadIdx.x >> 5] = sx; red_y[threadIdx.x >> 5] = sy; }
            __syncthreads();
            for (int w = 0; w < 8; w++) { mx += red_x[w]; my += red_y[w]; }
            mx /= (double)nfft; my /= (double)nfft;
        }
        T xr0 = 0, xi0 = 0, xr1 = 0, xi1 = 0;       // two interleaved partial sums per component
        int idx = 0;
        for (int n0 = 0; n0 < nfft; n0 += kDirectTile) {
            const int len = min(kDirectTile, nfft - n0);
            __syncthreads();
            for (int i = threadIdx.x; i < len; i += 256) {
                const cpx<double> v = welch_load<double>(sg, s0 + n0 + i);
                const double w = window_value(a.window_id, n0 + i, nfft);
                xs[i] = mk2<T>((T)((v.x - mx) * w), (T)((v.y - my) * w));
            }
            __syncthreads();
            if (k < nfft) {
                int i = 0;
                for (; i + 1 < len; i += 2) {
                    const cpx<T> x0 = xs[i], w0 = tw[idx];
                    idx += k; if (idx >= nfft) idx -= nfft;
                    const cpx<T> x1 = xs[i + 1], w1 = tw[idx];
                    idx += k; if (idx >= nfft) idx -= nfft;
                    xr0 = fma_t(-x0.y, w0.y, fma_t(x0.x, w0.x, xr0)); xi0 = fma_t(x0.y, w0.x, fma_t(x0.x, w0.y, xi0));
                    xr1 = fma_t(-x1.y, w1.y, fma_t(x1.x, w1.x, xr1)); xi1 = fma_t(x1.y, w1.x, fma_t(x1.x, w1.y, xi1));
                }
                if (i < len) {
                    const cpx<T> x0 = xs[i], w0 = tw[idx];
                    idx += k; if (idx >= nfft) idx -= nfft;
                    xr0 = fma_t(-x0.y, w0.y, fma_t(x0.x, w0.x, xr0)); xi0 = fma_t(x0.y, w0.x, fma_t(x0.x, w0.y, xi0));
                }
            }
        }
        const T xr = xr0 + xr1, xi = xi0 + xi1;
        acc += fma_t(xr, xr, xi * xi);
    }
    if (k < nfft) {
        const double v = sg.nseg > 0 ? 10.0 * log10((double)acc * sg.scale + 1e-30) : __longlong_as_double(0x7ff8000000000000LL);
        a.out_db[(size_t)blockIdx.y * nfft + (size_t)((k + nfft / 2) % nfft)] = v;
    }
}

}  // namespace sa
