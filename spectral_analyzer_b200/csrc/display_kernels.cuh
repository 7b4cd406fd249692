// display_kernels.cuh -- the steps either side of the hot path.
//
// N2  canvas_kernel: MainController.renderSpectrogram (S/controllers/MainController.java:1261-1291):
//     pixel (t, H-1-f) <- colour(waterfall[t][(int)(f / H * nfft)] - conversion), evaluated on the GPU
//     from the dB rows the spectrogram kernel just wrote, so that only the canvas crosses PCIe.
//     Beyond the reference's nearest-bin / one-frame-per-column pick it can reduce a block of
//     frames x bins per pixel (max, or mean of linear power) for whole-recording views.
// N3  iq_pack_kernel: IqData.getInterleavedBinary (S/data/IqData.java:160-187), float32 / int16 LE.
//     series_kernel : AnalysisDialogController.updateMagnitudeChart / updateFrequencyChart
//     (S/controllers/AnalysisDialogController.java:219-290): hypot -> EMA -> 20 log10, and
//     atan2 phase difference -> wrap -> Hz -> EMA (+ centre frequency), FP64 like the reference.
#pragma once
#include "spectrogram_kernel.cuh"

namespace sa {

enum { REDUCE_NEAREST = 0, REDUCE_MAX = 1, REDUCE_MEAN = 2 };

struct CanvasArgs {
    const float* db;        // [ncols * fpc][nfft] dB rows of this chunk (fft-shifted, spectrogram_kernel output)
    uint32_t*    out;       // [canvas_h][canvas_w] RGBA8, row 0 = top (+fs/2)
    int nfft, fpc;          // frames per canvas column
    int canvas_w, canvas_h;
    int col0, ncols;        // columns [col0, col0 + ncols) are produced from this chunk
    int reduce;
    int cmap;
    float inv_range, cmap_bias;
};

// CTA = 32 pixel rows x 8 frame groups: lane = row f (consecutive rows -> consecutive bin ranges, coalesced),
// warp g takes the column's frames g, g + 8, ... with four independent partial results in flight; the eight
// partials of a pixel are combined through shared memory.
constexpr int kCanvasRows = 32, kCanvasGroups = 8;

__global__ void __launch_bounds__(kCanvasRows * kCanvasGroups)
canvas_kernel(const CanvasArgs a) {
    __shared__ float part[kCanvasGroups][kCanvasRows];
    const int lane = threadIdx.x % kCanvasRows, g = threadIdx.x / kCanvasRows;
    const int f = blockIdx.x * kCanvasRows + lane;
    const int col = blockIdx.y;
    const bool live = f < a.canvas_h;
    // MainController.java:1280  bin = (int)((double) f / canvasH * nfft)
    const int b0 = live ? (int)((double)f / (double)a.canvas_h * (double)a.nfft) : 0;
    int b1 = live ? (int)((double)(f + 1) / (double)a.canvas_h * (double)a.nfft) : 1;
    if (b1 <= b0) b1 = b0 + 1;
    if (b1 > a.nfft) b1 = a.nfft;
    const float* rows = a.db + (size_t)col * a.fpc * a.nfft;
    float v = 0.f;
    if (a.reduce == REDUCE_NEAREST) {
        if (live && g == 0) v = rows[b0];               // first frame of the column, nearest bin (:1283)
    } else {
        const bool is_max = a.reduce == REDUCE_MAX;
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = is_max ? -3.0e38f : 0.f;
        if (live) {
            for (int fr = g; fr < a.fpc; fr += 4 * kCanvasGroups) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int fu = fr + u * kCanvasGroups;
                    if (fu < a.fpc) {
                        const float* r = rows + (size_t)fu * a.nfft;
                        for (int b = b0; b < b1; b++) {
                            const float x = r[b];
                            acc[u] = is_max ? fmaxf(acc[u], x) : acc[u] + exp2f(x * 0.33219280948873623f);
                        }
                    }
                }
            }
        }
        v = is_max ? fmaxf(fmaxf(acc[0], acc[1]), fmaxf(acc[2], acc[3])) : (acc[0] + acc[1]) + (acc[2] + acc[3]);
        part[g][lane] = v;
        __syncthreads();
        if (g == 0) {
#pragma unroll
            for (int k = 1; k < kCanvasGroups; k++) v = is_max ? fmaxf(v, part[k][lane]) : v + part[k][lane];
            if (!is_max) v = 3.01029995663981f * log2f(v / (float)(a.fpc * (b1 - b0)));   // mean of linear power, back to dB
        }
    }
    if (!live || g != 0) return;
    const uint32_t px = a.cmap == 1 ? colormap_px<1>(v, a.inv_range, a.cmap_bias) : colormap_px<0>(v, a.inv_range, a.cmap_bias);
    a.out[(size_t)(a.canvas_h - 1 - f) * a.canvas_w + a.col0 + col] = px;      // :1288 y flipped
}

// Second stage of the fused pooling: acc[col][bin] holds max / sum over the column's frames of |X|^2 (spectrogram
// epilogue, pool_mode); pixel (col, f) reduces the bins [bin(f), bin(f+1)) of its column, takes the level once and
// colours it.  One thread per pixel, consecutive threads = consecutive pixel rows (adjacent bin ranges: coalesced).
struct CanvasPowerArgs {
    const float* acc;       // [ncols][nfft]
    uint32_t*    out;
    int nfft, canvas_w, canvas_h, col0, ncols;
    int reduce, cmap, db_mode;
    float inv_count;        // MEAN: 1 / frames per column
    float inv_range, cmap_bias;
};

__global__ void __launch_bounds__(256)
canvas_power_kernel(const CanvasPowerArgs a) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int col = blockIdx.y;
    if (f >= a.canvas_h) return;
    const int b0 = (int)((double)f / (double)a.canvas_h * (double)a.nfft);          // MainController.java:1280
    int b1 = (int)((double)(f + 1) / (double)a.canvas_h * (double)a.nfft);
    if (b1 <= b0) b1 = b0 + 1;
    if (b1 > a.nfft) b1 = a.nfft;
    const float* r = a.acc + (size_t)col * a.nfft;
    float p = 0.f;
    if (a.reduce == REDUCE_MAX) { for (int b = b0; b < b1; b++) p = fmaxf(p, r[b]); }
    else { for (int b = b0; b < b1; b++) p += r[b]; p *= a.inv_count / (float)(b1 - b0); }
    const float db = a.db_mode == DBM_MAG_1E10 ? 6.02059991327962f * log2f(sqrtf(p) + 1e-10f) : 3.01029995663981f * log2f(p + 1e-20f);
    const uint32_t px = a.cmap == 1 ? colormap_px<1>(db, a.inv_range, a.cmap_bias) : colormap_px<0>(db, a.inv_range, a.cmap_bias);
    a.out[(size_t)(a.canvas_h - 1 - f) * a.canvas_w + a.col0 + col] = px;            // :1288 y flipped
}

// ---------------- N3 ----------------
enum { PACK_F32 = 0, PACK_I16 = 1 };

struct SeriesSig {
    const double* re;       // planar FP64 rows (downconverter output, Java double[2][n])
    const double* im;
    long long n;
    long long out_off;      // element offset of this signal in the packed / series outputs
};

// IqData.getInterleavedBinary: (float) x, or (short)(32767 * x) with Java's narrowing rules
// (double -> int truncates toward zero and saturates, NaN -> 0; int -> short keeps the low 16 bits).
__global__ void __launch_bounds__(256)
iq_pack_kernel(const SeriesSig* __restrict__ sigs, const int format, void* __restrict__ out) {
    const SeriesSig s = sigs[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (long long)gridDim.x * blockDim.x) {
        const double re = s.re[i], im = s.im[i];
        if (format == PACK_F32) {
            reinterpret_cast<float2*>(out)[s.out_off + i] = make_float2((float)re, (float)im);
        } else {
            const int a = __double2int_rz(32767.0 * re), b = __double2int_rz(32767.0 * im);   // saturating, NaN -> 0
            reinterpret_cast<uint32_t*>(out)[s.out_off + i] = ((uint32_t)a & 0xFFFFu) | ((uint32_t)b << 16);
        }
    }
}

// First-order recurrence y[i] = alpha x[i] + (1 - alpha) y[i-1], y[first] = x[first], as an affine scan:
// every thread folds a contiguous run of kSeriesRun samples into (A, B) with y_end = A y_in + B, the CTA scans
// the 256 pairs, and the run is replayed from its true start value.
constexpr int kSeriesThreads = 256;

struct SeriesArgs {
    const SeriesSig* sigs;
    double sample_rate, alpha_mag, alpha_freq, center_freq;
    double* out_mag_db;     // [out_off + i], i < n     : 20 log10(EMA(hypot))      (:226-240)
    double* out_freq;       // [out_off + i], 1 <= i < n: EMA(inst. freq) + centre  (:263-284); element 0 unused (NaN)
};

__device__ __forceinline__ double series_mag(const SeriesSig& s, long long i) { return hypot(s.re[i], s.im[i]); }
__device__ __forceinline__ double series_freq(const SeriesSig& s, long long i, double fs) {
    const double kPi = 3.14159265358979323846;
    double d = atan2(s.im[i], s.re[i]) - atan2(s.im[i - 1], s.re[i - 1]);
    if (d > kPi) d -= 2.0 * kPi; else if (d < -kPi) d += 2.0 * kPi;       // :268-273
    return d / (2.0 * kPi) * fs;
}

// WHICH 0: magnitude series (first index 0), 1: frequency series (first index 1)
template <int WHICH>
__device__ __forceinline__ void series_scan(const SeriesArgs& a, const SeriesSig& s, double* sh_a, double* sh_b) {
    const long long first = WHICH;
    const double alpha = WHICH == 0 ? a.alpha_mag : a.alpha_freq;
    const double beta = 1.0 - alpha;
    const long long n = s.n - first;                     // samples in the recurrence
    // element 0 of the frequency series is never produced by the Java loop (it starts at i = 1): NaN, also for a
    // one-sample signal, where nothing else is written
    if (WHICH == 1 && threadIdx.x == 0 && s.n > 0) a.out_freq[s.out_off] = __longlong_as_double(0x7ff8000000000000LL);
    if (n <= 0) return;
    const long long run = (n + kSeriesThreads - 1) / kSeriesThreads;
    const long long lo = first + (long long)threadIdx.x * run;
    const long long hi = min(lo + run, s.n);
    // fold the run: y_end = A * y_in + B  (the very first sample has y = x: A = 0)
    double A = 1.0, B = 0.0;
    for (long long i = lo; i < hi; i++) {
        const double x = WHICH == 0 ? series_mag(s, i) : series_freq(s, i, a.sample_rate);
        if (i == first) { A = 0.0; B = x; }
        else { A *= beta; B = __dadd_rn(__dmul_rn(alpha, x), __dmul_rn(beta, B)); }
    }
    sh_a[threadIdx.x] = A; sh_b[threadIdx.x] = B;
    __syncthreads();
    // value entering this thread's run: compose the pairs of all earlier threads (serial over 256 entries,
    // every thread does its own prefix: 256 FMAs, negligible next to the runs)
    double y = 0.0;
    for (int k = 0; k < (int)threadIdx.x; k++) y = sh_a[k] * y + sh_b[k];
    __syncthreads();
    double* out = WHICH == 0 ? a.out_mag_db : a.out_freq;
    for (long long i = lo; i < hi; i++) {
        const double x = WHICH == 0 ? series_mag(s, i) : series_freq(s, i, a.sample_rate);
        y = (i == first) ? x : __dadd_rn(__dmul_rn(alpha, x), __dmul_rn(beta, y));   // Java: no FMA contraction
        out[s.out_off + i] = WHICH == 0 ? 20.0 * log10(y) : y + a.center_freq;
    }
}

__global__ void __launch_bounds__(kSeriesThreads)
series_kernel(const SeriesArgs a) {
    __shared__ double sh_a[kSeriesThreads], sh_b[kSeriesThreads];
    const SeriesSig s = a.sigs[blockIdx.x];
    if (a.out_mag_db) series_scan<0>(a, s, sh_a, sh_b);
    __syncthreads();
    if (a.out_freq) series_scan<1>(a, s, sh_a, sh_b);
}

}  // namespace sa
