/*
 * sa_oracle.c -- CPU (FP64) restatement of the spectral_analyzer hot path.
 *
 * TEST INFRASTRUCTURE ONLY -- see sa_oracle.h.  PARITY UNPINNED (no JVM here, the
 * reference holds no golden vectors); every function cites the reference lines it follows.
 * Paths: S/ = src/main/java/net/kcundercover/spectral_analyzer/ in the reference repo.
 */
#include "sa_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* S/sigmf/Global.java:67-79 */
int ora_bytes_per_iq(int dtype) {
    switch (dtype) {
        case ORA_CF32: return 8;
        case ORA_CI16: return 4;
        case ORA_CU8:  return 2;
        case ORA_CI8:  return 2;
        case ORA_CF64: return 16;
        default:       return 0;
    }
}

int ora_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ora_window(int window_id, int n, double* w) {
    if (n <= 0) return -1;
    for (int i = 0; i < n; i++) {
        double x = 2.0 * M_PI * (double)i / (double)n;
        switch (window_id) {
            case ORA_WIN_RECT:     w[i] = 1.0; break;
            case ORA_WIN_HANN:     w[i] = 0.5 - 0.5 * cos(x); break;
            case ORA_WIN_HAMMING:  w[i] = 0.54 - 0.46 * cos(x); break;
            case ORA_WIN_BLACKMAN: w[i] = 0.42 - 0.5 * cos(x) + 0.08 * cos(2 * x); break;
            case ORA_WIN_BLACKMAN_HARRIS:
                w[i] = 0.35875 - 0.48829 * cos(x) + 0.14128 * cos(2 * x) - 0.01168 * cos(3 * x);
                break;
            default: return -1;
        }
    }
    return 0;
}

/* ByteBuffer.getShort/getFloat/getDouble honour the buffer order, which
 * S/sigmf/SigMfHelper.java:87-91 sets to LE iff datatype ends with "_le". */
static inline uint16_t rd16(const uint8_t* p, int be) {
    return be ? (uint16_t)((p[0] << 8) | p[1]) : (uint16_t)((p[1] << 8) | p[0]);
}
static inline uint32_t rd32(const uint8_t* p, int be) {
    return be ? ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]
              : ((uint32_t)p[3] << 24) | ((uint32_t)p[2] << 16) | ((uint32_t)p[1] << 8) | p[0];
}
static inline uint64_t rd64(const uint8_t* p, int be) {
    uint64_t v = 0;
    if (be) for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
    else    for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
    return v;
}

/* one IQ pair; S/services/SpectralService.java:42-63 and
 * S/services/ExtractDownConvertService.java:79-98 */
static inline void decode_one(const uint8_t* p, int dtype, int be, int strict, double* re, double* im) {
    switch (dtype) {
        case ORA_CI16: { /* SpectralService.java:44-45 */
            int16_t a = (int16_t)rd16(p, be), b = (int16_t)rd16(p + 2, be);
            *re = a / 32768.0; *im = b / 32768.0; break; }
        case ORA_CF32: { /* :48-49 */
            uint32_t a = rd32(p, be), b = rd32(p + 4, be); float fa, fb;
            memcpy(&fa, &a, 4); memcpy(&fb, &b, 4);
            *re = (double)fa; *im = (double)fb; break; }
        case ORA_CU8: { /* :51-54 */
            double a = (double)(p[0] & 0xFF), b = (double)(p[1] & 0xFF);
            *re = (a - 127.5) / 128; *im = (b - 127.5) / 128; break; }
        case ORA_CI8: { /* :56-59 */
            double a = (double)(int8_t)p[0], b = (double)(int8_t)p[1];
            *re = a / 128; *im = b / 128; break; }
        case ORA_CF64: { /* no branch in SpectralService (:60-63 -> 0.0); correct stride-16
                            decode otherwise (ExtractDownConvertService.java:79-81 minus its
                            stride bug) */
            if (strict) { *re = 0.0; *im = 0.0; break; }
            uint64_t a = rd64(p, be), b = rd64(p + 8, be);
            memcpy(re, &a, 8); memcpy(im, &b, 8); break; }
        default: *re = 0.0; *im = 0.0;
    }
}

int ora_decode(const uint8_t* buf, uint64_t cap_bytes, uint64_t start_byte, uint64_t count,
               int dtype, int big_endian, int strict_reference, double* re, double* im) {
    int bps = ora_bytes_per_iq(dtype);
    if (bps == 0) return -1;
    if (start_byte + count * (uint64_t)bps > cap_bytes) return -2; /* IndexOutOfBounds in Java */
    for (uint64_t i = 0; i < count; i++)
        decode_one(buf + start_byte + i * bps, dtype, big_endian, strict_reference, &re[i], &im[i]);
    return 0;
}

/* X[k] = sum_n x[n] exp(-2 pi i k n / N): commons-math3 3.6.1
 * FastFourierTransformer(DftNormalization.STANDARD).transform(.., FORWARD), called at
 * S/services/SpectralService.java:23,68.  Iterative radix-2 DIT, FP64, twiddles from
 * cos/sin per stage entry (no recurrence), bit-reversal first. */
int ora_fft(double* re, double* im, int n) {
    if (n <= 0 || (n & (n - 1))) return -1; /* MathIllegalArgumentException in Java */
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1;
        for (int k = 0; k < half; k++) {
            double ang = -2.0 * M_PI * (double)k / (double)len;
            double wr = cos(ang), wi = sin(ang);
            for (int b = k; b < n; b += len) {
                int o = b + half;
                double tr = wr * re[o] - wi * im[o], ti = wr * im[o] + wi * re[o];
                re[o] = re[b] - tr; im[o] = im[b] - ti;
                re[b] += tr;        im[b] += ti;
            }
        }
    }
    return 0;
}

/* a twiddle-table variant for the timed baseline: identical arithmetic, table built once */
typedef struct { int n; double* wr; double* wi; int* rev; } fft_plan;
static int plan_init(fft_plan* p, int n) {
    p->n = n;
    p->wr = (double*)malloc(sizeof(double) * n);
    p->wi = (double*)malloc(sizeof(double) * n);
    p->rev = (int*)malloc(sizeof(int) * n);
    if (!p->wr || !p->wi || !p->rev) return -1;
    /* table laid out per stage: entries [len/2 .. len) hold W_len^k, k = 0..len/2-1 */
    for (int len = 2; len <= n; len <<= 1)
        for (int k = 0; k < len / 2; k++) {
            double ang = -2.0 * M_PI * (double)k / (double)len;
            p->wr[len / 2 + k] = cos(ang); p->wi[len / 2 + k] = sin(ang);
        }
    p->rev[0] = 0;
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit; p->rev[i] = j;
    }
    return 0;
}
static void plan_free(fft_plan* p) { free(p->wr); free(p->wi); free(p->rev); }
static void plan_fft(const fft_plan* p, double* re, double* im) {
    int n = p->n;
    for (int i = 1; i < n; i++) {
        int j = p->rev[i];
        if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1;
        const double* wr = p->wr + half; const double* wi = p->wi + half;
        for (int b = 0; b < n; b += len)
            for (int k = 0; k < half; k++) {
                int a = b + k, o = a + half;
                double tr = wr[k] * re[o] - wi[k] * im[o], ti = wr[k] * im[o] + wi[k] * re[o];
                re[o] = re[a] - tr; im[o] = im[a] - ti;
                re[a] += tr;        im[a] += ti;
            }
    }
}

/* S/services/SpectralService.java:73-82 */
static inline double to_db(double xr, double xi, int db_mode) {
    if (db_mode == ORA_DB_MAG_1E10) return 20.0 * log10(hypot(xr, xi) + 1e-10);
    return 10.0 * log10(xr * xr + xi * xi + 1e-20);
}

int ora_compute_magnitudes(const uint8_t* buf, uint64_t cap_bytes, uint64_t start_byte, int nfft,
                           int dtype, int big_endian, int strict_reference, double* out) {
    if (nfft <= 0 || (nfft & (nfft - 1))) return -1;
    double* re = (double*)malloc(sizeof(double) * nfft);
    double* im = (double*)malloc(sizeof(double) * nfft);
    if (!re || !im) { free(re); free(im); return -3; }
    /* an unknown datatype decodes to zeros in the reference (:60-63); dtype ids are closed
       here, so only the strict cf64 case takes that path */
    int rc = ora_decode(buf, cap_bytes, start_byte, (uint64_t)nfft, dtype, big_endian,
                        strict_reference, re, im);
    if (rc == 0) {
        ora_fft(re, im, nfft);
        int half = nfft / 2;
        for (int i = 0; i < nfft; i++)                       /* :76-82 */
            out[(i + half) % nfft] = to_db(re[i], im[i], ORA_DB_MAG_1E10);
    }
    free(re); free(im);
    return rc;
}

/* S/controllers/MainController.java:980-999, generalised to hop/window/dB mode */
int ora_spectrogram(const uint8_t* buf, uint64_t cap_bytes, int dtype, int big_endian,
                    uint64_t start_sample, int nfft, uint64_t hop, int window_id,
                    uint64_t n_frames, int db_mode, double* out, int nthreads) {
    int bps = ora_bytes_per_iq(dtype);
    if (bps == 0 || nfft <= 0 || (nfft & (nfft - 1)) || hop == 0) return -1;
    double* w = (double*)malloc(sizeof(double) * nfft);
    fft_plan plan;
    if (!w || ora_window(window_id, nfft, w) || plan_init(&plan, nfft)) { free(w); return -1; }
    int half = nfft / 2;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        double* re = (double*)malloc(sizeof(double) * nfft);
        double* im = (double*)malloc(sizeof(double) * nfft);
#pragma omp for schedule(static)
        for (int64_t t = 0; t < (int64_t)n_frames; t++) {
            double* row = out + (uint64_t)t * nfft;
            uint64_t sample = start_sample + (uint64_t)t * hop;       /* :984 */
            uint64_t byte_off = sample * (uint64_t)bps;               /* :985 (64-bit here) */
            if (byte_off + (uint64_t)nfft * bps <= cap_bytes) {       /* :987 */
                for (int i = 0; i < nfft; i++) {
                    decode_one(buf + byte_off + (uint64_t)i * bps, dtype, big_endian, 0, &re[i], &im[i]);
                    if (window_id != ORA_WIN_RECT) { re[i] *= w[i]; im[i] *= w[i]; }
                }
                plan_fft(&plan, re, im);
                for (int i = 0; i < nfft; i++) row[(i + half) % nfft] = to_db(re[i], im[i], db_mode);
            } else {
                for (int i = 0; i < nfft; i++) row[i] = -150.0;       /* :996-997 */
            }
        }
        free(re); free(im);
    }
    plan_free(&plan); free(w);
    return 0;
}

/* getColorForMagnitude, S/controllers/MainController.java:926-957.  JavaFX Color keeps
 * float components and Color.interpolate works in float [JavaFX, not vendored]; the
 * PixelWriter packs channels with Math.round(c*255.0) = floor(c*255+0.5). */
static inline uint8_t chan(float c) { return (uint8_t)floor((double)c * 255.0 + 0.5); }
static inline float lerpf(float a, float b, double t) {
    if (t <= 0.0) return a;
    if (t >= 1.0) return b;
    float ft = (float)t;
    return a + (b - a) * ft;
}
static void color_for(double db, double min_db, double max_db, int cmap, uint8_t* px) {
    double n = (db - min_db) / (max_db - min_db);                     /* :929 */
    n = n < 0.0 ? 0.0 : (n > 1.0 ? 1.0 : n);                          /* :930 */
    if (n != n) n = 0.0;
    float r, g, b;
    if (cmap == ORA_CMAP_HEATMAP) {                                   /* :944-953 */
        if (n < 0.2) { r = g = b = 0.f; }
        else if (n < 0.5) { double u = (n - 0.2) / 0.3; r = lerpf(0.f, 1.f, u); g = 0.f; b = lerpf(1.f, 0.f, u); }
        else { double u = (n - 0.5) / 0.5; r = 1.f; g = lerpf(0.f, 1.f, u); b = 0.f; }
    } else {                                                          /* :939-942, :954-955 */
        r = g = b = lerpf(0.f, 1.f, n);
    }
    px[0] = chan(r); px[1] = chan(g); px[2] = chan(b); px[3] = 255;
}

/* dB/bin -> dB/Hz, MainController.java:1273-1274 */
static inline double render_conversion(double fs, int nfft) {
    return 10.0 * log10(fs / nfft) + 20.0 * log10((double)nfft);
}

int ora_render_rgba(const double* db, uint64_t n_frames, int nfft, double fs, double min_db,
                    double max_db, int cmap, uint8_t* rgba) {
    double conv = render_conversion(fs, nfft);
    for (uint64_t i = 0; i < n_frames * (uint64_t)nfft; i++)
        color_for(db[i] - conv, min_db, max_db, cmap, rgba + 4 * i);   /* :1283-1285 */
    return 0;
}

int ora_render_canvas(const double* db, uint64_t n_frames, int nfft, int canvas_h, double fs,
                      double min_db, double max_db, int cmap, uint8_t* rgba) {
    double conv = render_conversion(fs, nfft);
    for (uint64_t t = 0; t < n_frames; t++)                            /* :1276 */
        for (int f = 0; f < canvas_h; f++) {                           /* :1277 */
            int bin = (int)((double)f / canvas_h * nfft);              /* :1280 */
            double v = db[t * (uint64_t)nfft + bin] - conv;            /* :1283 */
            int y = canvas_h - 1 - f;                                  /* :1288 */
            color_for(v, min_db, max_db, cmap, rgba + 4 * ((uint64_t)y * n_frames + t));
        }
    return 0;
}

/* ---------- self-defined downconvert / Welch spec (JDSP not vendored) ---------- */

int ora_lowpass_taps(int down, double* taps) {
    if (down < 1) return -1;
    int nt = 8 * down + 1, mid = 4 * down;
    double fc = 0.5 / down, sum = 0.0;          /* cycles/sample */
    for (int k = 0; k < nt; k++) {
        double x = (double)(k - mid);
        double s = (k == mid) ? 2.0 * fc : sin(2.0 * M_PI * fc * x) / (M_PI * x);
        double w = 0.54 - 0.46 * cos(2.0 * M_PI * (double)k / (double)(nt - 1));   /* symmetric Hamming */
        taps[k] = s * w; sum += taps[k];
    }
    for (int k = 0; k < nt; k++) taps[k] /= sum;
    return 0;
}

/* ---- configurable spec: every choice JDSP makes that the reference does not pin ----
 * Output m of the conventional path:  z[m] = sum_{k<L} h[k] y[m*D + off - k],  y = mixed samples, y[n] = 0 outside
 * [0, count);  off = 0 (causal), (L-1)/2 (same: group delay compensated), L-1 (valid: only full overlaps).
 * Output count: valid -> (count-L)/D + 1 (0 if count < L); otherwise floor(count/D) or ceil(count/D).
 * Fast path: out[m] = (1/D) sum_{k<D} y[m*D + k] (delay modes do not apply; valid counts as floor). */
uint64_t ora_downconvert_length(uint64_t count, int down, int fast, const ora_analysis_cfg* cfg) {
    const int L = (cfg && cfg->taps) ? cfg->n_taps : 8 * down + 1;
    const int delay = cfg ? cfg->delay_mode : 0, len = cfg ? cfg->length_mode : 0;
    if (!fast && delay == 2) return count >= (uint64_t)L ? (count - (uint64_t)L) / (uint64_t)down + 1 : 0;
    return len == 1 ? (count + (uint64_t)down - 1) / (uint64_t)down : count / (uint64_t)down;
}

int ora_downconvert_ex(const uint8_t* buf, uint64_t cap_bytes, int dtype, int big_endian,
                       uint64_t start_sample, uint64_t count, double freq_off, int down, int fast,
                       const ora_analysis_cfg* cfg, double* out_re, double* out_im, uint64_t* out_len) {
    const int strict = cfg ? cfg->strict_reference : 0;
    if (down < 1) return -1;
    double* re = (double*)malloc(sizeof(double) * (count ? count : 1));
    double* im = (double*)malloc(sizeof(double) * (count ? count : 1));
    if (!re || !im) { free(re); free(im); return -3; }
    int rc = 0;
    if (strict && (dtype == ORA_CF64 || ora_bytes_per_iq(dtype) == 0)) {
        /* ExtractDownConvertService.java:60-67: bytesPerIQ falls to 8 for everything that is not ci16 / cu8 / ci8;
         * :79-81 cf64 then reads two doubles at that 8-byte stride (re = d[i], im = d[i+1]);
         * :93-96 any other datatype is read as two floats */
        const uint64_t start_byte = start_sample * 8u;
        const uint64_t need = dtype == ORA_CF64 ? (count ? (count + 1) * 8u : 0) : count * 8u;
        if (start_byte + need > cap_bytes) rc = -2;
        for (uint64_t i = 0; i < count && !rc; i++) {
            const uint8_t* p = buf + start_byte + i * 8u;
            if (dtype == ORA_CF64) {
                uint64_t a = rd64(p, big_endian), b = rd64(p + 8, big_endian);
                memcpy(&re[i], &a, 8); memcpy(&im[i], &b, 8);
            } else {
                decode_one(p, ORA_CF32, big_endian, 0, &re[i], &im[i]);
            }
        }
    } else {
        int bps = ora_bytes_per_iq(dtype);
        if (bps == 0) rc = -1;
        /* ExtractDownConvertService.java:67,74-100 (cf64 decoded with the correct stride) */
        else rc = ora_decode(buf, cap_bytes, start_sample * (uint64_t)bps, count, dtype, big_endian, 0, re, im);
    }
    if (rc) { free(re); free(im); return rc; }
    /* NCO mix; phase reduced mod 1 in FP64 before the trig call */
    for (uint64_t n = 0; n < count; n++) {
        double ph = fmod(freq_off * (double)n, 1.0);
        double c = cos(2.0 * M_PI * ph), s = -sin(2.0 * M_PI * ph);
        double a = re[n], b = im[n];
        re[n] = a * c - b * s; im[n] = a * s + b * c;
    }
    const uint64_t m_out = ora_downconvert_length(count, down, fast, cfg);
    if (fast) {                                                /* :104-106 "moving average" */
        for (uint64_t m = 0; m < m_out; m++) {
            double sr = 0.0, si = 0.0;
            for (int k = 0; k < down; k++) {
                const uint64_t n = m * (uint64_t)down + (uint64_t)k;
                if (n < count) { sr += re[n]; si += im[n]; }
            }
            out_re[m] = sr / down; out_im[m] = si / down;
        }
    } else {                                                   /* :109-112 "LPF - downconvert" */
        const int nt = (cfg && cfg->taps) ? cfg->n_taps : 8 * down + 1;
        double* h = (double*)malloc(sizeof(double) * (size_t)nt);
        if (cfg && cfg->taps) memcpy(h, cfg->taps, sizeof(double) * (size_t)nt); else ora_lowpass_taps(down, h);
        const int delay = cfg ? cfg->delay_mode : 0;
        const int64_t off = delay == 1 ? (nt - 1) / 2 : delay == 2 ? nt - 1 : 0;
        for (uint64_t m = 0; m < m_out; m++) {
            double sr = 0.0, si = 0.0;
            const int64_t n0 = (int64_t)(m * (uint64_t)down) + off;
            for (int k = 0; k < nt; k++) {
                const int64_t n = n0 - k;
                if (n < 0) break;
                if ((uint64_t)n >= count) continue;
                sr += h[k] * re[n]; si += h[k] * im[n];
            }
            out_re[m] = sr; out_im[m] = si;
        }
        free(h);
    }
    *out_len = m_out;
    free(re); free(im);
    return 0;
}

int ora_downconvert(const uint8_t* buf, uint64_t cap_bytes, int dtype, int big_endian,
                    uint64_t start_sample, uint64_t count, double freq_off, int down, int fast,
                    double* out_re, double* out_im, uint64_t* out_len) {
    return ora_downconvert_ex(buf, cap_bytes, dtype, big_endian, start_sample, count, freq_off, down, fast, NULL,
                              out_re, out_im, out_len);
}

/* Welch PSD, any nfft >= 1 (the reference's short-signal branch passes nfft = signal length,
 * AnalysisDialogController.java:304-307): power-of-two lengths through the radix-2 plan, everything else through
 * the direct O(N^2) DFT with exact twiddle indices.  Segment count 1 + (n - nfft)/hop; optional per-segment mean
 * removal (detrend = constant); density 1/(K fs sum w^2) or spectrum 1/(K (sum w)^2) scaling; two-sided; bin i
 * lands at (i + nfft/2) % nfft (numpy fftshift, also for odd nfft); freq[k] = (k - nfft/2) fs / nfft. */
int ora_psd_welch_ex(const double* re, const double* im, uint64_t n, double fs, int nfft,
                     uint64_t hop, int window_id, const ora_analysis_cfg* cfg, double* out_freq, double* out_db) {
    if (nfft <= 0 || (uint64_t)nfft > n || hop == 0) return -1;
    const int pow2 = (nfft & (nfft - 1)) == 0 && nfft >= 2;
    double* w = (double*)malloc(sizeof(double) * nfft);
    double* a = (double*)malloc(sizeof(double) * nfft);
    double* b = (double*)malloc(sizeof(double) * nfft);
    double* acc = (double*)calloc(nfft, sizeof(double));
    double* tr = NULL; double* ti = NULL; double* xr = NULL; double* xi = NULL;
    fft_plan plan;
    if (!w || !a || !b || !acc || ora_window(window_id, nfft, w)) return -3;
    if (pow2) { if (plan_init(&plan, nfft)) return -3; }
    else {
        tr = (double*)malloc(sizeof(double) * nfft); ti = (double*)malloc(sizeof(double) * nfft);
        xr = (double*)malloc(sizeof(double) * nfft); xi = (double*)malloc(sizeof(double) * nfft);
        if (!tr || !ti || !xr || !xi) return -3;
        for (int j = 0; j < nfft; j++) { tr[j] = cos(2.0 * M_PI * j / nfft); ti[j] = -sin(2.0 * M_PI * j / nfft); }
    }
    double sw = 0.0, sw2 = 0.0;
    for (int i = 0; i < nfft; i++) { sw += w[i]; sw2 += w[i] * w[i]; }
    const int detrend = cfg ? cfg->psd_detrend : 0, scaling = cfg ? cfg->psd_scaling : 0;
    uint64_t nseg = 1 + (n - (uint64_t)nfft) / hop;
    for (uint64_t s = 0; s < nseg; s++) {
        double mr = 0.0, mi = 0.0;
        if (detrend) {
            for (int i = 0; i < nfft; i++) { mr += re[s * hop + i]; mi += im[s * hop + i]; }
            mr /= nfft; mi /= nfft;
        }
        for (int i = 0; i < nfft; i++) { a[i] = (re[s * hop + i] - mr) * w[i]; b[i] = (im[s * hop + i] - mi) * w[i]; }
        if (pow2) {
            plan_fft(&plan, a, b);
        } else {
            for (int k = 0; k < nfft; k++) {
                double sr = 0.0, si = 0.0;
                int idx = 0;
                for (int i = 0; i < nfft; i++) {
                    sr += a[i] * tr[idx] - b[i] * ti[idx];
                    si += a[i] * ti[idx] + b[i] * tr[idx];
                    idx += k; if (idx >= nfft) idx -= nfft;
                }
                xr[k] = sr; xi[k] = si;
            }
            memcpy(a, xr, sizeof(double) * nfft); memcpy(b, xi, sizeof(double) * nfft);
        }
        for (int i = 0; i < nfft; i++) acc[i] += a[i] * a[i] + b[i] * b[i];
    }
    int half = nfft / 2;
    double scale = scaling == 1 ? 1.0 / ((double)nseg * sw * sw) : 1.0 / ((double)nseg * fs * sw2);
    for (int i = 0; i < nfft; i++) {
        int k = (i + half) % nfft;
        out_db[k] = 10.0 * log10(acc[i] * scale + 1e-30);
    }
    for (int k = 0; k < nfft; k++) out_freq[k] = ((double)k - half) * fs / nfft;
    if (pow2) plan_free(&plan);
    free(w); free(a); free(b); free(acc); free(tr); free(ti); free(xr); free(xi);
    return 0;
}

int ora_psd_welch(const double* re, const double* im, uint64_t n, double fs, int nfft,
                  uint64_t hop, int window_id, double* out_freq, double* out_db) {
    if (nfft <= 0 || (nfft & (nfft - 1))) return -1;
    return ora_psd_welch_ex(re, im, n, fs, nfft, hop, window_id, NULL, out_freq, out_db);
}

/* Java narrowing (short)(double): JLS 5.1.3 double -> int (NaN -> 0, saturate, else truncate
 * toward zero), then int -> short keeps the low 16 bits. */
static int16_t java_double_to_short(double v) {
    int32_t i;
    if (v != v) i = 0;
    else if (v >= 2147483647.0) i = INT32_MAX;
    else if (v <= -2147483648.0) i = INT32_MIN;
    else i = (int32_t)v;
    return (int16_t)(uint16_t)((uint32_t)i & 0xFFFFu);
}

/* IqData.getInterleavedBinary (IqData.java:160-187): format 0 "float32": putFloat((float) I),
 * putFloat((float) Q), little-endian (:163-171); format 1 "int16": putShort((short)(32767 * I)), ... (:173-183). */
int ora_iq_pack(const double* re, const double* im, uint64_t n, int format, uint8_t* out) {
    if (format != 0 && format != 1) return -1;                  /* IllegalArgumentException :185-186 */
    for (uint64_t i = 0; i < n; i++) {
        if (format == 0) {
            float a = (float)re[i], b = (float)im[i];
            uint32_t ua, ub;
            memcpy(&ua, &a, 4); memcpy(&ub, &b, 4);
            for (int k = 0; k < 4; k++) { out[8 * i + k] = (uint8_t)(ua >> (8 * k)); out[8 * i + 4 + k] = (uint8_t)(ub >> (8 * k)); }
        } else {
            uint16_t a = (uint16_t)java_double_to_short(32767 * re[i]);
            uint16_t b = (uint16_t)java_double_to_short(32767 * im[i]);
            out[4 * i] = (uint8_t)a; out[4 * i + 1] = (uint8_t)(a >> 8);
            out[4 * i + 2] = (uint8_t)b; out[4 * i + 3] = (uint8_t)(b >> 8);
        }
    }
    return 0;
}

/* AnalysisDialogController.updateMagnitudeChart (:219-251) and updateFrequencyChart (:256-290).
 * out_mag_db[i] = 20*log10(valueOld_i) (the chart skips non-finite values; they are kept here),
 * out_freq[i] = vOld_i + center_freq for i >= 1, out_freq[0] = NaN (the loop starts at 1). */
int ora_analysis_series(const double* re, const double* im, uint64_t n, double fs, double alpha_mag,
                        double alpha_freq, double center_freq, double* out_mag_db, double* out_freq) {
    double value_old = 0.0;
    if (out_mag_db)
        for (uint64_t i = 0; i < n; i++) {
            double abs_value = hypot(re[i], im[i]);                                  /* :231 */
            if (i == 0) value_old = abs_value;                                         /* :234-235 */
            else value_old = alpha_mag * abs_value + (1 - alpha_mag) * value_old;      /* :237 */
            out_mag_db[i] = 20 * log10(value_old);                                     /* :240 */
        }
    if (out_freq) {
        double v_old = 0.0;
        if (n > 0) out_freq[0] = NAN;
        for (uint64_t i = 1; i < n; i++) {
            double phase1 = atan2(im[i], re[i]);                                       /* :264 */
            double phase2 = atan2(im[i - 1], re[i - 1]);                               /* :265 */
            double d = phase1 - phase2;
            if (d > M_PI) d -= 2 * M_PI; else if (d < -M_PI) d += 2 * M_PI;            /* :268-273 */
            double inst = (d / (2 * M_PI)) * fs;                                       /* :275 */
            if (i == 1) v_old = inst; else v_old = (alpha_freq * inst) + (1 - alpha_freq) * v_old;   /* :276-280 */
            out_freq[i] = v_old + center_freq;                                         /* :281 */
        }
    }
    return 0;
}

