/*
 * sa_oracle.h -- CPU oracle for the spectral_analyzer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product path
 * (spectral_analyzer_b200/, include/) links, imports or calls this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the reference (GassiusODude/spectral_analyzer v0.5.0) is Java 21 +
 * commons-math3 3.6.1 + JDSP v1.3.1; no JVM exists in the build container, and the
 * reference ships no golden vectors or numeric tests
 * (src/test/.../SpectralAnalyzerApplicationTests.java:20-23 is a Spring contextLoads).
 * The spectrogram part is a line-by-line restatement of source that IS present
 * (SpectralService.java:33-85, MainController.java:980-999, :926-957, :1261-1291)
 * plus the textbook forward DFT that commons-math3 FastFourierTransformer(STANDARD)
 * implements.  The downconvert / Welch parts restate a self-defined spec because the
 * JDSP sources (Resampler, PowerSpectralDensity) are not vendored in the reference.
 *
 * All arithmetic is FP64, like the reference.
 */
#ifndef SA_ORACLE_H
#define SA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* datatype ids: SigMF core:datatype prefixes (Global.java:67-79) */
enum { ORA_CF32 = 0, ORA_CI16 = 1, ORA_CU8 = 2, ORA_CI8 = 3, ORA_CF64 = 4,
       ORA_OTHER = 5 /* no decode branch in the reference (ri16_le ...): zeros / cf32 fall-through, strict_reference only */ };
/* window ids (periodic / DFT-even definitions) */
enum { ORA_WIN_RECT = 0, ORA_WIN_HANN = 1, ORA_WIN_HAMMING = 2, ORA_WIN_BLACKMAN = 3, ORA_WIN_BLACKMAN_HARRIS = 4 };
/* dB modes */
enum { ORA_DB_MAG_1E10 = 0 /* 20*log10(|X|+1e-10), SpectralService.java:80-81 */,
       ORA_DB_POWER    = 1 /* 10*log10(|X|^2+1e-20) */ };
/* colormaps (MainController.java:939-956) */
enum { ORA_CMAP_GRAYSCALE = 0, ORA_CMAP_HEATMAP = 1 };

/* bytes per IQ pair (Global.java:67-79); 0 for unknown id */
int ora_bytes_per_iq(int dtype);

/* window coefficients, periodic form, FP64 */
int ora_window(int window_id, int n, double* w);

/* decode `count` IQ pairs starting at byte offset start_byte.
 * SpectralService.java:40-65 / ExtractDownConvertService.java:74-100.
 * strict_reference != 0 reproduces the reference's cf64 behaviour in
 * computeMagnitudes (no cf64 branch -> zeros, SpectralService.java:60-63). */
int ora_decode(const uint8_t* buf, uint64_t cap_bytes, uint64_t start_byte, uint64_t count,
               int dtype, int big_endian, int strict_reference, double* re, double* im);

/* SpectralService.computeMagnitudes (SpectralService.java:33-85): one frame,
 * rect window, unnormalised forward DFT, 20*log10(abs+1e-10), fftshift. */
int ora_compute_magnitudes(const uint8_t* buf, uint64_t cap_bytes, uint64_t start_byte, int nfft,
                           int dtype, int big_endian, int strict_reference, double* out);

/* MainController.updateDisplay frame loop (MainController.java:980-999) generalised:
 * frame t covers samples [start_sample + t*hop, +nfft); frames that would read past
 * cap_bytes become a row of -150.0 (:994-998).  window=RECT, hop=nfft, db_mode=MAG_1E10
 * is exactly the reference.  out is [n_frames][nfft] FP64.  nthreads<=0: all cores. */
int ora_spectrogram(const uint8_t* buf, uint64_t cap_bytes, int dtype, int big_endian,
                    uint64_t start_sample, int nfft, uint64_t hop, int window_id,
                    uint64_t n_frames, int db_mode, double* out, int nthreads);

/* renderSpectrogram + getColorForMagnitude (MainController.java:1261-1291, :926-957).
 * Full-resolution variant: one RGBA8 pixel per (frame, shifted bin), same layout as the
 * dB image; bytes are R,G,B,A with A=255; channel = floor(c*255+0.5). */
int ora_render_rgba(const double* db, uint64_t n_frames, int nfft, double fs, double min_db,
                    double max_db, int cmap, uint8_t* rgba);
/* Canvas variant: pixel(t, H-1-f) <- color(db[t][(int)((double)f/H*nfft)]) (:1276-1289);
 * rgba is [canvas_h][n_frames] row-major (y, x). */
int ora_render_canvas(const double* db, uint64_t n_frames, int nfft, int canvas_h, double fs,
                      double min_db, double max_db, int cmap, uint8_t* rgba);

/* ---- self-defined spec (JDSP not vendored; parity unpinned) ---- */
/* low-pass taps used by the conventional downconverter: Hamming-windowed sinc,
 * ntaps = 8*down+1, cutoff 0.5/down cycles/sample, unity DC gain. */
int ora_lowpass_taps(int down, double* taps /* 8*down+1 */);
/* ExtractDownConvertService.extractAndDownConvert (ExtractDownConvertService.java:54-117):
 * decode count samples from start_sample, mix y[n]=x[n]*exp(-2*pi*i*freq_off*n) (n from 0
 * at the first extracted sample), then
 *   fast=0: causal FIR (taps above, zero history) and keep every down-th: out[m]=z[m*down]
 *   fast=1: out[m] = mean(y[m*down .. m*down+down-1])
 * out_len = count/down (floor).  out_re/out_im hold out_len doubles each. */
int ora_downconvert(const uint8_t* buf, uint64_t cap_bytes, int dtype, int big_endian,
                    uint64_t start_sample, uint64_t count, double freq_off, int down, int fast,
                    double* out_re, double* out_im, uint64_t* out_len);
/* PowerSpectralDensity.calculatePsdWelch call site (AnalysisDialogController.java:303-313):
 * Hann (periodic) segments of length nfft, hop = nfft*(1-overlap) (overlap 0.75 -> nfft/4),
 * segments K = 1 + (n-nfft)/hop, two-sided, psd[k] = mean_K |FFT(w*x)|^2 / (fs*sum(w^2)),
 * fftshifted; out_freq[k] = (k - nfft/2)*fs/nfft ; out_db[k] = 10*log10(psd + 1e-30).
 * nfft must be a power of two and <= n. */
int ora_psd_welch(const double* re, const double* im, uint64_t n, double fs, int nfft,
                  uint64_t hop, int window_id, double* out_freq, double* out_db);

/* The choices JDSP makes inside Resampler.downConvert / calculatePsdWelch that the reference does not pin
 * (build.gradle:142, not vendored): filter taps, delay compensation, output length, PSD scaling and detrending.
 * NULL / all-zero = the self-defined default spec above.  strict_reference reproduces the decode fall-throughs of
 * ExtractDownConvertService.java:60-67,79-81,93-96 (cf64 read at an 8-byte stride; unknown datatypes read as cf32). */
typedef struct ora_analysis_cfg {
    const double* taps; int n_taps;   /* NULL: ora_lowpass_taps(down) */
    int delay_mode;                   /* 0 causal, 1 same ((L-1)/2 samples compensated), 2 valid (full overlaps only) */
    int length_mode;                  /* 0 floor(count/down), 1 ceil(count/down) */
    int psd_scaling;                  /* 0 density 1/(fs sum w^2), 1 spectrum 1/(sum w)^2 */
    int psd_detrend;                  /* 0 none, 1 constant (per-segment mean removed) */
    int strict_reference;
} ora_analysis_cfg;
uint64_t ora_downconvert_length(uint64_t count, int down, int fast, const ora_analysis_cfg* cfg);
int ora_downconvert_ex(const uint8_t* buf, uint64_t cap_bytes, int dtype, int big_endian,
                       uint64_t start_sample, uint64_t count, double freq_off, int down, int fast,
                       const ora_analysis_cfg* cfg, double* out_re, double* out_im, uint64_t* out_len);
/* any nfft >= 1 (the reference's short-signal branch: nfft = signal length, AnalysisDialogController.java:304-307) */
int ora_psd_welch_ex(const double* re, const double* im, uint64_t n, double fs, int nfft,
                     uint64_t hop, int window_id, const ora_analysis_cfg* cfg, double* out_freq, double* out_db);

/* IqData.getInterleavedBinary (S/data/IqData.java:160-187): format 0 float32, 1 int16, little-endian. */
int ora_iq_pack(const double* re, const double* im, uint64_t n, int format, uint8_t* out);
/* AnalysisDialogController.updateMagnitudeChart / updateFrequencyChart (:219-290). */
int ora_analysis_series(const double* re, const double* im, uint64_t n, double fs, double alpha_mag,
                        double alpha_freq, double center_freq, double* out_mag_db, double* out_freq);

/* in-place forward DFT, unnormalised, power-of-two n (radix-2, FP64) */
int ora_fft(double* re, double* im, int n);

int ora_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
