// display.cu -- C-ABI entry points of the rows next to the hot path: the display
// canvas (renderSpectrogram, MainController.java:1261-1291), the IqData packers (IqData.java:160-187)
// and the Analysis dialog's magnitude / instantaneous-frequency series
// (AnalysisDialogController.java:219-290).  Kernels in display_kernels.cuh.
#include "engine_internal.h"
#include "display_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace sa;



namespace {

constexpr uint64_t kCanvasChunkBytes = 64ull << 20;     // dB rows per chunk (stays mostly L2 resident)

int check_canvas(const sa_spectrogram_params* p, uint32_t w, uint32_t h, uint64_t fpc, int reduce) {
    if (!p) return set_error(SA_ERR_INVALID_ARG, "params is NULL");
    if (w == 0 || h == 0) return set_error(SA_ERR_INVALID_ARG, "empty canvas %ux%u", w, h);
    if (fpc == 0) return set_error(SA_ERR_INVALID_ARG, "frames_per_column is 0");
    if (reduce < SA_REDUCE_NEAREST || reduce > SA_REDUCE_MEAN) return set_error(SA_ERR_INVALID_ARG, "unknown reduce mode %d", reduce);
    if (!(p->sample_rate > 0.0)) return set_error(SA_ERR_INVALID_ARG, "canvas needs sample_rate > 0");
    if (!(p->max_db > p->min_db)) return set_error(SA_ERR_INVALID_ARG, "canvas needs max_db > min_db");
    if (p->colormap != SA_CMAP_GRAYSCALE && p->colormap != SA_CMAP_HEATMAP) return set_error(SA_ERR_INVALID_ARG, "unknown colormap %d", p->colormap);
    if (fpc * (uint64_t)p->nfft * 4 > (1ull << 30)) return set_error(SA_ERR_UNSUPPORTED, "one canvas column spans more than 1 GiB of dB rows");
    return SA_OK;
}

void fill_canvas_args(CanvasArgs& ca, const sa_spectrogram_params& p, uint32_t w, uint32_t h, uint64_t fpc, int reduce,
                      uint32_t* d_canvas) {
    memset(&ca, 0, sizeof(ca));
    ca.out = d_canvas;
    ca.nfft = (int)p.nfft; ca.fpc = (int)fpc;
    ca.canvas_w = (int)w; ca.canvas_h = (int)h;
    ca.reduce = reduce; ca.cmap = p.colormap;
    const double conv = 10.0 * std::log10(p.sample_rate / (double)p.nfft) + 20.0 * std::log10((double)p.nfft);
    ca.inv_range = (float)(1.0 / (p.max_db - p.min_db));
    ca.cmap_bias = (float)(-(conv + p.min_db) / (p.max_db - p.min_db));
}

// second stage of the fused pooling: acc (|X|^2 reduced over each column's frames) -> pixels of columns [col0, col0 + ncols)
int launch_canvas_power(Engine* eng, const CanvasArgs& ca, const sa_spectrogram_params& p, const float* d_acc, int col0,
                        int ncols, cudaStream_t stream) {
    CanvasPowerArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.acc = d_acc; pa.out = ca.out;
    pa.nfft = ca.nfft; pa.canvas_w = ca.canvas_w; pa.canvas_h = ca.canvas_h; pa.col0 = col0; pa.ncols = ncols;
    pa.reduce = ca.reduce; pa.cmap = ca.cmap; pa.db_mode = p.db_mode;
    pa.inv_count = 1.0f / (float)ca.fpc;
    pa.inv_range = ca.inv_range; pa.cmap_bias = ca.cmap_bias;
    void* args[] = { &pa };
    cudaError_t e = cudaLaunchKernel((const void*)&canvas_power_kernel, dim3((ca.canvas_h + 255) / 256, ncols), dim3(256), args, 0, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch canvas_power_kernel");
    eng->launches++;
    return SA_OK;
}

int launch_canvas(Engine* eng, CanvasArgs& ca, const float* d_db, int col0, int ncols, cudaStream_t stream) {
    ca.db = d_db; ca.col0 = col0; ca.ncols = ncols;
    void* args[] = { &ca };
    cudaError_t e = cudaLaunchKernel((const void*)&canvas_kernel, dim3((ca.canvas_h + kCanvasRows - 1) / kCanvasRows, ncols), dim3(kCanvasRows * kCanvasGroups), args, 0, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch canvas_kernel");
    eng->launches++;
    return SA_OK;
}

// in_bps: bytes per input IQ pair when the chunk's samples are staged too (host path), 0 for device-resident input:
// a chunk is bounded by the larger of its dB rows and its input samples (a view with a large hop reads far more
// than it writes)
uint64_t canvas_cols_per_chunk(const sa_spectrogram_params& p, uint64_t fpc, uint32_t w, uint64_t chunk_bytes, uint64_t in_bps = 0) {
    const uint64_t col_bytes = std::max<uint64_t>(fpc * (uint64_t)p.nfft * 4, fpc * p.hop * in_bps);
    // at most 65535 columns per launch (grid.y)
    return std::max<uint64_t>(1, std::min<uint64_t>(std::min<uint64_t>(w, 65535), chunk_bytes / col_bytes));
}

}  // namespace

extern "C" {

int32_t sa_render_canvas_device(sa_engine* engine, const void* d_iq, uint64_t iq_bytes, const sa_spectrogram_params* params,
                                uint32_t canvas_w, uint32_t canvas_h, uint64_t frames_per_column, int32_t reduce,
                                void* d_out_rgba, void* cuda_stream) {
    ENGINE_ENTER(engine);
    int rc = check_canvas(params, canvas_w, canvas_h, frames_per_column, reduce);
    if (rc) return rc;
    if (!d_out_rgba || (!d_iq && iq_bytes)) return set_error(SA_ERR_INVALID_ARG, "NULL buffer");
    sa_spectrogram_params q = *params;
    q.out_kind = SA_OUT_F32_DB;
    if (reduce == SA_REDUCE_NEAREST) {          // only the first frame of every column is shown: transform only those
        q.hop = params->hop * frames_per_column;
        frames_per_column = 1;
    }
    q.n_frames = (uint64_t)canvas_w * frames_per_column;
    int prec = 0;
    rc = check_spec_params(&q, &prec);
    if (rc) return rc;
    const uint64_t bps = spec_bytes_per_iq(q);
    if ((uintptr_t)d_iq % bps) return set_error(SA_ERR_INVALID_ARG, "d_iq must be aligned to %llu bytes", (unsigned long long)bps);
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    CanvasArgs ca;
    fill_canvas_args(ca, q, canvas_w, canvas_h, frames_per_column, reduce, (uint32_t*)d_out_rgba);
    static const char* fuse_env = getenv("SA_CANVAS_FUSED");
    const bool fuse_ok = !(fuse_env && atoi(fuse_env) == 0);
    if (reduce != SA_REDUCE_NEAREST && fuse_ok && Engine::can_pool(q, prec) && canvas_w <= 65535) {
        // MAX / MEAN: the frame reduction happens in the spectrogram epilogue (no dB rows are ever written); one launch
        // over all frames, then one small kernel over the canvas
        const size_t acc_bytes = (size_t)canvas_w * q.nfft * sizeof(float);
        rc = engine->ensure_scratch(3, acc_bytes);
        if (rc) return rc;
        cudaError_t e = cudaMemsetAsync(engine->scratch[3], 0, acc_bytes, stream);
        if (e != cudaSuccess) return cuda_fail(e, "clear pooling accumulator");
        rc = engine->launch_spectrogram(d_iq, iq_bytes / bps, q, prec, engine->scratch[3], stream, 2, reduce == SA_REDUCE_MAX ? 1 : 2,
                                        frames_per_column);
        if (rc) return rc;
        return launch_canvas_power(engine, ca, q, (const float*)engine->scratch[3], 0, (int)canvas_w, stream);
    }
    const uint64_t cpc = canvas_cols_per_chunk(q, frames_per_column, canvas_w, 4 * kCanvasChunkBytes);
    rc = engine->ensure_scratch(3, cpc * frames_per_column * q.nfft * 4);
    if (rc) return rc;
    for (uint64_t c0 = 0; c0 < canvas_w; c0 += cpc) {
        const uint64_t nc = std::min<uint64_t>(cpc, canvas_w - c0);
        sa_spectrogram_params r = q;
        r.start_sample = q.start_sample + c0 * frames_per_column * q.hop;
        r.n_frames = nc * frames_per_column;
        rc = engine->launch_spectrogram(d_iq, iq_bytes / bps, r, prec, engine->scratch[3], stream);
        if (rc) return rc;
        rc = launch_canvas(engine, ca, (const float*)engine->scratch[3], (int)c0, (int)nc, stream);
        if (rc) return rc;
    }
    return SA_OK;
}

int32_t sa_render_canvas(sa_engine* engine, const void* iq, uint64_t iq_bytes, const sa_spectrogram_params* params,
                         uint32_t canvas_w, uint32_t canvas_h, uint64_t frames_per_column, int32_t reduce, void* out_rgba) {
    ENGINE_ENTER(engine);
    int rc = check_canvas(params, canvas_w, canvas_h, frames_per_column, reduce);
    if (rc) return rc;
    if (!out_rgba || (!iq && iq_bytes)) return set_error(SA_ERR_INVALID_ARG, "NULL buffer");
    sa_spectrogram_params q = *params;
    q.out_kind = SA_OUT_F32_DB;
    if (reduce == SA_REDUCE_NEAREST) {          // only the first frame of every column is shown: transform only those
        q.hop = params->hop * frames_per_column;
        frames_per_column = 1;
    }
    q.n_frames = (uint64_t)canvas_w * frames_per_column;
    int prec = 0;
    rc = check_spec_params(&q, &prec);
    if (rc) return rc;
    const uint64_t bps = spec_bytes_per_iq(q);
    const uint64_t n_samples = iq_bytes / bps;
    const uint64_t cpc = canvas_cols_per_chunk(q, frames_per_column, canvas_w, kCanvasChunkBytes, bps);
    const uint64_t fpchunk = cpc * frames_per_column;
    // never more than the capture holds past the view's start
    const uint64_t readable = n_samples > q.start_sample ? n_samples - q.start_sample : 0;
    const uint64_t in_cap = std::max<uint64_t>(16, std::min<uint64_t>((fpchunk - 1) * q.hop + q.nfft, readable) * bps);
    const uint64_t out_cap = fpchunk * q.nfft * 4;
    const size_t canvas_bytes = (size_t)canvas_w * canvas_h * 4;
    rc = engine->ensure_scratch(3, canvas_bytes);
    if (rc) return rc;
    CanvasArgs ca;
    fill_canvas_args(ca, q, canvas_w, canvas_h, frames_per_column, reduce, (uint32_t*)engine->scratch[3]);
    // sparse frames (hop >= 2 nfft, e.g. the nearest-frame quick look of a long recording): only the frames'
    // own samples are packed into pinned memory and cross PCIe, not the gaps between them
    const uint64_t total_frames = (uint64_t)canvas_w * frames_per_column;
    const uint64_t frame_bytes = (uint64_t)q.nfft * bps;
    if (q.hop >= 2ull * q.nfft && total_frames * frame_bytes <= (256ull << 20) && canvas_w <= 65535) {
        uint64_t nr = 0;                                    // readable frames are a prefix (MainController.java:987)
        while (nr < total_frames && q.start_sample + nr * q.hop + q.nfft <= n_samples) nr++;
        Slot& s = engine->slots[0];
        rc = engine->ensure_slot(s, std::max<uint64_t>(nr * frame_bytes, 16), total_frames * q.nfft * 4);
        if (rc) return rc;
        rc = engine->ensure_staging(s, std::max<uint64_t>(nr * frame_bytes, 16), 0);
        if (rc) return rc;
        for (uint64_t t = 0; t < nr; t++)
            memcpy((char*)s.h_in + t * frame_bytes, (const char*)iq + (q.start_sample + t * q.hop) * bps, frame_bytes);
        cudaError_t e2 = cudaSuccess;
        // nothing may still be in flight on the staging buffer / the caller's memory when an error returns
        auto fail = [&](int code) { cudaStreamSynchronize(s.stream); return code; };
        if (nr) e2 = cudaMemcpyAsync(s.d_in, s.h_in, nr * frame_bytes, cudaMemcpyHostToDevice, s.stream);
        if (e2 != cudaSuccess) return fail(cuda_fail(e2, "H2D packed frames"));
        sa_spectrogram_params r = q;
        r.start_sample = 0; r.hop = q.nfft; r.n_frames = total_frames;
        rc = engine->launch_spectrogram(s.d_in, nr * q.nfft, r, prec, s.d_out, s.stream, 5);
        if (rc) return fail(rc);
        rc = launch_canvas(engine, ca, (const float*)s.d_out, 0, (int)canvas_w, s.stream);
        if (rc) return fail(rc);
        e2 = cudaMemcpyAsync(out_rgba, engine->scratch[3], canvas_bytes, cudaMemcpyDeviceToHost, s.stream);
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(s.stream);
        if (e2 != cudaSuccess) return fail(cuda_fail(e2, "sparse canvas"));
        return SA_OK;
    }
    // chunk c: H2D of its samples -> spectrogram -> canvas columns, on slot c % kSlots; only the canvas comes back
    static const char* fuse_env = getenv("SA_CANVAS_FUSED");
    const bool fused = reduce != SA_REDUCE_NEAREST && !(fuse_env && atoi(fuse_env) == 0) && Engine::can_pool(q, prec);
    const bool in_pinned = host_ptr_is_pinned(iq);
    uint64_t c = 0;
    cudaError_t e = cudaSuccess;
    for (uint64_t c0 = 0; c0 < canvas_w && rc == SA_OK; c0 += cpc, c++) {
        Slot& s = engine->slots[c % kSlots];
        rc = engine->ensure_slot(s, in_cap, out_cap);
        if (rc) break;
        rc = engine->ensure_staging(s, in_pinned ? 0 : in_cap, 0);
        if (rc) break;
        e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess) { rc = cuda_fail(e, "slot sync"); break; }
        const uint64_t nc = std::min<uint64_t>(cpc, canvas_w - c0);
        const uint64_t nf = nc * frames_per_column;
        const uint64_t s_begin = q.start_sample + c0 * frames_per_column * q.hop;
        uint64_t s_end = s_begin + (nf - 1) * q.hop + q.nfft;
        if (s_end > n_samples) s_end = n_samples;
        const uint64_t ns = s_end > s_begin ? s_end - s_begin : 0;
        if (ns) {
            const void* src = (const char*)iq + s_begin * bps;
            if (!in_pinned) { engine->host_copy(s.h_in, src, ns * bps); src = s.h_in; }
            e = cudaMemcpyAsync(s.d_in, src, ns * bps, cudaMemcpyHostToDevice, s.stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "H2D"); break; }
        }
        sa_spectrogram_params r = q;
        r.start_sample = 0;
        r.n_frames = nf;
        if (fused) {                       // the chunk's columns are reduced in the spectrogram epilogue (no dB rows)
            e = cudaMemsetAsync(s.d_out, 0, (size_t)nc * q.nfft * sizeof(float), s.stream);
            if (e != cudaSuccess) { rc = cuda_fail(e, "clear pooling accumulator"); break; }
            rc = engine->launch_spectrogram(s.d_in, ns, r, prec, s.d_out, s.stream, 5 + (int)(c % kSlots),
                                            reduce == SA_REDUCE_MAX ? 1 : 2, frames_per_column);
            if (rc) break;
            rc = launch_canvas_power(engine, ca, q, (const float*)s.d_out, (int)c0, (int)nc, s.stream);
            continue;
        }
        rc = engine->launch_spectrogram(s.d_in, ns, r, prec, s.d_out, s.stream, 5 + (int)(c % kSlots));
        if (rc) break;
        rc = launch_canvas(engine, ca, (const float*)s.d_out, (int)c0, (int)nc, s.stream);
    }
    // drain on every path: the staging buffers and the caller's samples must not be in flight when this returns
    for (int i = 0; i < kSlots; i++)
        if (engine->slots[i].stream) {
            e = cudaStreamSynchronize(engine->slots[i].stream);
            if (e != cudaSuccess && rc == SA_OK) rc = cuda_fail(e, "canvas pipeline drain");
        }
    if (rc) return rc;
    e = cudaMemcpy(out_rgba, engine->scratch[3], canvas_bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return cuda_fail(e, "D2H canvas");
    return SA_OK;
}

// ---------------- N3: IqData packers and analysis series ----------------

static int upload_sigs(Engine* eng, const std::vector<SeriesSig>& sigs, SeriesSig** d_sigs, cudaStream_t stream) {
    const size_t bytes = (sigs.size() * sizeof(SeriesSig) + 255) & ~(size_t)255;
    int rc = eng->ensure_scratch(4, bytes);
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(eng->scratch[4], sigs.data(), sigs.size() * sizeof(SeriesSig), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return cuda_fail(e, "upload signal list");
    *d_sigs = (SeriesSig*)eng->scratch[4];
    return SA_OK;
}

static int build_sigs(const double* d_rows, const uint64_t* row_offsets, const uint64_t* lengths, const uint64_t* out_offsets,
                      uint32_t n_sig, std::vector<SeriesSig>& sigs) {
    if (!d_rows || !row_offsets || !lengths || !out_offsets) return set_error(SA_ERR_INVALID_ARG, "NULL argument");
    sigs.resize(n_sig);
    for (uint32_t i = 0; i < n_sig; i++) {
        sigs[i].re = d_rows + row_offsets[i];
        sigs[i].im = sigs[i].re + lengths[i];
        sigs[i].n = (long long)lengths[i];
        sigs[i].out_off = (long long)out_offsets[i];
    }
    return SA_OK;
}

static int iq_pack_device(Engine* engine, const double* d_rows, const uint64_t* row_offsets, const uint64_t* lengths,
                          const uint64_t* out_offsets, uint32_t n_sig, int32_t format, void* d_out, cudaStream_t stream) {
    if (format != SA_PACK_F32 && format != SA_PACK_I16)      // IqData.java:185-186 IllegalArgumentException
        return set_error(SA_ERR_INVALID_ARG, "Unsupported binary format: %d", format);
    if (n_sig == 0) return SA_OK;
    if (!d_out) return set_error(SA_ERR_INVALID_ARG, "d_out is NULL");
    std::vector<SeriesSig> sigs;
    int rc = build_sigs(d_rows, row_offsets, lengths, out_offsets, n_sig, sigs);
    if (rc) return rc;
    SeriesSig* d_sigs = nullptr;
    rc = upload_sigs(engine, sigs, &d_sigs, stream);
    if (rc) return rc;
    uint64_t max_n = 0;
    for (const auto& s : sigs) max_n = std::max<uint64_t>(max_n, (uint64_t)s.n);
    const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((max_n + 255) / 256, 4096));
    int fmt = format;
    void* args[] = { &d_sigs, &fmt, &d_out };
    cudaError_t e = cudaLaunchKernel((const void*)&iq_pack_kernel, dim3(gx, n_sig), dim3(256), args, 0, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch iq_pack_kernel");
    engine->launches++;
    return SA_OK;
}

static int series_device(Engine* engine, const double* d_rows, const uint64_t* row_offsets, const uint64_t* lengths,
                         const uint64_t* out_offsets, uint32_t n_sig, double sample_rate, double alpha_mag,
                         double alpha_freq, double center_freq, double* d_out_mag_db, double* d_out_freq, cudaStream_t stream) {
    if (n_sig == 0) return SA_OK;
    if (!d_out_mag_db && !d_out_freq) return set_error(SA_ERR_INVALID_ARG, "both outputs are NULL");
    std::vector<SeriesSig> sigs;
    int rc = build_sigs(d_rows, row_offsets, lengths, out_offsets, n_sig, sigs);
    if (rc) return rc;
    SeriesSig* d_sigs = nullptr;
    rc = upload_sigs(engine, sigs, &d_sigs, stream);
    if (rc) return rc;
    SeriesArgs sa_;
    sa_.sigs = d_sigs; sa_.sample_rate = sample_rate; sa_.alpha_mag = alpha_mag; sa_.alpha_freq = alpha_freq;
    sa_.center_freq = center_freq; sa_.out_mag_db = d_out_mag_db; sa_.out_freq = d_out_freq;
    void* args[] = { &sa_ };
    cudaError_t e = cudaLaunchKernel((const void*)&series_kernel, dim3(n_sig), dim3(kSeriesThreads), args, 0, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch series_kernel");
    engine->launches++;
    return SA_OK;
}

// host-buffer forms: one signal, rows copied to the device, result copied back
static int stage_rows(Engine* eng, const double* re, const double* im, uint64_t n, size_t out_bytes, Slot** slot) {
    Slot& s = eng->slots[0];
    int rc = eng->ensure_slot(s, std::max<size_t>(2 * n * 8, 16), std::max<size_t>(out_bytes, 16));
    if (rc) return rc;
    cudaError_t e = cudaMemcpyAsync(s.d_in, re, n * 8, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync((double*)s.d_in + n, im, n * 8, cudaMemcpyHostToDevice, s.stream);
    if (e != cudaSuccess) { cudaStreamSynchronize(s.stream); return cuda_fail(e, "H2D rows"); }
    *slot = &s;
    return SA_OK;
}

int32_t sa_iq_pack_batch_device(sa_engine* engine, const double* d_rows, const uint64_t* row_offsets, const uint64_t* lengths,
                                const uint64_t* out_offsets, uint32_t n_sig, int32_t format, void* d_out, void* cuda_stream) {
    ENGINE_ENTER(engine);
    return iq_pack_device(engine, d_rows, row_offsets, lengths, out_offsets, n_sig, format, d_out, (cudaStream_t)cuda_stream);
}

int32_t sa_analysis_series_batch_device(sa_engine* engine, const double* d_rows, const uint64_t* row_offsets,
                                        const uint64_t* lengths, const uint64_t* out_offsets, uint32_t n_sig,
                                        double sample_rate, double alpha_mag, double alpha_freq, double center_freq,
                                        double* d_out_mag_db, double* d_out_freq, void* cuda_stream) {
    ENGINE_ENTER(engine);
    return series_device(engine, d_rows, row_offsets, lengths, out_offsets, n_sig, sample_rate, alpha_mag, alpha_freq,
                         center_freq, d_out_mag_db, d_out_freq, (cudaStream_t)cuda_stream);
}

int32_t sa_iq_pack(sa_engine* engine, const double* re, const double* im, uint64_t n, int32_t format, void* out) {
    ENGINE_ENTER(engine);
    if (format != SA_PACK_F32 && format != SA_PACK_I16)
        return set_error(SA_ERR_INVALID_ARG, "Unsupported binary format: %d", format);
    if (n == 0) return SA_OK;
    if (!re || !im || !out) return set_error(SA_ERR_INVALID_ARG, "NULL buffer");
    const size_t out_bytes = (size_t)n * (format == SA_PACK_F32 ? 8 : 4);
    Slot* s = nullptr;
    int rc = stage_rows(engine, re, im, n, out_bytes, &s);
    if (rc) return rc;
    const uint64_t zero = 0;
    rc = iq_pack_device(engine, (const double*)s->d_in, &zero, &n, &zero, 1, format, s->d_out, s->stream);
    if (rc) { cudaStreamSynchronize(s->stream); return rc; }      // the H2D of the caller's rows must have drained
    cudaError_t e = cudaMemcpyAsync(out, s->d_out, out_bytes, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { cudaStreamSynchronize(s->stream); return cuda_fail(e, "iq pack"); }
    return SA_OK;
}

int32_t sa_analysis_series(sa_engine* engine, const double* re, const double* im, uint64_t n, double sample_rate,
                           double alpha_mag, double alpha_freq, double center_freq, double* out_mag_db, double* out_freq) {
    ENGINE_ENTER(engine);
    if (n == 0) return SA_OK;
    if (!re || !im) return set_error(SA_ERR_INVALID_ARG, "NULL buffer");
    if (!out_mag_db && !out_freq) return set_error(SA_ERR_INVALID_ARG, "both outputs are NULL");
    Slot* s = nullptr;
    int rc = stage_rows(engine, re, im, n, (size_t)n * 16, &s);
    if (rc) return rc;
    const uint64_t zero = 0;
    double* d_mag = out_mag_db ? (double*)s->d_out : nullptr;
    double* d_frq = out_freq ? (double*)s->d_out + n : nullptr;
    rc = series_device(engine, (const double*)s->d_in, &zero, &n, &zero, 1, sample_rate, alpha_mag, alpha_freq, center_freq,
                       d_mag, d_frq, s->stream);
    if (rc) { cudaStreamSynchronize(s->stream); return rc; }      // the H2D of the caller's rows must have drained
    cudaError_t e = cudaSuccess;
    if (out_mag_db) e = cudaMemcpyAsync(out_mag_db, d_mag, n * 8, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess && out_freq) e = cudaMemcpyAsync(out_freq, d_frq, n * 8, cudaMemcpyDeviceToHost, s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) { cudaStreamSynchronize(s->stream); return cuda_fail(e, "analysis series"); }
    return SA_OK;
}

}  // extern "C"
