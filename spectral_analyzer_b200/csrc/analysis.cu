// analysis.cu -- annotation-analysis path behind the C-ABI: NCO downconvert + FIR decimate and
// Welch PSD (kernels in analysis_kernels.cuh).  Replaces ExtractDownConvertService.java:54-117,
// the batch loop AnnotationController.java:321-360 and the calculatePsdWelch call at
// AnalysisDialogController.java:308-312.
#include "engine_internal.h"
#include "analysis_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <tuple>

using namespace sa;


namespace {

const double kPi = 3.14159265358979323846264338327950288;

// Hamming-windowed sinc, ntaps = 8*down+1, cutoff 0.5/down cycles/sample, unity DC gain
void lowpass_taps(int down, std::vector<double>& h) {
    const int nt = 8 * down + 1, mid = 4 * down;
    const double fc = 0.5 / down;
    h.resize(nt);
    double sum = 0.0;
    for (int k = 0; k < nt; k++) {
        const double x = (double)(k - mid);
        const double s = (k == mid) ? 2.0 * fc : std::sin(2.0 * kPi * fc * x) / (kPi * x);
        const double w = 0.54 - 0.46 * std::cos(2.0 * kPi * (double)k / (double)(nt - 1));
        h[k] = s * w;
        sum += h[k];
    }
    for (int k = 0; k < nt; k++) h[k] /= sum;
}

struct WelchKernel { const void* fn; int n, cta, fpc; size_t smem; int p, np, radix[4]; };

template <int N> WelchKernel make_welch() {
    using G = Geo<float, N>;
    using PL = Plan<float, N>;
    WelchKernel k;
    k.fn = (const void*)&welch_accum_kernel<N>;
    k.n = N; k.cta = G::CTA; k.fpc = G::FPC; k.smem = G::SMEM_BYTES + G::TW_BYTES; k.p = G::P; k.np = PL::NP;
    for (int i = 0; i < 4; i++) k.radix[i] = PL::radix(i);
    return k;
}

const WelchKernel* find_welch(int n) {
    static const WelchKernel tab[] = { make_welch<64>(), make_welch<128>(), make_welch<256>(), make_welch<512>(),
                                       make_welch<1024>(), make_welch<2048>(), make_welch<4096>(),
                                       make_welch<8192>(), make_welch<16384>() };
    for (const auto& k : tab) if (k.n == n) return &k;
    return nullptr;
}

// which: 0 staged kernel with taps from global memory, 1 staged kernel with taps in the parameter bank, 2 wide
template <int DK> const void* dc_kernel_of(int which) {
    return which == 2 ? (const void*)&downconvert_wide_kernel<DK>
         : which == 1 ? (const void*)&downconvert_kernel<DK, true> : (const void*)&downconvert_kernel<DK, false>;
}
const void* dc_kernel(int dk, int which) {
    switch (dk) {
        case DK_CF32: return dc_kernel_of<DK_CF32>(which);
        case DK_CI16: return dc_kernel_of<DK_CI16>(which);
        case DK_C8:   return dc_kernel_of<DK_C8>(which);
        default:      return dc_kernel_of<DK_CF64>(which);
    }
}

size_t dc_smem_bytes(int down, int nb, int fast) {
    const int nblk = nb + (fast ? 0 : 7);
    const size_t n_stage = fast ? (size_t)nb * down : (size_t)nblk * down + 1;
    const size_t stage_phys = n_stage + n_stage / down + 2;
    return ((stage_phys + 1) & ~(size_t)1) * sizeof(float2) + (fast ? 0 : (size_t)nblk * 8 * sizeof(float2));
}

struct BatchPlan {
    std::vector<DcAnn> anns;
    std::vector<float> taps;            // concatenated per distinct down
    std::vector<long long> m_out;
    long long total_out = 0;            // doubles in the decimated-IQ output
};

int validate_anns(const sa_annotation* anns, uint32_t n_ann, uint64_t n_samples) {
    if (!anns && n_ann) return set_error(SA_ERR_INVALID_ARG, "annotations is NULL");
    for (uint32_t i = 0; i < n_ann; i++) {
        if (anns[i].down < 1) return set_error(SA_ERR_INVALID_ARG, "annotation %u: down %d < 1", i, anns[i].down);
        if (!std::isfinite(anns[i].freq_off)) return set_error(SA_ERR_INVALID_ARG, "annotation %u: freq_off not finite", i);
        if (anns[i].start_sample > n_samples || anns[i].count > n_samples - anns[i].start_sample)
            return set_error(SA_ERR_OUT_OF_RANGE, "annotation %u: [%llu, +%llu) exceeds %llu samples", i,
                             (unsigned long long)anns[i].start_sample, (unsigned long long)anns[i].count,
                             (unsigned long long)n_samples);
    }
    return SA_OK;
}

}  // namespace

namespace sa {

struct WelchJob { const double* re; const double* im; long long n; double fs; };

// Welch PSD of FP64 planar rows already on the device; d_out_psd is [jobs][nfft] doubles.
static int welch_device(Engine* eng, const std::vector<WelchJob>& jobs, uint32_t nfft, uint64_t hop, int window,
                        double* d_out_psd, cudaStream_t stream) {
    const WelchKernel* wk = find_welch((int)nfft);
    if (!wk) return set_error(SA_ERR_UNSUPPORTED, "no Welch kernel for nfft %u (power of two, 64..16384)", nfft);
    const uint32_t n_sig = (uint32_t)jobs.size();
    std::vector<double> w;
    host_window(window, (int)nfft, w);
    double sw2 = 0.0;
    for (double v : w) sw2 += v * v;
    std::vector<WelchSig> sigs(n_sig);
    long long max_seg = 0;
    for (uint32_t i = 0; i < n_sig; i++) {
        WelchSig& s = sigs[i];
        s.re = jobs[i].re; s.im = jobs[i].im; s.n = jobs[i].n;
        s.nseg = s.n >= (long long)nfft ? 1 + (s.n - nfft) / (long long)hop : 0;
        s.scale = s.nseg > 0 ? 1.0 / ((double)s.nseg * jobs[i].fs * sw2) : 0.0;
        max_seg = std::max(max_seg, s.nseg);
    }
    // enough CTAs for about two waves; every CTA walks its share of the segments
    const long long want = (4LL * eng->num_sms + n_sig - 1) / n_sig;
    const int nsplit = (int)std::max<long long>(1, std::min<long long>(want, (max_seg + wk->fpc - 1) / wk->fpc));
    const size_t sig_bytes = ((size_t)n_sig * sizeof(WelchSig) + 255) & ~(size_t)255;
    const size_t part_bytes = (size_t)n_sig * nsplit * wk->fpc * nfft * sizeof(float);
    int rc = eng->ensure_scratch(1, sig_bytes + part_bytes);
    if (rc) return rc;
    WelchSig* d_sigs = (WelchSig*)eng->scratch[1];
    cudaError_t e = cudaMemcpyAsync(d_sigs, sigs.data(), sigs.size() * sizeof(WelchSig), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return cuda_fail(e, "upload Welch plan");
    WelchArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.sigs = d_sigs;
    wa.hop = (long long)hop;
    SpecKernelInfo ki;
    memset(&ki, 0, sizeof(ki));
    ki.prec = SA_PREC_F32; ki.n = wk->n; ki.p = wk->p; ki.np = wk->np;
    for (int i = 0; i < 4; i++) ki.radix[i] = wk->radix[i];
    rc = eng->twiddle_table(ki, &wa.twiddle);
    if (rc) return rc;
    const void* wtab = nullptr;
    rc = eng->window_table(window, (int)nfft, SA_PREC_F32, &wtab);
    if (rc) return rc;
    wa.window = (const float*)wtab;
    wa.partial = (float*)((char*)eng->scratch[1] + sig_bytes);
    wa.nsplit = nsplit;
    wa.out_db = d_out_psd;
    int bps = 0;
    rc = eng->kernel_grid(wk->fn, wk->cta, wk->smem, &bps);
    if (rc) return rc;
    void* wargs[] = { &wa };
    e = cudaLaunchKernel(wk->fn, dim3(nsplit, n_sig), dim3(wk->cta), wargs, wk->smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch welch_accum_kernel");
    eng->launches++;
    int n = (int)nfft, slots = nsplit * wk->fpc;
    void* fargs[] = { &wa, &n, &slots };
    e = cudaLaunchKernel((const void*)&welch_finalize_kernel, dim3((n + 255) / 256, n_sig), dim3(256), fargs, 0, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch welch_finalize_kernel");
    eng->launches++;
    return SA_OK;
}

// Runs the downconverter (and optionally the Welch PSD) for a batch on DEVICE samples.
static int run_batch_device(Engine* eng, const void* d_iq, uint64_t n_samples, int dtype, int big_endian,
                            double sample_rate, const sa_annotation* anns, uint32_t n_ann, uint32_t psd_nfft,
                            uint64_t psd_hop, int psd_window, double* d_out_iq, const uint64_t* iq_offsets,
                            double* d_out_psd, cudaStream_t stream) {
    const int dk = dtype_kind(dtype);
    std::vector<DcAnn> plan(n_ann);
    std::vector<float> taps;
    std::map<int, int> taps_off;
    for (uint32_t i = 0; i < n_ann; i++) {
        DcAnn& a = plan[i];
        const int D = anns[i].down;
        a.start_sample = (long long)anns[i].start_sample;
        a.count = (long long)anns[i].count;
        const double fr = anns[i].freq_off - std::floor(anns[i].freq_off);    // frac in [0,1)
        a.phase_step = (unsigned long long)std::ldexp((long double)fr, 64);   // exact 64-bit phase increment
        a.m_out = a.count / D;
        a.out_off = (long long)iq_offsets[i];
        a.down = D;
        a.fast = anns[i].fast ? 1 : 0;
        if (!taps_off.count(D)) {
            // block layout: [pad so that ht is 16-byte aligned][h[0..8D]][ht[r*8+p] = h[D*p+r]]
            while ((taps.size() + 8 * (size_t)D + 1) % 4) taps.push_back(0.f);
            taps_off[D] = (int)taps.size();
            std::vector<double> h;
            lowpass_taps(D, h);
            for (double v : h) taps.push_back((float)v);
            for (int r = 0; r < D; r++) for (int p = 0; p < 8; p++) taps.push_back((float)h[D * p + r]);
        }
        a.taps_off = taps_off[D];
        a.qmagic = D > 1 ? (unsigned)(((1ull << 32) + (unsigned long long)D - 1) / (unsigned long long)D) : 0u;
        const bool wide = a.fast ? (D > kDcStage / 8) : (D > kDcMaxDown);
        if (wide) {
            a.nb = 0;                       // marks the warp-per-output kernel
        } else {
            const int nblk = std::max(8, std::min(kDcThreads, kDcStage / D));
            a.nb = a.fast ? nblk : nblk - 7;
        }
    }
    // launches are grouped by (kernel, fast, down): the device list is the plan sorted by that key
    std::vector<uint32_t> order(n_ann);
    for (uint32_t i = 0; i < n_ann; i++) order[i] = i;
    auto key = [&](uint32_t i) { return std::make_tuple(plan[i].nb == 0 ? 1 : 0, plan[i].fast, plan[i].down); };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return key(x) < key(y); });
    std::vector<DcAnn> sorted(n_ann);
    for (uint32_t i = 0; i < n_ann; i++) sorted[i] = plan[order[i]];
    const size_t ann_bytes = (sorted.size() * sizeof(DcAnn) + 255) & ~(size_t)255;
    const size_t taps_bytes = (taps.size() * sizeof(float) + 255) & ~(size_t)255;
    int rc = eng->ensure_scratch(0, ann_bytes + taps_bytes);
    if (rc) return rc;
    DcAnn* d_anns = (DcAnn*)eng->scratch[0];
    float* d_taps = (float*)((char*)eng->scratch[0] + ann_bytes);
    cudaError_t e = cudaMemcpyAsync(d_anns, sorted.data(), sorted.size() * sizeof(DcAnn), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_taps, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return cuda_fail(e, "upload annotation plan");

    DcArgs da;
    memset(&da, 0, sizeof(da));
    fill_load_params(da.lp, d_iq, dtype, big_endian);
    da.n_samples = (long long)n_samples;
    da.anns = d_anns;
    da.taps = d_taps;
    da.out = d_out_iq;
    DcTapParams tp;                      // copied into the launch by cudaLaunchKernel (kernel-parameter bank)
    tp.h_last = 0.f; tp.down = 0;
    void* args[] = { &da, &tp };
    for (uint32_t g0 = 0; g0 < n_ann;) {
        uint32_t g1 = g0 + 1;
        while (g1 < n_ann && key(order[g1]) == key(order[g0])) g1++;
        const DcAnn& first = sorted[g0];
        da.ann_base = (int)g0;
        if (first.nb == 0) {              // warp-per-output kernel: all wide annotations in one launch
            g1 = n_ann;
            long long tiles = 0;
            for (uint32_t i = g0; i < g1; i++) tiles = std::max<long long>(tiles, (sorted[i].m_out + 7) / 8);
            if (tiles > 0) {
                e = cudaLaunchKernel(dc_kernel(dk, 2), dim3((unsigned)tiles, g1 - g0), dim3(256), args, 0, stream);
                if (e != cudaSuccess) return cuda_fail(e, "launch downconvert_wide_kernel");
                eng->launches++;
            }
        } else {
            long long tiles = 0;
            for (uint32_t i = g0; i < g1; i++) tiles = std::max<long long>(tiles, (sorted[i].m_out + first.nb - 1) / first.nb);
            const int D = first.down;
            const bool ptaps = !first.fast && D <= kDcParamMaxDown;
            if (ptaps) {
                const float* h = taps.data() + first.taps_off;
                tp.h_last = h[8 * D];
                tp.down = D;
                memcpy(tp.ht, h + 8 * D + 1, sizeof(float) * 8 * (size_t)D);
            }
            const size_t smem = dc_smem_bytes(D, first.nb, first.fast);
            const void* fn = dc_kernel(dk, ptaps ? 1 : 0);
            if (tiles > 0) {
                e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return cuda_fail(e, "downconvert smem attribute");
                e = cudaLaunchKernel(fn, dim3((unsigned)tiles, g1 - g0), dim3(kDcThreads), args, smem, stream);
                if (e != cudaSuccess) return cuda_fail(e, "launch downconvert_kernel");
                eng->launches++;
            }
        }
        g0 = g1;
    }
    if (!d_out_psd) return SA_OK;
    std::vector<WelchJob> jobs(n_ann);
    for (uint32_t i = 0; i < n_ann; i++) {
        jobs[i].re = d_out_iq + plan[i].out_off;
        jobs[i].im = jobs[i].re + plan[i].m_out;
        jobs[i].n = plan[i].m_out;
        jobs[i].fs = sample_rate / (double)plan[i].down;
    }
    return welch_device(eng, jobs, psd_nfft, psd_hop, psd_window, d_out_psd, stream);
}

}  // namespace sa


static int check_psd(uint32_t nfft, uint64_t* hop, int32_t window) {
    if (nfft == 0 || (nfft & (nfft - 1))) return set_error(SA_ERR_INVALID_ARG, "psd nfft %u is not a power of two", nfft);
    if (nfft < 64 || nfft > 16384) return set_error(SA_ERR_UNSUPPORTED, "psd nfft %u outside 64..16384", nfft);
    if (*hop == 0) *hop = nfft / 4;           // 75 % overlap
    if (window < SA_WIN_RECT || window > SA_WIN_BLACKMAN_HARRIS) return set_error(SA_ERR_INVALID_ARG, "unknown window %d", window);
    return SA_OK;
}

extern "C" {

int32_t sa_lowpass_taps(int32_t down, double* taps) {
    if (down < 1 || !taps) return set_error(SA_ERR_INVALID_ARG, "down < 1 or taps NULL");
    std::vector<double> h;
    lowpass_taps(down, h);
    memcpy(taps, h.data(), h.size() * sizeof(double));
    return SA_OK;
}

int32_t sa_downconvert_psd_batch_device(sa_engine* engine, const void* d_iq, uint64_t iq_bytes, int32_t dtype,
                                        int32_t big_endian, double sample_rate, const sa_annotation* anns,
                                        uint32_t n_ann, uint32_t psd_nfft, uint64_t psd_hop, int32_t psd_window,
                                        double* d_out_iq, const uint64_t* iq_offsets, double* d_out_psd_db,
                                        void* cuda_stream) {
    ENGINE_ENTER(engine);
    const uint64_t bps = (uint64_t)sa_bytes_per_iq(dtype);
    if (!bps) return set_error(SA_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    if (n_ann == 0) return SA_OK;
    if (!d_out_iq || !iq_offsets) return set_error(SA_ERR_INVALID_ARG, "d_out_iq / iq_offsets is NULL (the PSD reads the decimated rows)");
    if ((uintptr_t)d_iq % bps) return set_error(SA_ERR_INVALID_ARG, "d_iq must be aligned to %llu bytes", (unsigned long long)bps);
    int rc = validate_anns(anns, n_ann, iq_bytes / bps);
    if (rc) return rc;
    if (d_out_psd_db) { rc = check_psd(psd_nfft, &psd_hop, psd_window); if (rc) return rc; }
    return run_batch_device(engine, d_iq, iq_bytes / bps, dtype, big_endian, sample_rate, anns, n_ann, psd_nfft, psd_hop,
                            psd_window, d_out_iq, iq_offsets, d_out_psd_db, (cudaStream_t)cuda_stream);
}

int32_t sa_downconvert_psd_batch(sa_engine* engine, const void* iq, uint64_t iq_bytes, int32_t dtype, int32_t big_endian,
                                 double sample_rate, const sa_annotation* anns, uint32_t n_ann, uint32_t psd_nfft,
                                 uint64_t psd_hop, int32_t psd_window, double* out_iq, const uint64_t* iq_offsets,
                                 double* out_psd_db) {
    ENGINE_ENTER(engine);
    const uint64_t bps = (uint64_t)sa_bytes_per_iq(dtype);
    if (!bps) return set_error(SA_ERR_INVALID_ARG, "unknown dtype %d", dtype);
    if (n_ann == 0) return SA_OK;
    if (!iq) return set_error(SA_ERR_INVALID_ARG, "iq is NULL");
    if (out_iq && !iq_offsets) return set_error(SA_ERR_INVALID_ARG, "iq_offsets is NULL");
    int rc = validate_anns(anns, n_ann, iq_bytes / bps);
    if (rc) return rc;
    if (out_psd_db) { rc = check_psd(psd_nfft, &psd_hop, psd_window); if (rc) return rc; }

    // Only the annotated spans cross PCIe: each annotation's samples are copied into a packed
    // device buffer and the plan is rebased onto it.
    Slot& s = engine->slots[0];
    std::vector<sa_annotation> local(anns, anns + n_ann);
    std::vector<uint64_t> dev_off(n_ann);
    uint64_t in_samples = 0, out_doubles = 0;
    for (uint32_t i = 0; i < n_ann; i++) {
        local[i].start_sample = in_samples;
        in_samples += anns[i].count;
        dev_off[i] = out_doubles;
        out_doubles += 2 * (anns[i].count / (uint64_t)anns[i].down);
    }
    const size_t psd_bytes = out_psd_db ? (size_t)n_ann * psd_nfft * sizeof(double) : 0;
    rc = engine->ensure_slot(s, std::max<size_t>(in_samples * bps, 16), std::max<size_t>(out_doubles * 8 + psd_bytes, 16));
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    uint64_t pos = 0;
    if (host_ptr_is_pinned(iq)) {
        for (uint32_t i = 0; i < n_ann && e == cudaSuccess; i++) {
            if (anns[i].count)
                e = cudaMemcpyAsync((char*)s.d_in + pos * bps, (const char*)iq + anns[i].start_sample * bps,
                                    anns[i].count * bps, cudaMemcpyHostToDevice, s.stream);
            pos += anns[i].count;
        }
    } else {
        // pageable capture (mmapped file): pieces of <= 32 MiB alternate between the pinned staging buffers of
        // slots 1 and 2; an event per buffer says when its previous H2D has drained
        const size_t piece = 32u << 20;
        Slot* st[2] = { &engine->slots[1], &engine->slots[2] };
        struct Events {                                      // destroyed on every exit path
            cudaEvent_t ev[2] = { nullptr, nullptr };
            ~Events() { for (auto x : ev) if (x) { cudaEventSynchronize(x); cudaEventDestroy(x); } }
        } evs;
        cudaEvent_t* ev = evs.ev;
        for (int k = 0; k < 2 && e == cudaSuccess; k++) {
            rc = engine->ensure_staging(*st[k], piece, 0);
            if (rc) return rc;
            e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
        }
        int k = 0;
        for (uint32_t i = 0; i < n_ann && e == cudaSuccess; i++) {
            const uint64_t total = anns[i].count * bps;
            for (uint64_t off = 0; off < total && e == cudaSuccess; off += piece, k ^= 1) {
                const size_t nb = (size_t)std::min<uint64_t>(piece, total - off);
                e = cudaEventSynchronize(ev[k]);                 // a never-recorded event is complete
                if (e != cudaSuccess) break;
                engine->host_copy(st[k]->h_in, (const char*)iq + anns[i].start_sample * bps + off, nb);
                e = cudaMemcpyAsync((char*)s.d_in + pos * bps + off, st[k]->h_in, nb, cudaMemcpyHostToDevice, s.stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev[k], s.stream);
            }
            pos += anns[i].count;
        }
    }
    // on failure nothing of the caller's may still be in flight when this returns
    auto fail = [&](int code) { cudaStreamSynchronize(s.stream); return code; };
    if (e != cudaSuccess) return fail(cuda_fail(e, "H2D annotation spans"));
    double* d_iq_out = (double*)s.d_out;
    double* d_psd = out_psd_db ? (double*)((char*)s.d_out + out_doubles * 8) : nullptr;
    rc = run_batch_device(engine, s.d_in, in_samples, dtype, big_endian, sample_rate, local.data(), n_ann, psd_nfft,
                          psd_hop, psd_window, d_iq_out, dev_off.data(), d_psd, s.stream);
    if (rc) return fail(rc);
    if (out_iq) {
        for (uint32_t i = 0; i < n_ann && e == cudaSuccess; i++) {
            const uint64_t m2 = 2 * (anns[i].count / (uint64_t)anns[i].down);
            if (m2) e = cudaMemcpyAsync(out_iq + iq_offsets[i], d_iq_out + dev_off[i], m2 * 8, cudaMemcpyDeviceToHost, s.stream);
        }
    }
    if (e == cudaSuccess && out_psd_db)
        e = cudaMemcpyAsync(out_psd_db, d_psd, psd_bytes, cudaMemcpyDeviceToHost, s.stream);
    if (e != cudaSuccess) return fail(cuda_fail(e, "D2H annotation results"));
    e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) return cuda_fail(e, "annotation batch");
    return SA_OK;
}

int32_t sa_downconvert(sa_engine* engine, const void* iq, uint64_t iq_bytes, int32_t dtype, int32_t big_endian,
                       uint64_t start_sample, uint64_t count, double freq_off, int32_t down, int32_t fast,
                       double* out_re, double* out_im, uint64_t* out_len) {
    if (!out_re || !out_im || !out_len) return set_error(SA_ERR_INVALID_ARG, "NULL output");
    if (down < 1) return set_error(SA_ERR_INVALID_ARG, "down %d < 1", down);
    const uint64_t m = count / (uint64_t)down;
    sa_annotation a;
    a.start_sample = start_sample; a.count = count; a.freq_off = freq_off; a.down = down; a.fast = fast;
    std::vector<double> tmp(std::max<uint64_t>(2 * m, 1));
    const uint64_t off = 0;
    int rc = sa_downconvert_psd_batch(engine, iq, iq_bytes, dtype, big_endian, 1.0, &a, 1, 0, 0, 0, tmp.data(), &off, nullptr);
    if (rc) return rc;
    memcpy(out_re, tmp.data(), m * 8);
    memcpy(out_im, tmp.data() + m, m * 8);
    *out_len = m;
    return SA_OK;
}

int32_t sa_psd_welch(sa_engine* engine, const double* re, const double* im, uint64_t n, double fs, uint32_t nfft,
                     uint64_t hop, int32_t window, double* out_freq, double* out_db) {
    ENGINE_ENTER(engine);
    if (!re || !im || !out_freq || !out_db) return set_error(SA_ERR_INVALID_ARG, "NULL buffer");
    int rc = check_psd(nfft, &hop, window);
    if (rc) return rc;
    if (n < nfft) return set_error(SA_ERR_INVALID_ARG, "signal (%llu samples) shorter than nfft %u", (unsigned long long)n, nfft);
    if (!(fs > 0.0)) return set_error(SA_ERR_INVALID_ARG, "fs must be > 0");
    // stage the FP64 rows as a one-annotation batch whose "decimated IQ" is the input itself
    Slot& s = engine->slots[0];
    rc = engine->ensure_slot(s, 16, (size_t)(2 * n + nfft) * 8);
    if (rc) return rc;
    double* d_rows = (double*)s.d_out;
    double* d_psd = d_rows + 2 * n;
    cudaError_t e = cudaMemcpyAsync(d_rows, re, n * 8, cudaMemcpyHostToDevice, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rows + n, im, n * 8, cudaMemcpyHostToDevice, s.stream);
    if (e != cudaSuccess) return cuda_fail(e, "H2D psd rows");
    std::vector<WelchJob> jobs(1);
    jobs[0].re = d_rows; jobs[0].im = d_rows + n; jobs[0].n = (long long)n; jobs[0].fs = fs;
    rc = welch_device(engine, jobs, nfft, hop, window, d_psd, s.stream);
    if (rc) return rc;
    e = cudaMemcpyAsync(out_db, d_psd, (size_t)nfft * 8, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
    if (e != cudaSuccess) return cuda_fail(e, "psd welch");
    for (uint32_t k = 0; k < nfft; k++) out_freq[k] = ((double)k - (double)(nfft / 2)) * fs / (double)nfft;
    return SA_OK;
}

}  // extern "C"
