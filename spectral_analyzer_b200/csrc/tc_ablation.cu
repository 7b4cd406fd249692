// tc_ablation.cu -- the tensor-core ablation north_star asks for ("tensor cores are used only if an ablation shows a
// DFT-as-GEMM stage beats the FP32 path within tolerance"), fused IN the kernel: the first radix-32
// pass of the 1024-point cf32 spectrogram as a tcgen05 GEMM on the 5th-generation tensor cores, everything else
// (decode, window, Stockham exchange, second radix-32 pass with twiddles, |X| -> dB, fft-shifted store) exactly as in
// the shipped FP32 kernel.  NOT the product path: sa_spectrogram* never launch this kernel; it is reachable only
// through sa_ablation_tc_spectrogram_device (tools/tc_ablation.py, tests/test_gpu_tc_ablation.py).
//
//   one CTA step = 4 frames = one 128 x 64 x 192 GEMM:   D[(frame, t)][(re|im, m)] = sum_k A[(frame, t)][k] B[(re|im, m)][k]
//     rows     r = 32 frame + t    : the thread that owns samples x[t + 32 q] of its warp's frame (M = 128)
//     columns  n = m | 32 + m      : real / imaginary part of output m of the 32-point DFT over q (N = 64)
//     K        k = q | 32 + q      : real / imaginary part of windowed input q (K = 64), three times: the FP32 values
//              are split x = hi + lo (hi = the 11 mantissa bits kind::tf32 reads, lo = x - hi, read the same way) and
//              D = A_hi B_hi + A_lo B_hi + A_hi B_lo  -- 22 mantissa bits, the 1e-5 power tolerance needs ~18
//     B        [[Fr, -Fi], [Fi, Fr]], F[m][q] = exp(-2 pi i m q / 32), split the same way on the host
//   A and B sit in shared memory in the canonical K-major no-swizzle layout (8 x 16-byte core matrices), the accumulator
//   in 64 TMEM columns; one thread issues the 24 tcgen05.mma (K = 8 each) and commits to an mbarrier; tcgen05.ld
//   32x32b hands thread (frame, t) its 64 accumulator columns = exactly the registers the FP32 pass 0 would have left.
#include "engine_internal.h"

namespace sa {

constexpr int kTcFrames = 4, kTcThreads = 128, kTcK = 64, kTcN = 64;
constexpr int kTcSegBytes = 128 * kTcK * 4;            // one 128 x 64 A segment (hi or lo): 32 KB
constexpr int kTcBSegBytes = kTcN * kTcK * 4;          // one 64 x 64 B segment: 16 KB
// canonical K-major no-swizzle layout: core matrix = 8 rows x 16 bytes (128 contiguous bytes); core matrices adjacent in
// K are kTcKStr bytes apart, groups of 8 rows kTcMnStr bytes apart
constexpr uint32_t kTcKStr = 128, kTcMnStr = (kTcK / 4) * 128;

struct TcArgs {
    SpecArgs s;
    const float* b_hi_lo;          // B_hi then B_lo, canonical layout (as they sit in shared memory)
    uint32_t lbo, sbo;             // the two stride fields of the shared-memory descriptors (which of kTcKStr / kTcMnStr goes where)
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    // UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in [0,14), leading byte offset >> 4
    // in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_NONE = 0 in [61,64)
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// mbarrier wait that traps instead of hanging the GPU if the commit never arrives (ablation code: fail loudly)
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar, uint32_t parity) {
    for (unsigned spins = 0; spins < (1u << 26); spins++) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// hi = the bits kind::tf32 reads (sign, exponent, 10 explicit mantissa bits), lo = the exact remainder
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

__global__ void __launch_bounds__(kTcThreads, 2)
tc_spectrogram_kernel(const TcArgs a) {
    using G = Geo<float, 1024>;
    constexpr int N = 1024, P = 32, TPF = 32;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* a_seg = smem;                                  // A_hi | A_lo (2 x 32 KB); the exchange buffers alias A_hi
    unsigned char* b_seg = smem + 2 * kTcSegBytes;                // B_hi | B_lo (2 x 16 KB)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, w = tid >> 5, t = tid & 31;
    // B once per CTA (already in canonical layout)
    for (int i = tid; i < 2 * kTcBSegBytes / 16; i += kTcThreads)
        reinterpret_cast<float4*>(b_seg)[i] = __ldg(reinterpret_cast<const float4*>(a.b_hi_lo) + i);
    if (tid == 0) mbar_init(smem_u32(&bar), 1);
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    fence_proxy_async();
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
    // N >> 3 at bit 17, M >> 4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_hi = smem_u32(a_seg), a_lo = a_hi + kTcSegBytes, b_hi = smem_u32(b_seg), b_lo = b_hi + kTcBSegBytes;
    const float* win = reinterpret_cast<const float*>(a.s.window);
    const cpx<float>* tw = reinterpret_cast<const cpx<float>*>(a.s.twiddle);
    const TwSeed<float> seed = load_tw_seed<float, N>(tw, t);
    cpx<float>* sm = reinterpret_cast<cpx<float>*>(a_seg) + (size_t)w * G::SM_ELEMS;      // this warp's exchange buffer
    const long long n_blocks = (a.s.n_frames + kTcFrames - 1) / kTcFrames;
    uint32_t parity = 0;
    const int r = tid;                                           // A row of this thread
    unsigned char* row_hi = a_seg + (r >> 3) * kTcMnStr + (r & 7) * 16;
    unsigned char* row_lo = row_hi + kTcSegBytes;
    // the raw samples of a step are loaded one step ahead (32 independent loads per thread, in flight under the
    // previous step's MMA, second pass and epilogue), as the shipped kernel's TMA staging does
    auto frame_ok = [&](long long fb_, long long& frame_, bool& in_grid_, bool& readable_) {
        frame_ = fb_ * kTcFrames + w;
        in_grid_ = frame_ < a.s.n_frames;
        readable_ = in_grid_ && (a.s.start_sample + frame_ * a.s.hop + N <= a.s.n_samples);
    };
    cpx<float> raw[P];
    auto load_raw = [&](long long fb_) {
        long long fr; bool ig, rd;
        frame_ok(fb_, fr, ig, rd);
        const long long s0_ = a.s.start_sample + fr * a.s.hop;
#pragma unroll
        for (int q = 0; q < P; q++) {
            raw[q] = mk2<float>(0.f, 0.f);
            if (rd) raw[q] = Loader<float, DK_CF32>::load<false>(a.s.lp, s0_ + t + TPF * q);
        }
    };
    float wreg[P];
#pragma unroll
    for (int q = 0; q < P; q++) wreg[q] = __ldg(&win[t + TPF * q]);
    if ((long long)blockIdx.x < n_blocks) load_raw(blockIdx.x);
    for (long long fb = blockIdx.x; fb < n_blocks; fb += gridDim.x) {
        long long frame; bool in_grid, readable;
        frame_ok(fb, frame, in_grid, readable);
        // ---- window, split, A rows into shared memory (k = q: real parts, k = 32 + q: imaginary parts)
        __syncthreads();                                         // the exchange buffers (aliasing A_hi) are free again
#pragma unroll
        for (int g = 0; g < 8; g++) {
            float xr[4], xi[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int q = 4 * g + j;
                xr[j] = raw[q].x * wreg[q]; xi[j] = raw[q].y * wreg[q];
            }
            float4 hr, lr, hi4, li4;
            split_tf32(xr[0], hr.x, lr.x); split_tf32(xr[1], hr.y, lr.y); split_tf32(xr[2], hr.z, lr.z); split_tf32(xr[3], hr.w, lr.w);
            split_tf32(xi[0], hi4.x, li4.x); split_tf32(xi[1], hi4.y, li4.y); split_tf32(xi[2], hi4.z, li4.z); split_tf32(xi[3], hi4.w, li4.w);
            *reinterpret_cast<float4*>(row_hi + g * kTcKStr) = hr;
            *reinterpret_cast<float4*>(row_hi + (8 + g) * kTcKStr) = hi4;
            *reinterpret_cast<float4*>(row_lo + g * kTcKStr) = lr;
            *reinterpret_cast<float4*>(row_lo + (8 + g) * kTcKStr) = li4;
        }
        fence_proxy_async();                                     // generic-proxy writes visible to the tensor core's async proxy
        __syncthreads();
        if (fb + gridDim.x < n_blocks) load_raw(fb + gridDim.x);
        // ---- 24 MMAs of K = 8 (two core matrices along K per step): hi*hi, lo*hi, hi*lo
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t acc = 0;
#pragma unroll 1
            for (int seg = 0; seg < 3; seg++) {
                const uint32_t ab = seg == 1 ? a_lo : a_hi, bb = seg == 2 ? b_lo : b_hi;
#pragma unroll 1
                for (int k = 0; k < kTcK / 8; k++) {
                    umma_tf32(tmem, umma_desc(ab + 2 * k * kTcKStr, a.lbo, a.sbo), umma_desc(bb + 2 * k * kTcKStr, a.lbo, a.sbo), idesc, acc);
                    acc = 1;
                }
            }
            umma_commit(smem_u32(&bar));                         // implies tcgen05.fence::before_thread_sync
        }
        mbar_wait_or_trap(smem_u32(&bar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- accumulator row (32 w + t): columns 0..31 = Re Y[m], 32..63 = Im Y[m]
        float yr[32], yi[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * w) << 16);
        tmem_ld32(taddr, yr);
        tmem_ld32(taddr + 32, yi);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        cpx<float> v[P];
#pragma unroll
        for (int m = 0; m < P; m++) v[m] = mk2<float>(yr[m], yi[m]);
        __syncthreads();                                         // every MMA has read A: the exchange may overwrite A_hi
        // ---- Stockham exchange of pass 0 (fft_core.cuh, NS = 1, R = 32, S = 1): write element 32 t + m, read t + 32 q
        {
            float4* dst = reinterpret_cast<float4*>(sm + pad_idx<G::R0, G::PADW>(t * 32));
#pragma unroll
            for (int m = 0; m < 32; m += 2) dst[m / 2] = make_float4(v[m].x, v[m].y, v[m + 1].x, v[m + 1].y);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = sm[pad_idx<G::R0, G::PADW>(t + TPF * q)];
        }
        fft_pass<float, N, 1, true, true>(v, t, sm, tw, nullptr, seed);
        if (!in_grid) continue;
        if (!readable) { store_fill<float, N>(a.s, frame, t); continue; }
        store_row<float, N>(a.s, frame, t, v);
    }
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

}  // namespace sa

using namespace sa;

extern "C" {

// Ablation entry point (not part of the drop-in boundary): cf32_le device samples, 1024-point, f32 dB rows out.
// layout_swap != 0 exchanges the two descriptor strides (tools/tc_ablation.py checks both against the FP32 kernel).
SA_API int32_t sa_ablation_tc_spectrogram_device(sa_engine* engine, const void* d_iq, uint64_t iq_bytes,
                                                 const sa_spectrogram_params* params, void* d_out, uint64_t out_bytes,
                                                 int32_t layout_swap, void* cuda_stream) {
    ENGINE_ENTER(engine);
    int prec = 0;
    int rc = check_spec_params(params, &prec);
    if (rc) return rc;
    if (params->dtype != SA_CF32 || params->big_endian || params->nfft != 1024 || params->out_kind != SA_OUT_F32_DB || prec != SA_PREC_F32)
        return set_error(SA_ERR_UNSUPPORTED, "the tensor-core ablation covers cf32_le, nfft 1024, f32 dB rows");
    if (!d_iq || !d_out) return set_error(SA_ERR_INVALID_ARG, "NULL buffer");
    if (out_bytes < params->n_frames * 1024ull * 4) return set_error(SA_ERR_SMALL_OUTPUT, "out_bytes too small");
    if (params->n_frames == 0) return SA_OK;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    TcArgs a;
    memset(&a, 0, sizeof(a));
    fill_load_params(a.s.lp, d_iq, params->dtype, 0);
    a.s.n_samples = (long long)(iq_bytes / 8);
    a.s.start_sample = (long long)params->start_sample;
    a.s.hop = (long long)params->hop;
    a.s.n_frames = (long long)params->n_frames;
    a.s.out = d_out;
    a.s.out_kind = params->out_kind;
    a.s.db_mode = params->db_mode;
    a.s.eof_fill = params->eof_fill_db;
    const SpecKernelInfo* k = find_spec_kernel(SA_PREC_F32, 1024, DK_CF32, 1, 0);
    if (!k) return set_error(SA_ERR_UNSUPPORTED, "1024-point plan missing");
    rc = engine->twiddle_table(*k, &a.s.twiddle);
    if (rc) return rc;
    rc = engine->window_table(params->window, 1024, SA_PREC_F32, &a.s.window);
    if (rc) return rc;
    // canonical K-major layout: core matrix = 8 rows x 16 bytes; K-adjacent core matrices 128 bytes apart, row groups
    // (K / 4) * 128 bytes apart
    const uint32_t kstride = kTcKStr, mnstride = kTcMnStr;
    a.lbo = layout_swap ? mnstride : kstride;
    a.sbo = layout_swap ? kstride : mnstride;
    // B = [[Fr, -Fi], [Fi, Fr]] split into hi / lo, element (n, k) at (n / 8) * mnstride + (k / 4) * kstride + (n % 8) * 16 + (k % 4) * 4
    const uint64_t key = (7ull << 40) | (layout_swap ? 1 : 0);
    auto it = engine->misc_tables.find(key);
    if (it == engine->misc_tables.end()) {
        std::vector<float> b(2 * kTcN * kTcK, 0.f);
        const double two_pi = 6.283185307179586476925286766559;
        for (int n = 0; n < kTcN; n++)
            for (int kk = 0; kk < kTcK; kk++) {
                const int m = n & 31, q = kk & 31;
                const double ang = -two_pi * (double)((m * q) & 31) / 32.0;
                const double fr = std::cos(ang), fi = std::sin(ang);
                const double val = (n < 32) ? (kk < 32 ? fr : -fi) : (kk < 32 ? fi : fr);
                const float x = (float)val;
                uint32_t bits;
                memcpy(&bits, &x, 4);
                bits &= 0xFFFFE000u;
                float hi;
                memcpy(&hi, &bits, 4);
                const float lo = (float)(val - (double)hi);
                const size_t off = ((size_t)(n / 8) * mnstride + (size_t)(kk / 4) * kstride + (size_t)(n % 8) * 16 + (size_t)(kk % 4) * 4) / 4;
                b[off] = hi;
                b[kTcN * kTcK + off] = lo;
            }
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, b.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(d, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) return cuda_fail(e, "ablation DFT matrix");
        engine->misc_tables[key] = d;
        it = engine->misc_tables.find(key);
    }
    a.b_hi_lo = (const float*)it->second;
    const size_t smem = 1024 + 2 * kTcSegBytes + 2 * kTcBSegBytes;
    int bps = 0;
    rc = engine->kernel_grid((const void*)&tc_spectrogram_kernel, kTcThreads, smem, &bps);
    if (rc) return rc;
    const long long n_blocks = ((long long)params->n_frames + kTcFrames - 1) / kTcFrames;
    const unsigned grid = (unsigned)std::min<long long>(n_blocks, (long long)bps * engine->num_sms);
    void* args[] = { &a };
    cudaError_t e = cudaLaunchKernel((const void*)&tc_spectrogram_kernel, dim3(grid), dim3(kTcThreads), args, smem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "launch tc_spectrogram_kernel");
    engine->launches++;
    engine->last_kernel = "tc_spectrogram_kernel (ablation: tcgen05 TF32x3 radix-32 pass)";
    return SA_OK;
}

}  // extern "C"
