// spectrogram_tma_kernel.cuh -- one-warp-per-frame spectrogram with TMA-staged frames.
//
// Same arithmetic and output as spectrogram_kernel (spectrogram_kernel.cuh) for the plans where a
// frame is owned by one warp (TPF == 32, i.e. the 1024-point FP32 plan), but the samples do not
// travel through LDG/L1: lane 0 of every warp issues ONE bulk asynchronous copy
// (cp.async.bulk global -> shared, completion on an mbarrier) that lands the warp's NEXT frame in
// the warp's own exchange buffer while pass 1 and the dB epilogue of the current frame run.  The
// raw frame and the Stockham exchange alias the same shared memory: the raw samples are consumed
// into registers before the exchange of that frame overwrites them, and the next copy is issued
// only after the exchange has been read back.
// Requires 16-byte aligned frame starts and sizes: the host takes this kernel when
// base, start_sample*bps and hop*bps are multiples of 16 and falls back to spectrogram_kernel.
#pragma once
#include "spectrogram_kernel.cuh"

#ifndef SA_TMA_EVICT_LAST
#define SA_TMA_EVICT_LAST 0
#endif
#ifndef SA_TMA_BLOCKS_PER_STEP
#define SA_TMA_BLOCKS_PER_STEP 1
#endif

namespace sa {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
#if SA_TMA_EVICT_LAST
    // overlapping frames read every sample twice: ask L2 to keep the lines until the second read
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T, int N, int DK, bool WIN>
__global__ void __launch_bounds__(Geo<T, N>::CTA, Geo<T, N>::MINB)
spectrogram_tma_kernel(const SpecArgs a) {
    using G = Geo<T, N>;
    using LD = Loader<T, DK>;
    using raw_t = typename LD::raw_t;
    constexpr int P = G::P, TPF = G::TPF, FPC = G::FPC;
    static_assert(TPF == 32 && Plan<T, N>::NP == 2, "one warp per frame, two passes");
    constexpr uint32_t FRAME_BYTES = N * sizeof(raw_t);
    static_assert(FRAME_BYTES <= G::SM_ELEMS * sizeof(cpx<T>), "raw frame must fit the exchange buffer");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int fl = threadIdx.x / TPF, t = threadIdx.x % TPF;
    cpx<T>* sm = reinterpret_cast<cpx<T>*>(smem_raw) + (size_t)fl * G::SM_ELEMS;
    const raw_t* raw = reinterpret_cast<const raw_t*>(sm);
    static_assert(G::TW_REC, "the TMA-staged kernel generates its pass-1 twiddles by recurrence");
    const cpx<T>* tw = reinterpret_cast<const cpx<T>*>(a.twiddle);
    const TwSeed<T> seed = load_tw_seed<T, N>(reinterpret_cast<const cpx<T>*>(a.twiddle), t);
    // one mbarrier per warp, after the exchange buffers, the twiddles and the window rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + G::SMEM_BYTES + (WIN ? G::WIN_BYTES : 0));
    const uint32_t bar = smem_u32(&bars[fl]);
    const uint32_t dst = smem_u32(sm);

    T win_reg[G::WIN_SMEM ? 1 : P];
    const T* win = win_reg;
    if constexpr (WIN) win = setup_window<T, N>(a, smem_raw + G::SMEM_BYTES, win_reg, t);

    if (t == 0) mbar_init(bar, 1);
    __syncwarp();
    fence_proxy_async();      // barrier init visible to the async proxy

    const long long n_blocks = (a.n_frames + FPC - 1) / FPC;
    const char* base = reinterpret_cast<const char*>(a.lp.base);
    auto frame_of = [&](long long fb) { return fb * FPC + fl; };
    auto readable_f = [&](long long frame) {
        return frame < a.n_frames && (a.start_sample + frame * a.hop + N <= a.n_samples);   // MainController.java:987
    };
    auto issue = [&](long long frame) {       // lane 0 only
        mbar_expect_tx(bar, FRAME_BYTES);
        tma_load_1d(dst, base + (a.start_sample + frame * a.hop) * (long long)sizeof(raw_t), FRAME_BYTES, bar);
    };

    // a CTA takes SA_TMA_BLOCKS_PER_STEP adjacent frame blocks before it strides by the grid
    constexpr long long K = SA_TMA_BLOCKS_PER_STEP;
    auto adv = [&](long long b) { return ((b + 1) % K != 0) ? b + 1 : b + 1 + ((long long)gridDim.x - 1) * K; };
    long long fb = (long long)blockIdx.x * K;
    if (fb < n_blocks && readable_f(frame_of(fb)) && t == 0) issue(frame_of(fb));
    uint32_t parity = 0;
    for (; fb < n_blocks; fb = adv(fb)) {
        const long long frame = frame_of(fb);
        const long long nfb = adv(fb);
        const long long next = frame_of(nfb);
        const bool next_readable = (nfb < n_blocks) && readable_f(next);
        if (frame >= a.n_frames) continue;            // warp-uniform; a later block cannot be in range either
        if (!readable_f(frame)) {                     // EOF row, no copy was issued for it
            store_fill<T, N>(a, frame, t);
            if (next_readable && t == 0) issue(next);
            continue;
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        cpx<T> v[P];
        if (a.lp.swap) {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = LD::template decode<true>(a.lp, raw[t + TPF * q]);
        } else {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = LD::template decode<false>(a.lp, raw[t + TPF * q]);
        }
        // pass 0 ends with: sync, exchange write (overwrites the raw frame), sync, exchange read
        fft_pass<T, N, 0, WIN, true>(v, t, sm, tw, win, seed);
        __syncwarp();                                 // every lane has read its exchange values back
        if (next_readable && t == 0) {
            fence_proxy_async();                      // generic-proxy accesses ordered before the async write
            issue(next);
        }
        fft_pass<T, N, 1, WIN, true>(v, t, sm, tw, win, seed);
        store_row<T, N>(a, frame, t, v);
    }
}

template <typename T, int N, int DK, bool WIN>
SpecKernelInfo make_spec_tma_info(int prec) {
    SpecKernelInfo k = make_spec_info<T, N, DK, WIN>(prec);
    k.fn = (const void*)&spectrogram_tma_kernel<T, N, DK, WIN>;
    k.smem = Geo<T, N>::SMEM_BYTES + (WIN ? Geo<T, N>::WIN_BYTES : 0) + Geo<T, N>::FPC * sizeof(uint64_t);
    k.tma = 1;
    return k;
}

}  // namespace sa
