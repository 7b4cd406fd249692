// spectrogram_split_kernel.cuh -- 2048-point FP32 spectrogram as TWO warp-private 1024-point transforms + one combine.
//
// Same arithmetic contract and output as spectrogram_kernel (SpectralService.java:33-85 in the frame loop of
// MainController.java:980-999).  Decimation in time by two: warp s of a frame's warp pair transforms the samples
// x[2n + s] with the one-warp-per-frame machinery of the 1024-point kernel (32 points per thread, radix 32 x 32, ONE
// exchange through the warp's own shared-memory buffer, __syncwarp only, pass-1 twiddles by the register recurrence),
// then the pair swaps the halves it does not finish (16 values per thread) and forms
//   X[k] = E[k] + W_2048^k O[k],   X[k + 1024] = E[k] - W_2048^k O[k],   k = t + 512 s + 32 i,  i < 16,
// with W_2048^k = W_2048^(t + 512 s) (per-lane seed) x W_64^i (compile-time).  Against the three-pass plan of
// spectrogram_mid_kernel.cuh (2 x 32 x 32: two full exchanges behind 64-thread barriers, pass-1 twiddles from a table)
// a frame crosses shared memory 1.5 times, three of its barriers are warp-local, and every twiddle is generated in
// registers.  The raw frame is staged by TMA (cp.async.bulk + mbarrier) into the pair's exchange buffers while the
// combine and the epilogue of the previous frame run, as in spectrogram_tma_kernel.cuh.
// Requires 16-byte aligned frames; no fused display pooling (those launches take the three-pass kernel).
// OPT-IN (SA_SPLIT=1), kept as a measured alternative: on B200 it executes the same number of instructions per point
// as the three-pass kernel (43.8, cu8 -> f32 dB) with 29 % fewer shared-memory wavefronts and two thirds fewer bank
// conflicts, and runs in the same time (2.02 vs 2.00 ms per 2^30 samples) -- the 2048-point kernels wait on the FMA
// pipe (math-pipe-throttle is the first stall reason of both), not on shared memory or barriers.
#pragma once
#include "spectrogram_tma_kernel.cuh"

namespace sa {

struct SplitGeo {
    static constexpr int N = 2048, P = 32, TPF = 64;
    static constexpr int CTA = 512, FPC = CTA / TPF;
    using G1 = Geo<float, 1024>;                               // the warp-private transform
    static constexpr int SUB_ELEMS = G1::SM_ELEMS;             // 1024 + 2 per 32
    static constexpr int WROW = P + 2;
    static constexpr size_t EX_BYTES = (size_t)FPC * 2 * SUB_ELEMS * sizeof(float2);
    static constexpr size_t WIN_BYTES = (size_t)TPF * WROW * sizeof(float);
    static constexpr size_t BAR_BYTES = (size_t)FPC * sizeof(uint64_t);
};

// sample 2 (t + 32 q) + s of the staged raw frame -> FP32 (same values as decode.cuh, bit for bit)
template <int DK, bool SWAP>
__device__ __forceinline__ float2 split_decode(const LoadParams& lp, const unsigned char* __restrict__ raw, const int t,
                                               const int s, const int q, const uint32_t c8_sel) {
    if constexpr (DK == DK_C8) {
        // both samples of the pair in one 32-bit word; byte -> bits 8..15 of a 2^23-exponent float, one exact FMA
        const uint32_t x = reinterpret_cast<const uint32_t*>(raw)[t + 32 * q] ^ lp.c8_flip;
        const float a = __uint_as_float(__byte_perm(x, 0x4B000000u, c8_sel));
        const float b = __uint_as_float(__byte_perm(x, 0x4B000000u, c8_sel + 0x10u));
        return make_float2(__fmaf_rn(a, 1.0f / 32768.0f, -lp.c8_c), __fmaf_rn(b, 1.0f / 32768.0f, -lp.c8_c));
    } else {
        using LD = Loader<float, DK>;
        return LD::template decode<SWAP>(lp, reinterpret_cast<const typename LD::raw_t*>(raw)[2 * t + s + 64 * q]);
    }
}

// Combine + epilogue of one warp (S = its half): e / o hold E[k_i] and O[k_i], k_i = t + 512 S + 32 i.
template <int MODE>
__device__ __forceinline__ void split_finish(const SpecArgs& a, const long long frame, const int k0, float2 (&e)[16],
                                             float2 (&o)[16], const float2 cw) {
    constexpr int N = SplitGeo::N;
    float2 x[32];                                     // x[i] = X[k_i + 1024] (row position k_i), x[16 + i] = X[k_i] (position k_i + 1024)
#pragma unroll
    for (int i = 0; i < 16; i++) {
        float2 y = make_float2(__fmaf_rn(-cw.y, o[i].y, cw.x * o[i].x), __fmaf_rn(cw.y, o[i].x, cw.x * o[i].y));
        float2 lo = e[i];
        bfly<float, 64>(lo, y, i);                    // (E + W_64^i y, E - W_64^i y)
        x[16 + i] = lo;
        x[i] = y;
    }
    float db[32];
    bins_to_db<float, 32, MODE>(x, db);
    const size_t row = (size_t)frame * N;
    if (a.out_kind == OUT_F32_DB) {
        float* out = reinterpret_cast<float*>(a.out) + row + k0;
#pragma unroll
        for (int j = 0; j < 32; j++) out[32 * (j % 16) + 1024 * (j / 16)] = db[j];
    } else if (a.out_kind == OUT_F64_DB) {
        double* out = reinterpret_cast<double*>(a.out) + row + k0;
#pragma unroll
        for (int j = 0; j < 32; j++) out[32 * (j % 16) + 1024 * (j / 16)] = (double)db[j];
    } else {
        uint32_t* out = reinterpret_cast<uint32_t*>(a.out) + row + k0;
        const float sc = a.inv_range, bi = a.cmap_bias;
        if (a.cmap == 1) {
#pragma unroll
            for (int j = 0; j < 32; j++) out[32 * (j % 16) + 1024 * (j / 16)] = colormap_px<1>(db[j], sc, bi);
        } else {
#pragma unroll
            for (int j = 0; j < 32; j++) out[32 * (j % 16) + 1024 * (j / 16)] = colormap_px<0>(db[j], sc, bi);
        }
    }
}

template <int DK, bool WIN>
__global__ void __launch_bounds__(SplitGeo::CTA, 1)
spectrogram_split_kernel(const SpecArgs a) {
    using G = SplitGeo;
    constexpr int N = G::N, P = G::P, TPF = G::TPF, FPC = G::FPC;
    constexpr uint32_t FRAME_BYTES = N * bytes_per_iq_kind<DK>();
    static_assert(FRAME_BYTES <= 2 * G::SUB_ELEMS * sizeof(float2), "raw frame must fit the pair's exchange buffers");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int fl = threadIdx.x / TPF, s = (threadIdx.x >> 5) & 1, t = threadIdx.x & 31;
    float2* smf = reinterpret_cast<float2*>(smem_raw) + (size_t)fl * 2 * G::SUB_ELEMS;    // the pair's two buffers
    float2* smw = smf + s * G::SUB_ELEMS;                                                 // this warp's
    float2* smo = smf + (1 - s) * G::SUB_ELEMS;                                           // the partner's
    const unsigned char* raw = reinterpret_cast<const unsigned char*>(smf);
    float* wsm = reinterpret_cast<float*>(smem_raw + G::EX_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + G::EX_BYTES + (WIN ? G::WIN_BYTES : 0));
    const uint32_t bar = smem_u32(&bars[fl]);
    const uint32_t dst = smem_u32(smf);

    if constexpr (WIN) {        // rows per (s, t) in first-stage pair order: factors of registers q and q + 16 adjacent
        const float* w = reinterpret_cast<const float*>(a.window);
        for (int i = threadIdx.x; i < N; i += G::CTA) {
            const int ss = i & 1, n = i >> 1, tt = n & 31, e = n >> 5;        // sample i = 2 (tt + 32 e) + ss
            const int slot = e < P / 2 ? 2 * e : 2 * (e - P / 2) + 1;
            wsm[(ss * 32 + tt) * G::WROW + slot] = __ldg(&w[i]);
        }
    }
    if (s == 0 && t == 0) mbar_init(bar, 1);
    __syncthreads();
    fence_proxy_async();
    const float* win = wsm + (s * 32 + t) * G::WROW;
    TwSeed<float> seed;             // W_1024^t = W_2048^(2t) and its 16th power
    float2 cw;                      // W_2048^(t + 512 s)
    {
        const float2* root = reinterpret_cast<const float2*>(a.aux);       // W_2048^j
        seed.om = __ldg(&root[2 * t]);
        seed.oh = __ldg(&root[(32 * t) & (N - 1)]);
        seed.q_lo = seed.om; seed.q_hi = seed.om;                          // (radix-64 re-seed: unused)
        cw = __ldg(&root[t + 512 * s]);
    }
    const uint32_t c8_sel = s ? 0x7424u : 0x7404u;

    const long long n_blocks = (a.n_frames + FPC - 1) / FPC;
    const char* base = reinterpret_cast<const char*>(a.lp.base);
    auto frame_of = [&](long long fb) { return fb * FPC + fl; };
    auto readable_f = [&](long long frame) {
        return frame < a.n_frames && (a.start_sample + frame * a.hop + N <= a.n_samples);   // MainController.java:987
    };
    auto issue = [&](long long frame) {       // one thread of the pair
        mbar_expect_tx(bar, FRAME_BYTES);
        tma_load_1d(dst, base + (a.start_sample + frame * a.hop) * (long long)bytes_per_iq_kind<DK>(), FRAME_BYTES, bar);
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(fl + 1), "n"(TPF) : "memory"); };
    const bool issuer = (s == 0 && t == 0);

    long long fb = blockIdx.x;
    if (fb < n_blocks && readable_f(frame_of(fb)) && issuer) issue(frame_of(fb));
    uint32_t parity = 0;
    for (; fb < n_blocks; fb += gridDim.x) {
        const long long frame = frame_of(fb);
        const long long next = frame_of(fb + gridDim.x);
        const bool next_readable = (fb + gridDim.x < n_blocks) && readable_f(next);
        if (frame >= a.n_frames) continue;            // uniform over the pair
        if (!readable_f(frame)) {                     // EOF row, no copy was issued for it
            store_fill<float, 1024>(a, 2 * frame + s, t);        // the row as two 1024-bin halves (Geo<float,1024>: 32 stores per lane)
            if (next_readable && issuer) issue(next);
            continue;
        }
        mbar_wait(bar, parity);
        parity ^= 1;
        float2 v[P];
        if (a.lp.swap) {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = split_decode<DK, true>(a.lp, raw, t, s, q, c8_sel);
        } else {
#pragma unroll
            for (int q = 0; q < P; q++) v[q] = split_decode<DK, false>(a.lp, raw, t, s, q, c8_sel);
        }
        // ---- warp-private 1024-point transform of x[2n + s]
        radix_fft<float, 32, 1, 0, P, WIN ? MUL_REAL : MUL_NONE, false>(v, win, nullptr, 0, seed);
        pair_sync();                                  // both warps have consumed the raw frame
        {
            float4* d = reinterpret_cast<float4*>(smw + 34 * t);          // outputs 32 t + m, two pad elements per 32
#pragma unroll
            for (int m = 0; m < P; m += 2) d[m / 2] = make_float4(v[m].x, v[m].y, v[m + 1].x, v[m + 1].y);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < P; q++) v[q] = smw[t + 34 * q];
        radix_fft<float, 32, 1, 0, P, MUL_REC, false>(v, nullptr, nullptr, 0, seed);     // v[q] = X_s[t + 32 q]
        // ---- swap halves: warp 0 finishes k = t + 32 i, warp 1 k = t + 512 + 32 i (i < 16)
        __syncwarp();                                 // every lane has read its exchange values back
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) smw[32 * i + t] = v[16 + i];     // E[t + 512 + 32 i]
        } else {
#pragma unroll
            for (int i = 0; i < 16; i++) smw[32 * i + t] = v[i];          // O[t + 32 i]
        }
        pair_sync();
        float2 r[16];
#pragma unroll
        for (int i = 0; i < 16; i++) r[i] = smo[32 * i + t];
        pair_sync();                                  // both buffers have been read: the pair's shared memory is free
        if (next_readable && issuer) {
            fence_proxy_async();
            issue(next);
        }
        const int k0 = t + 512 * s;
        if (s == 0) {
            float2 e[16];
#pragma unroll
            for (int i = 0; i < 16; i++) e[i] = v[i];
            if (a.db_mode == DBM_MAG_1E10) split_finish<DBM_MAG_1E10>(a, frame, k0, e, r, cw);
            else split_finish<DBM_POWER>(a, frame, k0, e, r, cw);
        } else {
            float2 o[16];
#pragma unroll
            for (int i = 0; i < 16; i++) o[i] = v[16 + i];
            if (a.db_mode == DBM_MAG_1E10) split_finish<DBM_MAG_1E10>(a, frame, k0, r, o, cw);
            else split_finish<DBM_POWER>(a, frame, k0, r, o, cw);
        }
    }
}

template <int DK, bool WIN>
SpecKernelInfo make_spec_split_info() {
    using G = SplitGeo;
    SpecKernelInfo k;
    k.fn = (const void*)&spectrogram_split_kernel<DK, WIN>;
    k.prec = 1; k.n = G::N; k.dk = DK; k.win = WIN ? 1 : 0;
    k.cta = G::CTA; k.fpc = G::FPC; k.minb = 1;
    k.smem = G::EX_BYTES + (WIN ? G::WIN_BYTES : 0) + G::BAR_BYTES;
    k.p = 32; k.np = 3; k.tma = 5;
    k.radix[0] = 32; k.radix[1] = 32; k.radix[2] = 2; k.radix[3] = 1;
    return k;
}

}  // namespace sa
