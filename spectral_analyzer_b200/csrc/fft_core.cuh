// fft_core.cuh -- register-resident radix-R butterflies and shared-memory Stockham passes.
//
// Replaces the FP64 radix-2 transform the reference calls per frame
// (FastFourierTransformer(STANDARD).transform(.., FORWARD), S/services/SpectralService.java:23,68):
// X[k] = sum_n x[n] exp(-2 pi i k n / N), unnormalised.
//
// Layout of one transform of N points over TPF = N/P threads (P points per thread):
//   at the start of every pass thread t holds elements  t + TPF*q  (q = 0..P-1) in v[q];
//   a pass of radix R runs S = P/R independent radix-R butterflies per thread on the register
//   sub-arrays v[s + m*S] (m = 0..R-1), preceded (for passes after the first) by the Stockham
//   twiddle W_N^{(j mod Ns) * m * N/(Ns*R)}, j = t + TPF*s, Ns = product of earlier radices;
//   the pass writes element (j/Ns)*Ns*R + (j mod Ns) + m*Ns to shared memory and every thread
//   reads back t + TPF*q.  After the last pass v[q] = X[t + TPF*q] (no final exchange), so both
//   the global loads of pass 0 and the global stores of the epilogue are coalesced.
// Butterflies are radix-2 DIT on bit-reversed registers with compile-time twiddles in FMA form
// (6 FMA per non-trivial butterfly, 4 ADD for W = 1 and W = -i).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sa {

template <typename T> struct V2;
template <> struct V2<float>  { using type = float2; };
template <> struct V2<double> { using type = double2; };
template <typename T> using cpx = typename V2<T>::type;

template <typename T> __device__ __forceinline__ cpx<T> mk2(T x, T y) { cpx<T> r; r.x = x; r.y = y; return r; }

__device__ __forceinline__ float  fma_t(float a, float b, float c)    { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return __fma_rn(a, b, c); }

// cos(2*pi*q/32), q = 0..8
__host__ __device__ constexpr double cos32_tab(int q) {
    return q == 0 ? 1.0
         : q == 1 ? 0.98078528040323044912618223613424
         : q == 2 ? 0.92387953251128675612818318939679
         : q == 3 ? 0.83146961230254523707878837761791
         : q == 4 ? 0.70710678118654752440084436210485
         : q == 5 ? 0.55557023301960222474283081394853
         : q == 6 ? 0.38268343236508977172845998403040
         : q == 7 ? 0.19509032201612826784828486847702
         : 0.0;
}
// W32^q = exp(-2*pi*i*q/32) for q = 0..16
__host__ __device__ constexpr double w32_re(int q) { return q <= 8 ? cos32_tab(q) : -cos32_tab(16 - q); }
__host__ __device__ constexpr double w32_im(int q) { return q <= 8 ? -cos32_tab(8 - q) : -cos32_tab(q - 8); }

// DIT butterfly: (a, b) <- (a + W b, a - W b), W = W32^q, q compile-time after unrolling.
template <typename T>
__device__ __forceinline__ void bfly(cpx<T>& a, cpx<T>& b, const int q) {
    if (q == 0) {
        T bx = b.x, by = b.y;
        b.x = a.x - bx; b.y = a.y - by;
        a.x = a.x + bx; a.y = a.y + by;
    } else if (q == 8) {            // W = -i : W b = (b.y, -b.x)
        T bx = b.x, by = b.y;
        b.x = a.x - by; b.y = a.y + bx;
        a.x = a.x + by; a.y = a.y - bx;
    } else {
        const T wr = (T)w32_re(q), wi = (T)w32_im(q);
        T ox = fma_t(-wi, b.y, fma_t(wr, b.x, a.x));
        T oy = fma_t( wi, b.x, fma_t(wr, b.y, a.y));
        b.x = fma_t((T)2, a.x, -ox);
        b.y = fma_t((T)2, a.y, -oy);
        a.x = ox; a.y = oy;
    }
}

// bit reversal of x within log2(n) bits, n <= 32; loop-free so it folds after unrolling
__host__ __device__ constexpr int bitrev_c(int x, int n) {
    const int r5 = ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
    return n == 32 ? r5 : n == 16 ? (r5 >> 1) : n == 8 ? (r5 >> 2) : n == 4 ? (r5 >> 3) : (r5 >> 4);
}

// one DIT stage of span LEN over a[0..R), then the next (template recursion: nvcc does not
// unroll a loop whose induction variable is shifted)
template <typename T, int R, int LEN>
__device__ __forceinline__ void dit_stages(cpx<T> (&a)[R]) {
#pragma unroll
    for (int b = 0; b < R; b += LEN) {
#pragma unroll
        for (int k = 0; k < LEN / 2; k++) bfly<T>(a[b + k], a[b + k + LEN / 2], k * (32 / LEN));
    }
    if constexpr (LEN < R) dit_stages<T, R, LEN * 2>(a);
}

// In-place radix-R DFT (natural in, natural out) on v[OFF + STR*m], m = 0..R-1.
template <typename T, int R, int STR, int OFF, int P>
__device__ __forceinline__ void radix_fft(cpx<T> (&v)[P]) {
    static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "radix");
    cpx<T> a[R];
#pragma unroll
    for (int m = 0; m < R; m++) a[bitrev_c(m, R)] = v[OFF + STR * m];
    dit_stages<T, R, 2>(a);
#pragma unroll
    for (int m = 0; m < R; m++) v[OFF + STR * m] = a[m];
}

template <typename T, int R, int S, int P, int I>
struct RadixAll {
    static __device__ __forceinline__ void run(cpx<T> (&v)[P]) {
        radix_fft<T, R, S, I, P>(v);
        if constexpr (I + 1 < S) RadixAll<T, R, S, P, I + 1>::run(v);
    }
};

// ---------------- plans ----------------
// radices per pass; every radix divides P and the product is N.
template <typename T, int N> struct Plan;
#define SA_PLAN(TY, NN, PP, NPASS, R0, R1, R2, R3)                                             \
    template <> struct Plan<TY, NN> {                                                          \
        static constexpr int P = PP, NP = NPASS;                                               \
        __host__ __device__ static constexpr int radix(int i) { return i == 0 ? R0 : i == 1 ? R1 : i == 2 ? R2 : R3; } \
    };
SA_PLAN(float,    64,  8, 2,  8,  8,  1, 1)
SA_PLAN(float,   128, 16, 2, 16,  8,  1, 1)
SA_PLAN(float,   256, 16, 2, 16, 16,  1, 1)
SA_PLAN(float,   512, 32, 2, 32, 16,  1, 1)
SA_PLAN(float,  1024, 32, 2, 32, 32,  1, 1)
SA_PLAN(float,  2048, 32, 3, 32, 32,  2, 1)
SA_PLAN(float,  4096, 32, 3, 32, 32,  4, 1)
SA_PLAN(float,  8192, 32, 3, 32, 32,  8, 1)
SA_PLAN(float, 16384, 32, 3, 32, 32, 16, 1)
SA_PLAN(double,   64,  8, 2,  8,  8,  1, 1)
SA_PLAN(double,  128, 16, 2, 16,  8,  1, 1)
SA_PLAN(double,  256, 16, 2, 16, 16,  1, 1)
SA_PLAN(double,  512, 16, 3, 16, 16,  2, 1)
SA_PLAN(double, 1024, 16, 3, 16, 16,  4, 1)
SA_PLAN(double, 2048, 16, 3, 16, 16,  8, 1)
SA_PLAN(double, 4096, 16, 3, 16, 16, 16, 1)
SA_PLAN(double, 8192, 16, 4, 16, 16, 16, 2)
#undef SA_PLAN

template <typename T, int N> struct Geo {
    using PL = Plan<T, N>;
    static constexpr int P = PL::P;
    static constexpr int TPF = N / P;                        // threads per frame
    static constexpr int CTA = TPF > 128 ? TPF : 128;        // threads per CTA
    static constexpr int FPC = CTA / TPF;                    // frames per CTA pass
    static constexpr int MINB = 512 / CTA;                   // CTAs per SM the register cap allows (128 regs)
    static constexpr int R0 = PL::radix(0);
    static constexpr int SM_ELEMS = N + N / R0;              // padded elements per frame
    static constexpr size_t SMEM_BYTES = (size_t)FPC * SM_ELEMS * sizeof(cpx<T>);
    static constexpr int TW_ELEMS = (PL::NP - 1) * N;        // twiddle table entries
    __host__ __device__ static constexpr int ns(int pass) { int r = 1; for (int i = 0; i < pass; i++) r *= PL::radix(i); return r; }
};

// padded shared-memory index: one element of padding per R0 elements keeps the stride-R0
// writes of pass 0 and the unit-stride reads conflict-free
template <int R0> __device__ __forceinline__ int pad_idx(int i) { return i + i / R0; }

template <int TPF> __device__ __forceinline__ void frame_sync() {
    if constexpr (TPF <= 32) __syncwarp(); else __syncthreads();
}

// One pass: optional Stockham twiddle, S radix-R butterflies, and (unless last) the exchange.
template <typename T, int N, int PASS>
__device__ __forceinline__ void fft_pass(cpx<T> (&v)[Plan<T, N>::P], const int t, cpx<T>* __restrict__ sm,
                                         const cpx<T>* __restrict__ tw) {
    using G = Geo<T, N>;
    using PL = Plan<T, N>;
    constexpr int P = G::P, TPF = G::TPF, R = PL::radix(PASS), S = P / R, NS = G::ns(PASS);
    if constexpr (PASS > 0) {
        const cpx<T>* twp = tw + (size_t)(PASS - 1) * N + t;
#pragma unroll
        for (int s = 0; s < S; s++) {
#pragma unroll
            for (int m = 1; m < R; m++) {
                const cpx<T> w = __ldg(&twp[(s * R + m) * TPF]);
                const cpx<T> x = v[s + m * S];
                v[s + m * S] = mk2<T>(fma_t(-w.y, x.y, w.x * x.x), fma_t(w.y, x.x, w.x * x.y));
            }
        }
    }
    RadixAll<T, R, S, P, 0>::run(v);
    if constexpr (PASS + 1 < PL::NP) {
        frame_sync<TPF>();   // every reader of the previous exchange (or previous frame) is done
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int j = t + TPF * s;
            const int base = (j / NS) * (NS * R) + (j % NS);
#pragma unroll
            for (int m = 0; m < R; m++) sm[pad_idx<G::R0>(base + m * NS)] = v[s + m * S];
        }
        frame_sync<TPF>();
#pragma unroll
        for (int q = 0; q < P; q++) v[q] = sm[pad_idx<G::R0>(t + TPF * q)];
    }
}

// Full transform of the registers of one frame.
template <typename T, int N>
__device__ __forceinline__ void fft_frame(cpx<T> (&v)[Plan<T, N>::P], const int t, cpx<T>* __restrict__ sm,
                                          const cpx<T>* __restrict__ tw) {
    using PL = Plan<T, N>;
    fft_pass<T, N, 0>(v, t, sm, tw);
    if constexpr (PL::NP > 1) fft_pass<T, N, 1>(v, t, sm, tw);
    if constexpr (PL::NP > 2) fft_pass<T, N, 2>(v, t, sm, tw);
    if constexpr (PL::NP > 3) fft_pass<T, N, 3>(v, t, sm, tw);
}

}  // namespace sa
