// fft_core.cuh -- register-resident radix-R butterflies and shared-memory Stockham passes.
//
// Replaces the FP64 radix-2 transform the reference calls per frame
// (FastFourierTransformer(STANDARD).transform(.., FORWARD), S/services/SpectralService.java:23,68):
// X[k] = sum_n x[n] exp(-2 pi i k n / N), unnormalised.
//
// Layout of one transform of N points over TPF = N/P threads (P points per thread):
//   at the start of every pass thread t holds elements  t + TPF*q  (q = 0..P-1) in v[q];
//   a pass of radix R runs S = P/R independent radix-R butterflies per thread on the register
//   sub-arrays v[s + m*S] (m = 0..R-1); passes after the first multiply by the Stockham twiddle
//   W_N^{(j mod Ns) * m * N/(Ns*R)}, j = t + TPF*s, Ns = product of earlier radices, fused into
//   the first butterfly stage (as is the window in pass 0);
//   the pass writes element (j/Ns)*Ns*R + (j mod Ns) + m*Ns to shared memory and every thread
//   reads back t + TPF*q.  After the last pass v[q] = X[t + TPF*q] (no final exchange), so both
//   the global loads of pass 0 and the global stores of the epilogue are coalesced.
// Butterflies are radix-2 DIT on bit-reversed registers with compile-time twiddles in FMA form
// (6 FMA per non-trivial butterfly, 4 ADD for W = 1 and W = -i).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// tuning knobs (kernel K2): resident warps per SM of the one-warp-per-frame FP32
// kernels, and the widest plan that keeps its window rows in shared memory instead of registers
#ifndef SA_WARPS_1WPF
#define SA_WARPS_1WPF 16
#endif
#ifndef SA_CTA_1WPF
#define SA_CTA_1WPF 512
#endif
#ifndef SA_CTA_F64
#define SA_CTA_F64 512      // FP64 plans whose frames fit a warp (nfft <= 512); cf64 256: 139 -> 147 Gsamples/s
#endif
#ifndef SA_WIN_SMEM_MAX_TPF
#define SA_WIN_SMEM_MAX_TPF 128
#endif
#ifndef SA_TW_SMEM_MAX_TPF
#define SA_TW_SMEM_MAX_TPF 32
#endif
#ifndef SA_PACKED_SMALL_RADIX
#define SA_PACKED_SMALL_RADIX 0
#endif
#ifndef SA_TW_RECURRENCE
#define SA_TW_RECURRENCE 1
#endif

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

// cos(2*pi*q/64), q = 0..16 (radix-64 passes)
__host__ __device__ constexpr double cos64_tab(int q) {
    return q == 0 ? 1.0
         : q == 1 ? 0.9951847266721968862448369531094799
         : q == 2 ? 0.980785280403230449126182236134239
         : q == 3 ? 0.95694033573220886493579788698027
         : q == 4 ? 0.9238795325112867561281831893967883
         : q == 5 ? 0.8819212643483550297127568636603884
         : q == 6 ? 0.8314696123025452370787883776179058
         : q == 7 ? 0.7730104533627369608109066097584698
         : q == 8 ? 0.707106781186547524400844362104849
         : q == 9 ? 0.6343932841636454982151716132254934
         : q == 10 ? 0.5555702330196022247428308139485329
         : q == 11 ? 0.4713967368259976485563876259052544
         : q == 12 ? 0.3826834323650897717284599840303989
         : q == 13 ? 0.2902846772544623676361923758173953
         : q == 14 ? 0.1950903220161282678482848684770222
         : q == 15 ? 0.09801714032956060199419556388864184
         : 0.0;
}
// W64^q = exp(-2*pi*i*q/64) for q = 0..32
__host__ __device__ constexpr double w64_re(int q) { return q <= 16 ? cos64_tab(q) : -cos64_tab(32 - q); }
__host__ __device__ constexpr double w64_im(int q) { return q <= 16 ? -cos64_tab(16 - q) : -cos64_tab(q - 16); }
// twiddle of base B (32 or 64): exponent q in units of 2*pi/B, q = 0..B/2
template <int B> __host__ __device__ constexpr double wB_re(int q) { return B == 64 ? w64_re(q) : w32_re(q); }
template <int B> __host__ __device__ constexpr double wB_im(int q) { return B == 64 ? w64_im(q) : w32_im(q); }

// DIT butterfly: (a, b) <- (a + W b, a - W b), W = W_B^q, q compile-time after unrolling.
template <typename T, int B = 32>
__device__ __forceinline__ void bfly(cpx<T>& a, cpx<T>& b, const int q) {
    if (q == 0) {
        T bx = b.x, by = b.y;
        b.x = a.x - bx; b.y = a.y - by;
        a.x = a.x + bx; a.y = a.y + by;
    } else if (q == B / 4) {        // W = -i : W b = (b.y, -b.x)
        T bx = b.x, by = b.y;
        b.x = a.x - by; b.y = a.y + bx;
        a.x = a.x + by; a.y = a.y - bx;
    } else {
        const T wr = (T)wB_re<B>(q), wi = (T)wB_im<B>(q);
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
    if (n == 64) return ((x & 1) << 5) | ((x & 2) << 3) | ((x & 4) << 1) | ((x & 8) >> 1) | ((x & 16) >> 3) | ((x & 32) >> 5);
    return n == 32 ? r5 : n == 16 ? (r5 >> 1) : n == 8 ? (r5 >> 2) : n == 4 ? (r5 >> 3) : (r5 >> 4);
}

// one DIT stage of span LEN over a[0..R), then the next (template recursion: nvcc does not
// unroll a loop whose induction variable is shifted)
template <typename T, int R, int LEN>
__device__ __forceinline__ void dit_stages(cpx<T> (&a)[R]) {
    constexpr int B = R > 32 ? 64 : 32;
#pragma unroll
    for (int b = 0; b < R; b += LEN) {
#pragma unroll
        for (int k = 0; k < LEN / 2; k++) bfly<T, B>(a[b + k], a[b + k + LEN / 2], k * (B / LEN));
    }
    if constexpr (LEN < R) dit_stages<T, R, LEN * 2>(a);
}

// ---- packed FP32 (Blackwell FFMA2 / FADD2 / FMUL2: two FP32 lanes per issue slot) ----
// A packed value is a 64-bit register pair (lo, hi).  With the pairing (a[i], a[i + R/2]) of the
// bit-reversed work array, the DIT stages of span 4 .. R/2 run the SAME butterfly (same twiddle) in both
// lanes, so each issue slot retires two butterfly operations; twiddle constants are scalar immediates
// that FFMA2 broadcasts.  The first stage (span 2, fused multipliers) and the last stage (span R, lanes
// interact) stay scalar and absorb the re-pairing for free.
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pack2(float lo, float hi) { pk2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(pk2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pk2 bcast2(float c) { return pack2(c, c); }
__device__ __forceinline__ pk2 fma2(pk2 a, pk2 b, pk2 c) { pk2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ pk2 add2(pk2 a, pk2 b) { pk2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pk2 sub2(pk2 a, pk2 b) { pk2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ pk2 mul2(pk2 a, pk2 b) { pk2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// negation written on the scalar halves so that ptxas folds it into the operand modifier of FFMA2/FADD2
__device__ __forceinline__ pk2 neg2(pk2 a) { float lo, hi; unpack2(a, lo, hi); return pack2(-lo, -hi); }

// two DIT butterflies at once: (a, b) <- (a + W b, a - W b), W = W32^q in both lanes
template <int B = 32>
__device__ __forceinline__ void bfly2(pk2& ar, pk2& ai, pk2& br, pk2& bi, const int q) {
    if (q == 0) {
        const pk2 xr = br, xi = bi;
        br = sub2(ar, xr); bi = sub2(ai, xi);
        ar = add2(ar, xr); ai = add2(ai, xi);
    } else if (q == B / 4) {        // W = -i : W b = (b.im, -b.re)
        const pk2 xr = br, xi = bi;
        br = sub2(ar, xi); bi = add2(ai, xr);
        ar = add2(ar, xi); ai = sub2(ai, xr);
    } else {
        const float wr = (float)wB_re<B>(q), wi = (float)wB_im<B>(q);
        const pk2 o_r = fma2(bcast2(-wi), bi, fma2(bcast2(wr), br, ar));
        const pk2 o_i = fma2(bcast2(wi), br, fma2(bcast2(wr), bi, ai));
        br = fma2(bcast2(2.0f), ar, neg2(o_r));
        bi = fma2(bcast2(2.0f), ai, neg2(o_i));
        ar = o_r; ai = o_i;
    }
}

// packed stages of span LEN .. H on H = R/2 packed positions
template <int H, int LEN>
__device__ __forceinline__ void dit_stages_packed(pk2 (&pr)[H], pk2 (&pi)[H]) {
    constexpr int B = H > 16 ? 64 : 32;          // H = R/2 packed positions
#pragma unroll
    for (int b = 0; b < H; b += LEN) {
#pragma unroll
        for (int k = 0; k < LEN / 2; k++) bfly2<B>(pr[b + k], pi[b + k], pr[b + k + LEN / 2], pi[b + k + LEN / 2], k * (B / LEN));
    }
    if constexpr (LEN < H) dit_stages_packed<H, LEN * 2>(pr, pi);
}

// stages of span 4 .. R on the bit-reversed work array (stage of span 2 already done)
template <typename T, int R>
__device__ __forceinline__ void dit_tail(cpx<T> (&a)[R]) {
    if constexpr (sizeof(T) == 4 && R >= 8) {
        constexpr int H = R / 2;
        pk2 pr[H], pi[H];
#pragma unroll
        for (int i = 0; i < H; i++) { pr[i] = pack2(a[i].x, a[i + H].x); pi[i] = pack2(a[i].y, a[i + H].y); }
        dit_stages_packed<H, 4>(pr, pi);
#pragma unroll
        for (int i = 0; i < H; i++) { unpack2(pr[i], a[i].x, a[i + H].x); unpack2(pi[i], a[i].y, a[i + H].y); }
#pragma unroll
        for (int k = 0; k < H; k++) bfly<T, (R > 32 ? 64 : 32)>(a[k], a[k + H], k * ((R > 32 ? 64 : 32) / R));     // last stage: lanes interact
    } else {
        if constexpr (R > 2) dit_stages<T, R, 4>(a);
    }
}

// Multiplier fused into the first DIT stage of a radix-R butterfly:
//   MUL_NONE  plain
//   MUL_REAL  x[m] *= window; wr[2m], wr[2m+1] hold the factors of elements m and m + R/2 (pass 0)
//   MUL_CPX   x[m] *= tw[m]  (m >= 1; tw[0] == 1)    (Stockham twiddle, later passes)
//   MUL_REC   like MUL_CPX, but tw[m] = om^m generated in registers: the pair (om^m, om^(m+R/2)) advances
//             by one packed complex multiply per butterfly (plans whose pass-1 twiddle is lane * m)
enum { MUL_NONE = 0, MUL_REAL = 1, MUL_CPX = 2, MUL_REC = 3 };

// per-lane seeds of the twiddle recurrence: om = W_N^t, oh = om^(R/2)
template <typename T> struct TwSeed { cpx<T> om, oh; cpx<T> q_lo, q_hi; };   // q_*: om^(R/4), om^(3R/4) (radix-64 re-seed)

template <typename T> struct TwPair { cpx<T> lo, hi; };     // twiddles of elements m and m + R/2
// TW_SMEM: the table was copied to shared memory (plain loads); otherwise read-only global loads
template <bool TW_SMEM> __device__ __forceinline__ TwPair<float> ldg_tw(const TwPair<float>* p) {
    float4 w;
    if constexpr (TW_SMEM) w = *reinterpret_cast<const float4*>(p);
    else w = __ldg(reinterpret_cast<const float4*>(p));
    TwPair<float> r; r.lo = make_float2(w.x, w.y); r.hi = make_float2(w.z, w.w); return r;
}
template <bool TW_SMEM> __device__ __forceinline__ TwPair<double> ldg_tw(const TwPair<double>* p) {
    TwPair<double> r;
    if constexpr (TW_SMEM) { r.lo = reinterpret_cast<const double2*>(p)[0]; r.hi = reinterpret_cast<const double2*>(p)[1]; }
    else { r.lo = __ldg(reinterpret_cast<const double2*>(p)); r.hi = __ldg(reinterpret_cast<const double2*>(p) + 1); }
    return r;
}

// In-place radix-R DFT (natural in, natural out) on v[OFF + STR*m], m = 0..R-1.  The first DIT
// stage pairs (m, m + R/2); an element-wise multiplier is folded into it:
//   real  : (wa a + wb b, wa a - wb b)            6 ops per component pair instead of 8
//   cpx   : t = ta a ; o0 = t + tb b ; o1 = 2t - o0   10 FMA-class ops instead of 12
template <typename T, int R, int STR, int OFF, int P, int MUL, bool TW_SMEM>
__device__ __forceinline__ void radix_fft(cpx<T> (&v)[P], const T* __restrict__ wr,
                                          const TwPair<T>* __restrict__ tw, const int tw_stride,
                                          const TwSeed<T>& seed) {
    static_assert(R == 2 || R == 4 || R == 8 || R == 16 || R == 32 || R == 64, "radix");
    if constexpr (SA_PACKED_SMALL_RADIX && sizeof(T) == 4 && R <= 4 && (MUL == MUL_REAL || MUL == MUL_NONE)) {
        // ablation, off by default (C2 2.55 -> 2.68 ms: the re-pairing costs more than the packed ops save):
        // small first-pass radices (spectrogram_mid_kernel): a complex value is already a packed (re, im) register
        // pair, so the window multiply and the W = 1 butterflies run as FMUL2 / FFMA2 / FADD2 with the real factor
        // as the scalar-broadcast operand -- 3 instead of 6 (2 instead of 4) instructions per butterfly
        pk2 o[R];
#pragma unroll
        for (int m = 0; m < R / 2; m++) {
            const cpx<T> x = v[OFF + STR * m], y = v[OFF + STR * (m + R / 2)];
            const pk2 X = pack2((float)x.x, (float)x.y), Y = pack2((float)y.x, (float)y.y);
            if constexpr (MUL == MUL_REAL) {
                const cpx<T> wp = *reinterpret_cast<const cpx<T>*>(wr + 2 * (OFF + STR * m));
                const pk2 t = mul2(X, bcast2((float)wp.x));
                o[2 * m] = fma2(Y, bcast2((float)wp.y), t);
                o[2 * m + 1] = fma2(Y, bcast2(-(float)wp.y), t);
            } else {
                o[2 * m] = add2(X, Y);
                o[2 * m + 1] = sub2(X, Y);
            }
        }
        if constexpr (R == 2) {
            float re, im;
            unpack2(o[0], re, im); v[OFF] = mk2<T>((T)re, (T)im);
            unpack2(o[1], re, im); v[OFF + STR] = mk2<T>((T)re, (T)im);
        } else {
            // o = (x0+x2, x0-x2, x1+x3, x1-x3); X0 = o0+o2, X2 = o0-o2, X1 = o1 - i o3, X3 = o1 + i o3
            float ar, ai, br, bi, re, im;
            unpack2(add2(o[0], o[2]), re, im); v[OFF] = mk2<T>((T)re, (T)im);
            unpack2(sub2(o[0], o[2]), re, im); v[OFF + 2 * STR] = mk2<T>((T)re, (T)im);
            unpack2(o[1], ar, ai); unpack2(o[3], br, bi);
            v[OFF + STR] = mk2<T>((T)(ar + bi), (T)(ai - br));
            v[OFF + 3 * STR] = mk2<T>((T)(ar - bi), (T)(ai + br));
        }
        return;
    }
    cpx<T> a[R];
    pk2 rec_re = 0, rec_im = 0;                 // (om^m, om^(m+R/2)) as packed re / im (MUL_REC)
    if constexpr (MUL == MUL_REC) { rec_re = pack2(1.0f, (float)seed.oh.x); rec_im = pack2(0.0f, (float)seed.oh.y); }
#pragma unroll
    for (int m = 0; m < R / 2; m++) {
        const cpx<T> x = v[OFF + STR * m], y = v[OFF + STR * (m + R / 2)];
        cpx<T> o0, o1;
        if constexpr (MUL == MUL_REAL) {
            // window row in pair order: factors of register e and e + P/2 are adjacent (e = OFF + STR*m < P/2)
            static_assert(MUL != MUL_REAL || (STR * R == P), "the window is applied in pass 0");
            const cpx<T> wp = *reinterpret_cast<const cpx<T>*>(wr + 2 * (OFF + STR * m));
            const T wa = wp.x, wb = wp.y;
            const T tx = wa * x.x, ty = wa * x.y;
            o0 = mk2<T>(fma_t(wb, y.x, tx), fma_t(wb, y.y, ty));
            o1 = mk2<T>(fma_t(-wb, y.x, tx), fma_t(-wb, y.y, ty));
        } else if constexpr (MUL == MUL_CPX || MUL == MUL_REC) {
            TwPair<T> w;
            if constexpr (MUL == MUL_REC) {
                float lr, hr, li, hi;
                unpack2(rec_re, lr, hr); unpack2(rec_im, li, hi);
                w.lo = mk2<T>((T)lr, (T)li); w.hi = mk2<T>((T)hr, (T)hi);
                if (R == 64 && m + 1 == R / 4) {      // long recurrence: restart from the exact pair half way
                    rec_re = pack2((float)seed.q_lo.x, (float)seed.q_hi.x);
                    rec_im = pack2((float)seed.q_lo.y, (float)seed.q_hi.y);
                } else if (m + 1 < R / 2) {           // advance both lanes: w <- w * om
                    const pk2 orr = bcast2((float)seed.om.x), oii = bcast2((float)seed.om.y);
                    const pk2 nre = fma2(neg2(oii), rec_im, mul2(orr, rec_re));
                    rec_im = fma2(oii, rec_re, mul2(orr, rec_im));
                    rec_re = nre;
                }
            } else {
                w = ldg_tw<TW_SMEM>(tw + (size_t)m * tw_stride);
            }
            cpx<T> t = x;
            if (m != 0) t = mk2<T>(fma_t(-w.lo.y, x.y, w.lo.x * x.x), fma_t(w.lo.y, x.x, w.lo.x * x.y));
            o0 = mk2<T>(fma_t(-w.hi.y, y.y, fma_t(w.hi.x, y.x, t.x)), fma_t(w.hi.y, y.x, fma_t(w.hi.x, y.y, t.y)));
            o1 = mk2<T>(fma_t((T)2, t.x, -o0.x), fma_t((T)2, t.y, -o0.y));
        } else {
            o0 = mk2<T>(x.x + y.x, x.y + y.y);
            o1 = mk2<T>(x.x - y.x, x.y - y.y);
        }
        a[bitrev_c(m, R)] = o0;
        a[bitrev_c(m, R) + 1] = o1;
    }
    dit_tail<T, R>(a);
#pragma unroll
    for (int m = 0; m < R; m++) v[OFF + STR * m] = a[m];
}

template <typename T, int R, int S, int P, int MUL, bool TW_SMEM, int I>
struct RadixAll {
    static __device__ __forceinline__ void run(cpx<T> (&v)[P], const T* __restrict__ wr,
                                               const TwPair<T>* __restrict__ tw, const int tpf,
                                               const TwSeed<T>& seed) {
        // twiddle pairs of sub-butterfly I sit at tw[(I*(R/2) + m) * tpf]
        radix_fft<T, R, S, I, P, MUL, TW_SMEM>(v, wr, tw + (size_t)I * (R / 2) * tpf, tpf, seed);
        if constexpr (I + 1 < S) RadixAll<T, R, S, P, MUL, TW_SMEM, I + 1>::run(v, wr, tw, tpf, seed);
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
    // threads per CTA: the FP32 plans whose frames fit a warp use ONE 16-warp CTA per SM: one copy of the
    // twiddle/window tables per SM, and the 16+ consecutive frames a CTA takes per step are one contiguous
    // stretch of the capture (DRAM page locality; the overlapping halves of neighbouring frames are fetched
    // together).  Measured on the 1024-point kernel: 128 -> 256 -> 512 threads = 0.865 -> 0.850 -> 0.801 ms
    static constexpr int CTA = (sizeof(T) == 4 && TPF <= 32) ? SA_CTA_1WPF
                             : (sizeof(T) == 8 && TPF <= 32) ? SA_CTA_F64 : (TPF > 128 ? TPF : 128);
    static constexpr int FPC = CTA / TPF;                    // frames per CTA pass
    // resident CTAs per SM the register cap is set for: one-warp-per-frame FP32 kernels run 5 warps per
    // scheduler (96 registers), everything else 4 (128 registers)
    static constexpr int MINB = (sizeof(T) == 4 && TPF <= 32) ? SA_WARPS_1WPF * 32 / CTA : 512 / CTA;
    static constexpr int WROW = P + 2;                       // window row per thread: P factors + 2 pad (bank spread)
    static constexpr bool WIN_SMEM = TPF <= SA_WIN_SMEM_MAX_TPF;             // window rows in shared memory, else in registers
    static constexpr size_t WIN_BYTES = WIN_SMEM ? (size_t)TPF * WROW * sizeof(T) : 0;
    static constexpr int R0 = PL::radix(0);
    static constexpr int PADW = sizeof(T) == 4 ? 2 : 1;      // pad elements per R0 (keeps 16-byte alignment)
    static constexpr int SM_ELEMS = N + (N / R0) * PADW;     // padded elements per frame
    static constexpr size_t SMEM_BYTES = (size_t)FPC * SM_ELEMS * sizeof(cpx<T>);
    // two-pass plans with one butterfly per thread in pass 1 (twiddle = lane * m) can generate the
    // twiddles by a register recurrence instead of loading them
    static constexpr bool TW_REC = SA_TW_RECURRENCE && sizeof(T) == 4 && PL::NP == 2 && PL::radix(1) == P;
    // one-warp-per-frame plans keep the Stockham twiddle table in shared memory as well
    static constexpr bool TW_SMEM = TPF <= SA_TW_SMEM_MAX_TPF;
    static constexpr size_t TW_BYTES = TW_SMEM ? (size_t)(PL::NP - 1) * N * sizeof(cpx<T>) : 0;
    static constexpr size_t EXTRA_WIN_OFF = SMEM_BYTES + TW_BYTES;   // window rows follow the twiddles
    static constexpr int TW_ELEMS = (PL::NP - 1) * N;        // twiddle table entries
    __host__ __device__ static constexpr int ns(int pass) { int r = 1; for (int i = 0; i < pass; i++) r *= PL::radix(i); return r; }
};

// padded shared-memory index: PADW elements (16 bytes) of padding per R0 elements keep the
// stride-R0 128-bit writes of pass 0 and the unit-stride reads conflict-free
template <int R0, int PADW> __device__ __forceinline__ int pad_idx(int i) { return i + (i / R0) * PADW; }

// BSYNC forces a CTA barrier (thread mappings where a frame's threads are spread over several warps)
template <int TPF, bool BSYNC = false> __device__ __forceinline__ void frame_sync() {
    if constexpr (TPF <= 32 && !BSYNC) __syncwarp(); else __syncthreads();
}

// One pass: S radix-R butterflies (window or Stockham twiddle fused into their first stage) and,
// unless it is the last pass, the exchange through shared memory.
template <typename T, int N, int PASS, bool WIN, bool REC, bool BSYNC = false>
__device__ __forceinline__ void fft_pass(cpx<T> (&v)[Plan<T, N>::P], const int t, cpx<T>* __restrict__ sm,
                                         const cpx<T>* __restrict__ tw, const T* __restrict__ win,
                                         const TwSeed<T>& seed) {
    using G = Geo<T, N>;
    using PL = Plan<T, N>;
    constexpr int P = G::P, TPF = G::TPF, R = PL::radix(PASS), S = P / R, NS = G::ns(PASS);
    if constexpr (PASS == 0) {
        RadixAll<T, R, S, P, WIN ? MUL_REAL : MUL_NONE, false, 0>::run(v, win, nullptr, TPF, seed);
    } else {
        const TwPair<T>* twp = reinterpret_cast<const TwPair<T>*>(tw) + (size_t)(PASS - 1) * (N / 2) + t;
        if constexpr (REC && G::TW_REC) RadixAll<T, R, S, P, MUL_REC, false, 0>::run(v, nullptr, nullptr, TPF, seed);
        else RadixAll<T, R, S, P, MUL_CPX, G::TW_SMEM, 0>::run(v, nullptr, twp, TPF, seed);
    }
    if constexpr (PASS + 1 < PL::NP) {
        frame_sync<TPF, BSYNC>();   // every reader of the previous exchange (or previous frame) is done
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int j = t + TPF * s;
            const int base = (j / NS) * (NS * R) + (j % NS);
            if constexpr (NS == 1 && sizeof(T) == 4) {
                // pass 0: R consecutive elements per butterfly -> 128-bit stores of element pairs
                float4* dst = reinterpret_cast<float4*>(sm + pad_idx<G::R0, G::PADW>(base));
#pragma unroll
                for (int m = 0; m < R; m += 2)
                    dst[m / 2] = make_float4(v[s + m * S].x, v[s + m * S].y, v[s + (m + 1) * S].x, v[s + (m + 1) * S].y);
            } else {
#pragma unroll
                for (int m = 0; m < R; m++) sm[pad_idx<G::R0, G::PADW>(base + m * NS)] = v[s + m * S];
            }
        }
        frame_sync<TPF, BSYNC>();
#pragma unroll
        for (int q = 0; q < P; q++) v[q] = sm[pad_idx<G::R0, G::PADW>(t + TPF * q)];
    }
}

// Seeds of the twiddle recurrence (kernels instantiated with REC), read once per thread from the pair table: pair (m'=1).lo = W_N^t,
// pair (m'=0).hi = W_N^(t*P/2).
template <typename T, int N>
__device__ __forceinline__ TwSeed<T> load_tw_seed(const cpx<T>* __restrict__ tw, const int t) {
    using G = Geo<T, N>;
    TwSeed<T> s;
    s.om = mk2<T>((T)1, (T)0); s.oh = s.om;
    if constexpr (G::TW_REC) {
        const TwPair<T>* p = reinterpret_cast<const TwPair<T>*>(tw);
        s.om = ldg_tw<false>(p + 1 * G::TPF + t).lo;
        s.oh = ldg_tw<false>(p + 0 * G::TPF + t).hi;
    }
    return s;
}

// Full transform of the registers of one frame; win (T[P], thread-private) multiplies the inputs
// when WIN is set.
template <typename T, int N, bool WIN, bool REC = false, bool BSYNC = false>
__device__ __forceinline__ void fft_frame(cpx<T> (&v)[Plan<T, N>::P], const int t, cpx<T>* __restrict__ sm,
                                          const cpx<T>* __restrict__ tw, const T* __restrict__ win,
                                          const TwSeed<T>& seed) {
    using PL = Plan<T, N>;
    fft_pass<T, N, 0, WIN, REC, BSYNC>(v, t, sm, tw, win, seed);
    if constexpr (PL::NP > 1) fft_pass<T, N, 1, WIN, REC, BSYNC>(v, t, sm, tw, win, seed);
    if constexpr (PL::NP > 2) fft_pass<T, N, 2, WIN, REC, BSYNC>(v, t, sm, tw, win, seed);
    if constexpr (PL::NP > 3) fft_pass<T, N, 3, WIN, REC, BSYNC>(v, t, sm, tw, win, seed);
}

}  // namespace sa
