// decode.cuh -- fused sample decode, one IQ pair per call.
//
// Device restatement of the per-sample decode of S/services/SpectralService.java:42-63 and
// S/services/ExtractDownConvertService.java:79-98:
//   ci16  short/32768.0          cf32  float widened        cf64  double (stride 16)
//   cu8   ((b&0xFF)-127.5)/128   ci8   byte/128
// The integer decodes are exactly representable in FP32 ((2u-255)/256, k/2^15, k/2^7), so
// the FP32 results are bit-identical to the reference's FP64 values.  Integers are converted
// on the FMA pipe with the 2^23 mantissa trick rather than I2F (quarter-rate on sm_100).
// Byte order (S/sigmf/SigMfHelper.java:87-91) is a warp-uniform runtime flag.
#pragma once
#include "fft_core.cuh"

namespace sa {

enum { DK_CF32 = 0, DK_CI16 = 1, DK_C8 = 2, DK_CF64 = 3 };

struct LoadParams {
    const void* base;     // sample 0 of the capture (device pointer, aligned to one IQ pair)
    int         swap;     // 1: big-endian data
    uint32_t    c8_flip;  // 0 for cu8, 0x80808080 for ci8 (sign bit flip -> offset binary)
    float       c8_off;   // 127.5/128 for cu8, 1.0 for ci8
    float       c8_c;     // 256 + c8_off (one-FMA decode of spectrogram_mid_kernel.cuh)
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// SWAP is a template flag so the byte-order test is hoisted out of the per-point loop
template <typename T, int DK> struct Loader;

template <typename T> struct Loader<T, DK_CF32> {
    using raw_t = uint2;
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> load(const LoadParams& lp, int64_t i) {
        return decode<SWAP>(lp, __ldg(reinterpret_cast<const uint2*>(lp.base) + i));
    }
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> decode(const LoadParams&, uint2 w) {
        if constexpr (SWAP) { w.x = bswap32(w.x); w.y = bswap32(w.y); }
        return mk2<T>((T)__uint_as_float(w.x), (T)__uint_as_float(w.y));
    }
};

template <typename T> struct Loader<T, DK_CI16> {
    using raw_t = uint32_t;
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> load(const LoadParams& lp, int64_t i) {
        return decode<SWAP>(lp, __ldg(reinterpret_cast<const uint32_t*>(lp.base) + i));
    }
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> decode(const LoadParams&, uint32_t w) {
        if constexpr (SWAP) w = __byte_perm(w, 0, 0x2301);
        if constexpr (sizeof(T) == 4) {
            w ^= 0x80008000u;                                   // two's complement -> offset binary
            const float a = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610));   // 2^23 + u_lo
            const float b = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632));   // 2^23 + u_hi
            // (2^23 + u - 2^23 - 32768) / 32768, exact in one FMA
            return mk2<T>(__fmaf_rn(a, 1.0f / 32768.0f, -257.0f), __fmaf_rn(b, 1.0f / 32768.0f, -257.0f));
        } else {
            return mk2<T>((T)(int16_t)(w & 0xFFFFu) / (T)32768.0, (T)(int16_t)(w >> 16) / (T)32768.0);
        }
    }
};

template <typename T> struct Loader<T, DK_C8> {
    using raw_t = uint16_t;
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> load(const LoadParams& lp, int64_t i) {
        return decode<SWAP>(lp, __ldg(reinterpret_cast<const uint16_t*>(lp.base) + i));
    }
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> decode(const LoadParams& lp, uint16_t w16) {
        uint32_t w = w16;
        w ^= lp.c8_flip;
        const float a = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440)) - 8388608.0f;   // byte 0
        const float b = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441)) - 8388608.0f;   // byte 1
        // (u - off*128)/128 : exact (u/128 and the constant are both multiples of 2^-8)
        return mk2<T>((T)__fmaf_rn(a, 1.0f / 128.0f, -lp.c8_off), (T)__fmaf_rn(b, 1.0f / 128.0f, -lp.c8_off));
    }
};

template <typename T> struct Loader<T, DK_CF64> {
    using raw_t = uint4;
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> load(const LoadParams& lp, int64_t i) {
        return decode<SWAP>(lp, __ldg(reinterpret_cast<const uint4*>(lp.base) + i));
    }
    template <bool SWAP>
    static __device__ __forceinline__ cpx<T> decode(const LoadParams&, uint4 w) {
        if constexpr (SWAP) {
            uint32_t a = bswap32(w.y), b = bswap32(w.x), c = bswap32(w.w), d = bswap32(w.z);
            w = make_uint4(a, b, c, d);
        }
        const double re = __hiloint2double((int)w.y, (int)w.x);
        const double im = __hiloint2double((int)w.w, (int)w.z);
        return mk2<T>((T)re, (T)im);
    }
};

}  // namespace sa
