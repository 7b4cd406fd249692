// fma_rate.cu -- FP32 pipe rates on sm_100a: scalar FFMA (register and immediate forms), FADD, packed FFMA2 / FADD2,
// and an FFMA + ALU (LOP3) mix.  Prints thread-operations per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long pk2;
__device__ __forceinline__ pk2 fma2(pk2 a, pk2 b, pk2 c) { pk2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ pk2 add2(pk2 a, pk2 b) { pk2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int ACC = 16, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, float x, float y, long long* clk) {
    float a[ACC];
    pk2 p[ACC];
    for (int i = 0; i < ACC; i++) { a[i] = threadIdx.x + i; p[i] = ((pk2)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 0.5f); }
    const pk2 xx = ((pk2)__float_as_uint(x) << 32) | __float_as_uint(x), yy = ((pk2)__float_as_uint(y) << 32) | __float_as_uint(y);
    unsigned lop = threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if (MODE == 0) a[i] = ffma(a[i], x, y);                         // 3 registers
            if (MODE == 1) a[i] = ffma(a[i], 1.0001f, 0.25f);               // immediates (the compiler may fold one)
            if (MODE == 2) a[i] = fadd(a[i], x);
            if (MODE == 3) p[i] = fma2(p[i], xx, yy);
            if (MODE == 4) p[i] = add2(p[i], xx);
            if (MODE == 5) { a[i] = ffma(a[i], x, y); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lop) : "r"(lop >> 3), "r"(it)); }
            if (MODE == 6) { p[i] = fma2(p[i], xx, yy); a[i] = ffma(a[i], x, y); }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ACC; i++) s += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + lop;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char* name, double ops_per_inner) {
    float* out; long long* clk;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&clk, 148 * sizeof(long long));
    k<MODE><<<148, 1024>>>(out, 1.0001f, 0.25f, clk);
    k<MODE><<<148, 1024>>>(out, 1.0001f, 0.25f, clk);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += h[i]; c /= 148;
    printf("%-34s %8.1f thread-ops/clk/SM  (%.0f clocks)\n", name, 1024.0 * ACC * ITERS * ops_per_inner / c, c);
    cudaFree(out); cudaFree(clk);
}

int main() {
    run<0>("FFMA  r,r,r", 1);
    run<1>("FFMA  r,imm,imm", 1);
    run<2>("FADD  r,r", 1);
    run<3>("FFMA2 (x2 lanes)", 2);
    run<4>("FADD2 (x2 lanes)", 2);
    run<5>("FFMA + LOP3 (FFMA counted)", 1);
    run<6>("FFMA2 + FFMA (3 lanes counted)", 3);
    return 0;
}
