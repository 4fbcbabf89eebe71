// Operand forms of the packed FP32 FMA (FFMA2, sm_100a): which ones keep the full FMA rate?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_forms ffma2_forms.cu && ./ffma2_forms
//   0  acc = acc * y + z          (y, z loop-invariant registers)          -- f32x2.cu's form
//   1  acc = t[j] * w[i] + acc    (three distinct register pairs, FIR form)
//   2  acc = c[j] * w[i] + acc    (tap from the constant bank -> uniform register operand)
//   3  scalar FFMA, acc = c[j] * w[i] + acc (constant-bank operand)        -- shipped decimator form
#include <cstdio>
#include <cuda_runtime.h>

__constant__ __align__(16) float2 c_t[64];

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long x;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iters) {
    unsigned long long acc[8], w[8], t[8];
    float facc[16], fw[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = pack(seed + i, seed - i);
        w[i] = pack(seed * 0.001f * (i + 1), seed * 0.002f);
        t[i] = pack(1.0f + seed * 1e-6f * i, 1.0f - seed * 1e-6f * i);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { facc[i] = seed + i; fw[i] = seed * 0.001f * (i + 1); }
    const unsigned long long* ct = reinterpret_cast<const unsigned long long*>(c_t);
    const float* cf = reinterpret_cast<const float*>(c_t);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) acc[i] = fma2(acc[i], t[0], t[1]);
                if (MODE == 1) acc[i] = fma2(t[j], w[(i + j) & 7], acc[i]);
                if (MODE == 2) acc[i] = fma2(ct[j], w[(i + j) & 7], acc[i]);
                if (MODE == 3) {
                    facc[2 * i] = fmaf(cf[2 * j], fw[(2 * i + j) & 15], facc[2 * i]);
                    facc[2 * i + 1] = fmaf(cf[2 * j + 1], fw[(2 * i + 1 + j) & 15], facc[2 * i + 1]);
                }
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float2 a = *reinterpret_cast<const float2*>(&acc[i]); s += a.x + a.y; }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += facc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 5000;
    k<MODE><<<148 * 8, 256>>>(out, 1.0f, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 148.0 * 8 * 256 * 128.0 * iters;   // scalar FMAs
    printf("%-34s %.3f ms  %.1f TFLOP/s\n", name, ms, 2 * fma / ms / 1e9);
    cudaFree(out);
}

int main() {
    float h[128];
    for (int i = 0; i < 128; ++i) h[i] = 1.0f + 1e-6f * i;
    cudaMemcpyToSymbol(c_t, h, sizeof(h));
    run<0>("FFMA2 acc*y+z (invariant y,z)");
    run<1>("FFMA2 t*w+acc (3 register pairs)");
    run<2>("FFMA2 c[]*w+acc (uniform operand)");
    run<3>("FFMA  c[]*w+acc (constant operand)");
    return 0;
}
