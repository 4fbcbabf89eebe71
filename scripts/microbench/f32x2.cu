// Throughput of packed FP32 (FADD2 / FFMA2, sm_100a) against scalar FADD / FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iters) {
    float a[16], b = seed * 1.0001f, c = seed * 0.5f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (MODE == 0) { a[i] = a[i] + b; a[i + 1] = a[i + 1] + c; }
            if (MODE == 1) { a[i] = fmaf(a[i], b, c); a[i + 1] = fmaf(a[i + 1], c, b); }
            if (MODE == 2 || MODE == 3) {
                unsigned long long x, y, z;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(c));
                asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c), "f"(b));
                if (MODE == 2) asm("add.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
                else asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 1.0f, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 1.0f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 8 * 256 * 16.0 * iters;   // scalar-equivalent operations
    printf("%-8s %.3f ms  %.2f T scalar-op/s\n", name, ms, ops / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("FADD");
    run<1>("FFMA");
    run<2>("FADD2");
    run<3>("FFMA2");
    return 0;
}
