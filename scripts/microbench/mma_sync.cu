// Throughput of the legacy warp-level tensor path on sm_100a: mma.sync m16n8k16 bf16 and m16n8k8 tf32,
// fp32 accumulate, independent accumulator chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync mma_sync.cu && ./mma_sync
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    unsigned a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
    unsigned b[2] = {0x3f803f80u, 0x3f803f80u + threadIdx.x};
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flop_per_mma) {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 4000;
    k<MODE><<<148 * 8, 256>>>(out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = 148.0 * 8 * 8 /*warps*/ * 8.0 * iters;
    printf("%-22s %.3f ms  %.1f TFLOP/s\n", name, ms, mmas * flop_per_mma / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("mma.sync bf16 m16n8k16", 2.0 * 16 * 8 * 16);
    run<1>("mma.sync tf32 m16n8k8", 2.0 * 16 * 8 * 8);
    return 0;
}
