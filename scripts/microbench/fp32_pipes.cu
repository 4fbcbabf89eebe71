// Microbenchmark: issue rate of FFMA / FADD / FMUL (3-register forms) vs packed f32x2 forms on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 4096

template <int MODE>
__global__ void bench(float* out, float a0, float b0) {
    float a = a0 + threadIdx.x * 1e-9f, b = b0;
    float r0 = 1.f, r1 = 2.f, r2 = 3.f, r3 = 4.f, r4 = 5.f, r5 = 6.f, r6 = 7.f, r7 = 8.f;
    unsigned long long p0 = 0x3f8000003f800000ull + threadIdx.x, p1 = p0 + 11, p2 = p0 + 23, p3 = p0 + 37, p4 = p0 + 41, p5 = p0 + 53, p6 = p0 + 67, p7 = p0 + 71;
    unsigned long long pa = (unsigned long long)__float_as_uint(a) << 32 | __float_as_uint(a);
    unsigned long long pb = (unsigned long long)__float_as_uint(b) << 32 | __float_as_uint(b);
    for (int i = 0; i < ITER; ++i) {
        if (MODE == 0) {  // FFMA 3-reg
            r0 = fmaf(r0, a, b); r1 = fmaf(r1, a, b); r2 = fmaf(r2, a, b); r3 = fmaf(r3, a, b);
            r4 = fmaf(r4, a, b); r5 = fmaf(r5, a, b); r6 = fmaf(r6, a, b); r7 = fmaf(r7, a, b);
        } else if (MODE == 1) {  // FADD
            r0 += a; r1 += a; r2 += a; r3 += a; r4 += a; r5 += a; r6 += a; r7 += a;
        } else if (MODE == 2) {  // FMUL
            r0 *= a; r1 *= a; r2 *= a; r3 *= a; r4 *= a; r5 *= a; r6 *= a; r7 *= a;
        } else if (MODE == 3) {  // fma.rn.f32x2
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p4) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p5) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p6) : "l"(pa), "l"(pb));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p7) : "l"(pa), "l"(pb));
        } else if (MODE == 4) {  // add.rn.f32x2
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p4) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p5) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p6) : "l"(pa));
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p7) : "l"(pa));
        } else if (MODE == 5) {  // mixed: 4 FADD + 4 FFMA (alu + fma pipes?)
            r0 += a; r1 = fmaf(r1, a, b); r2 += a; r3 = fmaf(r3, a, b);
            r4 += a; r5 = fmaf(r5, a, b); r6 += a; r7 = fmaf(r7, a, b);
        } else if (MODE == 6) {  // FFMA with immediate multiplier
            r0 = fmaf(r0, 1.0001f, b); r1 = fmaf(r1, 1.0001f, b); r2 = fmaf(r2, 1.0001f, b); r3 = fmaf(r3, 1.0001f, b);
            r4 = fmaf(r4, 1.0001f, b); r5 = fmaf(r5, 1.0001f, b); r6 = fmaf(r6, 1.0001f, b); r7 = fmaf(r7, 1.0001f, b);
        }
    }
    float s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
    unsigned long long q = p0 ^ p1 ^ p2 ^ p3 ^ p4 ^ p5 ^ p6 ^ p7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(q & 0xff);
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm, float* d_out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * blocks_per_sm;
    bench<MODE><<<grid, threads>>>(d_out, 1.0000001f, 1e-7f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench<MODE><<<grid, threads>>>(d_out, 1.0000001f, 1e-7f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instrs = (double)grid * threads / 32 * ITER * 8;
    const double per_sm_per_clk = warp_instrs / 148.0 / (ms * 1e-3 * 1.965e9);
    printf("%-22s threads/SM %5d  %.3f ms  warp-instr/clk/SM %.2f  (flop/clk/SM ~ %.0f)\n", name,
           threads * blocks_per_sm, ms, per_sm_per_clk,
           per_sm_per_clk * 32 * ((MODE == 3) ? 4 : (MODE == 4) ? 2 : (MODE == 0 || MODE == 6) ? 2 : (MODE == 5) ? 1.5 : 1));
}

int main() {
    float* d_out; cudaMalloc(&d_out, 148 * 2048 * sizeof(float));
    for (int w : {512, 1024, 2048}) {
        const int threads = 256, bps = w / 256;
        run<0>("FFMA 3-reg", threads, bps, d_out);
        run<6>("FFMA imm", threads, bps, d_out);
        run<1>("FADD", threads, bps, d_out);
        run<2>("FMUL", threads, bps, d_out);
        run<5>("FADD+FFMA mix", threads, bps, d_out);
        run<3>("fma.rn.f32x2", threads, bps, d_out);
        run<4>("add.rn.f32x2", threads, bps, d_out);
    }
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
