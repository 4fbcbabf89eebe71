// Factor-2 FIR decimator (381 taps): instruction-form variants of the inner product, timed alone.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o decimate_variants decimate_variants.cu && ./decimate_variants
//   A  scalar FFMA, taps as constant-bank operands, polyphase split (the shipped form)
//   B  packed FFMA2, taps as uniform-register operands (constant bank)
//   C  packed FFMA2, taps read from shared memory into registers (LDS.128 broadcast)
//   D  as C with 8 outputs per thread
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int kTile = 1024, kHalo = 96, kSpan = kTile + 192;
__constant__ float4 c_tap4[2][50];
__constant__ __align__(16) float2 c_tap2[192];

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int E_MAX, int PHASE>
__device__ __forceinline__ void fir_phase(const float* __restrict__ xs, int t, float (&acc)[4]) {
#pragma unroll
    for (int g = 0; g <= (E_MAX + 3) / 4; ++g) {
        const float4 q = *reinterpret_cast<const float4*>(xs + 4 * t + 4 * g);
        const float v[4] = {q.x, q.y, q.z, q.w};
        const float4 ta = c_tap4[PHASE][g], tb = c_tap4[PHASE][g + 1];
        const float tap[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int e = 4 * g + c - r;
                if (e >= 1 && e <= E_MAX) acc[r] = fmaf(tap[c - r + 3], v[c], acc[r]);
            }
    }
}

__global__ void __launch_bounds__(256) dec_a(const float* src, float* dst, int len_in) {
    __shared__ __align__(16) float xe[kSpan];
    __shared__ __align__(16) float xo[kSpan];
    const int len_out = (len_in + 1) >> 1;
    for (int mb = blockIdx.x * kTile; mb < len_out; mb += gridDim.x * kTile) {
        for (int q = threadIdx.x; q < kSpan; q += 256) {
            const int i = 2 * (mb - kHalo + q);
            float a = 0.0f, b = 0.0f;
            if (i >= 0 && i + 1 < len_in) { const float2 v = *reinterpret_cast<const float2*>(src + i); a = v.x; b = v.y; }
            xe[q] = a; xo[q] = b;
        }
        __syncthreads();
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        fir_phase<191, 0>(xe, threadIdx.x, acc);
        fir_phase<190, 1>(xo, threadIdx.x, acc);
        const int m = mb + 4 * threadIdx.x;
#pragma unroll
        for (int r = 0; r < 4; ++r) if (m + r < len_out) dst[m + r] = acc[r];
        __syncthreads();
    }
}

template <int MODE, int OUTS>   // MODE 1: constant-bank taps; 2: shared-memory taps
__global__ void __launch_bounds__(1024 / OUTS) dec_p(const float* src, float* dst, int len_in) {
    constexpr int T = 1024 / OUTS;
    __shared__ __align__(16) float2 xs[kSpan];
    __shared__ __align__(16) float2 ts[192 + 8];
    const int len_out = (len_in + 1) >> 1;
    if (MODE == 2) for (int q = threadIdx.x; q < 200; q += T) ts[q] = q < 192 ? c_tap2[q] : make_float2(0.f, 0.f);
    const unsigned long long* tapc = reinterpret_cast<const unsigned long long*>(c_tap2);
    for (int mb = blockIdx.x * kTile; mb < len_out; mb += gridDim.x * kTile) {
        for (int q = threadIdx.x; q < kSpan; q += T) {
            const int i = 2 * (mb - kHalo + q);
            float2 v = make_float2(0.0f, 0.0f);
            if (i >= 0 && i + 1 < len_in) v = *reinterpret_cast<const float2*>(src + i);
            xs[q] = v;
        }
        __syncthreads();
        unsigned long long acc[OUTS];
#pragma unroll
        for (int r = 0; r < OUTS; ++r) acc[r] = 0ull;
        const ulonglong2* row = reinterpret_cast<const ulonglong2*>(xs + OUTS * threadIdx.x);
        const ulonglong2* trow = reinterpret_cast<const ulonglong2*>(ts);
        // acc[r] += tap2[e] * xs[OUTS t + r + e], e = 1..191; position c = r + e walks 1 .. OUTS + 190
        constexpr int kGroups = (OUTS + 191 + 1) / 2;    // positions in pairs
        if (MODE == 1) {
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                const ulonglong2 v = row[g];
                const unsigned long long vv[2] = {v.x, v.y};
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int r = 0; r < OUTS; ++r) {
                        const int e = 2 * g + k - r;
                        if (e >= 1 && e < 192) acc[r] = fma2(tapc[e], vv[k], acc[r]);
                    }
            }
        } else {
            // taps outer (one LDS.128 broadcast per two tap pairs), samples in a rolling register
            // window w[j] = xs[OUTS t + 2 g + j], j = 0 .. OUTS + 1
            unsigned long long w[OUTS + 2];
#pragma unroll
            for (int j = 0; j < OUTS + 2; j += 2) { const ulonglong2 v = row[j / 2]; w[j] = v.x; w[j + 1] = v.y; }
#pragma unroll
            for (int g = 0; g < 96; ++g) {   // tap pairs e = 2g, 2g + 1
                const ulonglong2 tp = trow[g];
                if (g > 0) {
#pragma unroll
                    for (int r = 0; r < OUTS; ++r) acc[r] = fma2(tp.x, w[r], acc[r]);
                }
#pragma unroll
                for (int r = 0; r < OUTS; ++r) acc[r] = fma2(tp.y, w[r + 1], acc[r]);
                if (g < 95) {
#pragma unroll
                    for (int j = 0; j < OUTS; ++j) w[j] = w[j + 2];
                    const ulonglong2 v = row[g + 1 + OUTS / 2];
                    w[OUTS] = v.x; w[OUTS + 1] = v.y;
                }
            }
        }
        const int m = mb + OUTS * threadIdx.x;
#pragma unroll
        for (int r = 0; r < OUTS; ++r)
            if (m + r < len_out) {
                const float2 a = *reinterpret_cast<const float2*>(&acc[r]);
                dst[m + r] = a.x + a.y;
            }
        __syncthreads();
    }
}

// ---- padded layouts: thread t's window starts at OUTS t; a pad after every OUTS elements makes
// the 16-byte reads of eight neighbouring threads fall into distinct bank groups
__constant__ float c_tapf[2][208];   // c_tapf[phase][8 + e] = tap(e), zero outside 1..E_MAX

template <int OUTS>
__global__ void __launch_bounds__(1024 / OUTS) dec_s(const float* src, float* dst, int len_in) {
    constexpr int T = 1024 / OUTS;
    constexpr int PAD = OUTS == 4 ? 0 : 4;
    constexpr int PHYS = kSpan + PAD * (kSpan / OUTS);
    __shared__ __align__(16) float xe[PHYS];
    __shared__ __align__(16) float xo[PHYS];
    const int len_out = (len_in + 1) >> 1;
    for (int mb = blockIdx.x * kTile; mb < len_out; mb += gridDim.x * kTile) {
        for (int q = threadIdx.x; q < kSpan; q += T) {
            const int i = 2 * (mb - kHalo + q);
            float a = 0.0f, b = 0.0f;
            if (i >= 0 && i + 1 < len_in) { const float2 v = *reinterpret_cast<const float2*>(src + i); a = v.x; b = v.y; }
            const int ph = q + PAD * (q / OUTS);
            xe[ph] = a; xo[ph] = b;
        }
        __syncthreads();
        float acc[OUTS];
#pragma unroll
        for (int r = 0; r < OUTS; ++r) acc[r] = 0.0f;
#pragma unroll
        for (int phase = 0; phase < 2; ++phase) {
            const float* row = (phase == 0 ? xe : xo) + (OUTS + PAD) * threadIdx.x;
            const int e_max = phase == 0 ? 191 : 190;
#pragma unroll
            for (int g = 0; g <= (191 + OUTS - 1) / 4; ++g) {     // positions 4g .. 4g + 3
                const float4 q = *reinterpret_cast<const float4*>(row + 4 * g + PAD * ((4 * g) / OUTS));
                const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int r = 0; r < OUTS; ++r) {
                        const int e = 4 * g + c - r;
                        if (e >= 1 && e <= e_max) acc[r] = fmaf(c_tapf[phase][8 + e], v[c], acc[r]);
                    }
            }
        }
        const int m = mb + OUTS * threadIdx.x;
#pragma unroll
        for (int r = 0; r < OUTS; ++r) if (m + r < len_out) dst[m + r] = acc[r];
        __syncthreads();
    }
}

template <int OUTS>
__global__ void __launch_bounds__(1024 / OUTS) dec_pp(const float* src, float* dst, int len_in) {
    constexpr int T = 1024 / OUTS;
    constexpr int PAD = 2;
    constexpr int PHYS = kSpan + PAD * (kSpan / OUTS);
    __shared__ __align__(16) float2 xs[PHYS];
    const int len_out = (len_in + 1) >> 1;
    const unsigned long long* tapc = reinterpret_cast<const unsigned long long*>(c_tap2);
    for (int mb = blockIdx.x * kTile; mb < len_out; mb += gridDim.x * kTile) {
        for (int q = threadIdx.x; q < kSpan; q += T) {
            const int i = 2 * (mb - kHalo + q);
            float2 v = make_float2(0.0f, 0.0f);
            if (i >= 0 && i + 1 < len_in) v = *reinterpret_cast<const float2*>(src + i);
            xs[q + PAD * (q / OUTS)] = v;
        }
        __syncthreads();
        unsigned long long acc[OUTS];
#pragma unroll
        for (int r = 0; r < OUTS; ++r) acc[r] = 0ull;
        const float2* row = xs + (OUTS + PAD) * threadIdx.x;
#pragma unroll
        for (int g = 0; g < (OUTS + 192) / 2; ++g) {     // positions 2g, 2g + 1
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row + 2 * g + PAD * ((2 * g) / OUTS));
            const unsigned long long vv[2] = {v.x, v.y};
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int r = 0; r < OUTS; ++r) {
                    const int e = 2 * g + k - r;
                    if (e >= 1 && e < 192) acc[r] = fma2(tapc[e], vv[k], acc[r]);
                }
        }
        const int m = mb + OUTS * threadIdx.x;
#pragma unroll
        for (int r = 0; r < OUTS; ++r)
            if (m + r < len_out) {
                const float2 a = *reinterpret_cast<const float2*>(&acc[r]);
                dst[m + r] = a.x + a.y;
            }
        __syncthreads();
    }
}

template <typename F>
float time_it(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 10;
}

int main() {
    const int len_in = 2 * 1024 * 148 * 64;     // 19.4 M samples
    const int len_out = len_in / 2;
    std::vector<float> h(381), x(len_in);
    for (int k = 0; k < 381; ++k) {
        const double t = (k - 190) * 0.5 * M_PI;
        h[k] = static_cast<float>((k == 190 ? 0.5 : std::sin(t * 0.95) / (2 * t)) * (0.5 + 0.5 * std::cos(M_PI * (k - 190) / 191.0)));
    }
    srand(1);
    for (auto& v : x) v = rand() / static_cast<float>(RAND_MAX) - 0.5f;
    static float quads[2][50][4];
    for (int phase = 0; phase < 2; ++phase)
        for (int j = 0; j < 50; ++j)
            for (int k = 0; k < 4; ++k) {
                const int e = 4 * j - 3 + k, e_max = phase == 0 ? 191 : 190, idx = (phase == 0 ? 382 : 381) - 2 * e;
                quads[phase][j][k] = (e >= 1 && e <= e_max) ? h[idx] : 0.0f;
            }
    static float pairs[192][2];
    for (int e = 0; e < 192; ++e) {
        pairs[e][0] = e >= 1 ? h[382 - 2 * e] : 0.0f;
        pairs[e][1] = (e >= 1 && e <= 190) ? h[381 - 2 * e] : 0.0f;
    }
    cudaMemcpyToSymbol(c_tap4, quads, sizeof(quads));
    cudaMemcpyToSymbol(c_tap2, pairs, sizeof(pairs));
    static float flat[2][208];
    for (int e = 1; e <= 191; ++e) flat[0][8 + e] = h[382 - 2 * e];
    for (int e = 1; e <= 190; ++e) flat[1][8 + e] = h[381 - 2 * e];
    cudaMemcpyToSymbol(c_tapf, flat, sizeof(flat));
    float *src, *da, *db;
    cudaMalloc(&src, len_in * sizeof(float));
    cudaMalloc(&da, len_out * sizeof(float));
    cudaMalloc(&db, len_out * sizeof(float));
    cudaMemcpy(src, x.data(), len_in * sizeof(float), cudaMemcpyHostToDevice);
    const int grid = len_out / kTile;
    std::vector<float> ra(len_out), rb(len_out);
    auto check = [&](const char* name, float ms) {
        cudaMemcpy(rb.data(), db, len_out * sizeof(float), cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int i = 0; i < len_out; ++i) worst = std::max(worst, static_cast<double>(std::fabs(ra[i] - rb[i])));
        const double flops = 2.0 * 381 * len_out;
        printf("%-28s %.3f ms  %.1f TFLOP/s  max |diff to A| %.3g  (%s)\n", name, ms, flops / ms / 1e9, worst, cudaGetErrorString(cudaGetLastError()));
    };
    float ms = time_it([&] { dec_a<<<grid, 256>>>(src, da, len_in); });
    cudaMemcpy(ra.data(), da, len_out * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(db, da, len_out * sizeof(float), cudaMemcpyDeviceToDevice);
    check("A scalar FFMA const", ms);
    ms = time_it([&] { dec_p<1, 4><<<grid, 256>>>(src, db, len_in); });
    check("B FFMA2 const/UR 4 outs", ms);
    ms = time_it([&] { dec_p<1, 8><<<grid, 128>>>(src, db, len_in); });
    check("B8 FFMA2 const/UR 8 outs", ms);
    ms = time_it([&] { dec_p<2, 4><<<grid, 256>>>(src, db, len_in); });
    check("C FFMA2 smem taps 4 outs", ms);
    ms = time_it([&] { dec_p<2, 8><<<grid, 128>>>(src, db, len_in); });
    check("D FFMA2 smem taps 8 outs", ms);
    ms = time_it([&] { dec_s<4><<<grid, 256>>>(src, db, len_in); });
    check("S4 scalar flat taps 4 outs", ms);
    ms = time_it([&] { dec_s<8><<<grid, 128>>>(src, db, len_in); });
    check("S8 scalar padded 8 outs", ms);
    ms = time_it([&] { dec_s<16><<<grid, 64>>>(src, db, len_in); });
    check("S16 scalar padded 16 outs", ms);
    ms = time_it([&] { dec_pp<4><<<grid, 256>>>(src, db, len_in); });
    check("P4 FFMA2 padded 4 outs", ms);
    ms = time_it([&] { dec_pp<8><<<grid, 128>>>(src, db, len_in); });
    check("P8 FFMA2 padded 8 outs", ms);
    ms = time_it([&] { dec_pp<16><<<grid, 64>>>(src, db, len_in); });
    check("P16 FFMA2 padded 16 outs", ms);
    return 0;
}
