// Register-resident FFT building blocks (fp32, forward transform, e^(-i...)).
// All array indices are compile-time after unrolling, so every float2 v[32] lives in
// registers; twiddles of the 32-point transform fold into FFMA immediates.
#pragma once
#include "common.cuh"

namespace serb {

__host__ __device__ constexpr float cos32(int j) {
    constexpr float t[32] = {
        1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
        0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
        0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
        -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f,
        -1.0f, -0.98078528040323043f, -0.92387953251128685f, -0.83146961230254546f,
        -0.70710678118654768f, -0.55557023301960218f, -0.38268343236509034f, -0.19509032201612866f,
        0.0f, 0.1950903220161283f, 0.38268343236509f, 0.55557023301960184f,
        0.70710678118654735f, 0.83146961230254524f, 0.92387953251128652f, 0.98078528040323032f,
    };
    return t[j & 31];
}
__host__ __device__ constexpr float sin32(int j) {
    constexpr float t[32] = {
        0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
        0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
        1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
        0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f,
        0.0f, -0.19509032201612836f, -0.38268343236508967f, -0.55557023301960196f,
        -0.70710678118654746f, -0.83146961230254524f, -0.92387953251128652f, -0.98078528040323032f,
        -1.0f, -0.98078528040323043f, -0.92387953251128663f, -0.83146961230254546f,
        -0.70710678118654768f, -0.55557023301960218f, -0.38268343236509039f, -0.19509032201612872f,
    };
    return t[j & 31];
}
// cos/sin(2 pi j / 64), j = 0..31
__host__ __device__ constexpr float cos64(int j) {
    constexpr float t[32] = {
        1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220882f,
        0.92387953251128674f, 0.88192126434835505f, 0.83146961230254524f, 0.77301045336273699f,
        0.70710678118654757f, 0.63439328416364549f, 0.55557023301960229f, 0.47139673682599781f,
        0.38268343236508984f, 0.29028467725446233f, 0.19509032201612833f, 0.09801714032956077f,
        0.0f, -0.098017140329560645f, -0.19509032201612819f, -0.29028467725446216f,
        -0.38268343236508973f, -0.4713967368259977f, -0.55557023301960196f, -0.63439328416364538f,
        -0.70710678118654746f, -0.77301045336273699f, -0.83146961230254535f, -0.88192126434835494f,
        -0.92387953251128674f, -0.95694033573220882f, -0.98078528040323043f, -0.99518472667219682f,
    };
    return t[j & 31];
}
__host__ __device__ constexpr float sin64(int j) {
    constexpr float t[32] = {
        0.0f, 0.098017140329560604f, 0.19509032201612825f, 0.29028467725446233f,
        0.38268343236508978f, 0.47139673682599764f, 0.55557023301960218f, 0.63439328416364549f,
        0.70710678118654746f, 0.77301045336273699f, 0.83146961230254524f, 0.88192126434835494f,
        0.92387953251128674f, 0.95694033573220894f, 0.98078528040323043f, 0.99518472667219682f,
        1.0f, 0.99518472667219693f, 0.98078528040323043f, 0.95694033573220894f,
        0.92387953251128674f, 0.88192126434835505f, 0.83146961230254546f, 0.7730104533627371f,
        0.70710678118654757f, 0.63439328416364549f, 0.55557023301960218f, 0.47139673682599786f,
        0.38268343236508989f, 0.29028467725446239f, 0.19509032201612861f, 0.098017140329560826f,
    };
    return t[j & 31];
}

// complex add / subtract: one packed FADD2 on the device (IEEE per half, so bit-identical to the
// two scalar adds), which halves the issue slots of the butterflies
__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && !defined(SERB_SCALAR_BUTTERFLIES)
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && !defined(SERB_SCALAR_BUTTERFLIES)
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a * s on both halves: one packed FMUL2 on the device
__host__ __device__ __forceinline__ float2 cscale(float2 a, float s) {
#if defined(__CUDA_ARCH__) && !defined(SERB_SCALAR_BUTTERFLIES)
    unsigned long long d;
    const float2 ss = make_float2(s, s);
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&ss)));
    return *reinterpret_cast<float2*>(&d);
#else
    return make_float2(a.x * s, a.y * s);
#endif
}
// a * (c - i s)   (forward twiddle e^(-i theta), c = cos theta, s = sin theta)
__host__ __device__ __forceinline__ float2 cmul_conj_tw(float2 a, float c, float s) {
    return make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s));
}
// a * (-i)
__host__ __device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }

// 4-point forward DFT, natural order in and out
__host__ __device__ __forceinline__ void fft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 t0 = cadd(x0, x2), t1 = csub(x0, x2);
    const float2 t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1 = make_float2(t1.x + t3.y, t1.y - t3.x);   // t1 - i t3
    x3 = make_float2(t1.x - t3.y, t1.y + t3.x);   // t1 + i t3
}

// 8-point forward DFT, natural order in and out
__host__ __device__ __forceinline__ void fft8(float2& x0, float2& x1, float2& x2, float2& x3,
                                     float2& x4, float2& x5, float2& x6, float2& x7) {
    float2 e0 = x0, e1 = x2, e2 = x4, e3 = x6;
    float2 o0 = x1, o1 = x3, o2 = x5, o3 = x7;
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    constexpr float r = 0.70710678118654752f;
    // W8^1 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2
    const float2 w1 = make_float2((o1.x + o1.y) * r, (o1.y - o1.x) * r);
    const float2 w2 = mul_neg_i(o2);
    const float2 w3 = make_float2((o3.y - o3.x) * r, -(o3.x + o3.y) * r);
    x0 = cadd(e0, o0); x4 = csub(e0, o0);
    x1 = cadd(e1, w1); x5 = csub(e1, w1);
    x2 = cadd(e2, w2); x6 = csub(e2, w2);
    x3 = cadd(e3, w3); x7 = csub(e3, w3);
}

// 32-point forward DFT, natural order in and out: n = 4a + b, k = k2 + 8 k1.
__host__ __device__ __forceinline__ void fft32(float2 (&v)[32]) {
    // four 8-point DFTs over a (stride 4)
#pragma unroll
    for (int b = 0; b < 4; ++b)
        fft8(v[b], v[4 + b], v[8 + b], v[12 + b], v[16 + b], v[20 + b], v[24 + b], v[28 + b]);
    // v[4*k2 + b] now holds Y_b[k2]; twiddle by W32^(b k2)
#pragma unroll
    for (int k2 = 1; k2 < 8; ++k2) {
#pragma unroll
        for (int b = 1; b < 4; ++b)
            v[4 * k2 + b] = cmul_conj_tw(v[4 * k2 + b], cos32(b * k2), sin32(b * k2));
    }
    // eight 4-point DFTs over b; result X[k2 + 8 k1] lands in v[4*k2 + k1]
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) fft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    // un-permute to natural order (register renaming only)
    float2 t[32];
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2)
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) t[k2 + 8 * k1] = v[4 * k2 + k1];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = t[i];
}

}  // namespace serb
