// K1: framing + periodic-Hann window + 2048-point real FFT -> |X|, plus piptrack peaks.
//
// Replaces, for a ragged batch of clips, what the reference computes per clip with
//   np.abs(librosa.stft(y, n_fft=2048))                       ser/_internal/utils/dsp.py:100
//   librosa.piptrack (inside chroma_stft -> estimate_tuning)  ser/_internal/utils/dsp.py:113-118
// (librosa 0.11.0 semantics: SURVEY.md Appendix A.1, A.6).
//
// Layout: one CTA per 8 consecutive STFT columns of one clip (half a projection tile), two
// CTAs resident per SM.  The CTA's 5632 samples are staged once in shared memory by a TMA bulk copy (cp.async.bulk ->
// UBLKCP) signalling an mbarrier; zero padding of the centred STFT is a shared-memory
// fill.  Each warp then owns whole columns: a 2048-point real FFT is done as a 1024-point
// complex FFT split 32 x 32 across the 32 lanes -- two register-resident 32-point DFTs
// around one shared-memory transpose -- followed by the real-input split, which pairs
// lane l with lane 32-l through warp shuffles.  |X| goes to the L2-resident spill
// (row-major [column][1032]); peaks are compacted per column with a warp ballot.
#include "fft.cuh"
#include "kernels.h"

namespace serb {

constexpr int kStftWarps = 8;
constexpr int kStftThreads = kStftWarps * 32;
constexpr int kStftCols = 8;                                    // columns per CTA: one per warp
constexpr int kStftSamples = (kStftCols - 1) * kHop + kNFft;    // 5632 staged samples
constexpr int kBufPitch = 33;  // float2 pitch of the per-warp 32x32 transpose buffer


struct StftSmem {
    float wave[kStftSamples];                       // 22528 B
    float2 tw[32][32];                              // W_1024^(k1*n2): [k1][n2], 8192 B
    float2 buf[kStftWarps][32 * kBufPitch];         // per-warp transpose / |X| staging
    unsigned long long bar;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- staging: TMA bulk copy of the valid part, zero fill of the rest -------------------
__device__ __forceinline__ void stage_tile(StftSmem& sm, const float* __restrict__ wave,
                                           long long clip_start, int clip_len, int s0) {
    // shared index i <-> clip sample s0 + i ; valid when 0 <= s0 + i < clip_len
    const int lo = max(0, -s0);
    const int hi = min(kStftSamples, clip_len - s0);
    const float* src = wave + clip_start + s0;  // src[i] is the sample for shared index i
    const int tid = threadIdx.x;
    const bool aligned = (((clip_start + s0 + lo) & 3LL) == 0) && ((lo & 3) == 0);
    int bulk_end = lo;
    if (aligned && hi - lo >= 4) bulk_end = lo + ((hi - lo) & ~3);
    const uint32_t bytes = static_cast<uint32_t>(bulk_end - lo) * 4u;
    const uint32_t bar = smem_u32(&sm.bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0 && bytes > 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(&sm.wave[lo])), "l"(src + lo), "r"(bytes), "r"(bar)
            : "memory");
    }
    // everything the bulk copy does not cover: zero padding and the unaligned remainder
    for (int i = tid; i < lo; i += kStftThreads) sm.wave[i] = 0.0f;
    for (int i = bulk_end + tid; i < hi; i += kStftThreads) sm.wave[i] = __ldg(src + i);
    for (int i = max(hi, 0) + tid; i < kStftSamples; i += kStftThreads) sm.wave[i] = 0.0f;
    if (bytes > 0) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar)
                : "memory");
        }
    }
    __syncthreads();
}

// ---- one STFT column per warp -----------------------------------------------------------
// Writes |X[k]|, k = 0..1024, to row[] (global spill) and to the warp's staging buffer (as
// floats, aliasing buf, which is free once the transpose has been read back); returns the
// lane's running maximum of |X|.
__device__ __forceinline__ float column_fft(const float* __restrict__ frame, const float2 (*tw)[32],
                                            float2* __restrict__ buf, int lane,
                                            float wc0, float ws0, float wc1, float ws1,
                                            float tc, float ts, float* __restrict__ row) {
    float2 v[32];
    // load z[n] = x[2n] + i x[2n+1], n = 32 n1 + lane, times the periodic Hann window
    // w[j] = 0.5 - 0.5 cos(2 pi j / 2048), j = 64 n1 + 2 lane (+1):
    // cos(a + t) = cos a cos t - sin a sin t with a = 2 pi n1 / 32 (immediates), t per lane.
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
        const float2 x = *reinterpret_cast<const float2*>(frame + 64 * n1 + 2 * lane);
        const float ca = cos32(n1), sa = sin32(n1);
        const float w0 = fmaf(-0.5f * ca, wc0, fmaf(0.5f * sa, ws0, 0.5f));
        const float w1 = fmaf(-0.5f * ca, wc1, fmaf(0.5f * sa, ws1, 0.5f));
        v[n1] = make_float2(x.x * w0, x.y * w1);
    }
    fft32(v);  // over n1 -> index k1
    // twiddle W_1024^(lane * k1), transpose through shared memory
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
        float2 y = v[k1];
        if (k1 > 0) {
            const float2 w = tw[k1][lane];
            y = make_float2(fmaf(y.x, w.x, -y.y * w.y), fmaf(y.x, w.y, y.y * w.x));
        }
        buf[k1 * kBufPitch + lane] = y;
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = buf[lane * kBufPitch + n2];
    __syncwarp();
    fft32(v);  // over n2 -> Z[lane + 32 k2] in v[k2]

    // real-input split: X[k] = E + W_2048^k O, E = (Z[k] + conj Z[M-k]) / 2, O = -i (Z[k] - conj Z[M-k]) / 2
    const int src = (32 - lane) & 31;
    float* sbuf = reinterpret_cast<float*>(buf);
    float cmax = 0.0f;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        float px = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
        float py = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
        if (lane == 0) { px = v[(32 - k2) & 31].x; py = v[(32 - k2) & 31].y; }
        const float ax = v[k2].x, ay = v[k2].y;
        const float ex = 0.5f * (ax + px), ey = 0.5f * (ay - py);
        const float dx = 0.5f * (ax - px), dy = 0.5f * (ay + py);
        const float ox = dy, oy = -dx;
        // W_2048^(lane + 32 k2): angle = t(lane) + 2 pi k2 / 64
        const float c2 = cos64(k2), s2 = sin64(k2);
        const float c = fmaf(tc, c2, -ts * s2);
        const float s = fmaf(ts, c2, tc * s2);
        const float wx = fmaf(c, ox, s * oy);
        const float wy = fmaf(c, oy, -s * ox);
        const float xr = ex + wx, xi = ey + wy;
        const float mag = sqrtf(fmaf(xr, xr, xi * xi));
        row[lane + 32 * k2] = mag;
        sbuf[lane + 32 * k2] = mag;
        cmax = fmaxf(cmax, mag);
        if (k2 == 0 && lane == 0) {
            const float nr = ex - wx, ni = ey - wy;
            const float nyq = sqrtf(fmaf(nr, nr, ni * ni));
            row[1024] = nyq;
            sbuf[1024] = nyq;
            cmax = fmaxf(cmax, nyq);
        }
    }
    return cmax;
}

// ---- piptrack on one column (librosa.piptrack) -------------------
__device__ __forceinline__ void column_peaks(const float* __restrict__ sbuf, float colmax, int lane,
                                             const StftParams& p, long long col) {
    // S * (S > 0.1 * colmax): float32 product, as numpy computes it
    const float ref = __fmul_rn(0.1f, colmax);
    float2* out = p.peaks + col * p.peak_cap;
    int n_found = 0;
    for (int base = p.kmin; base < p.kmax; base += 32) {
        const int k = base + lane;
        bool is_peak = false;
        float sm1 = 0.f, s0 = 0.f, sp1 = 0.f;
        if (k < p.kmax && k >= 1 && k <= kNBins - 2) {
            sm1 = sbuf[k - 1]; s0 = sbuf[k]; sp1 = sbuf[k + 1];
            const float tm1 = sm1 > ref ? sm1 : 0.f;
            const float t0 = s0 > ref ? s0 : 0.f;
            const float tp1 = sp1 > ref ? sp1 : 0.f;
            is_peak = (t0 > tm1) && (t0 >= tp1);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, is_peak);
        if (is_peak) {
            // parabolic interpolation with the reference's mixed precision (numba stencil):
            // a = (x[+1] + x[-1])_f32 - 2 x[0] in f64, b = (x[+1] - x[-1])_f32 / 2 in f64
            const double a = static_cast<double>(__fadd_rn(sp1, sm1)) - 2.0 * static_cast<double>(s0);
            const double b = static_cast<double>(__fsub_rn(sp1, sm1)) / 2.0;
            float shift = 0.f;
            if (fabs(b) < fabs(a)) shift = static_cast<float>(-b / a);
            // np.gradient interior: (x[+1] - x[-1]) / 2 in float32; dskew = 0.5 * avg * shift
            const float avg = __fmul_rn(__fsub_rn(sp1, sm1), 0.5f);
            const float dskew = __fmul_rn(__fmul_rn(0.5f, avg), shift);
            const float mag = __fadd_rn(s0, dskew);
            // pitches = ((k + shift) * float(sr)) / n_fft in float64, stored as float32
            const double pitch64 = ((static_cast<double>(k) + static_cast<double>(shift)) * p.sr_over_nfft_num) /
                                   static_cast<double>(kNFft);
            const int slot = n_found + __popc(ballot & ((1u << lane) - 1u));
            if (slot < p.peak_cap) out[slot] = make_float2(mag, static_cast<float>(pitch64));
        }
        n_found += __popc(ballot);
    }
    if (lane == 0) p.peak_count[col] = min(n_found, p.peak_cap);
}

__global__ void __launch_bounds__(kStftThreads, 2) stft_kernel(StftParams p, int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StftSmem& sm = *reinterpret_cast<StftSmem*>(smem_raw);
    // blockIdx.x enumerates half tiles: two 8-column CTAs per 16-column projection tile
    const int tile = blockIdx.x >> 1;
    if (tile >= n_tiles) return;
    const int ci = find_clip_by_tile(p.clips, p.n_clips, tile);
    const ClipDev clip = p.clips[ci];
    const int t0 = (tile - clip.tile_base) * kColsPerTile + (blockIdx.x & 1) * kStftCols;
    if (t0 >= clip.n_cols) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // W_1024^(k1 * n2) table (accurate sincospi, once per CTA)
    for (int i = threadIdx.x; i < 1024; i += kStftThreads) {
        const int k1 = i >> 5, n2 = i & 31;
        float s, c;
        sincospif(static_cast<float>((k1 * n2) & 1023) * (2.0f / 1024.0f), &s, &c);
        sm.tw[k1][n2] = make_float2(c, -s);
    }
    stage_tile(sm, p.wave, clip.start, clip.length, t0 * kHop - kNFft / 2);
    {   // dsp.py:94 "Audio buffer is not finite everywhere." -> status bit 0, reported by the host entry
        int bad = 0;
        for (int i = threadIdx.x; i < kStftSamples; i += kStftThreads) bad |= !isfinite(sm.wave[i]);
        if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(p.status, 1);
    }

    // per-lane window / split angles
    float ws0, wc0, ws1, wc1, ts, tc;
    sincospif(static_cast<float>(2 * lane) * (2.0f / 2048.0f), &ws0, &wc0);
    sincospif(static_cast<float>(2 * lane + 1) * (2.0f / 2048.0f), &ws1, &wc1);
    sincospif(static_cast<float>(lane) * (2.0f / 2048.0f), &ts, &tc);

    float2* buf = sm.buf[warp];
    const int t = t0 + warp;
    if (t < clip.n_cols) {
        const long long col = static_cast<long long>(clip.col_base) + t;
        float cmax = column_fft(sm.wave + warp * kHop, sm.tw, buf, lane, wc0, ws0, wc1, ws1, tc, ts,
                                p.spill + col * kSpillStride);
        if (p.do_peaks) {
            cmax = warp_max(cmax);
            __syncwarp();
            column_peaks(reinterpret_cast<const float*>(buf), cmax, lane, p, col);
        }
    }
}

cudaError_t configure_stft() {
    return cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(sizeof(StftSmem)));
}

cudaError_t launch_stft(const StftParams& p, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0) return cudaSuccess;
    stft_kernel<<<2 * n_tiles, kStftThreads, sizeof(StftSmem), stream>>>(p, n_tiles);
    return cudaGetLastError();
}

}  // namespace serb
