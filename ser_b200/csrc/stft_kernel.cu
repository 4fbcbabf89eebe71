// K1: framing + periodic-Hann window + 2048-point real FFT -> |X|, plus piptrack peaks.
//
// Replaces, for a ragged batch of clips, what the reference computes per clip with
//   np.abs(librosa.stft(y, n_fft=2048))                       ser/_internal/utils/dsp.py:100
//   librosa.piptrack (inside chroma_stft -> estimate_tuning)  ser/_internal/utils/dsp.py:113-118
// (librosa 0.11.0 semantics: SURVEY.md Appendix A.1, A.6).
//
// Layout: one CTA per 8 consecutive STFT columns of one clip (half a projection tile), two
// CTAs resident per SM.  The CTA's 5632 samples and the two twiddle tables are staged in
// shared memory by TMA bulk copies (cp.async.bulk -> UBLKCP) signalling one mbarrier; the
// zero padding of the centred STFT is a shared-memory fill.  Each warp owns one column: the
// 2048-point real FFT is a 1024-point complex FFT split 32 x 32 across the 32 lanes -- two
// register-resident 32-point DFTs around one shared-memory transpose -- followed by the
// real-input split, which pairs lane l with lane 32-l through warp shuffles.  The CTA then
// writes its |X| block to the spill as [bin][8 columns] rows (32 bytes per thread, fully
// coalesced), the layout the projection kernel register-tiles over; peaks are compacted per
// column with a warp ballot.
#include "fft.cuh"
#include "kernels.h"

namespace serb {

constexpr int kStftWarps = 8;
constexpr int kStftThreads = kStftWarps * 32;
constexpr int kStftCols = kHalfTileCols;                        // columns per CTA: one per warp
constexpr int kStftSamples = (kStftCols - 1) * kHop + kNFft;    // 5632 staged samples
constexpr int kBufPitch = 33;  // float2 pitch of the per-warp 32x32 transpose buffer

struct StftSmem {
    float wave[kStftSamples];                       // 22528 B
    float2 tw[32][32];                              // W_1024^(k1*n2): [k1][n2]
    float2 tw2[1024];                               // W_2048^k, k < 1024 (real-input split)
    float2 buf[kStftWarps][32 * kBufPitch];         // per-warp transpose buffer, then |X| staging
    unsigned long long bar;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- staging: TMA bulk copies of the tables and the valid samples, zero fill of the rest ----
__device__ __forceinline__ void stage_tile(StftSmem& sm, const float* __restrict__ wave,
                                           const float2* __restrict__ tables,
                                           long long clip_start, int clip_len, int s0) {
    // shared index i <-> clip sample s0 + i ; valid when 0 <= s0 + i < clip_len
    const int lo = max(0, -s0);
    const int hi = min(kStftSamples, clip_len - s0);
    const float* src = wave + clip_start + s0;  // src[i] is the sample for shared index i
    const int tid = threadIdx.x;
    const bool aligned = (((clip_start + s0 + lo) & 3LL) == 0) && ((lo & 3) == 0);
    int bulk_end = lo;
    if (aligned && hi - lo >= 4) bulk_end = lo + ((hi - lo) & ~3);
    const uint32_t wave_bytes = static_cast<uint32_t>(bulk_end - lo) * 4u;
    constexpr uint32_t table_bytes = 2 * 1024 * sizeof(float2);
    const uint32_t bar = smem_u32(&sm.bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(wave_bytes + table_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(&sm.tw[0][0])), "l"(tables), "r"(table_bytes), "r"(bar) : "memory");
        if (wave_bytes > 0)
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(smem_u32(&sm.wave[lo])), "l"(src + lo), "r"(wave_bytes), "r"(bar) : "memory");
    }
    // everything the bulk copy does not cover: zero padding and the unaligned remainder
    for (int i = tid; i < lo; i += kStftThreads) sm.wave[i] = 0.0f;
    for (int i = bulk_end + tid; i < hi; i += kStftThreads) sm.wave[i] = __ldg(src + i);
    for (int i = max(hi, 0) + tid; i < kStftSamples; i += kStftThreads) sm.wave[i] = 0.0f;
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar) : "memory");
    }
    __syncthreads();
}

// ---- one STFT column per warp -----------------------------------------------------------
// Leaves |X[k]|, k = 0..1024, in the warp's staging buffer (floats aliasing buf, free once the
// transpose has been read back) and returns the lane's running maximum of |X|.
__device__ __forceinline__ float column_fft(const float* __restrict__ frame, const float2 (*tw)[32],
                                            const float2* __restrict__ tw2, float2* __restrict__ buf,
                                            int lane, float wc0, float ws0, float wc1, float ws1) {
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

    // real-input split, scaled by 2:  2 X[k] = (Z[k] + conj Z[M-k]) + W_2048^k (-i)(Z[k] - conj Z[M-k])
    const int src = (32 - lane) & 31;
    float* sbuf = reinterpret_cast<float*>(buf);
    float cmax = 0.0f;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        float px = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
        float py = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
        if (lane == 0) { px = v[(32 - k2) & 31].x; py = v[(32 - k2) & 31].y; }
        const float ax = v[k2].x, ay = v[k2].y;
        const float ex = ax + px, ey = ay - py;
        const float ox = ay + py, oy = px - ax;       // -i (A - conj P)
        const float2 w = tw2[lane + 32 * k2];          // (cos, sin) of 2 pi k / 2048
        const float wx = fmaf(w.x, ox, w.y * oy);
        const float wy = fmaf(w.x, oy, -w.y * ox);
        const float xr = ex + wx, xi = ey + wy;
        const float mag = 0.5f * sqrt_approx(fmaf(xr, xr, xi * xi));
        sbuf[lane + 32 * k2] = mag;
        cmax = fmaxf(cmax, mag);
        if (k2 == 0 && lane == 0) {
            const float nr = ex - wx, ni = ey - wy;
            const float nyq = 0.5f * sqrt_approx(fmaf(nr, nr, ni * ni));
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
    const int ci = p.tile_clip[tile];
    const ClipDev clip = p.clips[ci];
    const int t0 = (tile - clip.tile_base) * kColsPerTile + (blockIdx.x & 1) * kStftCols;
    float4* block = reinterpret_cast<float4*>(p.spill + static_cast<long long>(blockIdx.x) * kHalfTileFloats);
    if (t0 >= clip.n_cols) {
        // the projection kernel bulk-loads both halves of a tile: keep the unused half defined
        for (int i = threadIdx.x; i < kHalfTileFloats / 4; i += kStftThreads) block[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    stage_tile(sm, p.wave, p.tables, clip.start, clip.length, t0 * kHop - kNFft / 2);
    {   // dsp.py:94 "Audio buffer is not finite everywhere." -> status bit 0, reported by the host entry
        int bad = 0;
        for (int i = threadIdx.x; i < kStftSamples; i += kStftThreads) bad |= !isfinite(sm.wave[i]);
        if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(p.status, 1);
    }

    // per-lane window angles
    float ws0, wc0, ws1, wc1;
    sincospif(static_cast<float>(2 * lane) * (2.0f / 2048.0f), &ws0, &wc0);
    sincospif(static_cast<float>(2 * lane + 1) * (2.0f / 2048.0f), &ws1, &wc1);

    float2* buf = sm.buf[warp];
    float* sbuf = reinterpret_cast<float*>(buf);
    const int t = t0 + warp;
    if (t < clip.n_cols) {
        float cmax = column_fft(sm.wave + warp * kHop, sm.tw, sm.tw2, buf, lane, wc0, ws0, wc1, ws1);
        if (p.do_peaks) {
            cmax = warp_max(cmax);
            __syncwarp();
            column_peaks(sbuf, cmax, lane, p, static_cast<long long>(clip.col_base) + t);
        }
    } else {
        for (int i = lane; i < kNBins; i += 32) sbuf[i] = 0.0f;
    }
    __syncthreads();
    // spill block [bin][8 columns]: each thread gathers one bin from the 8 warp buffers
    // (conflict-free: consecutive lanes, consecutive bins) and stores 32 contiguous bytes
    for (int k = threadIdx.x; k < kNBins; k += kStftThreads) {
        float c[kStftCols];
#pragma unroll
        for (int w = 0; w < kStftCols; ++w) c[w] = reinterpret_cast<const float*>(sm.buf[w])[k];
        block[2 * k] = make_float4(c[0], c[1], c[2], c[3]);
        block[2 * k + 1] = make_float4(c[4], c[5], c[6], c[7]);
    }
}

// one thread per clip: tile_clip[tile] = clip index, for every 16-column tile of the clip
__global__ void expand_tiles_kernel(const ClipDev* __restrict__ clips, int n_clips, int* __restrict__ tile_clip) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clips) return;
    const int base = clips[c].tile_base;
    const int n = (clips[c].n_cols + kColsPerTile - 1) / kColsPerTile;
    for (int i = 0; i < n; ++i) tile_clip[base + i] = c;
}

cudaError_t launch_expand_tiles(const ClipDev* clips, int n_clips, int* tile_clip, cudaStream_t stream) {
    if (n_clips <= 0) return cudaSuccess;
    expand_tiles_kernel<<<(n_clips + 127) / 128, 128, 0, stream>>>(clips, n_clips, tile_clip);
    return cudaGetLastError();
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
