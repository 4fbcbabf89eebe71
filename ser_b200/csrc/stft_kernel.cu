// K1: framing + periodic-Hann window + 2048-point real FFT -> |X|, plus piptrack peaks.
//
// Replaces, for a ragged batch of clips, what the reference computes per clip with
//   np.abs(librosa.stft(y, n_fft=2048))                       ser/_internal/utils/dsp.py:100
//   librosa.piptrack (inside chroma_stft -> estimate_tuning)  ser/_internal/utils/dsp.py:113-118
// (librosa 0.11.0 semantics: SURVEY.md Appendix A.1, A.6).
//
// Persistent kernel: two CTAs of 8 warps per SM walk the list of half tiles (8 consecutive
// STFT columns of one clip).  A half tile's 5632 samples are staged in shared memory by a TMA
// bulk copy (cp.async.bulk -> UBLKCP) signalling an mbarrier; the copy for the NEXT half tile
// is issued as soon as every warp has pulled its frame into registers, so it lands while the
// FFTs run.  Zero padding of the centred STFT is a shared-memory fill.  Each warp owns one
// column: the 2048-point real FFT is a 1024-point complex FFT split 32 x 32 across the 32
// lanes -- two register-resident 32-point DFTs around one shared-memory transpose -- followed
// by the real-input split, which pairs lane l with lane 32-l through warp shuffles.  |X| goes
// to the spill (row-major [column][1032], coalesced 128-byte stores); peaks are compacted per
// column with a warp ballot.
#include "fft.cuh"
#include "kernels.h"

namespace serb {

constexpr int kStftWarps = 8;
constexpr int kStftThreads = kStftWarps * 32;
constexpr int kStftCols = kHalfTileCols;                        // columns per work item: one per warp
constexpr int kStftSamples = (kStftCols - 1) * kHop + kNFft;    // 5632 staged samples
constexpr int kBufPitch = 33;  // float2 pitch of the per-warp 32x32 transpose buffer

// What one work item needs; produced one iteration ahead by thread 0.
struct ItemDesc {
    const float* src;    // src[i] is the sample for shared index i
    int lo, hi;          // valid shared indices [lo, hi)
    int bulk_end;        // [lo, bulk_end) arrives by TMA, [bulk_end, hi) by plain loads
    int n_here;          // columns of this half tile that exist (0 = nothing to do)
    long long col0;      // spill / peak column of the first one
};

struct StftSmem {
    float wave[kStftSamples];                       // 22528 B
    float2 tw[32][32];                              // W_1024^(k1*n2): [k1][n2]
    float2 tw2[1024];                               // W_2048^k, k < 1024 (real-input split)
    float2 buf[kStftWarps][32 * kBufPitch];         // per-warp transpose buffer, then |X| staging
    ItemDesc desc[2];
    ItemDesc next;                                  // thread 0's look-ahead descriptor
    float4 win[32];                                 // per-lane window angles (cos, sin) x 2
    int bad;
    unsigned long long bar_wave, bar_tables;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ ItemDesc make_desc(const StftParams& p, int item, int n_items) {
    ItemDesc d;
    d.src = nullptr; d.lo = 0; d.hi = 0; d.bulk_end = 0; d.n_here = 0; d.col0 = 0;
    if (item >= n_items) return d;
    const int tile = item >> 1;
    const ClipDev clip = p.clips[p.tile_clip[tile]];
    const int t0 = (tile - clip.tile_base) * kColsPerTile + (item & 1) * kStftCols;
    if (t0 >= clip.n_cols) return d;
    const int s0 = t0 * kHop - kNFft / 2;       // shared index i <-> clip sample s0 + i
    d.n_here = min(kStftCols, clip.n_cols - t0);
    d.col0 = static_cast<long long>(clip.col_base) + t0;
    d.lo = max(0, -s0);
    d.hi = min(kStftSamples, clip.length - s0);
    d.src = p.wave + clip.start + s0;
    const bool aligned = (((clip.start + s0 + d.lo) & 3LL) == 0) && ((d.lo & 3) == 0);
    d.bulk_end = d.lo;
    if (aligned && d.hi - d.lo >= 4) d.bulk_end = d.lo + ((d.hi - d.lo) & ~3);
    return d;
}

// thread 0: arm the barrier and start the bulk copy of the item's aligned part (possibly empty)
__device__ __forceinline__ void issue_wave_copy(StftSmem& sm, const ItemDesc& d) {
    const uint32_t bar = smem_u32(&sm.bar_wave);
    const uint32_t bytes = static_cast<uint32_t>(d.bulk_end - d.lo) * 4u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    if (bytes > 0)
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(&sm.wave[d.lo])), "l"(d.src + d.lo), "r"(bytes), "r"(bar) : "memory");
}

// ---- one STFT column per warp -----------------------------------------------------------
// v[] holds the windowed frame z[32 n1 + lane].  Writes |X[k]|, k = 0..1024, to row[] (global
// spill) and to the warp's staging buffer; returns the lane's running maximum of |X|.
__device__ __forceinline__ float column_fft(float2 (&v)[32], const float2 (*tw)[32],
                                            const float2* __restrict__ tw2, float2* __restrict__ buf,
                                            int lane, float* __restrict__ row, float2* __restrict__ crow) {
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

    // real-input split:  2 X[k] = (Z[k] + conj Z[M-k]) + W_2048^k (-i)(Z[k] - conj Z[M-k]); the frame came in
    // halved (see the window), so the right-hand side is X[k] itself
    const int src = (32 - lane) & 31;
    float* sbuf = reinterpret_cast<float*>(buf);
    float cmax = 0.0f;
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) {
        float px = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
        float py = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
        if (lane == 0) { px = v[(32 - k2) & 31].x; py = v[(32 - k2) & 31].y; }
        // E = A + conj P, D = A - conj P as two packed adds; O = -i D = (D.y, -D.x)
        const float2 pc = make_float2(px, -py);
        const float2 e = cadd(v[k2], pc), d = csub(v[k2], pc);
        const float2 w = tw2[lane + 32 * k2];          // (cos, sin) of 2 pi k / 2048
        const float wx = fmaf(w.x, d.y, -(w.y * d.x));     // w.x O.x + w.y O.y
        const float wy = fmaf(w.x, -d.x, -(w.y * d.y));    // w.x O.y - w.y O.x
        const float2 x2 = cadd(e, make_float2(wx, wy));
        const float xr = x2.x, xi = x2.y;
        const float mag = sqrt_approx(fmaf(xr, xr, xi * xi));
        if (crow) crow[lane + 32 * k2] = x2;
        if (row) row[lane + 32 * k2] = mag;
        sbuf[lane + 32 * k2] = mag;
        cmax = fmaxf(cmax, mag);
        if (k2 == 0 && lane == 0) {
            const float nr = e.x - wx, ni = e.y - wy;
            const float nyq = sqrt_approx(fmaf(nr, nr, ni * ni));
            if (crow) crow[1024] = make_float2(nr, ni);
            if (row) row[1024] = nyq;
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
    const int n_items = 2 * n_tiles;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int item = blockIdx.x;
    if (item >= n_items) return;

    // ---- prologue: barriers, twiddle tables and the first item's samples by TMA ----
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_wave)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar_tables)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        constexpr uint32_t table_bytes = 2 * 1024 * sizeof(float2);
        const uint32_t bt = smem_u32(&sm.bar_tables);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bt), "r"(table_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(smem_u32(&sm.tw[0][0])), "l"(p.tables), "r"(table_bytes), "r"(bt) : "memory");
        sm.desc[0] = make_desc(p, item, n_items);
        issue_wave_copy(sm, sm.desc[0]);
        sm.next = make_desc(p, item + gridDim.x, n_items);
    }
    // per-lane window angles: w[j] = 0.5 - 0.5 cos(2 pi j / 2048), j = 64 n1 + 2 lane (+1)
    if (tid < 32) {
        float ws0, wc0, ws1, wc1;
        sincospif(static_cast<float>(2 * tid) * (2.0f / 2048.0f), &ws0, &wc0);
        sincospif(static_cast<float>(2 * tid + 1) * (2.0f / 2048.0f), &ws1, &wc1);
        sm.win[tid] = make_float4(wc0, ws0, wc1, ws1);
    }
    if (tid == 32) sm.bad = 0;
    __syncthreads();
    mbar_wait(smem_u32(&sm.bar_tables), 0);

    float2* buf = sm.buf[warp];
    for (int it = 0; item < n_items; ++it, item += gridDim.x) {
        const ItemDesc d = sm.desc[it & 1];
        // what the bulk copy does not cover: zero padding and the unaligned remainder
        for (int i = tid; i < d.lo; i += kStftThreads) sm.wave[i] = 0.0f;
        for (int i = d.bulk_end + tid; i < d.hi; i += kStftThreads) sm.wave[i] = __ldg(d.src + i);
        for (int i = max(d.hi, 0) + tid; i < kStftSamples; i += kStftThreads) sm.wave[i] = 0.0f;
        mbar_wait(smem_u32(&sm.bar_wave), it & 1);
        __syncthreads();

        // frame -> registers: z[n] = x[2n] + i x[2n+1], n = 32 n1 + lane, times the Hann window;
        // cos(a + t) = cos a cos t - sin a sin t with a = 2 pi n1 / 32 (immediates), t per lane
        float2 v[32];
        const bool active = warp < d.n_here;
        if (active) {
            const float* frame = sm.wave + warp * kHop;
            const float4 wa = sm.win[lane];
            const float wc0 = wa.x, ws0 = wa.y, wc1 = wa.z, ws1 = wa.w;
            int bad = 0;
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const float2 x = *reinterpret_cast<const float2*>(frame + 64 * n1 + 2 * lane);
                bad |= !isfinite(x.x) | !isfinite(x.y);   // dsp.py:94 finite check, folded into the load
                const float ca = cos32(n1), sa = sin32(n1);
                // HALF the Hann window: the real-input split below yields 2 X of what it is fed, so feeding
                // x w / 2 gives X itself and spares a multiply per bin for |X| and one for the complex spill.
                // Scaling by a power of two commutes with every rounding on the way: the same bits as before.
                const float w0 = fmaf(-0.25f * ca, wc0, fmaf(0.25f * sa, ws0, 0.25f));
                const float w1 = fmaf(-0.25f * ca, wc1, fmaf(0.25f * sa, ws1, 0.25f));
                v[n1] = make_float2(x.x * w0, x.y * w1);
            }
            if (bad) sm.bad = 1;
        }
        if (tid == 0) sm.desc[(it + 1) & 1] = sm.next;   // published by the barrier below
        __syncthreads();   // every warp holds its frame: the wave buffer may be refilled
        if (tid == 0) {
            if (item + gridDim.x < n_items) issue_wave_copy(sm, sm.next);
            sm.next = make_desc(p, item + 2 * gridDim.x, n_items);
        }
        if (active) {
            const long long col = d.col0 + warp;
            float cmax = column_fft(v, sm.tw, sm.tw2, buf, lane, p.spill ? p.spill + col * kSpillStride : nullptr,
                                    p.cspill ? p.cspill + col * kSpillStride : nullptr);
            if (p.do_peaks) {
                cmax = warp_max(cmax);
                __syncwarp();
                column_peaks(reinterpret_cast<const float*>(buf), cmax, lane, p, col);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid == 0 && sm.bad) atomicOr(p.status, 1);
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

static int g_stft_grid = 0;

cudaError_t configure_stft() {
    cudaError_t e = cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(StftSmem)));
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, per_sm = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stft_kernel, kStftThreads, sizeof(StftSmem))) != cudaSuccess) return e;
    g_stft_grid = sms * (per_sm > 0 ? per_sm : 1);
    return cudaSuccess;
}

cudaError_t launch_stft(const StftParams& p, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0) return cudaSuccess;
    const int grid = (2 * n_tiles < g_stft_grid || g_stft_grid <= 0) ? 2 * n_tiles : g_stft_grid;
    stft_kernel<<<grid, kStftThreads, sizeof(StftSmem), stream>>>(p, n_tiles);
    return cudaGetLastError();
}

}  // namespace serb
