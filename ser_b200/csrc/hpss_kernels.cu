// Tonnetz chain, first half: librosa.effects.harmonic(y) for a ragged batch of clips.
//
// Replaces ser/_internal/utils/dsp.py:139  `harmonic = librosa.effects.harmonic(prepared_audio)`
//   = istft(hpss(stft(y, 2048, hop 512))[0], length=len(y))        (librosa 0.11.0 semantics,
//   SURVEY.md Appendix A.7-A.8; oracle: oracle/shim/librosa/effects.py, core.py istft).
//
// hpss_perc_kernel  median of 31 along frequency (scipy.ndimage.median_filter, mode="reflect")
// hpss_harm_kernel  median of 31 along time, then the soft mask (power 2, split_zeros), written
//                   over the frequency median
// istft_kernel      (S * mask) * phase, inverse 2048-point real FFT per warp, Hann window -> frames
// ola_kernel        overlap-add in frame order / window sum-of-squares -> harmonic signal
//
// The medians are computed eight positions at a time.  The eight windows of 31 share a core of
// 24 values -- three aligned groups of eight, two of them shared with the next block: groups are
// sorted once, the merge of two groups serves two blocks, and a 16 + 8 merge pruned to the ranks
// the selection reads finishes a block (66.5 compare-exchanges instead of the 132 of a 24-value
// sort); the other seven values of each window form a small sorted set that slides by one replace
// per position; the median is the rank-15 element of the two sorted sets, seven max and seven
// min.  About 51 min/max/compare per output and no window fill, against 124 (plus the fill) for a
// sliding sorted window of 31.  The work is ALU-bound (FMNMX), not memory-bound.
#include <cfloat>

#include "fft.cuh"
#include "kernels.h"
#include "median_net.cuh"

namespace serb {

namespace {

constexpr int kMedHalf = 15;
constexpr int kMedBlock = 8;      // outputs per block
constexpr int kMedCore = 24;      // values common to the block's eight windows: offsets -8 .. 15
constexpr int kMedExt = 7;        // the rest of a window: offsets -15 .. -9 at first, 16 .. 22 at last

// scipy "reflect" (numpy "symmetric"): d c b a | a b c d | d c b a, period 2n
__device__ __forceinline__ int reflect_index(int i, int n) {
    const int period = 2 * n;
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - 1 - m;
}
// single reflection, valid for -n <= i < 2n
__device__ __forceinline__ int reflect_once(int i, int n) {
    if (i < 0) i = -1 - i;
    if (i >= n) i = 2 * n - 1 - i;
    return i;
}

// Rolling block median.  raw[k] holds the (boundary-extended) input at position f - 15 + k for the
// current block of eight outputs f .. f+7 (k = 0 .. 37).  Consecutive blocks share 30 of the 38
// values, so advancing costs eight loads and a register shift.  `one` is 1.0f passed at run time:
// the remove pass of the small set is a compare plus a PREDICATED MULTIPLY by it (exact) instead of
// a select, which keeps that work off the ALU pipe the min/max instructions saturate.
struct BlockMedian {
    float raw[38];

    template <typename Load>   // x(i): input at position (f - 8) + i, f = the first block's start
    __device__ __forceinline__ void prime(Load x) {
#pragma unroll
        for (int k = 0; k < 30; ++k) raw[k + 8] = x(k - 7);      // f - 15 + k: raw[0..29] after the first advance
    }
    // move to the next block (first call: to the first block) and fetch its last eight values
    template <typename Load>   // x(i): input at position f + i of the NEW block
    __device__ __forceinline__ void advance(Load x) {
#pragma unroll
        for (int k = 0; k < 30; ++k) raw[k] = raw[k + 8];
#pragma unroll
        for (int k = 30; k < 38; ++k) raw[k] = x(k - 15);
    }
    // the same with the eight new values already in registers (fetched one block ahead)
    __device__ __forceinline__ void advance(const float (&fresh)[kMedBlock]) {
#pragma unroll
        for (int k = 0; k < 30; ++k) raw[k] = raw[k + 8];
#pragma unroll
        for (int k = 0; k < kMedBlock; ++k) raw[30 + k] = fresh[k];
    }
    // The core of a block is three aligned groups of eight positions, and consecutive blocks share two
    // of them: every group is sorted once (19 compare-exchanges), the merge of a block's last two groups
    // (25) is shared with the next block, and each block merges that 16 with its third group for the
    // ranks 8 .. 15 only -- all the selection below reads -- (35): 66.5 compare-exchanges per block
    // against the 132 of sorting the 24 from scratch.  Same medians, bit for bit.
    float sa[8], sb[8];     // even block f: sorted groups f - 8 and f; odd block: sa = sorted group f (the next even block's first)
    float mm[16];           // odd block f: the two groups f - 8 and f merged by the even block before it
    int phase = 0;          // 0 first block, 1 even, 2 odd; the same on every thread
    __device__ __forceinline__ void medians(float one, float (&out)[kMedBlock]) {
#ifdef SERB_MEDIAN_SORT24
        float core24[kMedCore];
#pragma unroll
        for (int i = 0; i < kMedCore; ++i) core24[i] = raw[i + 7];    // offsets -8 .. 15
        sort_net24(core24);
        float mid[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mid[i] = core24[8 + i];
#else
        float mid[8], g[8], c[24];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = raw[23 + i];               // offsets 8 .. 15: the newest complete group
        sort_net8(g);
        if (phase != 2) {
            if (phase == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { sa[i] = raw[7 + i]; sb[i] = raw[15 + i]; }
                sort_net8(sa);
                sort_net8(sb);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { mm[i] = sb[i]; mm[8 + i] = g[i]; }
            merge_net8x8(mm);
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = mm[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[16 + i] = sa[i]; sa[i] = g[i]; }
            phase = 2;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = mm[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[16 + i] = g[i]; sb[i] = g[i]; }
            phase = 1;
        }
        middle_net16x8(c);
#pragma unroll
        for (int i = 0; i < 8; ++i) mid[i] = c[8 + i];
#endif
        float ext[kMedExt];
#pragma unroll
        for (int i = 0; i < kMedExt; ++i) ext[i] = raw[i];            // offsets -15 .. -9
        sort_net7(ext);
#pragma unroll
        for (int j = 0; j < kMedBlock; ++j) {
            // rank 15 of core (24, sorted) U ext (7, sorted): min over m of max(core[15 - m], ext[m - 1]),
            // core[8 .. 15] = mid[0 .. 7]
            float med = mid[7];
#pragma unroll
            for (int m = 1; m <= kMedExt; ++m) med = fminf(med, fmaxf(mid[7 - m], ext[m - 1]));
            out[j] = med;
            if (j + 1 < kMedBlock) {
                // window j + 1 drops position f + j - 15 (raw[j]) and gains position f + j + 16 (raw[31 + j])
#pragma unroll
                for (int i = 0; i < kMedExt - 1; ++i)
                    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %0, %1;\n\t@p mul.f32 %0, %2, %3;\n\t}"
                        : "+f"(ext[i]) : "f"(raw[j]), "f"(ext[i + 1]), "f"(one));
                ext[kMedExt - 1] = FLT_MAX;
#pragma unroll
                for (int i = kMedExt - 1; i >= 1; --i) ext[i] = fmaxf(ext[i - 1], fminf(ext[i], raw[31 + j]));
                ext[0] = fminf(ext[0], raw[31 + j]);
            }
        }
    }
};

}  // namespace

__device__ __forceinline__ float harm_mask(float h, float q) {
    // util.softmask(harm, perc, power=2, split_zeros=True): (h/z)^2 / ((h/z)^2 + (q/z)^2) with
    // z = max(h, q).  Dividing by z only guards the squares against under/overflow; scaling by
    // the power of two nearest 1/z does the same exactly and leaves one division.
    const float z = fmaxf(h, q);
    if (z < FLT_MIN) return 0.5f;
    const float s = __int_as_float(0x7e800000 - min(__float_as_int(z) & 0x7f800000, 0x7e000000));   // 2^(-exponent(z) - 1)
    const float a = h * s, b = q * s;
    const float ma = a * a, mb = b * b;
    return __fdividef(ma, ma + mb);
}

// ---- median along time -----------------------------------------------------------------
// one thread per (segment of seg_len columns, bin); bins are contiguous across the warp, so
// every load / store is a coalesced 128-byte row piece.  Runs after hpss_perc_kernel: with both
// medians in hand it overwrites the frequency median with the soft mask (12 bytes of traffic per
// bin; istft_kernel applies the mask to the complex spectrum as it loads it).
__global__ void __launch_bounds__(256, 2) hpss_harm_kernel(HpssParams p) {
    // 1025 bins = four blocks of 256 and one bin: the last block row takes that bin for 256 SEGMENTS per CTA
    // (one segment per thread) instead of one nearly empty CTA per segment holding half an SM's slots
    static_assert((kNBins - 1) % 256 == 0 && kNBins - 1 == 16 * 64, "the last bin is the only one outside whole 256-thread blocks");
    int seg_index = blockIdx.x;
    int f = blockIdx.y * blockDim.x + threadIdx.x;
    if (blockIdx.y == gridDim.y - 1) {
        seg_index = blockIdx.x * blockDim.x + threadIdx.x;
        f = kNBins - 1;
        if (seg_index >= static_cast<int>(gridDim.x)) return;
    }
    const int2 seg = p.segs[seg_index];
    const TonClip clip = p.clips[seg.x];
    const int T = clip.n_cols;
    const int t0 = seg.y;
    const int t1 = min(t0 + p.seg_len, T);
    const float* src = p.mag + static_cast<long long>(clip.col_base) * kSpillStride + f;
    float* perc = p.perc + static_cast<long long>(clip.col_base) * kSpillStride + f;   // in: median along frequency; out: the mask
    const float one = p.one;
    const bool wide = T > 30;           // one reflection reaches every offset of a block and of the look-ahead
    auto at = [&](int t) -> float {
        const int k = wide ? reflect_once(t, T) : reflect_index(t, T);
        return src[static_cast<long long>(k) * kSpillStride];
    };
    BlockMedian bm;
    bm.prime([&](int i) -> float { return at(t0 - 8 + i); });
    float fresh[kMedBlock];
#pragma unroll
    for (int k = 0; k < kMedBlock; ++k) fresh[k] = at(t0 + 15 + k);
    for (int t = t0; t < t1; t += kMedBlock) {
        bm.advance(fresh);
        // everything the block and its successor need from memory is requested before the sorting
        // work, so it is in registers by the time the medians are.  Away from the clip's end the
        // rows are addressed straight off one base pointer (immediate offsets): the reflection and
        // clamp arithmetic of the general path is integer work on the same ALU pipe the min/max
        // instructions saturate.
        const bool interior = (t + 2 * kMedBlock + 15 <= T) && (t + kMedBlock <= t1);    // warp-uniform (but for the last bin's CTAs)
        float pq[kMedBlock];
        if (interior) {
            const float* s_row = src + static_cast<long long>(t + kMedBlock + 15) * kSpillStride;
            const float* p_row = perc + static_cast<long long>(t) * kSpillStride;
#pragma unroll
            for (int k = 0; k < kMedBlock; ++k) fresh[k] = s_row[k * kSpillStride];
#pragma unroll
            for (int j = 0; j < kMedBlock; ++j) pq[j] = p_row[j * kSpillStride];
        } else {
#pragma unroll
            for (int k = 0; k < kMedBlock; ++k) fresh[k] = at(t + kMedBlock + 15 + k);
#pragma unroll
            for (int j = 0; j < kMedBlock; ++j) pq[j] = perc[static_cast<long long>(min(t + j, t1 - 1)) * kSpillStride];
        }
        float med[kMedBlock];
        bm.medians(one, med);
        if (interior) {
            float* p_row = perc + static_cast<long long>(t) * kSpillStride;
#pragma unroll
            for (int j = 0; j < kMedBlock; ++j) p_row[j * kSpillStride] = harm_mask(med[j], pq[j]);
        } else {
#pragma unroll
            for (int j = 0; j < kMedBlock; ++j)
                if (t + j < t1) perc[static_cast<long long>(t + j) * kSpillStride] = harm_mask(med[j], pq[j]);
        }
    }
}

// ---- median along frequency ------------------------------------------------------------
// One lane per (column, run of 64 bins): 16 runs per column (the last one also takes bin 1024), two
// columns per warp.  A lane walks its run eight bins at a time; the eight new values of a step and
// the eight medians are moved as two 16-byte accesses each (runs start on multiples of 8 bins),
// which matters because the lanes of a warp walk 32 different cache lines.
constexpr int kPercRun = 64;

// values at bins pos .. pos + 7 (pos a multiple of 8) of one column, scipy "reflect" at the ends
__device__ __forceinline__ void perc_load8(const float* __restrict__ src, int pos, float (&v)[8]) {
    if (pos >= 0 && pos + 7 < kNBins) {
        const float4 a = *reinterpret_cast<const float4*>(src + pos);
        const float4 b = *reinterpret_cast<const float4*>(src + pos + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = src[reflect_once(pos + k, kNBins)];
    }
}

__global__ void __launch_bounds__(256, 2) hpss_perc_kernel(HpssParams p, int n_cols) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    // 1025 bins = 16 runs of 64 and one bin.  The main warps take the runs (two columns per warp, eight
    // blocks per lane); the last bin of every column is one more block, done by tail warps with one
    // column per lane -- a ninth block on one lane in sixteen would idle the other fifteen for it.
    const int main_warps = (n_cols + 1) >> 1;
    int col, f0, f1;
    if (warp < main_warps) {
        col = 2 * warp + (lane >> 4);
        f0 = (lane & 15) * kPercRun;
        f1 = f0 + kPercRun;
    } else {
        col = (warp - main_warps) * 32 + lane;
        f0 = kNBins - 1;
        f1 = kNBins;
    }
    if (col >= n_cols) return;
    const float one = p.one;
    const float* src = p.mag + static_cast<long long>(col) * kSpillStride;
    float* dst = p.perc + static_cast<long long>(col) * kSpillStride;
    BlockMedian bm;
    float g[8], carry;
    // bins f0 - 16 .. f0 + 15 in four aligned groups: f0 - 15 .. f0 + 14 prime the window, f0 + 15 is
    // the first value of the first step
    perc_load8(src, f0 - 16, g);
#pragma unroll
    for (int k = 1; k < 8; ++k) bm.raw[8 + k - 1] = g[k];
    perc_load8(src, f0 - 8, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) bm.raw[15 + k] = g[k];
    perc_load8(src, f0, g);
#pragma unroll
    for (int k = 0; k < 8; ++k) bm.raw[23 + k] = g[k];
    perc_load8(src, f0 + 8, g);
#pragma unroll
    for (int k = 0; k < 7; ++k) bm.raw[31 + k] = g[k];
    carry = g[7];
    float fresh[kMedBlock];
    perc_load8(src, f0 + 16, g);
    fresh[0] = carry;
#pragma unroll
    for (int k = 0; k < 7; ++k) fresh[k + 1] = g[k];
    carry = g[7];
    for (int f = f0; f < f1; f += kMedBlock) {
        bm.advance(fresh);
        perc_load8(src, f + 24, g);          // bins f + 23 .. f + 30 feed the next step
        fresh[0] = carry;
#pragma unroll
        for (int k = 0; k < 7; ++k) fresh[k + 1] = g[k];
        carry = g[7];
        float med[kMedBlock];
        bm.medians(one, med);
        // the row pitch (1032) leaves room for a full 32-byte store at bin 1024
        *reinterpret_cast<float4*>(dst + f) = make_float4(med[0], med[1], med[2], med[3]);
        *reinterpret_cast<float4*>(dst + f + 4) = make_float4(med[4], med[5], med[6], med[7]);
    }
}

// ---- soft mask + inverse STFT frame ----------------------------------------------------
struct IstftSmem {
    float2 tw[32][32];                 // W_1024^(k1 n2)
    float2 tw2[1024];                  // (cos, sin) 2 pi k / 2048
    float2 buf[8][32 * 33];
    float4 win[32];
};

__global__ void __launch_bounds__(256, 2) istft_kernel(IstftParams p, int n_cols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IstftSmem& sm = *reinterpret_cast<IstftSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += 256) reinterpret_cast<float2*>(&sm.tw[0][0])[i] = p.tables[i];
    if (tid < 32) {
        float ws0, wc0, ws1, wc1;
        sincospif(static_cast<float>(2 * tid) * (2.0f / 2048.0f), &ws0, &wc0);
        sincospif(static_cast<float>(2 * tid + 1) * (2.0f / 2048.0f), &ws1, &wc1);
        sm.win[tid] = make_float4(wc0, ws0, wc1, ws1);
    }
    __syncthreads();
    float2* buf = sm.buf[warp];
    const int n_warps = gridDim.x * 8;
    for (int col = blockIdx.x * 8 + warp; col < n_cols; col += n_warps) {
        const float2* X = p.cspec + static_cast<long long>(col) * kSpillStride;
        const float* M = p.mask + static_cast<long long>(col) * kSpillStride;
        // masked spectrum (S * mask) * phase, k = 32 n1 + lane, kept in registers and mirrored in
        // shared memory so that the partner X[1024 - k] is one conflict-free load away
        float2 v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int k = 32 * n1 + lane;
            const float2 x = X[k];
            const float m = M[k];
            // rounded products: the reference holds (S * mask) * phase as complex64 before the inverse FFT
            v[n1] = make_float2(__fmul_rn(x.x, m), __fmul_rn(x.y, m));
            buf[n1 * 33 + lane] = v[n1];
        }
        float nyq = 0.0f;   // X[1024] (real)
        if (lane == 0) {
            nyq = __fmul_rn(X[1024].x, M[1024]);
            v[0].y = 0.0f;   // irfft ignores the imaginary part of the DC bin
        }
        __syncwarp();
        // conj(Z[k]), Z[k] = E[k] + i O[k]:  2E = A + conj P, 2O = (A - conj P) e^{+i 2 pi k / 2048}
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int kp = 1024 - (32 * n1 + lane);
            float2 pr = make_float2(nyq, 0.0f);
            if (kp < 1024) pr = buf[(kp >> 5) * 33 + (kp & 31)];
            const float ax = v[n1].x, ay = v[n1].y;
            const float ex = ax + pr.x, ey = ay - pr.y;      // A + conj P
            const float dx = ax - pr.x, dy = ay + pr.y;      // A - conj P
            const float2 w = sm.tw2[32 * n1 + lane];
            const float ox = dx * w.x - dy * w.y;            // (dx + i dy)(cos + i sin)
            const float oy = dx * w.y + dy * w.x;
            // 2Z = (ex - oy) + i (ey + ox); conjugate for the forward-FFT inverse
            v[n1] = make_float2(ex - oy, -(ey + ox));
        }
        __syncwarp();
        fft32(v);
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            float2 y = v[k1];
            if (k1 > 0) {
                const float2 w = sm.tw[k1][lane];
                y = make_float2(fmaf(y.x, w.x, -y.y * w.y), fmaf(y.x, w.y, y.y * w.x));
            }
            buf[k1 * 33 + lane] = y;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) v[n2] = buf[lane * 33 + n2];
        __syncwarp();
        fft32(v);
        // z[n] = conj(F[n]) / (2 * 1024), n = lane + 32 k2; x[2n] = Re z, x[2n+1] = Im z; Hann
        const float4 wa = sm.win[lane];
        float2* out = reinterpret_cast<float2*>(p.frames + static_cast<long long>(col) * kNFft);
        constexpr float scale = 1.0f / 2048.0f;
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) {
            const float ca = cos32(k2), sa = sin32(k2);
            const float w0 = fmaf(-0.5f * ca, wa.x, fmaf(0.5f * sa, wa.y, 0.5f));
            const float w1 = fmaf(-0.5f * ca, wa.z, fmaf(0.5f * sa, wa.w, 0.5f));
            out[lane + 32 * k2] = make_float2(v[k2].x * scale * w0, -v[k2].y * scale * w1);
        }
    }
}

// ---- overlap-add + window sum-of-squares ------------------------------------------------
// one thread per four consecutive output samples; the (at most four) frames are added in frame
// order in float32, the squared windows in float64 rounded to float32 after each add, as librosa
// does.  Where all four frames exist the window sum depends only on n mod 512 and comes from a
// table built with the same arithmetic (wss4); the clip's first and last hops take the general
// path sample by sample.
__device__ __forceinline__ float ola_sample(const float* __restrict__ frames, const double* __restrict__ hann_sq,
                                            int m, int t_lo, int t_hi) {
    float acc = 0.0f, wss = 0.0f;
    for (int t = t_lo; t <= t_hi; ++t) {
        const int j = m - t * kHop;
        acc += frames[static_cast<long long>(t) * kNFft + j];
        wss = static_cast<float>(static_cast<double>(wss) + hann_sq[j]);
    }
    return wss > FLT_MIN ? acc / wss : acc;
}

__global__ void __launch_bounds__(256) ola_kernel(OlaParams p) {
    const TonClip clip = p.clips[p.tile_clip[blockIdx.x]];
    const int tile = blockIdx.x - clip.tile_base;
    const int T = clip.n_cols;
    const float* frames = p.frames + static_cast<long long>(clip.col_base) * kNFft;
    float* y = p.yharm + clip.hoff;
    const int n_lo = tile * kColsPerTile * kHop;
    const int n_hi = min(n_lo + kColsPerTile * kHop, clip.length);
    for (int n = n_lo + 4 * threadIdx.x; n < n_hi; n += 4 * blockDim.x) {
        // the four samples n .. n + 3 see the same frames (n and the hop are multiples of four)
        const int m = n + kNFft / 2;
        const int t_hi = min(T - 1, m / kHop);
        const int t_lo = max(0, (m - (kNFft - 1) + kHop - 1) / kHop);
        if (t_hi - t_lo == 3 && n + 3 < n_hi) {
            const float* f = frames + static_cast<long long>(t_lo) * kNFft + (m - t_lo * kHop);
            float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float4 v = *reinterpret_cast<const float4*>(f + t * (kNFft - kHop));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            const float4 w = *reinterpret_cast<const float4*>(p.wss4 + (m & (kHop - 1)));
            float4 out;
            out.x = w.x > FLT_MIN ? acc.x / w.x : acc.x;
            out.y = w.y > FLT_MIN ? acc.y / w.y : acc.y;
            out.z = w.z > FLT_MIN ? acc.z / w.z : acc.z;
            out.w = w.w > FLT_MIN ? acc.w / w.w : acc.w;
            *reinterpret_cast<float4*>(y + n) = out;
        } else {
            for (int i = 0; i < 4 && n + i < n_hi; ++i) y[n + i] = ola_sample(frames, p.hann_sq, m + i, t_lo, t_hi);
        }
    }
}

// ---- inverse STFT fused with the overlap-add ---------------------------------------------
// One warp walks a RUN of consecutive columns of one clip and keeps the three unfinished hop
// blocks of the overlap-add in registers, so the windowed frames never go to memory (the split
// kernels above move 16 KB per column through HBM for them).  The inverse FFT leaves lane l with
// frame samples 64 k2 + 2 l + {0, 1}, k2 = 0 .. 31: the quarter of the frame (hop block) is a
// REGISTER index, so "add quarter q of this frame to block t + q" is plain register arithmetic.
// Block b = sum of quarter (b - t) of frames t = b - 3 .. b, added in frame order exactly as
// ola_kernel does: ((f[b-3] + f[b-2]) + f[b-1]) + f[b], so yharm is bit-identical to the split
// kernels'.  A run recomputes the three frames before its first column (3 / kIstftRun extra work).
constexpr int kIstftWarps = 8;
constexpr int kIstftStageX = 1026;     // float2 per staged spectrum row (1025 used; 16-byte multiple)
constexpr int kIstftStageM = 1028;     // floats per staged mask row
// per warp: the staged column (spectrum + mask, landed by a bulk copy) and the transpose buffer of
// the two-step FFT.  Eight warps per SM with 255 registers each: the transform (64 registers), the
// three unfinished overlap-add blocks (48) and the epilogue fit without spills.  Measured
// alternatives (c2 step): twelve warps of 168 registers that reuse the staged column as transpose
// buffer, keep two blocks in shared memory and prefetch later, 5.42 ms; twelve warps with plain
// global loads and spills, 6.55 ms; this layout, 4.62 ms; the split kernels, 3.38 + 1.50 ms.
struct IstftOlaWarp {
    struct { float2 x[kIstftStageX]; float m[kIstftStageM]; } in;
    float2 buf[32 * 33];
};
struct IstftOlaSmem {
    float2 tw[32][32];                 // W_1024^(k1 n2)
    float2 tw2[1024];                  // (cos, sin) 2 pi k / 2048
    float4 win[32];
    float2 wss4[kHop / 2];             // 1 / window sum of squares where four frames overlap, by n mod 512
    unsigned long long bar[kIstftWarps];
    IstftOlaWarp w[kIstftWarps];
};
static_assert(sizeof(IstftOlaSmem) <= 227 * 1024, "istft_ola_kernel shared memory");

__device__ __forceinline__ void istft_stage_column(IstftOlaWarp& w, unsigned long long* bar, const float2* X, const float* M) {
    const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    constexpr uint32_t bytes_x = 1025 * 8 + 8, bytes_m = 1025 * 4 + 12;       // 16-byte multiples inside the 1032-element rows
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes_x + bytes_m) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(w.in.x))), "l"(X), "r"(bytes_x), "r"(b) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(w.in.m))), "l"(M), "r"(bytes_m), "r"(b) : "memory");
}

__global__ void __launch_bounds__(kIstftWarps * 32, 1) istft_ola_kernel(IstftParams p, OlaParams o, const int2* __restrict__ runs, int n_runs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IstftOlaSmem& sm = *reinterpret_cast<IstftOlaSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 2048; i += kIstftWarps * 32) reinterpret_cast<float2*>(&sm.tw[0][0])[i] = p.tables[i];
    for (int i = tid; i < kHop / 2; i += kIstftWarps * 32) {
        const float2 w = reinterpret_cast<const float2*>(o.wss4)[i];
        sm.wss4[i] = make_float2(1.0f / w.x, 1.0f / w.y);       // reciprocals, see emit()
    }
    if (tid < 32) {
        float ws0, wc0, ws1, wc1;
        sincospif(static_cast<float>(2 * tid) * (2.0f / 2048.0f), &ws0, &wc0);
        sincospif(static_cast<float>(2 * tid + 1) * (2.0f / 2048.0f), &ws1, &wc1);
        sm.win[tid] = make_float4(wc0, ws0, wc1, ws1);
    }
    if (tid < kIstftWarps) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&sm.bar[tid]))));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    IstftOlaWarp& ws = sm.w[warp];
    float2* buf = ws.buf;
    const float4 wa = sm.win[lane];
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&sm.bar[warp]));
    uint32_t parity = 0;
    const int n_warps = gridDim.x * kIstftWarps;
    for (int item = blockIdx.x * kIstftWarps + warp; item < n_runs; item += n_warps) {
        const int2 run = runs[item];
        const TonClip clip = o.clips[run.x];
        const int T = clip.n_cols;
        const int t_first = run.y, t_last = min(run.y + kIstftRun, T) - 1;
        const int t_begin = max(0, t_first - 3);
        float* y = o.yharm + clip.hoff;
        const float2* Xc = p.cspec + static_cast<long long>(clip.col_base) * kSpillStride;
        const float* Mc = p.mask + static_cast<long long>(clip.col_base) * kSpillStride;
        // block b of the clip: padded samples m = 512 b + j, output n = m - 1024; lane owns
        // j = 64 k + 2 lane + {0, 1}, k = 0 .. 7
        auto emit = [&](int b, const float2 (&acc)[8]) {
            const int n0 = kHop * b - kNFft / 2 + 2 * lane;
            if (b < 2 || n0 - 2 * lane >= clip.length) return;
            const bool full = b >= 3 && b <= T - 1;             // all four frames exist: tabulated window sum
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int n = n0 + 64 * k;
                float2 w;
                if (full) {
                    // interior: the tabulated window sum (1.5 up to rounding) as a reciprocal -- sixteen
                    // true divisions per lane and column were 13 % of this kernel's issue slots; the
                    // product is within one ulp of the reference's quotient
                    const float2 r = sm.wss4[32 * k + lane];
                    const float2 out = make_float2(acc[k].x * r.x, acc[k].y * r.y);
                    if (n + 1 < clip.length) *reinterpret_cast<float2*>(y + n) = out;
                    else if (n < clip.length) y[n] = out.x;
                    continue;
                } else {
                    // librosa's window_sumsquare: float64 add, float32 store, frame by frame
                    w = make_float2(0.0f, 0.0f);
                    const int m = n + kNFft / 2;
                    for (int t = max(0, b - 3); t <= min(T - 1, b); ++t) {
                        w.x = static_cast<float>(static_cast<double>(w.x) + o.hann_sq[m - t * kHop]);
                        w.y = static_cast<float>(static_cast<double>(w.y) + o.hann_sq[m + 1 - t * kHop]);
                    }
                }
                float2 out;
                out.x = w.x > FLT_MIN ? acc[k].x / w.x : acc[k].x;
                out.y = w.y > FLT_MIN ? acc[k].y / w.y : acc[k].y;
                if (n + 1 < clip.length) *reinterpret_cast<float2*>(y + n) = out;
                else if (n < clip.length) y[n] = out.x;
            }
        };
        float2 a1[8], a2[8], a3[8];      // blocks t + 1, t + 2, t + 3 after frame t
#pragma unroll
        for (int k = 0; k < 8; ++k) a1[k] = a2[k] = a3[k] = make_float2(0.0f, 0.0f);
        if (lane == 0) istft_stage_column(ws, &sm.bar[warp], Xc + static_cast<long long>(t_begin) * kSpillStride,
                                          Mc + static_cast<long long>(t_begin) * kSpillStride);
        for (int t = t_begin; t <= t_last; ++t) {
            // wait for the column's spectrum and mask
            {
                uint32_t done = 0;
                while (!done)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                        "selp.u32 %0, 1, 0, p;\n\t}"
                        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
                parity ^= 1u;
            }
            // conj(Z[k]), Z[k] = E[k] + i O[k] of the masked spectrum A = (S * mask) * phase:
            // 2E = A + conj P, 2O = (A - conj P) e^{+i 2 pi k / 2048}, P = A[1024 - k]; k = 32 n1 + lane
            float2 v[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int k = 32 * n1 + lane, kp = 1024 - k;
                const float2 x = ws.in.x[k], xp = ws.in.x[kp];
                const float m = ws.in.m[k], mp = ws.in.m[kp];
                // rounded products: the reference holds (S * mask) * phase as complex64 before the inverse FFT
                float ax = __fmul_rn(x.x, m), ay = __fmul_rn(x.y, m);
                float px = __fmul_rn(xp.x, mp), py = __fmul_rn(xp.y, mp);
                if (k == 0) { ay = 0.0f; py = 0.0f; }      // irfft ignores the imaginary parts of the DC and Nyquist bins
                const float ex = ax + px, ey = ay - py;      // A + conj P
                const float dx = ax - px, dy = ay + py;      // A - conj P
                const float2 w = sm.tw2[k];
                const float ox = dx * w.x - dy * w.y;            // (dx + i dy)(cos + i sin)
                const float oy = dx * w.y + dy * w.x;
                v[n1] = make_float2(ex - oy, -(ey + ox));
            }
            __syncwarp();
            // the staged column is in registers: the next one lands while this one is transformed
            if (lane == 0 && t < t_last)
                istft_stage_column(ws, &sm.bar[warp], Xc + static_cast<long long>(t + 1) * kSpillStride,
                                   Mc + static_cast<long long>(t + 1) * kSpillStride);
            fft32(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) {
                float2 yv = v[k1];
                if (k1 > 0) {
                    const float2 w = sm.tw[k1][lane];
                    yv = make_float2(fmaf(yv.x, w.x, -yv.y * w.y), fmaf(yv.x, w.y, yv.y * w.x));
                }
                buf[k1 * 33 + lane] = yv;
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) v[n2] = buf[lane * 33 + n2];
            __syncwarp();
            fft32(v);
            // windowed frame samples 64 k2 + 2 lane + {0, 1} (the arithmetic of istft_kernel; the
            // rounded products are what ola_kernel adds, so no multiply may fuse into the adds below)
            constexpr float scale = 1.0f / 2048.0f;
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) {
                const float ca = cos32(k2), sa = sin32(k2);
                const float w0 = fmaf(-0.5f * ca, wa.x, fmaf(0.5f * sa, wa.y, 0.5f));
                const float w1 = fmaf(-0.5f * ca, wa.z, fmaf(0.5f * sa, wa.w, 0.5f));
                v[k2] = make_float2(__fmul_rn(__fmul_rn(v[k2].x, scale), w0), __fmul_rn(__fmul_rn(-v[k2].y, scale), w1));
            }
            if (t >= t_first) {
                float2 done[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) done[k] = make_float2(__fadd_rn(a1[k].x, v[k].x), __fadd_rn(a1[k].y, v[k].y));
                emit(t, done);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a1[k] = make_float2(__fadd_rn(a2[k].x, v[8 + k].x), __fadd_rn(a2[k].y, v[8 + k].y));
                a2[k] = make_float2(__fadd_rn(a3[k].x, v[16 + k].x), __fadd_rn(a3[k].y, v[16 + k].y));
                a3[k] = v[24 + k];
            }
        }
        if (t_last == T - 1) {     // the clip's tail: blocks that fewer than four frames reach
            emit(T, a1);
            emit(T + 1, a2);
            emit(T + 2, a3);
        }
    }
}

// ---- launchers ---------------------------------------------------------------------------
cudaError_t configure_hpss() {
    cudaError_t e = cudaFuncSetAttribute(istft_ola_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(IstftOlaSmem)));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(sizeof(IstftSmem)));
}

cudaError_t launch_hpss_harm(const HpssParams& p, int n_segs, cudaStream_t stream) {
    if (n_segs <= 0) return cudaSuccess;
    hpss_harm_kernel<<<dim3(n_segs, (kNBins + 255) / 256), 256, 0, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_hpss_perc(const HpssParams& p, int n_cols, cudaStream_t stream) {
    if (n_cols <= 0) return cudaSuccess;
    const int warps = (n_cols + 1) / 2 + (n_cols + 31) / 32;   // two columns per warp, then the last bin of 32 columns per warp
    hpss_perc_kernel<<<(warps + 7) / 8, 256, 0, stream>>>(p, n_cols);
    return cudaGetLastError();
}

cudaError_t launch_istft(const IstftParams& p, int n_cols, cudaStream_t stream) {
    if (n_cols <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = min((n_cols + 7) / 8, 2 * sms);
    istft_kernel<<<grid, 256, sizeof(IstftSmem), stream>>>(p, n_cols);
    return cudaGetLastError();
}

cudaError_t launch_istft_ola(const IstftParams& p, const OlaParams& o, const int2* runs, int n_runs, int n_sms,
                             cudaStream_t stream) {
    if (n_runs <= 0) return cudaSuccess;
    const int grid = min((n_runs + kIstftWarps - 1) / kIstftWarps, n_sms);
    istft_ola_kernel<<<grid, kIstftWarps * 32, sizeof(IstftOlaSmem), stream>>>(p, o, runs, n_runs);
    return cudaGetLastError();
}

cudaError_t launch_ola(const OlaParams& p, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0) return cudaSuccess;
    ola_kernel<<<n_tiles, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace serb
