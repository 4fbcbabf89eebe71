// Tonnetz chain, second half: librosa.feature.tonnetz(y=harmonic, sr) for a ragged batch.
//
// Replaces ser/_internal/utils/dsp.py:140-143
//   tonnetz(y) = Phi @ l1norm(chroma_cqt(y)),  chroma_cqt = linf_norm(cq_to_chroma @ |cqt(y)|)
// with cqt = 252 bins, 36 per octave, 7 octaves, hop 512, fmin = C1 * 2^(tuning/36), computed
// the way librosa 0.11.0 does (oracle/shim/librosa/core.py vqt):
// per octave a rectangular-window STFT of the (recursively half-band decimated) signal times a
// sparsified FFT-domain wavelet basis.
//
// decimate2_kernel   factor-2 FIR decimation (389-tap restatement of libsoxr's HQ low-pass, x sqrt 2):
//                    the two polyphase branches in the halves of packed f32x2 FMAs (FFMA2),
//                    8 outputs per thread, bank-group padded window, fully unrolled
// decimate_any_kernel  early downsampling by 4 / 8 (sample rates >= 64 kHz), plain
// cqt_kernel<R>      n_fft = 64 R real FFT per column (32 / R columns per warp, register DFTs
//                    around one shared-memory transpose), sparse complex rows, |.| / sqrt(len)
// tonnetz_kernel     cq_to_chroma fold, L-inf and L1 normalisation, 6 x 12 projection (fp64),
//                    mean over columns
#include <cfloat>

#include "fft.cuh"
#include "kernels.h"

namespace serb {

// Factor-2 decimator taps (x sqrt 2) paired for packed FP32 FMAs: with K = kDecTaps2 taps,
// c_tap2[e] = (h[K + 1 - 2e], h[K - 2e]) multiplies the sample pair (x[2q], x[2q + 1]) at pair
// distance e = 1 .. (K + 1) / 2 from the output (zero outside the K taps).
constexpr int kDecHalo = (kDecTaps2 + 3) / 4;     // polyphase samples staged before the tile
constexpr int kTapPairs = 2 * kDecHalo;
static_assert(kDecTaps2 % 4 == 1, "the pairing below assumes a tap count of 1 mod 4");
// The same taps (x sqrt 2) in float64 for clips shorter than kDecExactBelow samples: their decimated
// levels are a handful of samples of filter tails whose RATIOS the chroma normalisation reads, so the
// float32 accumulation error (relative to the largest term) would come out amplified; the oracle's
// float64 convolution rounded to float32 is reproduced instead (cost: nothing, these clips are tiny).
__constant__ double c_tap_d[kDecTaps2];
__constant__ __align__(16) float2 c_tap2[kTapPairs];

namespace {

__device__ __forceinline__ int level_length(int len0, int level) {
    int n = len0;
    for (int i = 0; i < level; ++i) n = (n + 1) >> 1;
    return n;
}
// level -1 = the full-rate harmonic signal; level 0 aliases it when there is no early downsampling
__device__ __forceinline__ const float* level_ptr(const CqtParams& p, const TonClip& c, int level) {
    if (level < 0 || (level == 0 && p.early_factor == 1)) return p.yharm + c.hoff;
    return p.yoct + p.level_base[level] + (c.off0 >> level);
}

}  // namespace

// ---- factor-2 decimation -----------------------------------------------------------------
constexpr int kDecTile = 1024;               // outputs per CTA
constexpr int kDecSpan = kDecTile + kTapPairs;   // polyphase samples staged per phase

// d = a * b + c on both halves of a register pair (FFMA2: one issue slot for two FMAs)
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// src_level -1: yharm -> level 0 (early downsampling by 2); otherwise level l -> l + 1
//   out[m] = sum_k h[k] x[2 m + (K - 1) / 2 - k]
//          = sum_{e >= 1} h[K + 1 - 2e] x[2 (m + e - kDecHalo)] + h[K - 2e] x[2 (m + e - kDecHalo) + 1]:
// the two polyphase branches ride in the two halves of one packed accumulator (the signal is
// read as natural (even, odd) sample pairs) and are added once at the end.  Thread t owns
// kDecOuts consecutive outputs; its window starts at pair kDecOuts t, and kDecPad pairs of
// padding after every kDecOuts pairs put the 16-byte reads of eight neighbouring threads into
// distinct bank groups (80-byte thread stride): without it the shared-memory wavefronts, not
// the FMA pipe, bound this kernel (scripts/microbench/decimate_variants.cu).
constexpr int kDecOuts = 8;
constexpr int kDecThreads = kDecTile / kDecOuts;
constexpr int kDecPad = 2;
constexpr int kDecPhys = kDecSpan + kDecPad * ((kDecSpan + kDecOuts - 1) / kDecOuts);

// exact_only: the clips of at least kDecExactBelow samples are left to decimate2_mma_kernel
__global__ void __launch_bounds__(kDecThreads) decimate2_kernel(CqtParams p, int src_level, int exact_only) {
    __shared__ __align__(16) float2 xs[kDecPhys];
    const TonClip clip = p.clips[blockIdx.x];
    if (exact_only && clip.length >= kDecExactBelow) return;
    const int len_in = src_level < 0 ? clip.length : level_length(clip.len0, src_level);
    const int len_out = (len_in + 1) >> 1;
    const float* src = level_ptr(p, clip, src_level);
    float* dst = p.yoct + p.level_base[src_level + 1] + (clip.off0 >> (src_level + 1));
    const unsigned long long* tap2 = reinterpret_cast<const unsigned long long*>(c_tap2);
    if (clip.length < kDecExactBelow) {
        // out[m] = sum_k h[k] x[2 m + (K - 1) / 2 - k] in float64, rounded once
        for (int m = blockIdx.y * kDecThreads + threadIdx.x; m < len_out; m += gridDim.y * kDecThreads) {
            const int centre = 2 * m + (kDecTaps2 - 1) / 2;
            const int k_lo = max(0, centre - (len_in - 1)), k_hi = min(kDecTaps2 - 1, centre);
            double acc = 0.0;
            for (int k = k_lo; k <= k_hi; ++k) acc = fma(c_tap_d[k], static_cast<double>(src[centre - k]), acc);
            dst[m] = static_cast<float>(acc);
        }
        return;
    }
    // grid.y is capped at 65535 tiles: longer signals walk the tiles with a grid stride
    for (int mb = blockIdx.y * kDecTile; mb < len_out; mb += gridDim.y * kDecTile) {
        for (int q = threadIdx.x; q < kDecSpan; q += kDecThreads) {
            const int i = 2 * (mb - kDecHalo + q);
            float2 v = make_float2(0.0f, 0.0f);
            if (i >= 0 && i + 1 < len_in) {
                v = *reinterpret_cast<const float2*>(src + i);
            } else {
                if (i >= 0 && i < len_in) v.x = src[i];
                if (i + 1 >= 0 && i + 1 < len_in) v.y = src[i + 1];
            }
            xs[q + kDecPad * (q / kDecOuts)] = v;
        }
        __syncthreads();
        // acc[r] += tap2[e] * xs[kDecOuts t + r + e], e = 1 .. kTapPairs - 1
        unsigned long long acc[kDecOuts];
#pragma unroll
        for (int r = 0; r < kDecOuts; ++r) acc[r] = 0ull;
        const float2* row = xs + (kDecOuts + kDecPad) * threadIdx.x;
#pragma unroll
        for (int g = 0; g < (kDecOuts + kTapPairs) / 2; ++g) {     // window positions 2g, 2g + 1
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row + 2 * g + kDecPad * ((2 * g) / kDecOuts));
            const unsigned long long vv[2] = {v.x, v.y};
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int r = 0; r < kDecOuts; ++r) {
                    const int e = 2 * g + k - r;
                    if (e >= 1 && e < kTapPairs) acc[r] = fma2(tap2[e], vv[k], acc[r]);
                }
        }
        const int m = mb + kDecOuts * threadIdx.x;
#pragma unroll
        for (int r = 0; r < kDecOuts; ++r)
            if (m + r < len_out) {
                const float2 a = *reinterpret_cast<const float2*>(&acc[r]);
                dst[m + r] = a.x + a.y;
            }
        __syncthreads();
    }
}

// yharm -> level 0 for early factors 4 and 8 (taps pre-scaled by sqrt(factor))
__global__ void __launch_bounds__(256) decimate_any_kernel(CqtParams p) {
    const TonClip clip = p.clips[blockIdx.x];
    const int f = p.early_factor;
    const int half = (p.n_early_taps - 1) / 2;
    const float* src = p.yharm + clip.hoff;
    float* dst = p.yoct + p.level_base[0] + clip.off0;
    for (int m = blockIdx.y * 256 + threadIdx.x; m < clip.len0; m += gridDim.y * 256) {
        const int centre = f * m + half;
        const int k_lo = max(0, centre - (clip.length - 1));
        const int k_hi = min(p.n_early_taps - 1, centre);
        float acc = 0.0f;
        for (int k = k_lo; k <= k_hi; ++k) acc = fmaf(p.early_taps[k], src[centre - k], acc);
        dst[m] = acc;
    }
}

// ---- constant-Q response of one octave ---------------------------------------------------
__device__ __forceinline__ void fft16(float2* v) {
    // n = 4a + b, k = k2 + 4 k1
#pragma unroll
    for (int b = 0; b < 4; ++b) fft4(v[b], v[4 + b], v[8 + b], v[12 + b]);
#pragma unroll
    for (int k2 = 1; k2 < 4; ++k2)
#pragma unroll
        for (int b = 1; b < 4; ++b)
            v[4 * k2 + b] = cmul_conj_tw(v[4 * k2 + b], cos32(2 * b * k2), sin32(2 * b * k2));
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) fft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
    float2 t[16];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2)
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) t[k2 + 4 * k1] = v[4 * k2 + k1];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = t[i];
}

template <int R>
__device__ __forceinline__ void fft_small(float2* v) {
    if constexpr (R == 4) fft4(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 8) fft8(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    else fft16(v);
}

constexpr size_t kCqtMaxSmem = 227 * 1024;   // dynamic shared memory every instantiation is opted in to
constexpr int kCqtWarps = 8;
constexpr int kCqtBufPitch = 34;              // float2 pitch of the transpose buffer: 16-byte aligned rows 272 bytes apart
constexpr int kCqtValPitch = kCqRowCap + 1;   // float2 pitch of a staged basis row (bank spread)

struct CqtSmemHead {
    float2 buf[kCqtWarps][32 * kCqtBufPitch]; // per-warp transpose buffer, then the column spectra
    float2 vals[kCqRows][kCqtValPitch];       // sparse basis rows of this (tuning, octave)
    CqRow rows[kCqRows];
    float mags[kCqtWarps][2 * kCqRows];       // scaled magnitudes of two of the warp's columns, by bin within the octave
                                              // (two CTAs per SM leave room for no more)
    int cmax;                                 // widest staged row, rounded up to a multiple of four
    int bin_lo, bin_hi;                       // smallest and largest first bin of the staged rows
};

// One CTA = one clip, one octave, `cols_per_block` consecutive columns.  The signal span those
// columns touch is staged once in shared memory (zero padded at the clip ends), together with
// the 36 sparse basis rows of the clip's tuning; after one barrier every warp works alone:
// G = 32 / R columns per iteration, R complex points per lane per column (N = 32 R = n_fft / 2).
//
// SHARED (octaves whose hop is at most 16 samples): neighbouring frames overlap by more than 98 %,
// and step A of the transform -- the R-point DFT over n1 of z[ct + 32 n1 + l] -- depends on the
// absolute position m = ct + l only.  The CTA computes D~[m][k1] = W_N^(m k1) sum_n1 z[m + 32 n1]
// W_R^(n1 k1) ONCE for its span (a few hundred m instead of 32 per column), and a column is
//   Z[k1 + R k2] = W_N^(-ct k1) sum_l W_32^(l k2) D~[ct + l][k1]:
// 32 loads, one 32-point DFT and one constant phase per lane.  No frame loads, no step A, no
// transpose: the shared-memory traffic of the transform drops from 12 to 4 KB per column and
// its arithmetic by two fifths.
template <int R, bool SHARED>
__global__ void __launch_bounds__(kCqtWarps * 32, 2) cqt_kernel(CqtParams p, int oct_first) {
    constexpr int G = 32 / R;
    constexpr int N = 32 * R;
    extern __shared__ __align__(16) unsigned char cqt_smem_raw[];
    CqtSmemHead& sm = *reinterpret_cast<CqtSmemHead*>(cqt_smem_raw);
    float* sig_s = reinterpret_cast<float*>(cqt_smem_raw + sizeof(CqtSmemHead));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // consecutive octaves of equal FFT size share a launch (blockIdx.z)
    const int octave = oct_first + static_cast<int>(blockIdx.z);
    const int cols_per_block = p.cq_cols_per_block[octave];
    const TonClip clip = p.clips[blockIdx.x];
    const int t_block = blockIdx.y * cols_per_block;
    if (t_block >= clip.cq_cols) return;
    if (tid == 0) { sm.cmax = 4; sm.bin_lo = N; sm.bin_hi = 0; }
    __syncthreads();
    const int n_here = min(cols_per_block, clip.cq_cols - t_block);
    const int tuning = p.tuning_idx[blockIdx.x];
    const float* sig = level_ptr(p, clip, octave);
    const int len = level_length(clip.len0, octave);
    const int hop = p.hop0 >> octave;
    // ---- stage the span and the basis ----
    const int s0 = t_block * hop - N;                       // signal index of sig_s[0]
    const int span = (n_here - 1) * hop + 2 * N;
    // asynchronous copies (LDGSTS): every request of the CTA is in flight at once; samples outside
    // the clip are zero-filled by a zero source size
    for (int i = tid; i < span; i += kCqtWarps * 32) {
        const int j = s0 + i;
        const bool inside = j >= 0 && j < len;
        const float* src = inside ? sig + j : sig;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(sig_s + i))),
                     "l"(src), "r"(inside ? 4 : 0) : "memory");
    }
    const size_t bank = (static_cast<size_t>(tuning) * kCqOctaves + octave) * kCqRows;
    for (int i = tid; i < kCqRows * kCqRowCap; i += kCqtWarps * 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&sm.vals[i / kCqRowCap][i % kCqRowCap]))),
                     "l"(p.vals + bank * kCqRowCap + i) : "memory");
    if (tid < kCqRows) {
        const CqRow row = p.rows[bank + tid];
        sm.rows[tid] = row;
        atomicMax(&sm.cmax, (row.count + 3) & ~3);
        atomicMin(&sm.bin_lo, row.start);
        atomicMax(&sm.bin_hi, row.start);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const float2* twa = p.twiddles + (N - 128);            // W_N^j = (cos, -sin)
    const float2* twb = p.twiddles + 1920 + (N - 128);     // (cos, sin) 2 pi k / (2N)
    float2* buf = sm.buf[warp];
    const int cmax = min(sm.cmax, kCqRowCap);
    // the rows only read bins [bin_lo, bin_hi + cmax): the real-input split is done for the
    // register indices k2 (bins k1 + R k2) that overlap that range and skipped for the rest
    const int k2_lo = sm.bin_lo / R, k2_hi = min(31, (sm.bin_hi + cmax - 1) / R);
    const int g2 = lane / R, k1 = lane % R;
    const int src = (k1 == 0) ? lane : g2 * R + (R - k1);
    // SHARED: D~[k1][m] behind the staged span, row pitch = 1 mod 16 so that the sixteen k1 of a
    // half warp read sixteen different bank pairs.  The table covers `sub_cols` columns at a time
    // (the CTA's columns in sub-blocks, two CTA barriers each), so that the CTA keeps the long
    // column run that amortises the staging of the basis.
    const int h2 = hop >> 1;
    const int sub_cols = SHARED ? p.cq_sub_cols[octave] : cols_per_block;
    const int dpitch = ((sub_cols - 1) * h2 + 32 + 15) / 16 * 16 + 1;
    float2* dt = reinterpret_cast<float2*>(sig_s + (((cols_per_block - 1) * hop + 2 * N + 3) & ~3));
    for (int sb = 0; sb < n_here; sb += sub_cols) {
    const int sub_end = min(sb + sub_cols, n_here);
    if constexpr (SHARED) {
        if (sb) __syncthreads();                       // the previous sub-block's readers are done
        const int m_range = (sub_end - sb - 1) * h2 + 32;
        const float* base = sig_s + 2 * sb * h2;       // complex sample 0 of the sub-block
        for (int m = tid; m < m_range; m += kCqtWarps * 32) {
            float2 y[R];
#pragma unroll
            for (int n1 = 0; n1 < R; ++n1) y[n1] = *reinterpret_cast<const float2*>(base + 2 * (m + 32 * n1));
            fft_small<R>(y);
#pragma unroll
            for (int q = 0; q < R; ++q) {
                float2 z = y[q];
                if (q > 0) {
                    const float2 w = twa[(m * q) & (N - 1)];
                    z = make_float2(fmaf(z.x, w.x, -z.y * w.y), fmaf(z.x, w.y, z.y * w.x));
                }
                dt[q * dpitch + m] = z;
            }
        }
        __syncthreads();
    }
    for (int lc0 = sb + warp * G; lc0 < sub_end; lc0 += kCqtWarps * G) {
        float2 v[32];
        if constexpr (SHARED) {
            // lane = (column g2, k1): the 32 table entries of this column's span, then the DFT over l
            const int ct = (min(lc0 + g2, sub_end - 1) - sb) * h2;     // relative to the sub-block's table
            const float2* drow = dt + k1 * dpitch + ct;
#pragma unroll
            for (int l = 0; l < 32; ++l) v[l] = drow[l];
        } else {
        // frames (rectangular window): z[n] = x[2n] + i x[2n+1], n = 32 n1 + lane

#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int lc = min(lc0 + g, n_here - 1);       // surplus columns repeat the last one (not stored)
            const float* frame = sig_s + lc * hop;
            if (hop & 1) {     // bottom octave at 192 kHz and above: frames start on odd samples
#pragma unroll
                for (int n1 = 0; n1 < R; ++n1)
                    v[g * R + n1] = make_float2(frame[2 * (32 * n1 + lane)], frame[2 * (32 * n1 + lane) + 1]);
            } else {
#pragma unroll
                for (int n1 = 0; n1 < R; ++n1)
                    v[g * R + n1] = *reinterpret_cast<const float2*>(frame + 2 * (32 * n1 + lane));
            }
        }
        // step A: R-point DFTs over n1 (per column), twiddle W_N^(lane k1), transpose
        if constexpr (R == 32) {
            fft32(v);
        } else {
#pragma unroll
            for (int g = 0; g < G; ++g) fft_small<R>(v + g * R);
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int q = 0; q < R; ++q) {
                float2 y = v[g * R + q];
                if (q > 0) {
                    const float2 w = twa[lane * q];
                    y = make_float2(fmaf(y.x, w.x, -y.y * w.y), fmaf(y.x, w.y, y.y * w.x));
                }
                buf[(g * R + q) * kCqtBufPitch + lane] = y;
            }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2 += 2) {     // conflict-free 16-byte loads
            const float4 q2 = *reinterpret_cast<const float4*>(&buf[lane * kCqtBufPitch + n2]);
            v[n2] = make_float2(q2.x, q2.y);
            v[n2 + 1] = make_float2(q2.z, q2.w);
        }
        __syncwarp();
        }
        // step B: 32-point DFT over n2; lane = (column g2, k1): v[k2] = Z[k1 + R k2]
        fft32(v);
        if constexpr (SHARED) {
            // the column's constant phase W_N^(-ct k1) = conj(twa[ct k1])
            const int ct = (min(lc0 + g2, sub_end - 1) - sb) * h2;
            const float2 w = twa[(ct * k1) & (N - 1)];
            if (k1 > 0) {
#pragma unroll
                for (int k2 = 0; k2 < 32; ++k2)
                    v[k2] = make_float2(fmaf(v[k2].x, w.x, v[k2].y * w.y), fmaf(v[k2].y, w.x, -v[k2].x * w.y));
            }
        }
        float2* xs = buf + g2 * (N + 1);
        // in groups of four register indices (one warp-uniform branch per group, so that the
        // twiddle loads and shuffles of a group are in flight together); the up to three surplus
        // bins at either end are written and never read
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) {
            if ((k2 | 3) < k2_lo || (k2 & ~3) > k2_hi) continue;        // warp-uniform, same for a group of four
            float px = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
            float py = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
            if (k1 == 0) { px = v[(32 - k2) & 31].x; py = v[(32 - k2) & 31].y; }
            // E = A + conj P, D = A - conj P as two packed adds; O = -i D = (D.y, -D.x)
            const float2 pc = make_float2(px, -py);
            const float2 e = cadd(v[k2], pc), d = csub(v[k2], pc);
            const int k = k1 + R * k2;
            const float2 w = twb[k];
            const float wx = fmaf(w.x, d.y, -(w.y * d.x));     // w.x O.x + w.y O.y
            const float wy = fmaf(w.x, -d.x, -(w.y * d.y));    // w.x O.y - w.y O.x
            xs[k] = cscale(cadd(e, make_float2(wx, wy)), 0.5f);
        }
        // Nyquist bin X[N] = Re Z[0] - Im Z[0], only if a row reaches it (sample rates whose top
        // wavelet sits right under the Nyquist frequency)
        if (sm.bin_hi + cmax > N && k1 == 0) xs[N] = make_float2(v[0].x - v[0].y, 0.0f);
        __syncwarp();
        // sparse basis rows: C[r] = sum_c B[r][c] X[start + c].  The staged rows are zero padded, so
        // every lane runs the same trip count (the widest row of this basis, a multiple of four).
        // Rows 0..31: lane = row, all G columns of the warp at once (one basis load feeds G
        // products: shared-memory bandwidth is what bounds this kernel).  Rows 32..35: lane =
        // (row, column).
        float mval[G];      // rows 0..31: this lane's row, every column
        int mbin;
        float mag4;         // rows 32..35: this lane's (row, column), valid on the lanes with part == 0
        int bin4, g4;
        {
            const CqRow row = sm.rows[lane];
            const float2* b = sm.vals[lane];
            const float2* x = buf + row.start;
            float cr[G], ci[G];
#pragma unroll
            for (int g = 0; g < G; ++g) { cr[g] = 0.0f; ci[g] = 0.0f; }
#pragma unroll 1
            for (int c = 0; c < cmax; c += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float2 bv = b[c + u];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float2 xv = x[g * (N + 1) + c + u];
                        cr[g] = fmaf(bv.x, xv.x, fmaf(-bv.y, xv.y, cr[g]));
                        ci[g] = fmaf(bv.x, xv.y, fmaf(bv.y, xv.x, ci[g]));
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float mag;
                asm("sqrt.approx.f32 %0, %1;" : "=f"(mag) : "f"(fmaf(cr[g], cr[g], ci[g] * ci[g])));
                mval[g] = mag * row.scale;
                if (p.cqmag && lc0 + g < n_here) p.cqmag[(clip.cq_base + t_block + lc0 + g) * kCqBins + row.bin] = mval[g];
            }
            mbin = row.bin % kCqRows;
        }
        {
            // rows 32..35 for all G columns: every (row, column) is shared by P = 8 / G lanes that
            // take every P-th basis value and add up with shuffles, so the whole warp stays busy
            constexpr int P = 8 / G;
            const int part = lane % P, rg = lane / P;
            const int r = 32 + (rg & 3), g = rg >> 2;
            const CqRow row = sm.rows[r];
            const float2* b = sm.vals[r];
            const float2* x = buf + g * (N + 1) + row.start;
            float cr = 0.0f, ci = 0.0f;
#pragma unroll 4
            for (int c = part; c < cmax; c += P) {
                const float2 bv = b[c], xv = x[c];
                cr = fmaf(bv.x, xv.x, fmaf(-bv.y, xv.y, cr));
                ci = fmaf(bv.x, xv.y, fmaf(bv.y, xv.x, ci));
            }
#pragma unroll
            for (int o = P / 2; o > 0; o >>= 1) {
                cr += __shfl_xor_sync(0xffffffffu, cr, o);
                ci += __shfl_xor_sync(0xffffffffu, ci, o);
            }
            asm("sqrt.approx.f32 %0, %1;" : "=f"(mag4) : "f"(fmaf(cr, cr, ci * ci)));
            mag4 *= row.scale;
            bin4 = part == 0 ? row.bin % kCqRows : -1;
            g4 = g;
            if (p.cqmag && part == 0 && lc0 + g < n_here) p.cqmag[(clip.cq_base + t_block + lc0 + g) * kCqBins + row.bin] = mag4;
        }
        __syncwarp();
        // this octave's share of the chroma fold (filters.cq_to_chroma, 36 bins per octave: chroma c <-
        // bins 3 c - 1, 3 c, 3 c + 1 of the octave, the first wrapping to bin 35): 12 floats per column
        // instead of 36 magnitudes, so the 252-wide constant-Q matrix never goes to memory
        {
            const int oct_slot = sm.rows[0].bin / kCqRows;
            float* ms = sm.mags[warp];
#pragma unroll
            for (int g0 = 0; g0 < G; g0 += 2) {            // two columns per pass
#pragma unroll
                for (int g = g0; g < g0 + 2 && g < G; ++g) ms[(g - g0) * kCqRows + mbin] = mval[g];
                if (bin4 >= 0 && g4 >= g0 && g4 < g0 + 2) ms[(g4 - g0) * kCqRows + bin4] = mag4;
                __syncwarp();
                if (lane < 24) {
                    const int g = g0 + lane / 12, c = lane % 12;
                    if (g < G && lc0 + g < n_here) {
                        const float* m = ms + (g - g0) * kCqRows;
                        const float sum = (m[(3 * c + kCqRows - 1) % kCqRows] + m[3 * c]) + m[3 * c + 1];
                        p.cq_chroma[(static_cast<size_t>(clip.cq_base) + t_block + lc0 + g) * (kCqOctaves * 12) + oct_slot * 12 + c] = sum;
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }
    }
}

// ---- n_fft = 1024 / 512 octaves with the rows mapped lane = column ------------------------------
// Same transform as cqt_kernel<R, SHARED>; what differs is the product with the sparse rows.  There
// the lane is a ROW: every lane walks its own row of the basis and its own window of the spectrum
// (all loads distinct: 45 % of the kernel's shared-memory wavefronts, 2.1 of its 8.5 ms per c2 step,
// every row padded to the widest).  Here the CTA's eight warps transform CI = 16 (n_fft 1024) or 32
// (512) columns, leave the bins the rows read ([bin_lo, bin_lo + n_bins), about 100 of 513) in
// shared memory column-major, and after a CTA barrier the lane is a COLUMN: CI consecutive threads
// hold one row set (CqSetBank: three or two consecutive rows over the union of their bins), so one
// spectrum load feeds every row of the set, the basis values are warp-uniform 16-byte loads, and
// every row runs its own length.  Magnitudes meet in a [CI columns][36] buffer; after a second
// barrier they are folded into the octave's 12 chroma shares while the other warps already
// transform the next CI columns.
//
// Column c of warp w's G = 32 / R columns lives at Xw[c * kCcPitch + (bin - x_base)] inside warp w's
// own transpose region (free between its exchange and the next iteration's); the regions are
// 8 G bytes mod 128 apart and kCcPitch = 1 mod 16, so 16 consecutive lanes read 16 different
// 8-byte bank slots.
// float2 between two columns of a warp: the split writes whole groups of four register indices (4 R bins),
// and at most 128 bins starting anywhere touch 128 / (4 R) + 1 such groups; 193 and 161 are 1 mod 16
template <int R> constexpr int kCcPitch = (128 / (4 * R) + 1) * 4 * R + 1;
constexpr int kCcMagPitch = kCqRows + 1;
// float2 per warp: the 32 x 34 transpose buffer + the bank shift of 8 G bytes; the shared-stage variant has
// no transpose and keeps the G columns only (G x pitch x 8 bytes = 8 G mod 128 as well)
template <int R, bool SHARED> constexpr int kCcRegion = SHARED ? (32 / R) * kCcPitch<R> : 32 * kCqtBufPitch + 32 / R;
static_assert(4 * kCcPitch<8> <= 32 * kCqtBufPitch && 2 * kCcPitch<16> <= 32 * kCqtBufPitch, "the columns' bins fit the warp's transpose region");
static_assert((kCcRegion<16, true> * 8) % 128 == 16 && (kCcRegion<16, false> * 8) % 128 == 16 &&
              (kCcRegion<8, true> * 8) % 128 == 32 && (kCcRegion<8, false> * 8) % 128 == 32, "bank shift between the warps' regions");

template <int R, bool SHARED>
struct CqtcHead {
    float2 buf[kCqtWarps * kCcRegion<R, SHARED>];
    CqSetBank bank;
    float mags[kCqtWarps * (32 / R)][kCcMagPitch];
};

template <int R, bool SHARED>
__global__ void __launch_bounds__(kCqtWarps * 32, 2) cqtc_kernel(CqtParams p, int oct_first) {
    constexpr int G = 32 / R, N = 32 * R;
    constexpr int CI = kCqtWarps * G;                      // columns per CTA iteration: 16 (n_fft 1024) or 32 (512)
    constexpr int kSetsPerThread = kCqSets * CI / (kCqtWarps * 32);
    constexpr int kPitch = kCcPitch<R>;
    extern __shared__ __align__(16) unsigned char cqt_smem_raw[];
    CqtcHead<R, SHARED>& sm = *reinterpret_cast<CqtcHead<R, SHARED>*>(cqt_smem_raw);
    float* sig_s = reinterpret_cast<float*>(cqt_smem_raw + sizeof(CqtcHead<R, SHARED>));
    constexpr int kRegion = kCcRegion<R, SHARED>;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // blockIdx.y walks the launch's octaves one after the other, each with its own number of column blocks
    // (an octave whose blocks hold 128 columns does not launch the 18 blocks a 16-column octave needs)
    int oct_in_launch = 0, block_lo = 0;
    while (static_cast<int>(blockIdx.y) >= p.cq_block_end[oct_in_launch]) block_lo = p.cq_block_end[oct_in_launch++];
    const int octave = oct_first + oct_in_launch;
    const int cols_per_block = p.cq_cols_per_block[octave];
    const TonClip clip = p.clips[blockIdx.x];
    const int t_block = (static_cast<int>(blockIdx.y) - block_lo) * cols_per_block;
    if (t_block >= clip.cq_cols) return;
    const int n_here = min(cols_per_block, clip.cq_cols - t_block);
    const int tuning = p.tuning_idx[blockIdx.x];
    const float* sig = level_ptr(p, clip, octave);
    const int len = level_length(clip.len0, octave);
    const int hop = p.hop0 >> octave;
    // ---- stage the span and the rows ----
    const int s0 = t_block * hop - N;
    const int span = (n_here - 1) * hop + 2 * N;
    if (((s0 | span) & 3) == 0 && (reinterpret_cast<uintptr_t>(sig) & 15) == 0) {
        // 16 bytes per request where the four samples lie inside the clip, else sample by sample
        for (int i = 4 * tid; i < span; i += 4 * kCqtWarps * 32) {
            const int j = s0 + i;
            const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(sig_s + i));
            if (j >= 0 && j + 3 < len) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(sig + j) : "memory");
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool inside = j + u >= 0 && j + u < len;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4 * u), "l"(inside ? sig + j + u : sig),
                                 "r"(inside ? 4 : 0) : "memory");
                }
            }
        }
    } else {
        for (int i = tid; i < span; i += kCqtWarps * 32) {
            const int j = s0 + i;
            const bool inside = j >= 0 && j < len;
            const float* src = inside ? sig + j : sig;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(sig_s + i))),
                         "l"(src), "r"(inside ? 4 : 0) : "memory");
        }
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.set_banks + static_cast<size_t>(tuning) * kCqOctaves + octave);
        for (int i = tid; i < static_cast<int>(sizeof(CqSetBank) / 16); i += kCqtWarps * 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(reinterpret_cast<uint4*>(&sm.bank) + i))),
                         "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const float2* twa = p.twiddles + (N - 128);            // W_N^j = (cos, -sin)
    const float2* twb = p.twiddles + 1920 + (N - 128);     // (cos, sin) 2 pi k / (2N)
    float2* buf = sm.buf + warp * kRegion;
    const int bin_lo = sm.bank.bin_lo, n_bins = sm.bank.n_bins;
    const int k2_lo = bin_lo / R, k2_hi = min(31, (bin_lo + n_bins - 1) / R);
    const int x_base = R * (k2_lo & ~3);                   // bin held at index 0 of a column
    const int g2 = lane / R, k1 = lane % R;
    const int src = (k1 == 0) ? lane : g2 * R + (R - k1);
    // this thread's part of the row product: lane = column, CI consecutive threads share a row set (a half
    // warp holds one set at n_fft 1024; a warp walks two sets one after the other at n_fft 512)
    const int set0 = (tid / CI) * kSetsPerThread, col = tid % CI;
    const float2* xcol0 = sm.buf + (col / G) * kRegion + (col % G) * kPitch + (bin_lo - x_base);
    const int oct_slot = sm.bank.sets[0].bin[0] / kCqRows;
    // first-stage table of the shared variant (see cqt_kernel)
    const int h2 = hop >> 1;
    const int sub_cols = SHARED ? p.cq_sub_cols[octave] : cols_per_block;
    const int dpitch = ((sub_cols - 1) * h2 + 32 + 15) / 16 * 16 + 1;
    float2* dt = reinterpret_cast<float2*>(sig_s + (((cols_per_block - 1) * hop + 2 * N + 3) & ~3));
    for (int sb = 0; sb < n_here; sb += sub_cols) {
        const int sub_end = min(sb + sub_cols, n_here);
        if constexpr (SHARED) {
            if (sb) __syncthreads();                       // the previous sub-block's table readers are done
            const int m_range = (sub_end - sb - 1) * h2 + 32;
            const float* base = sig_s + 2 * sb * h2;
            for (int m = tid; m < m_range; m += kCqtWarps * 32) {
                float2 y[R];
#pragma unroll
                for (int n1 = 0; n1 < R; ++n1) y[n1] = *reinterpret_cast<const float2*>(base + 2 * (m + 32 * n1));
                fft_small<R>(y);
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    float2 z = y[q];
                    if (q > 0) {
                        const float2 w = twa[(m * q) & (N - 1)];
                        z = make_float2(fmaf(z.x, w.x, -z.y * w.y), fmaf(z.x, w.y, z.y * w.x));
                    }
                    dt[q * dpitch + m] = z;
                }
            }
            __syncthreads();
        }
        // every warp runs every iteration (CTA barriers inside): surplus columns repeat the last one
        for (int it0 = sb; it0 < sub_end; it0 += kCqtWarps * G) {
            const int lc0 = it0 + warp * G;
            float2 v[32];
            if constexpr (SHARED) {
                const int ct = (min(lc0 + g2, sub_end - 1) - sb) * h2;
                const float2* drow = dt + k1 * dpitch + ct;
#pragma unroll
                for (int l = 0; l < 32; ++l) v[l] = drow[l];
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int lc = min(lc0 + g, n_here - 1);
                    const float* frame = sig_s + lc * hop;
                    if (hop & 1) {
#pragma unroll
                        for (int n1 = 0; n1 < R; ++n1)
                            v[g * R + n1] = make_float2(frame[2 * (32 * n1 + lane)], frame[2 * (32 * n1 + lane) + 1]);
                    } else {
#pragma unroll
                        for (int n1 = 0; n1 < R; ++n1)
                            v[g * R + n1] = *reinterpret_cast<const float2*>(frame + 2 * (32 * n1 + lane));
                    }
                }
#pragma unroll
                for (int g = 0; g < G; ++g) fft_small<R>(v + g * R);
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        float2 y = v[g * R + q];
                        if (q > 0) {
                            const float2 w = twa[lane * q];
                            y = make_float2(fmaf(y.x, w.x, -y.y * w.y), fmaf(y.x, w.y, y.y * w.x));
                        }
                        buf[(g * R + q) * kCqtBufPitch + lane] = y;
                    }
                __syncwarp();
#pragma unroll
                for (int n2 = 0; n2 < 32; n2 += 2) {
                    const float4 q2 = *reinterpret_cast<const float4*>(&buf[lane * kCqtBufPitch + n2]);
                    v[n2] = make_float2(q2.x, q2.y);
                    v[n2 + 1] = make_float2(q2.z, q2.w);
                }
                __syncwarp();
            }
            fft32(v);
            if constexpr (SHARED) {
                const int ct = (min(lc0 + g2, sub_end - 1) - sb) * h2;
                const float2 w = twa[(ct * k1) & (N - 1)];
                // only the register indices the split reads: k2 of the needed groups and their mirrors 31 - k2
                if (k1 > 0) {
#pragma unroll
                    for (int k2 = 0; k2 < 32; ++k2) {
                        const int m2 = 31 - k2;
                        if (((k2 | 3) < k2_lo || (k2 & ~3) > k2_hi) && ((m2 | 3) < k2_lo || (m2 & ~3) > k2_hi)) continue;
                        v[k2] = make_float2(fmaf(v[k2].x, w.x, v[k2].y * w.y), fmaf(v[k2].y, w.x, -v[k2].x * w.y));
                    }
                }
            }
            // real-input split for the bins the rows read, into this warp's column-major pair
            float2* xw = buf + g2 * kPitch - x_base;
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) {
                if ((k2 | 3) < k2_lo || (k2 & ~3) > k2_hi) continue;        // warp-uniform, same for a group of four
                float px = __shfl_sync(0xffffffffu, v[31 - k2].x, src);
                float py = __shfl_sync(0xffffffffu, v[31 - k2].y, src);
                if (k1 == 0) { px = v[(32 - k2) & 31].x; py = v[(32 - k2) & 31].y; }
                const float2 pc = make_float2(px, -py);
                const float2 e = cadd(v[k2], pc), d = csub(v[k2], pc);
                const int k = k1 + R * k2;
                const float2 w = twb[k];
                const float wx = fmaf(w.x, d.y, -(w.y * d.x));
                const float wy = fmaf(w.x, -d.x, -(w.y * d.y));
                xw[k] = cscale(cadd(e, make_float2(wx, wy)), 0.5f);   // surplus bins of the group are written and never read
            }
            // Nyquist bin X[N] = Re Z[0] - Im Z[0], only if a row reaches it
            if (bin_lo + n_bins > N && k1 == 0) xw[N] = make_float2(v[0].x - v[0].y, 0.0f);   // N - x_base < kPitch
            __syncthreads();                               // A: the 16 columns' bins are in place
            const int n_valid = min(CI, sub_end - it0);
#pragma unroll
            for (int si = 0; si < kSetsPerThread; ++si) {
                const CqSet set = sm.bank.sets[set0 + si];
                const float2* xcol = xcol0 + set.u0;
                const float4* b4 = reinterpret_cast<const float4*>(sm.bank.vals + set.off);
                float mag[3];
                if (warp < 2) {
                    // three rows per set, stored [bin][4]
                    float cr[3] = {0.f, 0.f, 0.f}, ci[3] = {0.f, 0.f, 0.f};
                    for (int b = 0; b < set.ulen; ++b) {
                        const float2 xv = xcol[b];
                        const float4 b01 = b4[2 * b], b23 = b4[2 * b + 1];
                        cr[0] = fmaf(b01.x, xv.x, fmaf(-b01.y, xv.y, cr[0])); ci[0] = fmaf(b01.x, xv.y, fmaf(b01.y, xv.x, ci[0]));
                        cr[1] = fmaf(b01.z, xv.x, fmaf(-b01.w, xv.y, cr[1])); ci[1] = fmaf(b01.z, xv.y, fmaf(b01.w, xv.x, ci[1]));
                        cr[2] = fmaf(b23.x, xv.x, fmaf(-b23.y, xv.y, cr[2])); ci[2] = fmaf(b23.x, xv.y, fmaf(b23.y, xv.x, ci[2]));
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) asm("sqrt.approx.f32 %0, %1;" : "=f"(mag[q]) : "f"(fmaf(cr[q], cr[q], ci[q] * ci[q])));
                } else {
                    float cr[2] = {0.f, 0.f}, ci[2] = {0.f, 0.f};
                    for (int b = 0; b < set.ulen; ++b) {
                        const float2 xv = xcol[b];
                        const float4 b01 = b4[b];
                        cr[0] = fmaf(b01.x, xv.x, fmaf(-b01.y, xv.y, cr[0])); ci[0] = fmaf(b01.x, xv.y, fmaf(b01.y, xv.x, ci[0]));
                        cr[1] = fmaf(b01.z, xv.x, fmaf(-b01.w, xv.y, cr[1])); ci[1] = fmaf(b01.z, xv.y, fmaf(b01.w, xv.x, ci[1]));
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) asm("sqrt.approx.f32 %0, %1;" : "=f"(mag[q]) : "f"(fmaf(cr[q], cr[q], ci[q] * ci[q])));
                    mag[2] = 0.0f;
                }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q < set.nrows) {
                        const float mv = mag[q] * set.scale[q];
                        sm.mags[col][set.bin[q] % kCqRows] = mv;
                        if (p.cqmag && col < n_valid) p.cqmag[(clip.cq_base + t_block + it0 + col) * kCqBins + set.bin[q]] = mv;
                    }
                }
            }
            __syncthreads();                               // B: every row has read the bins (the transpose regions are free
                                                           // again) and the 16 x 36 magnitudes are complete
            // this octave's share of the chroma fold (filters.cq_to_chroma, 36 bins per octave: chroma c <- bins
            // 3 c - 1, 3 c, 3 c + 1 of the octave, the first wrapping to bin 35); the next iteration's rows write
            // sm.mags only after its barrier A, which these threads reach after the fold
            for (int i = tid; i < 12 * n_valid; i += kCqtWarps * 32) {
                const int g = i / 12, c = i % 12;
                const float* m = sm.mags[g];
                const float sum = (m[(3 * c + kCqRows - 1) % kCqRows] + m[3 * c]) + m[3 * c + 1];
                p.cq_chroma[(static_cast<size_t>(clip.cq_base) + t_block + it0 + g) * (kCqOctaves * 12) + oct_slot * 12 + c] = sum;
            }
        }
    }
}

// ---- chroma fold + tonnetz ---------------------------------------------------------------
// one CTA per (clip, kTonTile columns): partial sums of the six tonnetz rows in float64, then a
// per-clip reduction in tile order (tonnetz_final_kernel), so long clips spread over many CTAs
// and every clip's result is independent of the batch around it
__global__ void __launch_bounds__(128) tonnetz_kernel(CqtParams p) {
    __shared__ __align__(16) float mags[8][kCqOctaves * 12];
    __shared__ double phi[6][12];
    __shared__ double part[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, hl = lane & 15;      // one column per half warp
    const int slot = 2 * warp + half;
    const TonClip clip = p.clips[blockIdx.x];
    const int t_lo = blockIdx.y * kTonTile;
    if (t_lo >= clip.cq_cols) return;
    const int t_hi = min(t_lo + kTonTile, clip.cq_cols);
    if (threadIdx.x < 72) {
        // librosa.feature.tonnetz: phi = R * cos(pi * V), V = outer(scale, 0..11), even rows - 0.5
        const int q = threadIdx.x / 12, c = threadIdx.x % 12;
        const double scale = (q < 2) ? 7.0 / 6.0 : (q < 4) ? 3.0 / 2.0 : 2.0 / 3.0;
        double vv = scale * static_cast<double>(c);
        if ((q & 1) == 0) vv -= 0.5;
        phi[q][c] = ((q < 4) ? 1.0 : 0.5) * cospi(vv);
    }
    __syncthreads();
    // chroma c = sum over the octaves (lowest first) of cqt_kernel's per-octave folds
    double acc = 0.0;   // lanes 0..5 of each half: running sum of tonnetz row `hl` over the half's columns
    for (int tb = t_lo + 2 * warp; tb < t_hi; tb += 8) {
        const int t = tb + half;
        const bool valid = t < t_hi;
        // one row of 7 x 12 chroma shares = 21 16-byte pieces
        const float4* row = reinterpret_cast<const float4*>(p.cq_chroma + (static_cast<size_t>(clip.cq_base) + (valid ? t : tb)) * (kCqOctaves * 12));
#pragma unroll
        for (int i = 0; i < 2; ++i)
            if (hl + 16 * i < kCqOctaves * 3) reinterpret_cast<float4*>(mags[slot])[hl + 16 * i] = row[hl + 16 * i];
        __syncwarp();
        float ch = 0.0f;
        if (hl < 12) {
#pragma unroll
            for (int o = 0; o < kCqOctaves; ++o) ch += mags[slot][12 * o + hl];
        }
        __syncwarp();
        // util.normalize(norm=inf) then util.normalize(norm=1): float32 values, float64 lengths
        float mx = ch;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        // lanes 12..15 of a half hold 0, which never wins a max of non-negative values
        double len_inf = static_cast<double>(mx);
        if (len_inf < static_cast<double>(FLT_MIN)) len_inf = 1.0;
        const float cn = static_cast<float>(static_cast<double>(ch) / len_inf);
        double l1 = static_cast<double>(cn);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        if (l1 < static_cast<double>(FLT_MIN)) l1 = 1.0;
        const float c1 = static_cast<float>(static_cast<double>(cn) / l1);
        double proj = 0.0;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const double cv = static_cast<double>(__shfl_sync(0xffffffffu, c1, 16 * half + c));
            if (hl < 6) proj = fma(phi[hl][c], cv, proj);
        }
        if (valid) acc += proj;
    }
    if (hl < 6) part[slot][hl] = acc;
    __syncthreads();
    if (threadIdx.x < 6) {
        double total = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) total += part[i][threadIdx.x];
        p.ton_part[(static_cast<size_t>(clip.part_base) + blockIdx.y) * 6 + threadIdx.x] = total;
    }
}

__global__ void __launch_bounds__(192) tonnetz_final_kernel(CqtParams p) {
    const int c = blockIdx.x * 32 + threadIdx.x / 6, d = threadIdx.x % 6;
    if (c >= p.n_clips) return;
    const TonClip clip = p.clips[c];
    const int tiles = (clip.cq_cols + kTonTile - 1) / kTonTile;
    double total = 0.0;
    for (int i = 0; i < tiles; ++i) total += p.ton_part[(static_cast<size_t>(clip.part_base) + i) * 6 + d];
    p.out[static_cast<size_t>(clip.out_row) * p.dim + p.off_tonnetz + d] =
        static_cast<float>(total / static_cast<double>(clip.cq_cols));
}

// ---- launchers ---------------------------------------------------------------------------
// per device (called from serb_ctx_create with the device current): opt every instantiation in
// to the largest dynamic shared memory a launch can ask for, once, so that concurrent contexts
// never race an attribute change against a launch
cudaError_t configure_cqt(const float* taps2_scaled, const double* taps2_scaled_f64) {
    cudaError_t e0 = cudaMemcpyToSymbol(c_tap_d, taps2_scaled_f64, sizeof(double) * kDecTaps2);
    if (e0 != cudaSuccess) return e0;
    static float pairs[kTapPairs][2];
    for (int e = 0; e < kTapPairs; ++e) {
        const int k0 = kDecTaps2 + 1 - 2 * e, k1 = kDecTaps2 - 2 * e;          // even / odd sample of the pair
        pairs[e][0] = (e >= 1 && k0 >= 0 && k0 < kDecTaps2) ? taps2_scaled[k0] : 0.0f;
        pairs[e][1] = (e >= 1 && k1 >= 0 && k1 < kDecTaps2) ? taps2_scaled[k1] : 0.0f;
    }
    cudaError_t e = cudaMemcpyToSymbol(c_tap2, pairs, sizeof(pairs));
    if (e != cudaSuccess) return e;
    const int smem = static_cast<int>(kCqtMaxSmem);
    if ((e = cudaFuncSetAttribute(cqt_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqt_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqt_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqt_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqtc_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqtc_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqtc_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(cqtc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(cqt_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

constexpr int kCqtSpanBudget = 8384;    // staged signal floats per CTA (two CTAs per SM)
constexpr int kCqtSharedMaxHop = 16;    // octaves with a hop up to this share the first FFT stage (n_fft 1024 only)

static bool cqt_octave_shared(const CqtParams& p, int octave) {
    const int hop = p.hop0 >> octave;
    return p.n_fft[octave] == 1024 && (hop & 1) == 0 && hop <= kCqtSharedMaxHop && !p.cqt_no_shared;
}

// columns per CTA and dynamic shared memory of one octave's launch
template <int R>
static void cqt_octave_shape(const CqtParams& p, int octave, int& cols_per_block, int& sub_cols, size_t& bytes) {
    constexpr int G = 32 / R;
    constexpr int N = 32 * R;
    const int hop = p.hop0 >> octave;
    const int per_iter = kCqtWarps * G;
    // as many whole iterations as fit the span budget, at least one, at most eight
    int iters = (kCqtSpanBudget - 2 * N + hop) / (per_iter * hop);
    iters = max(1, min(8, iters));
    cols_per_block = per_iter * iters;
    const size_t span = static_cast<size_t>(cols_per_block - 1) * hop + 2 * N;
    bytes = sizeof(CqtSmemHead) + span * sizeof(float);
    if (cqt_octave_shared(p, octave)) {
        // the first-stage table takes what the span leaves of the budget: R rows of
        // ((sub - 1) hop / 2 + 32, rounded up to 1 mod 16) float2 for `sub` columns at a time
        const size_t span_bytes = ((span + 3) & ~size_t(3)) * sizeof(float);
        auto table = [&](int sub) { return static_cast<size_t>(R) * (((sub - 1) * (hop / 2) + 32 + 15) / 16 * 16 + 1) * sizeof(float2); };
        int sub = per_iter;
        for (int c = per_iter; c <= cols_per_block; c += per_iter)
            if (span_bytes + table(c) <= kCqtSpanBudget * sizeof(float)) sub = c;
        sub_cols = sub;
        bytes = sizeof(CqtSmemHead) + span_bytes + table(sub);
    }
}

// cqtc_kernel: columns per CTA, columns per first-stage table and dynamic shared memory of one octave
constexpr size_t kCqtcBudget = (227 * 1024 - 2048) / 2;     // per CTA, two CTAs per SM

static bool cqtc_fft_size(int n_fft) { return n_fft == 1024 || n_fft == 512; }
static bool cqtc_octave_shared(const CqtParams& p, int octave) {
    const int hop = p.hop0 >> octave;
    return cqtc_fft_size(p.n_fft[octave]) && (hop & 1) == 0 && hop <= p.cqtc_shared_max_hop && !p.cqt_no_shared;
}

template <int R, bool SHARED>
static void cqtc_octave_shape(const CqtParams& p, int octave, int& cols_per_block, int& sub_cols, size_t& bytes) {
    constexpr int N = 32 * R, per_iter = kCqtWarps * (32 / R);
    const int hop = p.hop0 >> octave;
    const size_t room = kCqtcBudget - sizeof(CqtcHead<R, SHARED>);
    auto span_bytes = [&](int cols) { return ((static_cast<size_t>(cols - 1) * hop + 2 * N + 3) & ~size_t(3)) * sizeof(float); };
    auto table = [&](int sub) { return static_cast<size_t>(R) * (((sub - 1) * (hop / 2) + 32 + 15) / 16 * 16 + 1) * sizeof(float2); };
    for (int iters = 128 / per_iter; iters >= 1; --iters) {
        const int cols = per_iter * iters;
        if (!SHARED) {
            if (span_bytes(cols) <= room || iters == 1) {
                cols_per_block = cols; sub_cols = cols; bytes = sizeof(CqtcHead<R, SHARED>) + span_bytes(cols);
                return;
            }
            continue;
        }
        int sub = 0;
        for (int c = per_iter; c <= cols; c += per_iter)
            if (span_bytes(cols) + table(c) <= room) sub = c;
        if (sub > 0 || iters == 1) {
            if (sub == 0) sub = per_iter;
            cols_per_block = cols; sub_cols = sub; bytes = sizeof(CqtcHead<R, SHARED>) + span_bytes(cols) + table(sub);
            return;
        }
    }
}

template <int R, bool SHARED>
static cudaError_t launch_cqtc_group(CqtParams p, int first, int count, cudaStream_t stream) {
    size_t max_bytes = 0;
    int blocks = 0;
    for (int o = first; o < first + count; ++o) {
        size_t bytes;
        cqtc_octave_shape<R, SHARED>(p, o, p.cq_cols_per_block[o], p.cq_sub_cols[o], bytes);
        max_bytes = max(max_bytes, bytes);
        blocks += (p.max_cq_cols + p.cq_cols_per_block[o] - 1) / p.cq_cols_per_block[o];
        p.cq_block_end[o - first] = blocks;
    }
    for (int i = count; i < kCqOctaves; ++i) p.cq_block_end[i] = 0x7fffffff;
    if (blocks > 65535 && count > 1) {
        // hours-long clips: one octave per launch keeps grid.y inside its limit
        for (int o = first; o < first + count; ++o) {
            const cudaError_t e = launch_cqtc_group<R, SHARED>(p, o, 1, stream);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    if (max_bytes > kCqtMaxSmem || blocks > 65535) return cudaErrorInvalidConfiguration;
    dim3 grid(p.n_clips, blocks, 1);
    cqtc_kernel<R, SHARED><<<grid, kCqtWarps * 32, max_bytes, stream>>>(p, first);
    return cudaGetLastError();
}

// octaves [first, first + count) share the FFT size 64 R and the kernel variant: one launch,
// blockIdx.z = octave - first
template <int R, bool SHARED>
static cudaError_t launch_cqt_group(CqtParams p, int first, int count, cudaStream_t stream) {
    size_t max_bytes = 0;
    int min_cols = 1 << 30;
    for (int o = first; o < first + count; ++o) {
        size_t bytes;
        cqt_octave_shape<R>(p, o, p.cq_cols_per_block[o], p.cq_sub_cols[o], bytes);
        max_bytes = max(max_bytes, bytes);
        min_cols = min(min_cols, p.cq_cols_per_block[o]);
    }
    if (max_bytes > kCqtMaxSmem) return cudaErrorInvalidConfiguration;
    dim3 grid(p.n_clips, (p.max_cq_cols + min_cols - 1) / min_cols, count);
    cqt_kernel<R, SHARED><<<grid, kCqtWarps * 32, max_bytes, stream>>>(p, first);
    return cudaGetLastError();
}

// one factor-2 stage: the tensor-core kernel for the long clips (when its operand table is there)
// plus the float64 path for the clips below kDecExactBelow samples, else the FFMA2 kernel for all
static cudaError_t launch_decimate2(const CqtParams& p, int src_level, int max_len_in, cudaStream_t stream, long long* n) {
    const int max_out = (max_len_in + 1) >> 1;
    const dim3 grid(p.n_clips, min(65535, (max_out + kDecTile - 1) / kDecTile));
    if (p.dec_toeplitz == nullptr) {
        decimate2_kernel<<<grid, kDecThreads, 0, stream>>>(p, src_level, 0);
        ++*n;
        return cudaGetLastError();
    }
    if (p.n_dec_exact < p.n_clips) {
        cudaError_t e = launch_decimate2_mma(p, src_level, max_len_in, p.dec_toeplitz, p.n_sms, stream);
        if (e != cudaSuccess) return e;
        ++*n;
    }
    if (p.n_dec_exact > 0) {
        // the short clips only: their levels are at most kDecExactBelow / 2 outputs long
        const dim3 small(p.n_clips, min(static_cast<int>(grid.y), (kDecExactBelow / 2 + kDecThreads - 1) / kDecThreads));
        decimate2_kernel<<<small, kDecThreads, 0, stream>>>(p, src_level, 1);
        ++*n;
    }
    return cudaGetLastError();
}

cudaError_t launch_decimations(const CqtParams& p, cudaStream_t stream, long long* launches) {
    if (p.n_clips <= 0) return cudaSuccess;
    long long n = 0;
    cudaError_t e;
    if (p.early_factor == 2) {
        if ((e = launch_decimate2(p, -1, p.max_length, stream, &n)) != cudaSuccess) return e;
    } else if (p.early_factor > 2) {
        const int tiles = min(4096, (p.max_len0 + 255) / 256);
        decimate_any_kernel<<<dim3(p.n_clips, tiles), 256, 0, stream>>>(p);
        ++n;
    }
    int len = p.max_len0;
    for (int level = 0; level + 1 < kCqOctaves; ++level) {
        if ((e = launch_decimate2(p, level, len, stream, &n)) != cudaSuccess) return e;
        len = (len + 1) >> 1;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

cudaError_t launch_cqt_octaves(const CqtParams& p, cudaStream_t stream, long long* launches) {
    if (p.n_clips <= 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    long long n = 0;
    // maximal runs of octaves with one FFT size and one kernel variant (at the common sample rates:
    // the top octaves with per-column transforms, the bottom ones sharing the first stage)
    for (int first = 0; first < kCqOctaves;) {
        const bool cols = p.set_banks != nullptr && cqtc_fft_size(p.n_fft[first]);      // lane = column rows (cqtc_kernel)
        const bool shared = cols ? cqtc_octave_shared(p, first) : cqt_octave_shared(p, first);
        int count = 1;
        while (first + count < kCqOctaves && p.n_fft[first + count] == p.n_fft[first] &&
               (cols ? cqtc_octave_shared(p, first + count) : cqt_octave_shared(p, first + count)) == shared)
            ++count;
        if (cols) {
            if (p.n_fft[first] == 1024)
                e = shared ? launch_cqtc_group<16, true>(p, first, count, stream) : launch_cqtc_group<16, false>(p, first, count, stream);
            else
                e = shared ? launch_cqtc_group<8, true>(p, first, count, stream) : launch_cqtc_group<8, false>(p, first, count, stream);
            if (e != cudaSuccess) return e;
            ++n;
            first += count;
            continue;
        }
        switch (p.n_fft[first]) {
            case 256: e = launch_cqt_group<4, false>(p, first, count, stream); break;
            case 512: e = launch_cqt_group<8, false>(p, first, count, stream); break;
            case 1024: e = shared ? launch_cqt_group<16, true>(p, first, count, stream)
                                  : launch_cqt_group<16, false>(p, first, count, stream); break;
            case 2048: e = launch_cqt_group<32, false>(p, first, count, stream); break;
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        ++n;
        first += count;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

cudaError_t launch_tonnetz(const CqtParams& p, cudaStream_t stream) {
    if (p.n_clips <= 0) return cudaSuccess;
    tonnetz_kernel<<<dim3(p.n_clips, (p.max_cq_cols + kTonTile - 1) / kTonTile), 128, 0, stream>>>(p);
    tonnetz_final_kernel<<<(p.n_clips + 31) / 32, 192, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace serb
