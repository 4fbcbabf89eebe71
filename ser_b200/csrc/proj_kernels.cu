// K2 tuning estimation, K3 mel + chroma projections, K4 per-clip pooling + DCT.
//
// Reference arithmetic replaced (librosa 0.11.0 via ser/_internal/utils/dsp.py):
//   K2  librosa.estimate_tuning / pitch_tuning                dsp.py:113-118 (inside chroma_stft)
//   K3  librosa.feature.melspectrogram + power_to_db          dsp.py:106-111, 120-125
//       librosa.feature.chroma_stft (filterbank product, L-inf column normalisation)
//   K4  np.mean(..., axis=1) per group, top_db clip against the clip-global maximum,
//       scipy.fftpack.dct(type=2, norm="ortho")[:40]           dsp.py:107, 114, 121
// (SURVEY.md Appendix A.2-A.6).
#include "common.cuh"
#include "kernels.h"
#include <cfloat>

namespace serb {

// =========================================================================================
// K2: tuning.  One CTA per clip.  Exact median of the peak magnitudes by 4-pass radix
// select on the float bit patterns, then the 100-bin residual histogram of the peaks at or
// above the median and its first arg-max.
// =========================================================================================
constexpr int kTuneThreads = 256;


__device__ __forceinline__ unsigned key_of(float v) {
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float value_of(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kTuneCache = 6144;    // peak magnitudes cached in shared memory (24 KB: eight CTAs per SM)

// The two middle order statistics (0-based ranks rank_lo <= rank_hi <= rank_lo + 1) in one 4-pass
// radix select; every thread returns the same pair.  While both ranks still fall in the same
// bucket one histogram serves both.  Keys come from the shared-memory cache when the clip's peaks
// fit, else from the global peak lists.
__device__ float2 select_middle(const TuneParams& p, const ClipDev& clip, long long rank_lo, long long rank_hi,
                                unsigned (*hist)[256], unsigned* shared_prefix, long long* shared_rank,
                                const float* cache, int n_cached) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = kTuneThreads / 32;
    unsigned prefix0 = 0, prefix1 = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const unsigned hi_mask = (pass == 0) ? 0u : (0xffffffffu << (shift + 8));
        const bool same = prefix0 == prefix1;
        for (int i = threadIdx.x; i < 512; i += kTuneThreads) (&hist[0][0])[i] = 0;
        __syncthreads();
        auto count = [&](unsigned key) {
            const unsigned b = (key >> shift) & 255u;
            if ((key & hi_mask) == (prefix0 & hi_mask)) atomicAdd(&hist[0][b], 1u);
            if (!same && (key & hi_mask) == (prefix1 & hi_mask)) atomicAdd(&hist[1][b], 1u);
        };
        if (cache) {
            for (int i = threadIdx.x; i < n_cached; i += kTuneThreads) count(key_of(cache[i]));
        } else {
            for (int t = warp; t < clip.n_cols; t += n_warps) {
                const long long col = static_cast<long long>(clip.col_base) + t;
                const int cnt = p.peak_count[col];
                const float2* src = p.peaks + col * p.peak_cap;
                for (int i = lane; i < cnt; i += 32) count(key_of(src[i].x));
            }
        }
        __syncthreads();
        if (warp < 2) {
            // warp j finds the bucket of rank j: eight buckets per lane, a warp scan of the lane
            // sums, then the owning lane walks its eight (the smallest b with r < h[0] + .. + h[b])
            const int j = warp;
            const unsigned* h = hist[(j == 1 && !same) ? 1 : 0];
            const long long r = (j == 0) ? rank_lo : rank_hi;
            unsigned c[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[k] = h[8 * lane + k]; sum += c[k]; }
            unsigned incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, r < static_cast<long long>(incl));
            const int owner = ballot ? __ffs(ballot) - 1 : 31;
            if (lane == owner) {
                long long rr = r - static_cast<long long>(incl - sum);
                unsigned b = 0;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    if (b == static_cast<unsigned>(k) && rr >= static_cast<long long>(c[k])) { rr -= c[k]; ++b; }
                }
                shared_prefix[j] = ((j == 0) ? prefix0 : prefix1) | ((8u * lane + b) << shift);
                shared_rank[j] = rr;
            }
        }
        __syncthreads();
        prefix0 = shared_prefix[0];
        prefix1 = shared_prefix[1];
        rank_lo = shared_rank[0];
        rank_hi = shared_rank[1];
        __syncthreads();
    }
    return make_float2(value_of(prefix0), value_of(prefix1));
}

__global__ void __launch_bounds__(kTuneThreads, 8) tuning_kernel(TuneParams p) {
    __shared__ float cache[kTuneCache];
    __shared__ unsigned hist[2][256];
    __shared__ unsigned s_prefix[2];
    __shared__ long long s_rank[2];
    __shared__ long long s_total;
    __shared__ int s_cursor;
    __shared__ int counts[100];
    const ClipDev clip = p.clips[blockIdx.x];
    if (clip.n_cols > kTuneLongCols) return;      // handled by the multi-CTA path below
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = kTuneThreads / 32;

    if (threadIdx.x == 0) { s_total = 0; s_cursor = 0; }
    for (int i = threadIdx.x; i < 100; i += kTuneThreads) counts[i] = 0;
    __syncthreads();
    long long local = 0;
    for (int t = threadIdx.x; t < clip.n_cols; t += kTuneThreads)
        local += p.peak_count[static_cast<long long>(clip.col_base) + t];
    if (local) atomicAdd(reinterpret_cast<unsigned long long*>(&s_total), static_cast<unsigned long long>(local));
    __syncthreads();
    const long long n = s_total;
    if (n == 0) {
        // pitch_tuning on an empty set returns 0.0 == np.linspace(-0.5, 0.5, 101)[50]
        if (threadIdx.x == 0) p.tuning_idx[blockIdx.x] = 50;
        return;
    }
    const bool cached = n <= kTuneCache;
    if (cached) {
        // gather the magnitudes once; their order is irrelevant to the order statistics
        for (int t = warp; t < clip.n_cols; t += n_warps) {
            const long long col = static_cast<long long>(clip.col_base) + t;
            const int cnt = p.peak_count[col];
            if (cnt == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_cursor, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            const float2* src = p.peaks + col * p.peak_cap;
            for (int i = lane; i < cnt; i += 32) cache[base + i] = src[i].x;
        }
        __syncthreads();
    }
    const float* keys = cached ? cache : nullptr;
    const int n_cached = static_cast<int>(cached ? n : 0);
    // np.median: mean of the two middle order statistics (float32 arithmetic) when n is even
    const float2 mid = select_middle(p, clip, (n - 1) / 2, n / 2, hist, s_prefix, s_rank, keys, n_cached);
    const float thr = (n & 1) ? mid.x : __fmul_rn(__fadd_rn(mid.x, mid.y), 0.5f);
    const float bpo = static_cast<float>(p.bins_per_octave);
    for (int t = warp; t < clip.n_cols; t += n_warps) {
        const long long col = static_cast<long long>(clip.col_base) + t;
        const int cnt = p.peak_count[col];
        const float2* src = p.peaks + col * p.peak_cap;
        for (int i = lane; i < cnt; i += 32) {
            const float2 pk = src[i];
            if (!(pk.x >= thr) || !(pk.y > 0.f)) continue;
            // residual = mod(bpo * log2(f / 27.5), 1.0) in float32, folded to [-0.5, 0.5)
            const float q = __fdiv_rn(pk.y, 27.5f);
            const float l2 = static_cast<float>(log2(static_cast<double>(q)));
            float r = fmodf(__fmul_rn(bpo, l2), 1.0f);
            if (r < 0.f) r += 1.0f;
            if (r >= 0.5f) r = __fsub_rn(r, 1.0f);
            // np.histogram with explicit float64 edges: edges[i] <= r < edges[i+1]
            const double rd = static_cast<double>(r);
            int bin = static_cast<int>(floor((rd + 0.5) * 100.0));
            bin = max(0, min(99, bin));
            while (bin > 0 && rd < p.edges[bin]) --bin;
            while (bin < 99 && rd >= p.edges[bin + 1]) ++bin;
            atomicAdd(&counts[bin], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0, best_count = counts[0];
        for (int i = 1; i < 100; ++i)
            if (counts[i] > best_count) { best_count = counts[i]; best = i; }
        p.tuning_idx[blockIdx.x] = best;
    }
}

// ---- long clips: the same selection spread over many CTAs ---------------------------------
// A clip with more than kTuneLongCols columns would keep one CTA busy for milliseconds, so its
// order statistics come from a 3-pass (11 + 11 + 10 bit) radix select whose histograms are built
// by one CTA per 64 columns, followed by a residual histogram built the same way.  Both middle
// ranks are tracked at once.  Integer counting: the result is identical to the one-CTA kernel.
constexpr int kTuneTileCols = 64;
constexpr int kTuneBuckets = 2048;

__device__ __forceinline__ void tune_pass_bits(int pass, int& shift, unsigned& mask, unsigned& hi_mask) {
    shift = (pass == 0) ? 21 : (pass == 1) ? 10 : 0;
    mask = (pass == 2) ? 1023u : 2047u;
    hi_mask = (pass == 0) ? 0u : (0xffffffffu << (pass == 1 ? 21 : 10));
}

__global__ void __launch_bounds__(256) tune_long_hist_kernel(TuneParams p, int pass) {
    __shared__ unsigned hist[2][kTuneBuckets];
    const int li = blockIdx.x;
    const ClipDev clip = p.clips[p.long_clips[li]];
    const int t_lo = blockIdx.y * kTuneTileCols;
    if (t_lo >= clip.n_cols) return;
    const int t_hi = min(t_lo + kTuneTileCols, clip.n_cols);
    TuneLongState* st = p.long_state + li;
    int shift; unsigned mask, hi_mask;
    tune_pass_bits(pass, shift, mask, hi_mask);
    const unsigned pre0 = st->prefix[0], pre1 = st->prefix[1];
    for (int i = threadIdx.x; i < 2 * kTuneBuckets; i += 256) (&hist[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = t_lo + warp; t < t_hi; t += 8) {
        const long long col = static_cast<long long>(clip.col_base) + t;
        const int cnt = p.peak_count[col];
        const float2* src = p.peaks + col * p.peak_cap;
        for (int i = lane; i < cnt; i += 32) {
            const unsigned key = key_of(src[i].x);
            const unsigned b = (key >> shift) & mask;
            if (((key ^ pre0) & hi_mask) == 0) atomicAdd(&hist[0][b], 1u);
            if (((key ^ pre1) & hi_mask) == 0) atomicAdd(&hist[1][b], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * kTuneBuckets; i += 256) {
        const unsigned c = (&hist[0][0])[i];
        if (c) atomicAdd(&st->hist[0][0] + i, c);
    }
}

// one CTA per long clip: locate both ranks in the pass histograms, extend the prefixes, clear
__global__ void __launch_bounds__(256) tune_long_pick_kernel(TuneParams p, int pass) {
    TuneLongState* st = p.long_state + blockIdx.x;
    int shift; unsigned mask, hi_mask;
    tune_pass_bits(pass, shift, mask, hi_mask);
    if (threadIdx.x < 2) {
        const int j = threadIdx.x;
        if (pass == 0) {
            long long n = 0;
            for (int b = 0; b < kTuneBuckets; ++b) n += st->hist[j][b];
            st->n = n;
            st->rank[j] = (j == 0) ? (n - 1) / 2 : n / 2;     // the two middle order statistics
        }
        long long r = st->rank[j];
        unsigned b = 0;
        if (st->n > 0) {
            for (; b < mask; ++b) {
                const long long c = st->hist[j][b];
                if (r < c) break;
                r -= c;
            }
        }
        st->prefix[j] |= b << shift;
        st->rank[j] = r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * kTuneBuckets; i += 256) (&st->hist[0][0])[i] = 0;
    if (pass == 2 && threadIdx.x < 100) st->counts[threadIdx.x] = 0;
}

__device__ __forceinline__ int residual_bin(float pitch, float bpo, const double* __restrict__ edges) {
    // residual = mod(bpo * log2(f / 27.5), 1.0) in float32, folded to [-0.5, 0.5)
    const float q = __fdiv_rn(pitch, 27.5f);
    const float l2 = static_cast<float>(log2(static_cast<double>(q)));
    float r = fmodf(__fmul_rn(bpo, l2), 1.0f);
    if (r < 0.f) r += 1.0f;
    if (r >= 0.5f) r = __fsub_rn(r, 1.0f);
    // np.histogram with explicit float64 edges: edges[i] <= r < edges[i+1]
    const double rd = static_cast<double>(r);
    int bin = static_cast<int>(floor((rd + 0.5) * 100.0));
    bin = max(0, min(99, bin));
    while (bin > 0 && rd < edges[bin]) --bin;
    while (bin < 99 && rd >= edges[bin + 1]) ++bin;
    return bin;
}

__global__ void __launch_bounds__(256) tune_long_resid_kernel(TuneParams p) {
    __shared__ int counts[100];
    const int li = blockIdx.x;
    const ClipDev clip = p.clips[p.long_clips[li]];
    const int t_lo = blockIdx.y * kTuneTileCols;
    if (t_lo >= clip.n_cols) return;
    const int t_hi = min(t_lo + kTuneTileCols, clip.n_cols);
    TuneLongState* st = p.long_state + li;
    if (st->n == 0) return;
    // np.median: mean of the two middle order statistics in float32 (equal when n is odd)
    const float thr = __fmul_rn(__fadd_rn(value_of(st->prefix[0]), value_of(st->prefix[1])), 0.5f);
    if (threadIdx.x < 100) counts[threadIdx.x] = 0;
    __syncthreads();
    const float bpo = static_cast<float>(p.bins_per_octave);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = t_lo + warp; t < t_hi; t += 8) {
        const long long col = static_cast<long long>(clip.col_base) + t;
        const int cnt = p.peak_count[col];
        const float2* src = p.peaks + col * p.peak_cap;
        for (int i = lane; i < cnt; i += 32) {
            const float2 pk = src[i];
            if (!(pk.x >= thr) || !(pk.y > 0.f)) continue;
            atomicAdd(&counts[residual_bin(pk.y, bpo, p.edges)], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 100 && counts[threadIdx.x]) atomicAdd(&st->counts[threadIdx.x], counts[threadIdx.x]);
}

__global__ void tune_long_argmax_kernel(TuneParams p) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= p.n_long) return;
    TuneLongState* st = p.long_state + li;
    int best = 50;      // pitch_tuning on an empty set returns 0.0 == np.linspace(-0.5, 0.5, 101)[50]
    if (st->n > 0) {
        best = 0;
        int best_count = st->counts[0];
        for (int i = 1; i < 100; ++i)
            if (st->counts[i] > best_count) { best_count = st->counts[i]; best = i; }
    }
    p.tuning_idx[p.long_clips[li]] = best;
    st->prefix[0] = st->prefix[1] = 0;     // ready for the next use of the slot
}

cudaError_t launch_tuning(const TuneParams& p, int n_clips, cudaStream_t stream, long long* launches) {
    if (n_clips <= 0) return cudaSuccess;
    tuning_kernel<<<n_clips, kTuneThreads, 0, stream>>>(p);
    long long n = 1;
    if (p.n_long > 0) {
        const dim3 grid(p.n_long, (p.max_long_cols + kTuneTileCols - 1) / kTuneTileCols);
        for (int pass = 0; pass < 3; ++pass) {
            tune_long_hist_kernel<<<grid, 256, 0, stream>>>(p, pass);
            tune_long_pick_kernel<<<p.n_long, 256, 0, stream>>>(p, pass);
        }
        tune_long_resid_kernel<<<grid, 256, 0, stream>>>(p);
        tune_long_argmax_kernel<<<(p.n_long + 127) / 128, 128, 0, stream>>>(p);
        n += 8;
    }
    if (launches) *launches += n;
    return cudaGetLastError();
}

// =========================================================================================
// K3: mel (sparse Slaney triangles) and chroma (dense 12 x 1025, bank picked by the clip's
// tuning) projections as shared-memory fp32 contractions.
//
// Persistent, warp-specialised kernel: one CTA of 16 warps per SM walks a contiguous range of
// 16-column tiles.  Thread 0 stages each tile's |X| rows with TMA bulk copies (one per column,
// cp.async.bulk -> UBLKCP, mbarrier completion) and reloads the clip's chroma bank into shared
// memory when the tuning changes.  All warps transpose the staged rows to [bin][column]
// (pitch 20 floats, so 16-byte loads over 4 columns are bank-conflict free); the copy of the
// NEXT tile is issued right after the transpose and lands while the products run.  Warps 0-7
// then do the chroma product (split-K over 20 bin slices, 4 chroma x 4 columns register tile,
// deterministic slice reduction, per-column L-inf normalisation) while warps 8-15 do mel
// (thread = (column, 8 bands); the 16 lanes of a half warp read one 64-byte bin row, and bands
// are dealt short-with-long so every thread sums about the same number of non-zeros),
// power_to_db and the log-mel write-out ([tile][band][16 columns]).
// =========================================================================================
constexpr int kProjWarps = 16;
constexpr int kProjThreads = kProjWarps * 32;
constexpr int kChromaThreads = 256;                 // warps 0..7 (240 active): 20 bin slices x 3 chroma groups x 4 column groups
constexpr int kMelThreads = kProjThreads - kChromaThreads;   // warps 8..15: 16 columns x 16 band sets
constexpr int kChromaSlices = 20;
constexpr int kSliceBins = 52;                      // 20 * 52 = 1040 >= 1025
constexpr int kRowPitch = 1028;   // floats per staged column row (16-byte multiple)
constexpr int kTPitch = 20;       // floats per transposed bin row: 16 columns + 4 pad

struct ProjSmem {
    float raw[kColsPerTile][kRowPitch];    // TMA landing zone: |X| rows of one tile, 65792 B
    float t[(kNBins + 3) * kTPitch];       // the same tile as [bin][column] + 3 zero rows, 82240 B
    float w[kNBins * 12];                  // chroma bank of the current clip, [bin][12], 49200 B
    float red[kChromaSlices * 192];        // chroma split-K partials, 15360 B
    float chr[192];
    float melw[2560];                      // sparse mel weights, 4-padded per band (<= 2560)
    int mstart[128], mcount[128], moffset[129];
    float wmax[8];
    unsigned long long bar_tile, bar_bank;
};

__device__ __forceinline__ uint32_t proj_smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void proj_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// thread 0: start the bulk copies of one tile's existing column rows
__device__ __forceinline__ void proj_issue_tile(ProjSmem& sm, const ProjParams& p, int tile, const ClipDev& clip) {
    const int t0 = (tile - clip.tile_base) * kColsPerTile;
    const int n_valid = min(kColsPerTile, clip.n_cols - t0);
    const long long col0 = static_cast<long long>(clip.col_base) + t0;
    const uint32_t bar = proj_smem_u32(&sm.bar_tile);
    constexpr uint32_t row_bytes = kRowPitch * sizeof(float);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(row_bytes * n_valid) : "memory");
    for (int j = 0; j < n_valid; ++j) {
        const float* src = p.spill + (col0 + j) * kSpillStride;
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(proj_smem_u32(&sm.raw[j][0])), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
    }
}

__global__ void __launch_bounds__(kProjThreads, 1) proj_kernel(ProjParams p, int n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ProjSmem& sm = *reinterpret_cast<ProjSmem*>(smem_raw);
    const int tid = threadIdx.x;
    // contiguous tile range of this CTA (consecutive tiles mostly share a clip and hence a bank)
    const int tile_lo = static_cast<int>(static_cast<long long>(n_tiles) * blockIdx.x / gridDim.x);
    const int tile_hi = static_cast<int>(static_cast<long long>(n_tiles) * (blockIdx.x + 1) / gridDim.x);
    if (tile_lo >= tile_hi) return;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(proj_smem_u32(&sm.bar_tile)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(proj_smem_u32(&sm.bar_bank)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (p.do_mel) {
        for (int i = tid; i < p.mel_nnz; i += kProjThreads) sm.melw[i] = p.mel_weights[i];
        for (int i = tid; i < 128; i += kProjThreads) { sm.mstart[i] = p.mel_start[i]; sm.mcount[i] = p.mel_count[i]; }
        for (int i = tid; i < 129; i += kProjThreads) sm.moffset[i] = p.mel_offset[i];
    }
    for (int i = tid; i < 3 * kTPitch; i += kProjThreads) sm.t[kNBins * kTPitch + i] = 0.0f;   // padding rows
    __syncthreads();
    // tile -> clip -> (descriptor, tuning) is a chain of three dependent global loads (about 15 % of
    // a tile's 12 k cycles when every tile waits for it), so it runs two tiles ahead: the clip index
    // of tile + 2 and the descriptor and tuning of tile + 1 are requested at the top of a tile and
    // consumed at its end
    int ci = p.tile_clip[tile_lo];
    ClipDev clip = p.clips[ci];
    int want_bank = p.do_chroma ? p.tuning_idx[ci] : -1;
    int ci_next = (tile_lo + 1 < tile_hi) ? p.tile_clip[tile_lo + 1] : ci;
    if (tid == 0) proj_issue_tile(sm, p, tile_lo, clip);

    int bank_loaded = -1;     // tuning index whose bank sits in sm.w (uniform across the CTA)
    int bank_phase = 0;
    for (int tile = tile_lo, it = 0; tile < tile_hi; ++tile, ++it) {
        const int ci_next2 = (tile + 2 < tile_hi) ? p.tile_clip[tile + 2] : ci_next;
        const ClipDev clip_next = p.clips[ci_next];
        const int bank_next = p.do_chroma ? p.tuning_idx[ci_next] : -1;
        const int t0 = (tile - clip.tile_base) * kColsPerTile;
        const int n_valid = min(kColsPerTile, clip.n_cols - t0);
        const bool new_bank = want_bank != bank_loaded;

        if (tid == 0 && new_bank) {
            // sm.w was released by the barrier that ended the previous tile
            constexpr uint32_t bank_bytes = kNBins * 12 * sizeof(float);
            const uint32_t bb = proj_smem_u32(&sm.bar_bank);
            const float* src = p.chroma_banks + static_cast<size_t>(want_bank) * (kNBins * 12);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(bank_bytes) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(proj_smem_u32(&sm.w[0])), "l"(src), "r"(bank_bytes), "r"(bb) : "memory");
        }
        proj_mbar_wait(proj_smem_u32(&sm.bar_tile), it & 1);

        // ---- transpose raw[column][bin] -> t[bin][column]: conflict-free loads and 16-byte stores ----
        {
            // thread = (bin lane, column quad); 4 quads x 128 bin lanes
            const int q = tid >> 7;
            const float* r0 = &sm.raw[4 * q][0];
            const bool v0 = 4 * q + 0 < n_valid, v1 = 4 * q + 1 < n_valid, v2 = 4 * q + 2 < n_valid, v3 = 4 * q + 3 < n_valid;
            for (int f = tid & 127; f < kNBins; f += 128) {
                float4 v;
                v.x = v0 ? r0[f] : 0.0f;
                v.y = v1 ? r0[kRowPitch + f] : 0.0f;
                v.z = v2 ? r0[2 * kRowPitch + f] : 0.0f;
                v.w = v3 ? r0[3 * kRowPitch + f] : 0.0f;
                *reinterpret_cast<float4*>(&sm.t[f * kTPitch + 4 * q]) = v;
            }
        }
        __syncthreads();
        if (tid == 0 && tile + 1 < tile_hi) proj_issue_tile(sm, p, tile + 1, clip_next);   // lands during the products

        if (tid < kChromaThreads) {
            // ================= chroma: raw[c][t] = sum_f W[c][f] |X|[f][t] =================
            if (p.do_chroma) {
                if (new_bank) proj_mbar_wait(proj_smem_u32(&sm.bar_bank), bank_phase & 1);
                if (tid < 240) {
                    const int cg = tid % 3, tg = (tid / 3) & 3, ks = tid / 12;
                    float acc[4][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
                    const int f_lo = ks * kSliceBins, f_hi = min(kNBins, f_lo + kSliceBins);
                    const float* xs = &sm.t[4 * tg];
                    const float* ws = &sm.w[4 * cg];
#pragma unroll 4
                    for (int f = f_lo; f < f_hi; ++f) {
                        const float4 w = *reinterpret_cast<const float4*>(ws + f * 12);
                        const float4 x = *reinterpret_cast<const float4*>(xs + f * kTPitch);
                        acc[0][0] = fmaf(w.x, x.x, acc[0][0]); acc[0][1] = fmaf(w.x, x.y, acc[0][1]);
                        acc[0][2] = fmaf(w.x, x.z, acc[0][2]); acc[0][3] = fmaf(w.x, x.w, acc[0][3]);
                        acc[1][0] = fmaf(w.y, x.x, acc[1][0]); acc[1][1] = fmaf(w.y, x.y, acc[1][1]);
                        acc[1][2] = fmaf(w.y, x.z, acc[1][2]); acc[1][3] = fmaf(w.y, x.w, acc[1][3]);
                        acc[2][0] = fmaf(w.z, x.x, acc[2][0]); acc[2][1] = fmaf(w.z, x.y, acc[2][1]);
                        acc[2][2] = fmaf(w.z, x.z, acc[2][2]); acc[2][3] = fmaf(w.z, x.w, acc[2][3]);
                        acc[3][0] = fmaf(w.w, x.x, acc[3][0]); acc[3][1] = fmaf(w.w, x.y, acc[3][1]);
                        acc[3][2] = fmaf(w.w, x.z, acc[3][2]); acc[3][3] = fmaf(w.w, x.w, acc[3][3]);
                    }
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) sm.red[(ks * 12 + 4 * cg + a) * 16 + 4 * tg + b] = acc[a][b];
                }
                named_barrier(1, kChromaThreads);
                float mine = 0.f;
                if (tid < 192) {
#pragma unroll
                    for (int k = 0; k < kChromaSlices; ++k) mine += sm.red[k * 192 + tid];
                    sm.chr[tid] = mine;  // [c][t]
                }
                named_barrier(1, kChromaThreads);
                // util.normalize(norm=inf, axis=-2): divide each column by its maximum (float64 quotient);
                // one division per thread, then the columns of the tile are summed in order
                if (tid < 192) {
                    const int t = tid & 15;
                    float length = 0.f;
#pragma unroll
                    for (int c = 0; c < 12; ++c) length = fmaxf(length, fabsf(sm.chr[c * 16 + t]));
                    const double len = (length < FLT_MIN) ? 1.0 : static_cast<double>(length);
                    float v = static_cast<float>(static_cast<double>(mine) / len);
                    if (t >= n_valid) v = 0.0f;
                    // sum over the 16 columns held by 16 consecutive lanes, in column order
                    float total = 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) total += __shfl_sync(0xffffffffu, v, (threadIdx.x & 16) + j);
                    if (t == 0) p.tile_chroma[static_cast<long long>(tile) * 12 + (tid >> 4)] = total;
                }
            }
        } else if (p.do_mel) {
            // ================= mel power + log-mel =================
            const int mt = tid - kChromaThreads;
            const int col = mt & 15, bs = mt >> 4;      // column, band set (two band sets per warp)
            const bool col_ok = col < n_valid;
            float* lm_tile = p.logmel + static_cast<long long>(tile) * (128 * kColsPerTile);
            float lmax = -FLT_MAX;
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
                // bands dealt short-with-long: {bs, 31-bs, 32+bs, 63-bs, 64+bs, 95-bs, 96+bs, 127-bs}
                const int m = (j & 1) ? (32 * (j >> 1) + 31 - bs) : (32 * (j >> 1) + bs);
                const float4* w = reinterpret_cast<const float4*>(sm.melw + sm.moffset[m]);
                const float4* w_end = w + (sm.mcount[m] >> 2);     // counts are padded to multiples of 4
                const float* xs = &sm.t[sm.mstart[m] * kTPitch + col];
                // four independent partial sums (bins i mod 4) so the FMAs pipeline, combined as
                // (a0 + a1) + (a2 + a3); the next group's operands are fetched before the current
                // group is consumed
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (w < w_end) {
                    float4 wi = *w;
                    float x0 = xs[0], x1 = xs[kTPitch], x2 = xs[2 * kTPitch], x3 = xs[3 * kTPitch];
                    for (++w, xs += 4 * kTPitch; w < w_end; ++w, xs += 4 * kTPitch) {
                        const float4 wn = *w;
                        const float y0 = xs[0], y1 = xs[kTPitch], y2 = xs[2 * kTPitch], y3 = xs[3 * kTPitch];
                        // power = |X| * |X| in float32 (np.abs(D) ** 2.0)
                        a0 = fmaf(wi.x, x0 * x0, a0); a1 = fmaf(wi.y, x1 * x1, a1);
                        a2 = fmaf(wi.z, x2 * x2, a2); a3 = fmaf(wi.w, x3 * x3, a3);
                        wi = wn; x0 = y0; x1 = y1; x2 = y2; x3 = y3;
                    }
                    a0 = fmaf(wi.x, x0 * x0, a0); a1 = fmaf(wi.y, x1 * x1, a1);
                    a2 = fmaf(wi.z, x2 * x2, a2); a3 = fmaf(wi.w, x3 * x3, a3);
                }
                const float acc = (a0 + a1) + (a2 + a3);
                // power_to_db(ref=1, amin=1e-10): 10 * log10(max(1e-10, S)); log10 via the MUFU log2
                // (absolute error < 1e-6 dB, far below float32 resolution at these magnitudes)
                const float lmv = 3.01029995663981195f * __log2f(fmaxf(1e-10f, acc));
                lm_tile[m * kColsPerTile + col] = lmv;
                if (col_ok) lmax = fmaxf(lmax, lmv);
                // tile sum over the 16 columns (half warp), fixed tree order; missing columns hold 0
                float tsum = acc;
                tsum += __shfl_xor_sync(0xffffffffu, tsum, 8);
                tsum += __shfl_xor_sync(0xffffffffu, tsum, 4);
                tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
                tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
                if (col == 0) p.tile_mel[static_cast<long long>(tile) * 128 + m] = tsum;
            }
            lmax = warp_max(lmax);
            if ((mt & 31) == 0) sm.wmax[mt >> 5] = lmax;
            named_barrier(2, kMelThreads);
            if (mt == 0) {
                float v = sm.wmax[0];
#pragma unroll
                for (int g = 1; g < 8; ++g) v = fmaxf(v, sm.wmax[g]);
                p.tile_lmax[tile] = v;
            }
        }
        if (new_bank) { bank_loaded = want_bank; bank_phase += 1; }
        ci = ci_next; clip = clip_next; want_bank = bank_next; ci_next = ci_next2;
        __syncthreads();   // both groups are done with sm.t (and sm.w)
    }
}

static int g_proj_grid = 0;

cudaError_t configure_proj() {
    cudaError_t e = cudaFuncSetAttribute(proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(sizeof(ProjSmem)));
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    g_proj_grid = sms;
    return cudaSuccess;
}

cudaError_t launch_proj(const ProjParams& p, int n_tiles, cudaStream_t stream) {
    if (n_tiles <= 0) return cudaSuccess;
    const int grid = (n_tiles < g_proj_grid || g_proj_grid <= 0) ? n_tiles : g_proj_grid;
    proj_kernel<<<grid, kProjThreads, sizeof(ProjSmem), stream>>>(p, n_tiles);
    return cudaGetLastError();
}

// =========================================================================================
// K4: per-clip pooling.  mean over columns of every group, MFCC via the DCT of the mean of
// the top_db-clipped log-mel (DCT and mean commute), output assembled in the reference's
// group order.
// =========================================================================================

constexpr int kPoolSlices = 4;
constexpr int kPoolThreads = 128 * kPoolSlices;

__global__ void __launch_bounds__(kPoolThreads) pool_kernel(PoolParams p) {
    __shared__ double part[kPoolSlices][128];
    __shared__ double meanlog[128];
    __shared__ float s_thr;
    const ClipDev clip = p.clips[blockIdx.x];
    const int tid = threadIdx.x;
    const int m = tid & 127, slice = tid >> 7;
    const int n_tiles = (clip.n_cols + kColsPerTile - 1) / kColsPerTile;
    const double inv_t = 1.0 / static_cast<double>(clip.n_cols);
    float* out = p.out + static_cast<long long>(clip.out_row) * p.dim;

    if (p.off_mfcc >= 0) {
        if (tid < 32) {
            float v = -FLT_MAX;
            for (int i = tid; i < n_tiles; i += 32) v = fmaxf(v, p.tile_lmax[clip.tile_base + i]);
            v = warp_max(v);
            // np.maximum(log_spec, log_spec.max() - top_db) with top_db = 80.0, float32
            if (tid == 0) s_thr = __fsub_rn(v, 80.0f);
        }
        __syncthreads();
        const float thr = s_thr;
        // log-mel is stored [tile][band][16 columns]; each slice takes every 4th tile of the clip,
        // sums its columns in order, and the slices are then added in order
        double acc = 0.0;
        const float* src = p.logmel + static_cast<long long>(clip.tile_base) * (128 * kColsPerTile) + m * kColsPerTile;
        for (int tl = slice; tl < n_tiles; tl += kPoolSlices) {
            const float4* row = reinterpret_cast<const float4*>(src + static_cast<long long>(tl) * (128 * kColsPerTile));
            const float4 q0 = row[0], q1 = row[1], q2 = row[2], q3 = row[3];
            const float v[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w,
                                 q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
            const int n_here = min(kColsPerTile, clip.n_cols - tl * kColsPerTile);
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < n_here) acc += static_cast<double>(fmaxf(v[j], thr));
        }
        part[slice][m] = acc;
        __syncthreads();
        if (tid < 128) {
            double total = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < kPoolSlices; ++s2) total += part[s2][tid];
            meanlog[tid] = total * inv_t;
        }
        __syncthreads();
        if (tid < 40) {
            const double* d = p.dct + tid * 128;
            double v = 0.0;
            for (int k = 0; k < 128; ++k) v = fma(d[k], meanlog[k], v);
            out[p.off_mfcc + tid] = static_cast<float>(v);
        }
    }
    if (p.off_mel >= 0 && tid < 128) {
        double acc = 0.0;
        for (int i = 0; i < n_tiles; ++i) acc += static_cast<double>(p.tile_mel[static_cast<long long>(clip.tile_base + i) * 128 + tid]);
        out[p.off_mel + tid] = static_cast<float>(acc * inv_t);
    }
    if (p.off_chroma >= 0 && tid >= 128 && tid < 140) {
        const int c = tid - 128;
        double acc = 0.0;
        for (int i = 0; i < n_tiles; ++i) acc += static_cast<double>(p.tile_chroma[static_cast<long long>(clip.tile_base + i) * 12 + c]);
        out[p.off_chroma + c] = static_cast<float>(acc * inv_t);
    }
    if (p.off_contrast >= 0 && tid >= 160 && tid < 167) out[p.off_contrast + tid - 160] = 0.0f;
}

cudaError_t launch_pool(const PoolParams& p, int n_clips, cudaStream_t stream) {
    if (n_clips <= 0) return cudaSuccess;
    pool_kernel<<<n_clips, kPoolThreads, 0, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace serb
