// Generic path for clips shorter than 2048 samples (n_fft = len(padded clip) < 2048).
//
// ser/_internal/utils/dsp.py:93-96 pads a clip to >= 512 samples and sets
// n_fft = min(len, 2048); for len < 2048 the reference then runs an STFT of arbitrary
// (possibly odd) length n_fft with hop n_fft // 4 for chroma (dsp.py:100,113-118) and a
// second one with hop 512 for mel / mfcc (dsp.py:106-111,120-125).  These are rare (tail
// windows of a recording, SURVEY.md section 7 "Ragged shapes"), so one CTA per clip does
// everything with a direct float64 DFT and on-the-fly filterbanks; no attempt at speed.
#include "common.cuh"
#include "kernels.h"
#include <cfloat>

namespace serb {

constexpr int kShortThreads = 256;
constexpr int kShortMaxN = 2048;
constexpr int kShortMaxBins = 1025;
constexpr int kShortColsA = 8;   // chroma STFT columns (always 5) -- padded
constexpr int kShortColsB = 4;   // mel STFT columns (2..4)
constexpr int kShortPeakCap = 5 * 520;



struct ShortSmem {
    double2 tw[kShortMaxN];
    float y[kShortMaxN];
    float xw[kShortMaxN];
    float sa[kShortColsA][kShortMaxBins];
    float pb[kShortColsB][kShortMaxBins];
    float2 pk[kShortPeakCap];
    float wc[12][kShortMaxBins];
    double meanlog[128];
    float logmel[kShortColsB][128];
    float mel[kShortColsB][128];
    float raw[12][kShortColsA];
    float colmax[kShortColsA];
    int counts[100];
    int n_peaks;
    int tuning;
    float thr;
    float med[2];
    float gmax;
};

// |X[k]| of one windowed frame held in sm.xw, direct DFT in float64, rounded like complex64
__device__ void short_dft_mag(ShortSmem& sm, int nf, int nb, float* __restrict__ out) {
    for (int k = threadIdx.x; k < nb; k += kShortThreads) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        for (int n = 0; n < nf; ++n) {
            const double x = static_cast<double>(sm.xw[n]);
            const double2 w = sm.tw[idx];
            re = fma(x, w.x, re);
            im = fma(x, w.y, im);
            idx += k;
            if (idx >= nf) idx -= nf;
        }
        const float r32 = static_cast<float>(re), i32 = static_cast<float>(im);
        out[k] = static_cast<float>(sqrt(static_cast<double>(r32) * r32 + static_cast<double>(i32) * i32));
    }
}

__global__ void __launch_bounds__(kShortThreads, 1) short_kernel(ShortParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ShortSmem& sm = *reinterpret_cast<ShortSmem*>(smem_raw);
    const ShortClip clip = p.clips[blockIdx.x];
    const int tid = threadIdx.x;
    const int nf = max(clip.length, 512);     // padded length == n_fft
    const int nb = 1 + nf / 2;
    const int pad = nf / 2;
    const double sr = static_cast<double>(p.sample_rate);

    for (int i = tid; i < nf; i += kShortThreads) {
        sm.y[i] = (i < clip.length) ? p.wave[clip.start + i] : 0.0f;
        double s, c;
        sincospi(2.0 * static_cast<double>(i) / static_cast<double>(nf), &s, &c);
        sm.tw[i] = make_double2(c, -s);
    }
    if (tid == 0) sm.n_peaks = 0;
    for (int i = tid; i < 100; i += kShortThreads) sm.counts[i] = 0;
    __syncthreads();
    {
        int bad = 0;
        for (int i = tid; i < nf; i += kShortThreads) bad |= !isfinite(sm.y[i]);
        if (__syncthreads_or(bad) && tid == 0) atomicOr(p.status, 1);
    }

    // ---- STFT A: hop = n_fft // 4, centred, zero padded ----
    const int hop_a = nf / 4;
    const int cols_a = 1 + (nf + 2 * pad - nf) / hop_a;
    for (int t = 0; t < cols_a && t < kShortColsA; ++t) {
        for (int n = tid; n < nf; n += kShortThreads) {
            const int src = t * hop_a + n - pad;
            const float x = (src >= 0 && src < nf) ? sm.y[src] : 0.0f;
            const double w = 0.5 - 0.5 * sm.tw[n].x;
            sm.xw[n] = static_cast<float>(w * static_cast<double>(x));  // float64 product (window is float64)
        }
        __syncthreads();
        short_dft_mag(sm, nf, nb, sm.sa[t]);
        __syncthreads();
    }
    // NOTE: the reference multiplies in float64 and feeds the float64 product to the FFT; xw
    // is rounded to float32 here, a relative 6e-8 perturbation of each sample.

    // ---- piptrack + tuning (chroma only) ----
    if (p.off_chroma >= 0) {
        const int nfft_pip = 2 * (nb - 1);    // _spectrogram re-derives n_fft from S when n_fft=2048 mismatches
        const double fval = 1.0 / (static_cast<double>(nfft_pip) * (1.0 / sr));
        const double f_limit = fmin(4000.0, sr / 2.0);
        if (tid < cols_a) {
            float m = 0.f;
            for (int k = 0; k < nb; ++k) m = fmaxf(m, sm.sa[tid][k]);
            sm.colmax[tid] = m;
        }
        __syncthreads();
        for (int i = tid; i < cols_a * nb; i += kShortThreads) {
            const int t = i / nb, k = i - t * nb;
            const double fk = static_cast<double>(k) * fval;
            if (k < 1 || k > nb - 2 || !(fk >= 150.0) || !(fk < f_limit)) continue;
            const float ref = __fmul_rn(0.1f, sm.colmax[t]);
            const float sm1 = sm.sa[t][k - 1], s0 = sm.sa[t][k], sp1 = sm.sa[t][k + 1];
            const float tm1 = sm1 > ref ? sm1 : 0.f, t0 = s0 > ref ? s0 : 0.f, tp1 = sp1 > ref ? sp1 : 0.f;
            if (!((t0 > tm1) && (t0 >= tp1))) continue;
            const double a = static_cast<double>(__fadd_rn(sp1, sm1)) - 2.0 * static_cast<double>(s0);
            const double b = static_cast<double>(__fsub_rn(sp1, sm1)) / 2.0;
            float shift = 0.f;
            if (fabs(b) < fabs(a)) shift = static_cast<float>(-b / a);
            const float avg = __fmul_rn(__fsub_rn(sp1, sm1), 0.5f);
            const float mag = __fadd_rn(s0, __fmul_rn(__fmul_rn(0.5f, avg), shift));
            const float pitch = static_cast<float>(((static_cast<double>(k) + static_cast<double>(shift)) * sr) /
                                                   static_cast<double>(nfft_pip));
            if (pitch > 0.f) {
                const int slot = atomicAdd(&sm.n_peaks, 1);
                if (slot < kShortPeakCap) sm.pk[slot] = make_float2(mag, pitch);
            }
        }
        __syncthreads();
        const int n = min(sm.n_peaks, kShortPeakCap);
        int tuning = 50;
        if (n > 0) {
            // exact median by rank counting (n <= 2600)
            const int r_lo = (n & 1) ? n / 2 : n / 2 - 1, r_hi = n / 2;
            for (int i = tid; i < n; i += kShortThreads) {
                const float v = sm.pk[i].x;
                int rank = 0;
                for (int j = 0; j < n; ++j) {
                    const float u = sm.pk[j].x;
                    rank += (u < v) || (u == v && j < i);
                }
                if (rank == r_lo) sm.med[0] = v;
                if (rank == r_hi) sm.med[1] = v;
            }
            __syncthreads();
            const float thr = (n & 1) ? sm.med[1] : __fmul_rn(__fadd_rn(sm.med[0], sm.med[1]), 0.5f);
            for (int i = tid; i < n; i += kShortThreads) {
                const float2 pk = sm.pk[i];
                if (!(pk.x >= thr)) continue;
                const float q = __fdiv_rn(pk.y, 27.5f);
                const float l2 = static_cast<float>(log2(static_cast<double>(q)));
                float r = fmodf(__fmul_rn(12.0f, l2), 1.0f);
                if (r < 0.f) r += 1.0f;
                if (r >= 0.5f) r = __fsub_rn(r, 1.0f);
                const double rd = static_cast<double>(r);
                int bin = max(0, min(99, static_cast<int>(floor((rd + 0.5) * 100.0))));
                while (bin > 0 && rd < p.edges[bin]) --bin;
                while (bin < 99 && rd >= p.edges[bin + 1]) ++bin;
                atomicAdd(&sm.counts[bin], 1);
            }
            __syncthreads();
            if (tid == 0) {
                int best = 0, bc = sm.counts[0];
                for (int i = 1; i < 100; ++i) if (sm.counts[i] > bc) { bc = sm.counts[i]; best = i; }
                sm.tuning = best;
            }
            __syncthreads();
            tuning = sm.tuning;
        }
        if (tid == 0 && p.tuning_idx) p.tuning_idx[blockIdx.x] = tuning;

        // ---- chroma filterbank for (sr, n_fft = nf, tuning) built in place (filters.chroma) ----
        const double tun = p.edges[tuning];
        const double ref440 = (440.0 * exp2(tun / 12.0)) / 16.0;
        const double step = sr / static_cast<double>(nf);
        for (int k = tid; k < nb; k += kShortThreads) {
            // frqbins[k], frqbins[k+1] over the full n_fft-long axis (k + 1 < nf always for k < nb when nf >= 4)
            double fb_k, fb_k1;
            const double fb_1 = 12.0 * log2((1.0 * step) / ref440);
            fb_k = (k == 0) ? (fb_1 - 18.0) : 12.0 * log2((static_cast<double>(k) * step) / ref440);
            double bw = 1.0;
            if (k + 1 < nf) {
                fb_k1 = 12.0 * log2((static_cast<double>(k + 1) * step) / ref440);
                bw = fmax(fb_k1 - fb_k, 1.0);
            }
            double col[12];
            double sumsq = 0.0;
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                double d = (fb_k - static_cast<double>(c) + 6.0) + 120.0;
                double r = fmod(d, 12.0);
                if (r < 0) r += 12.0;
                d = r - 6.0;
                const double z = 2.0 * d / bw;
                col[c] = exp(-0.5 * (z * z));
                sumsq += col[c] * col[c];
            }
            double length = sqrt(sumsq);
            if (length < DBL_MIN) length = 1.0;
            const double oz = (fb_k / 12.0 - 5.0) / 2.0;
            const double octw = exp(-0.5 * (oz * oz));
#pragma unroll
            for (int c = 0; c < 12; ++c) sm.wc[(c + 9) % 12][k] = static_cast<float>((col[c] / length) * octw);
        }
        __syncthreads();
        if (tid < 12 * cols_a) {
            const int c = tid / cols_a, t = tid - c * cols_a;
            double acc = 0.0;
            for (int k = 0; k < nb; ++k) acc = fma(static_cast<double>(sm.wc[c][k]), static_cast<double>(sm.sa[t][k]), acc);
            sm.raw[c][t] = static_cast<float>(acc);
        }
        __syncthreads();
        if (tid < cols_a) {
            float length = 0.f;
            for (int c = 0; c < 12; ++c) length = fmaxf(length, fabsf(sm.raw[c][tid]));
            const double len = (length < FLT_MIN) ? 1.0 : static_cast<double>(length);
            for (int c = 0; c < 12; ++c) sm.raw[c][tid] = static_cast<float>(static_cast<double>(sm.raw[c][tid]) / len);
        }
        __syncthreads();
        if (tid < 12) {
            double acc = 0.0;
            for (int t = 0; t < cols_a; ++t) acc += static_cast<double>(sm.raw[tid][t]);
            p.out[static_cast<long long>(clip.out_row) * p.dim + p.off_chroma + tid] =
                static_cast<float>(acc / static_cast<double>(cols_a));
        }
        __syncthreads();
    }

    // ---- STFT B (hop 512) -> mel power, log-mel, MFCC ----
    if (p.off_mel >= 0 || p.off_mfcc >= 0) {
        const int cols_b = 1 + (nf + 2 * pad - nf) / 512;
        for (int t = 0; t < cols_b && t < kShortColsB; ++t) {
            for (int n = tid; n < nf; n += kShortThreads) {
                const int src = t * 512 + n - pad;
                const float x = (src >= 0 && src < nf) ? sm.y[src] : 0.0f;
                const double w = 0.5 - 0.5 * sm.tw[n].x;
                sm.xw[n] = static_cast<float>(w * static_cast<double>(x));
            }
            __syncthreads();
            short_dft_mag(sm, nf, nb, sm.pb[t]);
            __syncthreads();
            for (int k = tid; k < nb; k += kShortThreads) { const float v = sm.pb[t][k]; sm.pb[t][k] = v * v; }
            __syncthreads();
        }
        const double val = 1.0 / (static_cast<double>(nf) * (1.0 / sr));
        for (int i = tid; i < 128 * cols_b; i += kShortThreads) {
            const int m = i % 128, t = i / 128;
            const double f0 = p.mel_points[m], f1 = p.mel_points[m + 1], f2 = p.mel_points[m + 2];
            const double fd0 = f1 - f0, fd1 = f2 - f1, enorm = 2.0 / (f2 - f0);
            // bins with f0 < k*val < f2
            int k_lo = static_cast<int>(floor(f0 / val)) - 1, k_hi = static_cast<int>(ceil(f2 / val)) + 1;
            k_lo = max(k_lo, 0);
            k_hi = min(k_hi, nb - 1);
            float acc = 0.f;
            for (int k = k_lo; k <= k_hi; ++k) {
                const double fk = static_cast<double>(k) * val;
                const double tri = fmax(0.0, fmin(-(f0 - fk) / fd0, (f2 - fk) / fd1));
                const float w = static_cast<float>(static_cast<double>(static_cast<float>(tri)) * enorm);
                acc = fmaf(w, sm.pb[t][k], acc);
            }
            sm.mel[t][m] = acc;
            sm.logmel[t][m] = 10.0f * log10f(fmaxf(1e-10f, acc));
        }
        __syncthreads();
        if (tid == 0) {
            float g = -FLT_MAX;
            for (int t = 0; t < cols_b; ++t)
                for (int m = 0; m < 128; ++m) g = fmaxf(g, sm.logmel[t][m]);
            sm.gmax = g;
        }
        __syncthreads();
        if (tid < 128) {
            const float thr = __fsub_rn(sm.gmax, 80.0f);
            double acc = 0.0, macc = 0.0;
            for (int t = 0; t < cols_b; ++t) {
                acc += static_cast<double>(fmaxf(sm.logmel[t][tid], thr));
                macc += static_cast<double>(sm.mel[t][tid]);
            }
            sm.meanlog[tid] = acc / static_cast<double>(cols_b);
            if (p.off_mel >= 0)
                p.out[static_cast<long long>(clip.out_row) * p.dim + p.off_mel + tid] =
                    static_cast<float>(macc / static_cast<double>(cols_b));
        }
        __syncthreads();
        if (p.off_mfcc >= 0 && tid < 40) {
            double v = 0.0;
            for (int m = 0; m < 128; ++m) v = fma(p.dct[tid * 128 + m], sm.meanlog[m], v);
            p.out[static_cast<long long>(clip.out_row) * p.dim + p.off_mfcc + tid] = static_cast<float>(v);
        }
    }
    if (p.off_contrast >= 0 && tid < 7)
        p.out[static_cast<long long>(clip.out_row) * p.dim + p.off_contrast + tid] = 0.0f;
}

cudaError_t configure_short() {
    return cudaFuncSetAttribute(short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(sizeof(ShortSmem)));
}

cudaError_t launch_short(const ShortParams& p, int n_clips, cudaStream_t stream) {
    if (n_clips <= 0) return cudaSuccess;
    short_kernel<<<n_clips, kShortThreads, sizeof(ShortSmem), stream>>>(p);
    return cudaGetLastError();
}

}  // namespace serb
