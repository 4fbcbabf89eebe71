// Host-side constant tables (see filterbanks.h).  Pure C++ / libm, float64 like numpy.
#include "filterbanks.h"

#include <algorithm>
#include <cmath>

namespace serb {

double tuning_edge(int idx) {
    // np.linspace(-0.5, 0.5, 101): y = arange(101) * step + start, step = 1.0 / 100
    if (idx >= 100) return 0.5;
    volatile double step = 1.0 / 100.0;
    volatile double prod = static_cast<double>(idx) * step;  // volatile: no fused multiply-add
    return prod + (-0.5);
}

namespace {

constexpr double kFsp = 200.0 / 3.0;
constexpr double kMinLogHz = 1000.0;
const double kMinLogMel = kMinLogHz / kFsp;  // 15
const double kLogStep = std::log(6.4) / 27.0;

double hz_to_mel(double f) {
    if (f >= kMinLogHz) return kMinLogMel + std::log(f / kMinLogHz) / kLogStep;
    return f / kFsp;
}

double mel_to_hz(double m) {
    if (m >= kMinLogMel) return kMinLogHz * std::exp(kLogStep * (m - kMinLogMel));
    return kFsp * m;
}

}  // namespace

void mel_points(int sample_rate, std::vector<double>& out) {
    const int n = kNMels + 2;
    const double fmax = static_cast<double>(sample_rate) / 2.0;
    const double min_mel = hz_to_mel(0.0);
    const double max_mel = hz_to_mel(fmax);
    const double step = (max_mel - min_mel) / static_cast<double>(n - 1);
    out.resize(n);
    for (int i = 0; i < n; ++i) {
        volatile double prod = static_cast<double>(i) * step;
        double mel = prod + min_mel;
        if (i == n - 1) mel = max_mel;
        out[i] = mel_to_hz(mel);
    }
}

void mel_filterbank(int sample_rate, int n_fft, std::vector<float>& w) {
    const int n_bins = 1 + n_fft / 2;
    std::vector<double> mel_f;
    mel_points(sample_rate, mel_f);
    // np.fft.rfftfreq(n, d = 1/sr): k * (1 / (n * d))
    const double d = 1.0 / static_cast<double>(sample_rate);
    const double val = 1.0 / (static_cast<double>(n_fft) * d);
    w.assign(static_cast<size_t>(kNMels) * n_bins, 0.0f);
    for (int m = 0; m < kNMels; ++m) {
        const double fd0 = mel_f[m + 1] - mel_f[m];
        const double fd1 = mel_f[m + 2] - mel_f[m + 1];
        const double enorm = 2.0 / (mel_f[m + 2] - mel_f[m]);
        for (int k = 0; k < n_bins; ++k) {
            const double freq = static_cast<double>(k) * val;
            const double lower = -(mel_f[m] - freq) / fd0;
            const double upper = (mel_f[m + 2] - freq) / fd1;
            const double tri = std::max(0.0, std::min(lower, upper));
            const float tri32 = static_cast<float>(tri);  // weights[] is float32 ...
            // ... and `weights *= enorm[:, None]` multiplies in float64, rounds to float32
            w[static_cast<size_t>(m) * n_bins + k] =
                static_cast<float>(static_cast<double>(tri32) * enorm);
        }
    }
}

void chroma_filterbank(int sample_rate, int n_fft, double tuning, std::vector<float>& w) {
    const int n_bins = 1 + n_fft / 2;
    const int n = n_fft;
    const double a440 = 440.0 * std::pow(2.0, tuning / 12.0);
    const double ref = a440 / 16.0;
    const double step = static_cast<double>(sample_rate) / static_cast<double>(n_fft);
    std::vector<double> frqbins(n), binwidth(n);
    for (int k = 1; k < n; ++k) {
        const double freq = static_cast<double>(k) * step;
        frqbins[k] = 12.0 * std::log2(freq / ref);
    }
    frqbins[0] = frqbins[1] - 1.5 * 12.0;
    for (int k = 0; k + 1 < n; ++k) binwidth[k] = std::max(frqbins[k + 1] - frqbins[k], 1.0);
    binwidth[n - 1] = 1.0;
    w.assign(static_cast<size_t>(kNChroma) * n_bins, 0.0f);
    for (int k = 0; k < n_bins && k < n; ++k) {
        double col[kNChroma];
        double sumsq = 0.0;
        for (int c = 0; c < kNChroma; ++c) {
            double dd = frqbins[k] - static_cast<double>(c);
            dd = (dd + 6.0) + 120.0;
            double r = std::fmod(dd, 12.0);
            if (r < 0) r += 12.0;
            dd = r - 6.0;
            const double z = 2.0 * dd / binwidth[k];
            col[c] = std::exp(-0.5 * (z * z));
            sumsq += col[c] * col[c];
        }
        double length = std::pow(sumsq, 0.5);
        if (length < 2.2250738585072014e-308) length = 1.0;
        const double oz = (frqbins[k] / 12.0 - 5.0) / 2.0;
        const double octw = std::exp(-0.5 * (oz * oz));
        for (int c = 0; c < kNChroma; ++c) {
            const double v = (col[c] / length) * octw;
            // np.roll(wts, -3, axis=0): new[c'] = old[(c' + 3) % 12]
            const int dst = (c + kNChroma - 3) % kNChroma;
            w[static_cast<size_t>(dst) * n_bins + k] = static_cast<float>(v);
        }
    }
}

void dct_matrix(std::vector<double>& d) {
    const int N = kNMels;
    d.resize(static_cast<size_t>(kNMfcc) * N);
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < kNMfcc; ++k) {
        const double s = (k == 0) ? std::sqrt(1.0 / (4.0 * N)) : std::sqrt(1.0 / (2.0 * N));
        for (int n = 0; n < N; ++n)
            d[static_cast<size_t>(k) * N + n] = 2.0 * s * std::cos(pi * k * (2.0 * n + 1.0) / (2.0 * N));
    }
}

void hann_periodic(int n, std::vector<double>& w) {
    w.resize(n);
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < n; ++i) w[i] = 0.5 - 0.5 * std::cos(2.0 * pi * i / static_cast<double>(n));
}

void mel_sparse(const std::vector<float>& dense, int n_bins, MelSparse& out) {
    out.start.assign(kNMels, 0);
    out.count.assign(kNMels, 0);
    out.offset.assign(kNMels + 1, 0);
    out.weights.clear();
    for (int m = 0; m < kNMels; ++m) {
        int first = -1, last = -1;
        for (int k = 0; k < n_bins; ++k) {
            if (dense[static_cast<size_t>(m) * n_bins + k] != 0.0f) {
                if (first < 0) first = k;
                last = k;
            }
        }
        out.offset[m] = static_cast<int32_t>(out.weights.size());   // always a multiple of 4
        if (first >= 0) {
            out.start[m] = first;
            // padded with zero weights to a multiple of 4 so the kernel reads 16-byte groups; the
            // padding may index up to 3 bins past the last one (the kernel keeps 3 zero rows there)
            const int n = last - first + 1;
            const int padded = (n + 3) / 4 * 4;
            out.count[m] = padded;
            for (int k = first; k < first + padded; ++k)
                out.weights.push_back(k <= last ? dense[static_cast<size_t>(m) * n_bins + k] : 0.0f);
        }
    }
    out.offset[kNMels] = static_cast<int32_t>(out.weights.size());
}

}  // namespace serb
