// Host-side tables of the tonnetz chain (see cqt_tables.h).  float64 math with float32 /
// complex64 rounding at the points where librosa 0.11.0 rounds.
#include "cqt_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <complex>

#include "filterbanks.h"

namespace serb {

namespace {

constexpr double kPi = 3.14159265358979323846;
constexpr double kHannBandwidth = 1.50018310546875;   // librosa.filters.WINDOW_BANDWIDTHS["hann"]

struct Wavelets {
    double freqs[kCqtBins];
    double alpha[kCqtBins];
    double cutoff;
};

// freqs / alpha of the 252 bins for one tuning (librosa.vqt: fmin * 2^(tuning/bpo), equal
// temperament; filters._relative_bandwidth) and the top wavelet's cutoff (wavelet_lengths)
void wavelets_for_tuning(int tuning_idx, Wavelets& w) {
    const double tuning = tuning_edge(tuning_idx);
    const double c1 = 440.0 * std::pow(2.0, (24.0 - 69.0) / 12.0);
    const double fmin = c1 * std::pow(2.0, tuning / static_cast<double>(kCqtBpo));
    for (int o = 0; o < kCqtOctaves; ++o)
        for (int j = 0; j < kCqtBpo; ++j) {
            const double ratio = std::pow(2.0, static_cast<double>(j) / kCqtBpo);
            w.freqs[o * kCqtBpo + j] = (std::pow(2.0, static_cast<double>(o)) * ratio) * fmin;
        }
    double logf[kCqtBins], bpo[kCqtBins];
    for (int i = 0; i < kCqtBins; ++i) logf[i] = std::log2(w.freqs[i]);
    bpo[0] = 1.0 / (logf[1] - logf[0]);
    bpo[kCqtBins - 1] = 1.0 / (logf[kCqtBins - 1] - logf[kCqtBins - 2]);
    for (int i = 1; i + 1 < kCqtBins; ++i) bpo[i] = 2.0 / (logf[i + 1] - logf[i - 1]);
    double cutoff = 0.0;
    for (int i = 0; i < kCqtBins; ++i) {
        const double p = std::pow(2.0, 2.0 / bpo[i]);
        w.alpha[i] = (p - 1.0) / (p + 1.0);
        const double q = 1.0 / w.alpha[i];
        cutoff = std::max(cutoff, w.freqs[i] * (1.0 + 0.5 * kHannBandwidth / q));
    }
    w.cutoff = cutoff;
}

int early_count(double nyquist, double cutoff) {
    const int c1 = std::max(0, static_cast<int>(std::ceil(std::log2(nyquist / cutoff)) - 1) - 1);
    const int c2 = std::max(0, 9 - kCqtOctaves + 1);   // hop 512 = 2^9
    return std::min(c1, c2);
}

// per-octave FFT size: power of two >= the longest wavelet of the octave
void octave_nfft(const Wavelets& w, double sr_eff, int* n_fft) {
    for (int i = 0; i < kCqtOctaves; ++i) {
        const double my_sr = sr_eff / std::pow(2.0, static_cast<double>(i));
        double max_len = 0.0;
        for (int j = 0; j < kCqtBpo; ++j) {
            const int b = kCqtBins - kCqtBpo * (i + 1) + j;
            max_len = std::max(max_len, (1.0 / w.alpha[b]) * my_sr / w.freqs[b]);
        }
        n_fft[i] = static_cast<int>(std::pow(2.0, std::ceil(std::log2(max_len))));
    }
}

// iterative radix-2 FFT; tw[k] = exp(-2 pi i k / n), k < n / 2
void fft_inplace(std::vector<std::complex<double>>& a, const std::vector<std::complex<double>>& tw) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t stride = n / len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * tw[k * stride];
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

// sparsified complex64 FFT-domain basis of one octave: dense[36][n_bins] (re, im) float32,
// scaled by sqrt(sr_eff / my_sr); lengths[j] of the octave's wavelets at my_sr
void octave_basis(const Wavelets& w, double sr_eff, int octave, int n_fft, std::vector<float>& dense) {
    const int n_bins = 1 + n_fft / 2;
    const double my_sr = sr_eff / std::pow(2.0, static_cast<double>(octave));
    const double oct_scale = std::sqrt(sr_eff / my_sr);
    dense.assign(static_cast<size_t>(kCqtBpo) * n_bins * 2, 0.0f);
    std::vector<std::complex<double>> buf(n_fft), tw(n_fft / 2);
    for (int k = 0; k < n_fft / 2; ++k) {
        const double ang = -2.0 * kPi * static_cast<double>(k) / static_cast<double>(n_fft);
        tw[k] = std::complex<double>(std::cos(ang), std::sin(ang));
    }
    std::vector<float> mags(n_bins), sorted(n_bins);
    for (int j = 0; j < kCqtBpo; ++j) {
        const int b = kCqtBins - kCqtBpo * (octave + 1) + j;
        const double freq = w.freqs[b];
        const double ilen = (1.0 / w.alpha[b]) * my_sr / freq;
        const long n_lo = static_cast<long>(std::floor(-ilen / 2.0));
        const long n_hi = static_cast<long>(std::floor(ilen / 2.0));
        const int count = static_cast<int>(n_hi - n_lo);
        std::vector<std::complex<double>> sig(count);
        // periodic Hann of `count` points: scipy general_cosine over linspace(-pi, pi, count + 1)
        const double step = (2.0 * kPi) / static_cast<double>(count);
        double l1 = 0.0;
        for (int t = 0; t < count; ++t) {
            const double n = static_cast<double>(n_lo + t);
            const double angle = (((n * 2.0) * kPi) * freq) / my_sr;
            const double fac = static_cast<double>(t) * step + (-kPi);
            const double win = 0.5 + 0.5 * std::cos(fac);
            sig[t] = std::complex<double>(std::cos(angle), std::sin(angle)) * win;
            l1 += std::abs(sig[t]);
        }
        if (l1 < 2.2250738585072014e-308) l1 = 1.0;
        std::fill(buf.begin(), buf.end(), std::complex<double>(0.0, 0.0));
        const int lpad = (n_fft - count) / 2;
        const double len_scale = ilen / static_cast<double>(n_fft);
        for (int t = 0; t < count; ++t) {
            const std::complex<double> v = sig[t] / l1;
            // asarray(..., complex64), then `basis *= lengths / n_fft` rounds to complex64 again
            const float re32 = static_cast<float>(v.real()), im32 = static_cast<float>(v.imag());
            const float re = static_cast<float>(static_cast<double>(re32) * len_scale);
            const float im = static_cast<float>(static_cast<double>(im32) * len_scale);
            buf[lpad + t] = std::complex<double>(re, im);
        }
        fft_inplace(buf, tw);
        double l1m = 0.0;
        for (int k = 0; k < n_bins; ++k) {
            const float re = static_cast<float>(buf[k].real()), im = static_cast<float>(buf[k].imag());
            buf[k] = std::complex<double>(re, im);   // complex64 result
            mags[k] = static_cast<float>(std::hypot(static_cast<double>(re), static_cast<double>(im)));
            l1m += mags[k];
        }
        // util.sparsify_rows(quantile=0.01): float32 magnitudes, float32 cumulative sum
        sorted = mags;
        std::sort(sorted.begin(), sorted.end());
        const float norm = static_cast<float>(l1m);
        float cum = 0.0f, thresh = sorted[n_bins - 1];
        for (int k = 0; k < n_bins; ++k) {
            cum += sorted[k] / norm;
            if (!(cum < 0.01f)) { thresh = sorted[k]; break; }
        }
        float* row = dense.data() + static_cast<size_t>(j) * n_bins * 2;
        for (int k = 0; k < n_bins; ++k) {
            if (mags[k] >= thresh) {
                row[2 * k] = static_cast<float>(buf[k].real() * oct_scale);
                row[2 * k + 1] = static_cast<float>(buf[k].imag() * oct_scale);
            }
        }
    }
}

}  // namespace


namespace {

// Kuhn's augmenting-path matching: lane -> row with adjacency `allowed`
bool try_lane(int lane, const std::vector<std::vector<int>>& allowed, std::vector<int>& row_of_lane,
              std::vector<int>& lane_of_row, std::vector<char>& seen) {
    for (int r : allowed[lane]) {
        if (seen[r]) continue;
        seen[r] = 1;
        if (lane_of_row[r] < 0 || try_lane(lane_of_row[r], allowed, row_of_lane, lane_of_row, seen)) {
            row_of_lane[lane] = r;
            lane_of_row[r] = lane;
            return true;
        }
    }
    return false;
}

// Reorders the 36 rows of one octave for the kernel's lane = row mapping.  The kernel's lanes read
// X[start + c] at the same c: a half-warp (16 lanes, 8-byte words) is conflict-free when the 16
// first bins are distinct modulo 16.  Row r may start up to a few bins early (zero coefficients in
// front); among all lane assignments with first bin = lane (mod 16) the one with the smallest
// widest row is taken (bottleneck matching).  Entries 32..35 (handled four lanes at a time) keep
// their bins.  Falls back to the natural order if no assignment fits the row capacity.
void lane_order_octave(CqtBank& bank, int octave) {
    CqtRow* rows = bank.rows.data() + static_cast<size_t>(octave) * kCqtBpo;
    float* vals = bank.vals.data() + static_cast<size_t>(octave) * kCqtBpo * kCqtRowCap * 2;
    int max_count = 0;
    for (int r = 0; r < kCqtBpo; ++r) max_count = std::max(max_count, rows[r].count);
    auto shift_of = [&](int lane, int r) { return ((rows[r].start - (lane & 15)) % 16 + 16) % 16; };
    for (int tau = max_count; tau <= kCqtRowCap; ++tau) {
        std::vector<std::vector<int>> allowed(32);
        for (int lane = 0; lane < 32; ++lane)
            for (int r = 0; r < kCqtBpo; ++r)
                if (rows[r].count > 0 && rows[r].count + shift_of(lane, r) <= tau && rows[r].start - shift_of(lane, r) >= 0)
                    allowed[lane].push_back(r);
        std::vector<int> row_of_lane(32, -1), lane_of_row(kCqtBpo, -1);
        int matched = 0;
        for (int lane = 0; lane < 32; ++lane) {
            std::vector<char> seen(kCqtBpo, 0);
            if (try_lane(lane, allowed, row_of_lane, lane_of_row, seen)) ++matched;
        }
        if (matched < 32) continue;
        std::vector<CqtRow> new_rows(kCqtBpo);
        std::vector<float> new_vals(static_cast<size_t>(kCqtBpo) * kCqtRowCap * 2, 0.0f);
        int spare = 32;
        for (int r = 0; r < kCqtBpo; ++r) {
            const int slot = lane_of_row[r] >= 0 ? lane_of_row[r] : spare++;
            const int d = lane_of_row[r] >= 0 ? shift_of(slot, r) : 0;
            new_rows[slot] = rows[r];
            new_rows[slot].start = rows[r].start - d;
            new_rows[slot].count = rows[r].count + d;
            for (int k = 0; k < rows[r].count; ++k) {
                new_vals[(static_cast<size_t>(slot) * kCqtRowCap + d + k) * 2] = vals[(static_cast<size_t>(r) * kCqtRowCap + k) * 2];
                new_vals[(static_cast<size_t>(slot) * kCqtRowCap + d + k) * 2 + 1] = vals[(static_cast<size_t>(r) * kCqtRowCap + k) * 2 + 1];
            }
        }
        std::copy(new_rows.begin(), new_rows.end(), rows);
        std::copy(new_vals.begin(), new_vals.end(), vals);
        return;
    }
}

}  // namespace

void cqt_plan(int sample_rate, CqtPlan& plan) {
    plan = CqtPlan();
    plan.sample_rate = sample_rate;
    const double sr = static_cast<double>(sample_rate);
    const double nyquist = sr / 2.0;
    int n_bad = 0;
    bool first = true;
    for (int t = 0; t < kNTunings; ++t) {
        Wavelets w;
        wavelets_for_tuning(t, w);
        if (w.cutoff > nyquist) { ++n_bad; continue; }
        const int count = early_count(nyquist, w.cutoff);
        int n_fft[kCqtOctaves];
        octave_nfft(w, sr / std::pow(2.0, static_cast<double>(count)), n_fft);
        if (first) {
            plan.early_factor = 1 << count;
            plan.hop0 = 512 >> count;
            for (int i = 0; i < kCqtOctaves; ++i) plan.n_fft[i] = n_fft[i];
            first = false;
        } else {
            bool same = plan.early_factor == (1 << count);
            for (int i = 0; i < kCqtOctaves; ++i) same = same && plan.n_fft[i] == n_fft[i];
            if (!same) {
                plan.status = 2;
                plan.message = "tonnetz: the constant-Q plan at this sample rate depends on the tuning estimate (unsupported)";
                return;
            }
        }
    }
    if (n_bad == kNTunings) {
        plan.status = 1;
        plan.message = "Wavelet basis would exceed the Nyquist frequency. Try reducing the number of frequency bins.";
        return;
    }
    if (n_bad > 0) {
        plan.status = 2;
        plan.message = "tonnetz: the constant-Q Nyquist check at this sample rate depends on the tuning estimate (unsupported)";
        return;
    }
    for (int i = 0; i < kCqtOctaves; ++i) {
        if (plan.n_fft[i] < 128 || plan.n_fft[i] > 2048) {
            plan.status = 2;
            plan.message = "tonnetz: constant-Q FFT size outside 128..2048 at this sample rate (unsupported)";
            return;
        }
    }
}

void cqt_basis_dense(const CqtPlan& plan, int tuning_idx, int octave, std::vector<float>& out) {
    Wavelets w;
    wavelets_for_tuning(tuning_idx, w);
    const double sr_eff = static_cast<double>(plan.sample_rate) / static_cast<double>(plan.early_factor);
    octave_basis(w, sr_eff, octave, plan.n_fft[octave], out);
}

bool cqt_bank(const CqtPlan& plan, int tuning_idx, CqtBank& bank, bool lane_order) {
    Wavelets w;
    wavelets_for_tuning(tuning_idx, w);
    const double sr_eff = static_cast<double>(plan.sample_rate) / static_cast<double>(plan.early_factor);
    bank.rows.assign(static_cast<size_t>(kCqtOctaves) * kCqtBpo, CqtRow{0, 0, 0.0f, 0});
    bank.vals.assign(static_cast<size_t>(kCqtOctaves) * kCqtBpo * kCqtRowCap * 2, 0.0f);
    std::vector<float> dense;
    bool ok = true;
    for (int i = 0; i < kCqtOctaves; ++i) {
        const int n_bins = 1 + plan.n_fft[i] / 2;
        octave_basis(w, sr_eff, i, plan.n_fft[i], dense);
        for (int j = 0; j < kCqtBpo; ++j) {
            const float* row = dense.data() + static_cast<size_t>(j) * n_bins * 2;
            int first = -1, last = -1;
            for (int k = 0; k < n_bins; ++k)
                if (row[2 * k] != 0.0f || row[2 * k + 1] != 0.0f) { if (first < 0) first = k; last = k; }
            const int b = kCqtBins - kCqtBpo * (i + 1) + j;
            CqtRow& r = bank.rows[static_cast<size_t>(i) * kCqtBpo + j];
            r.bin = b;
            // V /= sqrt(lengths), lengths = Q * sr_eff / freqs over all 252 bins
            r.scale = static_cast<float>(1.0 / std::sqrt((1.0 / w.alpha[b]) * sr_eff / w.freqs[b]));
            if (first < 0) continue;
            int count = last - first + 1;
            if (count > kCqtRowCap) { ok = false; count = kCqtRowCap; }
            r.start = first;
            r.count = count;
            float* dst = bank.vals.data() + (static_cast<size_t>(i) * kCqtBpo + j) * kCqtRowCap * 2;
            for (int k = 0; k < count; ++k) { dst[2 * k] = row[2 * (first + k)]; dst[2 * k + 1] = row[2 * (first + k) + 1]; }
        }
        if (lane_order) lane_order_octave(bank, i);
    }
    return ok;
}

bool cqt_set_banks(const CqtPlan& plan, int tuning_idx, CqtSetBank* out7) {
    CqtBank bank;
    if (!cqt_bank(plan, tuning_idx, bank, /*lane_order=*/false)) return false;
    for (int i = 0; i < kCqtOctaves; ++i) {
        CqtSetBank& sb = out7[i];
        std::memset(&sb, 0, sizeof(sb));
        const CqtRow* rows = bank.rows.data() + static_cast<size_t>(i) * kCqtBpo;
        const float* vals = bank.vals.data() + static_cast<size_t>(i) * kCqtBpo * kCqtRowCap * 2;
        int lo = 1 << 30, hi = 0;
        for (int j = 0; j < kCqtBpo; ++j)
            if (rows[j].count > 0) { lo = std::min(lo, rows[j].start); hi = std::max(hi, rows[j].start + rows[j].count); }
        if (hi <= lo) { lo = 0; hi = 1; }
        if (hi - lo > kCqtSetMaxBins) return false;
        sb.bin_lo = lo;
        sb.n_bins = hi - lo;
        int off = 0, row = 0;
        for (int s = 0; s < kCqtSets; ++s) {
            const int nrows = s < 4 ? 3 : 2, rpad = s < 4 ? 4 : 2;
            int u0 = 1 << 30, u1 = 0;
            for (int q = 0; q < nrows; ++q)
                if (rows[row + q].count > 0) {
                    u0 = std::min(u0, rows[row + q].start);
                    u1 = std::max(u1, rows[row + q].start + rows[row + q].count);
                }
            if (u1 <= u0) { u0 = lo; u1 = lo; }
            CqtSet& set = sb.sets[s];
            set.u0 = static_cast<int16_t>(u0 - lo);
            set.ulen = static_cast<int16_t>(u1 - u0);
            set.off = static_cast<int16_t>(off);
            set.nrows = static_cast<int16_t>(nrows);
            if (off + (u1 - u0) * rpad > kCqtSetValCap) return false;
            for (int q = 0; q < 4; ++q) { set.bin[q] = -1; set.scale[q] = 0.0f; }
            for (int q = 0; q < nrows; ++q) {
                const CqtRow& r = rows[row + q];
                set.bin[q] = static_cast<int16_t>(r.bin);
                set.scale[q] = r.scale;
                for (int k = 0; k < r.count; ++k) {
                    const size_t at = static_cast<size_t>(off) + static_cast<size_t>(r.start + k - u0) * rpad + q;
                    sb.vals[2 * at] = vals[(static_cast<size_t>(row + q) * kCqtRowCap + k) * 2];
                    sb.vals[2 * at + 1] = vals[(static_cast<size_t>(row + q) * kCqtRowCap + k) * 2 + 1];
                }
            }
            off += (u1 - u0) * rpad;
            off = (off + 1) & ~1;          // sets start on 16-byte boundaries
            row += nrows;
        }
    }
    return true;
}

// libsoxr's cubic fits of the Kaiser beta (restated; see oracle/shim/librosa/core.py _LSX_BETA_ROWS)
static double lsx_kaiser_beta(double att, double tr_bw) {
    static const double rows[10][4] = {
        {-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + 0.001},
        {-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + 0.002},
        {-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + 0.003},
        {-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + 0.006},
        {8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + 0.015},
        {9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + 0.025},
        {-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + 0.05},
        {-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + 0.085},
        {1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + 0.1},
        {-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + 0.18},
    };
    const double realm = std::log(tr_bw / 0.0005) / std::log(2.0);
    const int i0 = std::min(std::max(static_cast<int>(realm), 0), 9);
    const int i1 = std::min(std::max(1 + static_cast<int>(realm), 0), 9);
    const double b0 = ((rows[i0][0] * att + rows[i0][1]) * att + rows[i0][2]) * att + rows[i0][3];
    const double b1 = ((rows[i1][0] * att + rows[i1][1]) * att + rows[i1][2]) * att + rows[i1][3];
    return b0 + (b1 - b0) * (realm - static_cast<int>(realm));
}

void decimation_taps(int factor, std::vector<double>& taps) {
    // libsoxr "HQ" low-pass for an integer decimation, restated (oracle/shim/librosa/core.py
    // _soxr_hq_decimation_filter has the derivation): 389 taps, beta 13.04 for factor 2
    const double bits = 20.0;
    const double rej = bits * 20.0 * std::log10(2.0);
    const double att = (bits + 1.0) * 20.0 * std::log10(2.0);
    const double pass_end = 1.0 - 0.05 / ((1.6e-6 * rej - 7.5e-4) * rej + 0.646);
    const double fp = pass_end / static_cast<double>(factor), fs = 1.0 / static_cast<double>(factor);
    const double tr_bw = std::min(0.5 * (fs - fp), 0.5 * fs);
    const double fc = fs - tr_bw;
    const double beta = lsx_kaiser_beta(att, tr_bw * 0.5 / fc);
    const double a = ((0.0007528358 - 1.577737e-05 * beta) * beta + 0.6248022) * beta + 0.06186902;
    const int numtaps = (static_cast<int>(std::ceil(a / tr_bw + 1.0)) + 2) / 4 * 4 + 1;
    const int m = numtaps - 1;
    taps.resize(numtaps);
    const double i0_beta = std::cyl_bessel_i(0.0, beta);
    for (int i = 0; i < numtaps; ++i) {
        const double z = static_cast<double>(i) - 0.5 * m;
        const double x = z * kPi;
        const double sinc = (x == 0.0) ? fc : std::sin(fc * x) / x;
        const double y = z / (0.5 * m + 0.5);
        taps[i] = sinc * std::cyl_bessel_i(0.0, beta * std::sqrt(std::max(0.0, 1.0 - y * y))) / i0_beta;
    }
}

void hann_squared_2048(std::vector<double>& w) {
    // scipy get_window("hann", 2048, fftbins=True) squared (filters.window_sumsquare)
    const int n = 2048;
    w.resize(n);
    const double step = (2.0 * kPi) / static_cast<double>(n);
    for (int i = 0; i < n; ++i) {
        const double fac = static_cast<double>(i) * step + (-kPi);
        const double h = 0.5 + 0.5 * std::cos(fac);
        w[i] = h * h;
    }
}

}  // namespace serb
