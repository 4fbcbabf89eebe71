// Host-side construction of the tables behind the tonnetz chain: the constant-Q plan and
// sparse FFT-domain wavelet bases of librosa.cqt as chroma_cqt calls it (252 bins, 36 per
// octave, fmin = C1 * 2^(tuning/36), filter_scale 1, norm 1, sparsity 0.01, Hann), and the
// linear-phase decimation filters standing in for soxr_hq.  Reference call site:
// ser/_internal/utils/dsp.py:138-143 (librosa.feature.tonnetz(y=harmonic, sr)); librosa
// 0.11.0 semantics per SURVEY.md Appendix A.9-A.10.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace serb {

constexpr int kCqtOctaves = 7;
constexpr int kCqtBpo = 36;
constexpr int kCqtBins = kCqtOctaves * kCqtBpo;   // 252
constexpr int kCqtRowCap = 32;                    // widest run of non-zero bins a basis row may span

struct CqtPlan {
    int sample_rate = 0;
    int status = 0;            // 0 ok, 1 top wavelet exceeds Nyquist (librosa ParameterError), 2 unsupported
    std::string message;
    int early_factor = 1;      // 2^k early downsampling (librosa __early_downsample)
    int hop0 = 512;            // hop of the top octave after early downsampling
    int n_fft[kCqtOctaves];    // FFT size per octave (top octave first)
};

// 16-byte row descriptor: bins [start, start + count) of the octave's rfft, then
// |sum| * scale is the entry of the scaled constant-Q magnitude
struct CqtRow {
    int32_t start;
    int32_t count;
    float scale;               // 1 / sqrt(wavelet length)
    int32_t bin;               // output bin 0..251 (low to high frequency)
};

struct CqtBank {
    std::vector<CqtRow> rows;  // [7][36], octave-major (top octave first); inside an octave in the
                               // kernel's lane order (see cqt_bank), each row naming its output bin
    std::vector<float> vals;   // [7][36][kCqtRowCap][2] complex64, zero padded
};

// The plan must not depend on the tuning estimate (it fixes buffer shapes before the tuning is
// known); status 2 reports the sample rates where it would.
void cqt_plan(int sample_rate, CqtPlan& plan);
// basis for tuning = -0.5 + 0.01 * tuning_idx.  Returns false if a row is wider than kCqtRowCap.
// lane_order: rows of an octave are permuted (and started a few bins early, zero filled) so that
// the kernel's lane = row reads of the spectrum are shared-memory bank-conflict free; otherwise
// rows stay in bin order.
bool cqt_bank(const CqtPlan& plan, int tuning_idx, CqtBank& bank, bool lane_order = true);
// Column-mapped layout of the same rows for cqtc_kernel (lane = column): the 36 rows of an octave in
// frequency order, cut into 16 sets of consecutive rows -- four sets of three (the narrow low rows),
// twelve sets of two -- each stored over the union of its rows' bins as [bin][2 or 4 rows] complex64
// (zero where a row does not reach), so that one spectrum value feeds every row of the set and the
// basis values are warp-uniform loads.  u0 is relative to bin_lo, the lowest bin any row reads.
constexpr int kCqtSets = 16;
constexpr int kCqtSetValCap = 1536;               // complex values per (tuning, octave) record
constexpr int kCqtSetMaxBins = 128;               // bins [bin_lo, bin_lo + n_bins) the kernel keeps per column
struct CqtSet {
    int16_t u0, ulen, off, nrows;                 // first bin - bin_lo, bins, first value in vals[], rows (2 or 3)
    int16_t bin[4];                               // output bins 0..251 (-1: padding row)
    float scale[4];                               // 1 / sqrt(wavelet length)
};
struct CqtSetBank {
    int32_t bin_lo, n_bins, reserved[2];
    CqtSet sets[kCqtSets];
    float vals[kCqtSetValCap * 2];
};
// one record per octave (out7[0..6]); false if the basis does not fit the layout's capacities
bool cqt_set_banks(const CqtPlan& plan, int tuning_idx, CqtSetBank* out7);
// dense basis of one octave, [36][1 + n_fft/2] complex64 (tests)
void cqt_basis_dense(const CqtPlan& plan, int tuning_idx, int octave, std::vector<float>& out);

// Kaiser-windowed sinc to soxr "HQ" spec (pass-band 0.913, 21 bits), odd length, unit DC gain
void decimation_taps(int factor, std::vector<double>& taps);

// hann(2048, periodic)^2 in float64 (window sum-of-squares of the inverse STFT)
void hann_squared_2048(std::vector<double>& w);

}  // namespace serb
