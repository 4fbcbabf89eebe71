// Host-side construction of the constant tables the kernels read (float64 math, like numpy).
// Follows librosa 0.11.0 filters.mel / filters.chroma / scipy.fftpack.dct semantics as
// restated in SURVEY.md Appendix A.3-A.5; the reference reaches them through
// ser/_internal/utils/dsp.py:106-125.
#pragma once
#include <cstdint>
#include <vector>

namespace serb {

constexpr int kNMels = 128;
constexpr int kNMfcc = 40;
constexpr int kNChroma = 12;
constexpr int kNTunings = 100;

// np.linspace(-0.5, 0.5, 101)[idx]: the left edge of residual-histogram bin idx
double tuning_edge(int idx);

// Slaney mel points: mel_frequencies(n_mels + 2, fmin=0, fmax=sr/2) -> 130 doubles
void mel_points(int sample_rate, std::vector<double>& out);

// Dense Slaney-normalised mel filterbank, row-major [128 x (1 + n_fft/2)], float32
void mel_filterbank(int sample_rate, int n_fft, std::vector<float>& w);

// Dense chroma filterbank for one tuning, row-major [12 x (1 + n_fft/2)], float32
void chroma_filterbank(int sample_rate, int n_fft, double tuning, std::vector<float>& w);

// DCT-II "ortho" rows 0..39 over 128 inputs, row-major [40 x 128], float64
void dct_matrix(std::vector<double>& d);

// periodic Hann: 0.5 - 0.5 cos(2 pi n / N)
void hann_periodic(int n, std::vector<double>& w);

// CSR of the non-zero mel weights per band: band m covers bins [start[m], start[m] + count[m])
struct MelSparse {
    std::vector<int32_t> start;   // [128]
    std::vector<int32_t> count;   // [128]
    std::vector<int32_t> offset;  // [129] into weights
    std::vector<float> weights;   // nnz
};
void mel_sparse(const std::vector<float>& dense, int n_bins, MelSparse& out);

}  // namespace serb
