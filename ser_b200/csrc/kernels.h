// Parameter blocks and launchers of the kernels in libser_b200 (shared by api.cu and the kernels).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "common.cuh"

namespace serb {

// ---- K1 stft_kernel.cu -------------------------------------------------------------------
struct StftParams {
    const float* wave;
    const ClipDev* clips;
    int n_clips;
    const int* tile_clip;    // [tiles] clip index of every 16-column tile
    const float2* tables;    // W_1024^(k1 n2) [32][32] then W_2048^k [1024] as (cos, sin)
    float* spill;            // [cols][kSpillStride]  |X|
    int do_peaks;            // piptrack wanted (chroma enabled)
    int kmin, kmax;          // bins with 150 <= f < min(4000, sr/2):  kmin <= k < kmax
    int peak_cap;            // slots per column
    float2* peaks;           // [cols][peak_cap] (mag, pitch)
    int* peak_count;         // [cols]
    double sr_over_nfft_num; // float(sr)
    int* status;             // bit 0: a non-finite sample was staged
};
cudaError_t configure_stft();
cudaError_t launch_expand_tiles(const ClipDev* clips, int n_clips, int* tile_clip, cudaStream_t stream);
cudaError_t launch_stft(const StftParams& p, int n_tiles, cudaStream_t stream);

// ---- K2/K3/K4 proj_kernels.cu ------------------------------------------------------------
struct TuneParams {
    const ClipDev* clips;
    const float2* peaks;      // [cols][peak_cap] (mag, pitch)
    const int* peak_count;    // [cols]
    int peak_cap;
    int bins_per_octave;      // 12 (chroma_stft) or 36 (chroma_cqt)
    const double* edges;      // np.linspace(-0.5, 0.5, 101)
    int* tuning_idx;          // [clips] -> 0..99 ; 50 (tuning 0.0) when no pitch was found
};
cudaError_t launch_tuning(const TuneParams& p, int n_clips, cudaStream_t stream);

struct ProjParams {
    const ClipDev* clips;
    int n_clips;
    const int* tile_clip;      // [tiles]
    const float* spill;        // [cols][kSpillStride]
    int do_mel;                // mel or mfcc requested
    const int* mel_start;      // [128]
    const int* mel_count;      // [128]
    const int* mel_offset;     // [129]
    const float* mel_weights;  // nnz
    int mel_nnz;
    float* logmel;             // [tiles][128][16]
    float* tile_mel;           // [tiles][128]  sum of mel power over the tile's columns
    float* tile_lmax;          // [tiles]       max of logmel over the tile
    int do_chroma;
    const float* chroma_banks; // [100][1025][12]
    const int* tuning_idx;     // [clips]
    float* tile_chroma;        // [tiles][12]   sum of normalised chroma over the tile's columns
};
cudaError_t configure_proj();
cudaError_t launch_proj(const ProjParams& p, int n_tiles, cudaStream_t stream);

struct PoolParams {
    const ClipDev* clips;
    const float* logmel;       // [tiles][128][16]
    const float* tile_mel;     // [tiles][128]
    const float* tile_lmax;    // [tiles]
    const float* tile_chroma;  // [tiles][12]
    const double* dct;         // [40][128]
    float* out;                // [rows][dim]
    int dim;
    int off_mfcc, off_chroma, off_mel, off_contrast;  // -1 when the group is disabled
};
cudaError_t launch_pool(const PoolParams& p, int n_clips, cudaStream_t stream);

// ---- short_kernel.cu ---------------------------------------------------------------------
struct ShortClip {
    long long start;
    int length;
    int out_row;
};
struct ShortParams {
    const float* wave;
    const ShortClip* clips;
    int sample_rate;
    const double* mel_points;  // [130]
    const double* edges;       // [101]
    const double* dct;         // [40][128]
    float* out;
    int dim;
    int off_mfcc, off_chroma, off_mel, off_contrast;
    int* tuning_idx;           // [n_short]
    int* status;
};
cudaError_t configure_short();
cudaError_t launch_short(const ShortParams& p, int n_clips, cudaStream_t stream);

// ---- mlp_kernel.cu -----------------------------------------------------------------------
struct MlpParams {
    int n_in, n_hidden, n_out, n_classes, out_activation;
    const double* mean;
    const double* scale;
    const double* w1;   // [n_in][n_hidden]
    const double* b1;
    const double* w2;   // [n_hidden][n_out]
    const double* b2;
    const float* x32;   // exactly one of x32 / x64 is non-null, [n][n_in]
    const double* x64;
    long long n;
    double* proba;      // [n][n_classes]
    int* label;         // [n]
};
size_t mlp_smem_bytes(int n_in, int n_hidden, int n_out);
cudaError_t configure_mlp(int n_in, int n_hidden, int n_out);
cudaError_t launch_mlp(const MlpParams& p, cudaStream_t stream);
cudaError_t launch_prepare_pcm16(const short* d_pcm, long long n, int* d_scratch_max, float* d_out,
                                 cudaStream_t stream);

}  // namespace serb
