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
    float* spill;            // [cols][kSpillStride]  |X|, or NULL when only the peaks are wanted
    float2* cspill;          // [cols][kSpillStride]  X (complex), or NULL (tonnetz chain: HPSS needs the phase)
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
constexpr int kTuneLongCols = 1024;   // clips with more STFT columns take the multi-CTA tuning path
struct TuneLongState {
    unsigned hist[2][2048];   // pass histograms for the two middle ranks
    unsigned prefix[2];       // selected key bits so far (zero between uses)
    long long rank[2];
    long long n;              // peaks of the clip
    int counts[100];          // residual histogram
};
struct TuneParams {
    const ClipDev* clips;
    const float2* peaks;      // [cols][peak_cap] (mag, pitch)
    const int* peak_count;    // [cols]
    int peak_cap;
    int bins_per_octave;      // 12 (chroma_stft) or 36 (chroma_cqt)
    const double* edges;      // np.linspace(-0.5, 0.5, 101)
    int* tuning_idx;          // [clips] -> 0..99 ; 50 (tuning 0.0) when no pitch was found
    const int* long_clips;    // [n_long] indices (into clips) of the clips longer than kTuneLongCols
    int n_long;
    int max_long_cols;
    TuneLongState* long_state;   // [n_long], zero-initialised prefixes
};
cudaError_t launch_tuning(const TuneParams& p, int n_clips, cudaStream_t stream, long long* launches);

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

// ---- tonnetz chain: hpss_kernels.cu, cqt_kernels.cu ----------------------------------------
constexpr int kCqOctaves = 7;
constexpr int kCqRows = 36;       // bins per octave
constexpr int kCqBins = 252;
constexpr int kCqRowCap = 32;     // complex values stored per basis row
constexpr int kDecTaps2 = 389;    // libsoxr HQ 2:1 low-pass, restated (cqt_tables.cpp decimation_taps); 1 mod 4
constexpr int kDecExactBelow = 2048;   // clips shorter than this decimate in float64 (cqt_kernels.cu)

struct TonClip {
    long long hoff;      // harmonic signal of the clip: yharm[hoff .. hoff + length)
    long long off0;      // level-0 constant-Q signal: yoct[off0 .. off0 + len0); level l at level_base[l] + (off0 >> l)
    int length;          // samples
    int n_cols;          // STFT columns, 1 + length / 512
    int col_base;        // first STFT column inside the chunk scratch
    int tile_base;       // first 16-column tile inside the chunk
    int len0;            // ceil(length / early_factor)
    int cq_cols;         // constant-Q columns kept (min over octaves, librosa __trim_stack)
    int cq_base;         // first row of the clip in cqmag
    int out_row;
    int part_base;       // first kTonTile-column partial-sum slot of the clip (tonnetz reduction)
    int pad_;
};
constexpr int kTonTile = 256;     // constant-Q columns per tonnetz partial sum

struct HpssParams {
    const TonClip* clips;
    const int2* segs;        // (clip index, first column) per time-median work item
    int seg_len;             // columns per time-median work item
    float one;               // 1.0f, passed at run time so the predicated multiply-by-one of the median update stays a multiply
    const float* mag;        // [cols][kSpillStride] |X|
    float* perc;             // [cols][kSpillStride] median along frequency
    float2* cspec;           // [cols][kSpillStride] X (not touched by the median kernels)
};
struct IstftParams {
    const float2* cspec;     // [cols][kSpillStride] X
    const float* mask;       // [cols][kSpillStride] soft mask (hpss_harm_kernel's output)
    const float2* tables;    // FFT twiddles (same layout as StftParams::tables)
    float* frames;           // [cols][2048] windowed inverse-FFT frames
};
struct OlaParams {
    const TonClip* clips;
    const int* tile_clip;    // [tiles]
    const float* frames;
    const double* hann_sq;   // [2048] squared periodic Hann, float64
    const float* wss4;       // [512] window sum of squares where four frames overlap (n mod 512)
    float* yharm;
};
cudaError_t configure_hpss();
cudaError_t launch_hpss_harm(const HpssParams& p, int n_segs, cudaStream_t stream);
cudaError_t launch_hpss_perc(const HpssParams& p, int n_cols, cudaStream_t stream);
cudaError_t launch_istft(const IstftParams& p, int n_cols, cudaStream_t stream);
cudaError_t launch_ola(const OlaParams& p, int n_tiles, cudaStream_t stream);
// inverse STFT and overlap-add in one kernel: runs = (clip, first column) per kIstftRun columns
constexpr int kIstftRun = 48;
cudaError_t launch_istft_ola(const IstftParams& p, const OlaParams& o, const int2* runs, int n_runs, int n_sms,
                             cudaStream_t stream);

struct CqRow { int start; int count; float scale; int bin; };
// column-mapped rows of one (tuning, octave) for cqtc_kernel: same layout as CqtSetBank (cqt_tables.h)
constexpr int kCqSets = 16;
constexpr int kCqSetValCap = 1536;
constexpr int kCqSetMaxBins = 128;
struct CqSet { short u0, ulen, off, nrows; short bin[4]; float scale[4]; };
struct CqSetBank { int bin_lo, n_bins, reserved[2]; CqSet sets[kCqSets]; float2 vals[kCqSetValCap]; };
struct CqtParams {
    const TonClip* clips;
    int n_clips;
    const int* tuning_idx;       // [clips] 0..99 (36 bins per octave)
    const float* yharm;          // full-rate harmonic signals
    float* yoct;                 // decimated signals, all levels
    long long level_base[kCqOctaves];
    int early_factor;            // 1, 2, 4, 8
    int hop0;                    // hop of level 0
    int n_fft[kCqOctaves];
    int cq_cols_per_block[kCqOctaves];   // filled by the launcher
    int cq_sub_cols[kCqOctaves];         // columns per first-stage table of the shared-stage kernel (launcher)
    int cq_block_end[kCqOctaves];        // cqtc_kernel: blockIdx.y < cq_block_end[i] belongs to the launch's i-th octave (launcher)
    const float* early_taps;     // [n_early_taps] (scaled by sqrt(early_factor)); NULL if factor 1
    int n_early_taps;
    const CqRow* rows;           // [100][7][36]
    const float2* vals;          // [100][7][36][kCqRowCap]
    const CqSetBank* set_banks;  // [100][7] column-mapped rows (n_fft 1024 octaves take cqtc_kernel); NULL: lane = row kernels only
    const float2* twiddles;      // W_N^j (cos, -sin), j < N, for N = 128, 256, 512, 1024 back to back, then
                                 // (cos, sin) 2 pi k / (2N), k < N, for the same N
    float* cqmag;                // [cq rows][252] scaled magnitudes: debug output only, NULL on the product path
    float* cq_chroma;            // [cq rows][7 octaves][12]: every octave's share of the chroma fold
    double* ton_part;            // [partial slots][6] tonnetz sums per kTonTile columns
    float* out;                  // [rows][dim]
    int dim, off_tonnetz;
    int max_len0;                // longest level-0 signal in the chunk
    int max_cq_cols;             // most constant-Q columns of one clip
    int max_length;              // longest full-rate signal in the chunk
    int n_dec_exact;             // clips shorter than kDecExactBelow samples (float64 decimation)
    int cqt_no_shared;           // 1: every octave takes the per-column transform (SERB_CQT=percolumn, A/B tests)
    int cqtc_shared_max_hop;    // cqtc_kernel: octaves with a hop up to this share the first FFT stage (SERB_CQT_SHARED_MAXHOP)
    const void* dec_toeplitz;    // bf16 Toeplitz operand of decimate2_mma_kernel (decimate_mma_table)
    int n_sms;
};
cudaError_t configure_cqt(const float* taps2_scaled, const double* taps2_scaled_f64);   // factor-2 taps (x sqrt 2) to constant memory
cudaError_t launch_decimations(const CqtParams& p, cudaStream_t stream, long long* launches);
// decimate_mma.cu: the factor-2 decimation on tcgen05 (clips of at least kDecExactBelow samples)
size_t decimate_mma_table_bytes();
void decimate_mma_table(const double* taps2_scaled, unsigned char* out);
cudaError_t configure_decimate_mma();
cudaError_t launch_decimate2_mma(const CqtParams& p, int src_level, int max_len_in, const void* d_toeplitz, int n_sms,
                                 cudaStream_t stream);
cudaError_t launch_cqt_octaves(const CqtParams& p, cudaStream_t stream, long long* launches);
cudaError_t launch_tonnetz(const CqtParams& p, cudaStream_t stream);

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
// mode 0: float64 mean, 1: float64 mean + std (out [n][2 dim]), 2: float32 mean widened to float64
cudaError_t launch_pool_stats(const float* d_emb, int dim, const int* d_lo, const int* d_hi, long long n_windows,
                              int mode, double* d_out, cudaStream_t stream);
cudaError_t launch_prepare_pcm16(const short* d_pcm, long long n, int* d_scratch_max, float* d_out,
                                 cudaStream_t stream);

// ---- pcm_kernels.cu (N1: PCM16 files prepared on the device) --------------------------------
struct PcmFile {
    long long pcm_off;    // first int16 of the file in the staged PCM buffer (multiple of 8)
    long long wave_off;   // first float of the prepared mono signal in the waveform buffer (multiple of 4)
    long long frames;     // samples per channel
    int channels;         // interleaved channels (1..256)
    int pad_;
};
// files [file_lo, file_hi): x / 32768, channel mean, per-file peak normalisation -> d_wave
cudaError_t launch_pcm_prepare_files(const short* d_pcm, const PcmFile* d_files, int file_lo, int file_hi,
                                     long long max_frames, int* d_peak_bits, float* d_wave, cudaStream_t stream,
                                     long long* launches);

// blocks x 256 threads x 16 chains x iters FMAs
cudaError_t launch_fp32_peak(float* d_scratch, int blocks, int iters, cudaStream_t stream);

}  // namespace serb
