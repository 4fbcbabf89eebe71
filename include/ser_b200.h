/*
 * ser_b200.h -- C ABI of libser_b200.so: the B200-native (sm_100a) fast-profile
 * acoustic front-end + MLP of jsugg/ser.
 *
 * Plain pointers and sizes only; no torch / numpy types.  Every function returns an
 * int status unless noted:
 *     0   ok
 *    <0   invalid argument     -> the Python mirror raises ValueError
 *    >0   CUDA / runtime error -> the Python mirror raises RuntimeError (the reference
 *         maps RuntimeError to FastInferenceExecutionError,
 *         ser/_internal/runtime/fast_public_boundary.py:408-411)
 * and `serb_last_error(ctx)` returns the message.  One context per device; a context
 * owns its FFT twiddles, filterbanks (mel per sample rate, 100 tuning-indexed chroma banks
 * per sample rate, DCT matrix), scratch and streams.  Calls on one context are
 * serialised internally; use one context per thread / device for concurrency.
 * No function touches the Python GIL, so a ctypes caller may release it.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   serb_features_*      ser/_internal/utils/dsp.py:67-151 `extract_feature_from_signal`
 *                        (one call per clip there; a ragged batch of clips here) and its
 *                        callers ser/_internal/repr/handcrafted.py:65-107 `encode_sequence`
 *                        (sliding windows = overlapping clips of one buffer) and :124-137
 *                        `extract_vector`.
 *   serb_mlp_*           sklearn Pipeline(StandardScaler, MLPClassifier) `.predict` /
 *                        `.predict_proba` as built at
 *                        ser/_internal/models/training_support.py:87-106 and called at
 *                        ser/_internal/models/fast_path.py:48,181.
 *   serb_infer_host      ser/_internal/models/fast_path.py:147-226 arithmetic
 *                        (features -> labels + probabilities), timestamps and segment
 *                        merging stay on the host.
 *   serb_prepare_pcm16_host
 *                        ser/_internal/utils/audio_utils.py:28-60 `_prepare_audio_buffer`
 *                        for mono PCM16 input (x/32768, whole-file peak normalise).
 */
#ifndef SER_B200_H
#define SER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct serb_ctx serb_ctx;

/* feature-group bits, in output order (ser/_internal/utils/dsp.py:106-144) */
#define SERB_FLAG_MFCC      1u   /* 40  */
#define SERB_FLAG_CHROMA    2u   /* 12  */
#define SERB_FLAG_MEL       4u   /* 128 */
#define SERB_FLAG_CONTRAST  8u   /* 7, identically 0 on the reference path */
#define SERB_FLAG_TONNETZ   16u  /* 6   */
#define SERB_FLAG_ALL       31u

/* status codes */
#define SERB_OK                    0
#define SERB_ERR_INVALID_ARG      -1
#define SERB_ERR_SAMPLE_RATE      -2   /* "Sample rate must be a positive integer." */
#define SERB_ERR_EMPTY            -3   /* "Audio contains no samples." */
#define SERB_ERR_NOT_FINITE       -4   /* "Audio buffer is not finite everywhere." */
#define SERB_ERR_NYQUIST          -5   /* librosa ParameterError: band exceeds Nyquist (sr <= 12800 with contrast) */
#define SERB_ERR_NO_MODEL         -6
#define SERB_ERR_UNSUPPORTED      -7
#define SERB_ERR_CUDA              1
#define SERB_ERR_NO_DEVICE         2

/* MLP output activation (sklearn `out_activation_`) */
#define SERB_OUT_SOFTMAX   0
#define SERB_OUT_LOGISTIC  1

const char* serb_version(void);
/* number of CUDA devices visible, or 0 */
int serb_device_count(void);

int  serb_ctx_create(int device_ordinal, serb_ctx** out_ctx);
void serb_ctx_destroy(serb_ctx* ctx);
/* message of the last failing call on ctx (or of the last failing serb_ctx_create when ctx is NULL) */
const char* serb_last_error(const serb_ctx* ctx);

/* 40/12/128/7/6 summed over the set bits */
int serb_feature_dim(uint32_t flag_bits);

/*
 * Ragged batch feature extraction.  Clip i is wave[starts[i] : starts[i] + lengths[i]]
 * (clips may overlap: the sliding windows of encode_sequence share one buffer).
 * out is [n_clips x serb_feature_dim(flag_bits)] float32, row-major.
 *
 *   _device: d_wave and d_out are device pointers on ctx's device; starts/lengths are
 *            host arrays; stream is a cudaStream_t (NULL = ctx's own stream).  Returns
 *            after enqueueing; the caller synchronises the stream.
 *   _host:   everything is host memory (pinned or pageable); the call stages the
 *            waveform to the device in pieces, overlapping copy and compute, and returns
 *            when h_out is complete.
 */
int serb_features_device(serb_ctx* ctx, const float* d_wave, int64_t n_wave,
                         const int64_t* starts, const int64_t* lengths, int64_t n_clips,
                         int32_t sample_rate, uint32_t flag_bits, float* d_out, void* stream);
int serb_features_host(serb_ctx* ctx, const float* h_wave, int64_t n_wave,
                       const int64_t* starts, const int64_t* lengths, int64_t n_clips,
                       int32_t sample_rate, uint32_t flag_bits, float* h_out);

/*
 * The same for clips that live in separate host arrays (one per decoded file, as the training
 * loader holds them): h_clips[i] points at lengths[i] float32 samples.  The library copies them
 * piecewise while earlier chunks compute; the caller builds no packed buffer.
 * (ser/_internal/data/data_loader.py:485-529, one extract_vector call per file there.)
 */
int serb_features_host_clips(serb_ctx* ctx, const float* const* h_clips, const int64_t* lengths,
                             int64_t n_clips, int32_t sample_rate, uint32_t flag_bits, float* h_out);

/*
 * serb_features_device only enqueues; this synchronises `stream` (NULL = ctx's own) and reports what
 * the chain found: SERB_ERR_NOT_FINITE ("Audio buffer is not finite everywhere.", dsp.py:94-95)
 * when any staged sample was NaN/Inf -- the rows are garbage then -- else SERB_OK.  The host
 * entries do this check themselves.  Scratch is per context: launch chains issued on different
 * streams of one context are ordered after one another by the library.
 */
int serb_features_device_check(serb_ctx* ctx, void* stream);

/*
 * The same ragged batch straight from PCM16 files: file f is h_files[f], a buffer
 * of file_frames[f] frames x file_channels[f] interleaved int16 (file_channels NULL = all mono).
 * The device does what ser/_internal/utils/audio_utils.py:28-60 `_prepare_audio_buffer` does after
 * librosa.load / soundfile decode PCM16 -- float32 x / 32768, channel mean, whole-FILE peak
 * normalisation x / max|x| (an all-zero file stays zero) -- bit-identically, so the host ships
 * 2 bytes per sample.  Clip i is frames [clip_starts[i], clip_starts[i] + clip_lengths[i]) of
 * file clip_file[i]; clip_file must be non-decreasing.  Files contiguous in host memory are
 * copied as one transfer; copies of later files overlap the kernels of earlier chunks.
 * serb_infer_host_pcm16 adds the classifier (h_features may be NULL).
 */
int serb_features_host_pcm16(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames,
                             const int32_t* file_channels, int64_t n_files, const int64_t* clip_file,
                             const int64_t* clip_starts, const int64_t* clip_lengths, int64_t n_clips,
                             int32_t sample_rate, uint32_t flag_bits, float* h_out);
int serb_infer_host_pcm16(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames,
                          const int32_t* file_channels, int64_t n_files, const int64_t* clip_file,
                          const int64_t* clip_starts, const int64_t* clip_lengths, int64_t n_clips,
                          int32_t sample_rate, uint32_t flag_bits, float* h_features, double* h_proba,
                          int32_t* h_label_index);

/*
 * Classifier weights (float64, row-major as sklearn stores them):
 * mean/scale [n_in], w1 [n_in x n_hidden], b1 [n_hidden], w2 [n_hidden x n_out], b2 [n_out].
 * n_out is the width of the output layer (1 for sklearn's binary logistic case).
 */
int serb_mlp_load(serb_ctx* ctx, int32_t n_in, int32_t n_hidden, int32_t n_out,
                  const double* mean, const double* scale,
                  const double* w1, const double* b1, const double* w2, const double* b2,
                  int32_t out_activation);
/* number of probability columns predict writes: n_out, or 2 for the binary logistic case */
int serb_mlp_n_classes(const serb_ctx* ctx);

/*
 * Fused scaler + MLP forward in float64: x [n x n_in] -> proba [n x n_classes],
 * label_index [n] = sklearn's predict() index into classes_.  x is float64 on the host
 * entry (what Pipeline.predict receives) and float32 on the device entry (the feature
 * rows produced by serb_features_device, widened on the fly as fast_path.py:180 does).
 */
int serb_mlp_predict_host(serb_ctx* ctx, const double* h_x, int64_t n,
                          double* h_proba, int32_t* h_label_index);
int serb_mlp_predict_device(serb_ctx* ctx, const float* d_x, int64_t n,
                            double* d_proba, int32_t* d_label_index, void* stream);

/*
 * features + predict in one call with host buffers (the bench's end-to-end path).
 * h_features may be NULL when the caller only wants predictions.
 */
int serb_infer_host(serb_ctx* ctx, const float* h_wave, int64_t n_wave,
                    const int64_t* starts, const int64_t* lengths, int64_t n_clips,
                    int32_t sample_rate, uint32_t flag_bits,
                    float* h_features, double* h_proba, int32_t* h_label_index);

/*
 * Segmented pooling of frame embeddings [n_frames x dim] float32 over frame ranges [lo[w], hi[w]):
 * mode 0 float64 mean, mode 1 float64 mean then std (ddof 0), out [n_windows x 2 dim]
 * (ser/_internal/pool/stats_pool.py:15-43 mean_std_pool), mode 2 float32 mean widened to float64
 * (ser/_internal/repr/handcrafted.py:109-122 HandcraftedBackend.pool).  Row-order summation:
 * bit-identical to numpy.  The host computes the ranges from the timestamps
 * (ser/_internal/repr/backend.py:81-111 overlap_frame_mask selects a contiguous run).
 */
int serb_pool_frames_host(serb_ctx* ctx, const float* h_embeddings, int64_t n_frames, int32_t dim,
                          const int32_t* h_lo, const int32_t* h_hi, int64_t n_windows, int32_t mode,
                          double* h_out);

/* mono PCM16 -> float32 (x / 32768) peak-normalised over the whole buffer, on the device */
int serb_prepare_pcm16_host(serb_ctx* ctx, const int16_t* h_pcm, int64_t n, float* h_out);
int serb_prepare_pcm16_device(serb_ctx* ctx, const int16_t* d_pcm, int64_t n, float* d_out, void* stream);
/* the preparation step of serb_features_host_pcm16 on its own: file f (file_frames[f] frames x
 * file_channels[f] interleaved int16, file_channels NULL = mono) -> h_out[f][file_frames[f]] float32,
 * exactly `_prepare_audio_buffer(decode(file))` (audio_utils.py:28-60) */
int serb_prepare_pcm16_files_host(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames,
                                  const int32_t* file_channels, int64_t n_files, float* const* h_out);

/* ---- introspection (tests, profiling); not part of the drop-in surface ---- */

/* kind 0: mel [128 x (1+n_fft/2)], kind 1: chroma [12 x (1+n_fft/2)] for tuning index 0..99
 * (tuning = -0.5 + 0.01*index), kind 2: DCT-II ortho [40 x 128] as float64 in out (cast),
 * kind 3: periodic Hann [n_fft].  Host-side computation only; needs no GPU. */
int serb_debug_filterbank(int32_t kind, int32_t sample_rate, int32_t n_fft, int32_t tuning_index,
                          float* out);
/* magnitude STFT |X| of one clip (n_fft = 2048, hop 512, centred): out [n_cols x 1025] */
int serb_debug_stft_host(serb_ctx* ctx, const float* h_wave, int64_t n, float* h_out, int64_t n_cols);
/* per-clip tuning index (0..99) chosen by the last serb_features_* call; -1 if chroma was off */
int serb_debug_last_tuning(serb_ctx* ctx, int32_t* h_out, int64_t n_clips);
/* intermediates of the tonnetz chain for ONE clip (any pointer may be NULL): h_yharm
 * [max(n, 512)] = librosa.effects.harmonic(y); tuning_index = estimate_tuning(harmonic, 36 bins per
 * octave) as a bin 0..99; h_cqmag [cq_cols x 252] = |cqt| / sqrt(length), rows = columns; h_tonnetz6. */
int serb_debug_tonnetz_stages(serb_ctx* ctx, const float* h_wave, int64_t n, int32_t sample_rate,
                              float* h_yharm, int32_t* tuning_index, float* h_cqmag,
                              int64_t cq_capacity_rows, int32_t* cq_cols, float* h_tonnetz6);
/* constant-Q plan of the tonnetz chain at one sample rate: out10 = {status (0 ok, 1 Nyquist,
 * 2 unsupported), early downsampling factor, top-octave hop, n_fft of octaves 0..6}. Host only. */
int serb_debug_cqt_plan(int32_t sample_rate, int32_t* out10);
/* sparsified FFT-domain wavelet basis of one octave (0 = top) for tuning index 0..99:
 * out_basis [36 x (1 + n_fft/2) x 2] complex64 (may be NULL), out_scale36 = 1/sqrt(length) per
 * row (may be NULL).  Host only. */
int serb_debug_cqt_basis(int32_t sample_rate, int32_t tuning_index, int32_t octave, float* out_basis,
                         float* out_scale36);
/* the same basis as cqtc_kernel holds it (rows in sets over the union of their bins, csrc/cqt_tables.h
 * CqtSetBank), expanded back to [36 x (1 + n_fft/2) x 2]: must equal serb_debug_cqt_basis.  Host only;
 * SERB_ERR_UNSUPPORTED when the basis does not fit the layout (the lane = row kernels run then). */
int serb_debug_cqt_set_basis(int32_t sample_rate, int32_t tuning_index, int32_t octave, float* out_basis,
                             float* out_scale36);
/* decimation filter (soxr_hq stand-in) for an integer factor 2..8; returns the tap count
 * (> 0) and, when out is non-NULL and capacity suffices, the taps.  Host only. */
int serb_debug_decimation_taps(int32_t factor, double* out, int32_t capacity);
/* FP32 ceiling of this GPU right now: dependent-free FFMA chains on every SM, best of four timed
 * launches, in TFLOP/s (the denominator of bench.py's whole-step FP32 fraction) */
int serb_debug_fp32_peak(serb_ctx* ctx, double* tflops);
/* kernels launched by this context since creation */
int64_t serb_debug_launch_count(const serb_ctx* ctx);
/* per-kernel CUDA-event timing: kinds 0 stft, 1 tuning, 2 proj, 3 pool, 4 short, 5 mlp, 6 hpss_harm,
 * 7 hpss_perc, 8 istft, 9 ola, 10 decimations, 11 constant-Q octaves, 12 tonnetz, 13 PCM16 preparation.
 * set_profile(1) brackets every launch with an event pair (and resets the totals);
 * kernel_ms returns the accumulated device time and launch count of one kind. */
int serb_debug_set_profile(serb_ctx* ctx, int32_t enabled);
int serb_debug_kernel_ms(serb_ctx* ctx, int32_t kind, double* total_ms, int64_t* n_launches);
/* device-side milliseconds of the last serb_features_device/_host compute chain (CUDA events) */
float serb_debug_last_compute_ms(serb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SER_B200_H */
