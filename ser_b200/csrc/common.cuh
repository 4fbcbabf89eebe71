// Shared device-side definitions for libser_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace serb {

constexpr int kNFft = 2048;          // librosa default, dsp.py:96 n_fft = min(len, 2048)
constexpr int kHop = 512;            // n_fft // 4 (stft) == melspectrogram's hop_length
constexpr int kNBins = 1025;         // 1 + n_fft / 2
constexpr int kSpillStride = 1032;   // |X| spill row pitch in floats: [column][1032], 16-byte multiple
constexpr int kColsPerTile = 16;     // STFT columns per projection tile
constexpr int kHalfTileCols = 8;     // STFT columns per STFT work item (half a projection tile)

// One clip of the ragged batch, chunk-relative bookkeeping included.
struct ClipDev {
    long long start;     // first sample in the waveform buffer
    int length;          // samples (>= 2048 on the main path)
    int n_cols;          // 1 + length / 512
    int col_base;        // first column of this clip inside the chunk's scratch
    int tile_base;       // first tile of this clip inside the chunk
    int out_row;         // row of the output matrix
    int pad_;
};


__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace serb
