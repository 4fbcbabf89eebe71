// Fused StandardScaler + MLP(300, relu) + softmax / logistic forward in float64.
//
// Replaces sklearn Pipeline.predict / predict_proba on the fast path
// (ser/_internal/models/fast_path.py:48,181; model built at
// ser/_internal/models/training_support.py:87-106):
//   z = (x - mean_) / scale_ ; h = relu(z W1 + b1) ; o = h W2 + b2 ;
//   softmax(o - max o)  |  [1 - expit(o), expit(o)] ;  label = first arg-max (or o > 0.5).
// float64 keeps near-tie labels identical to sklearn's; the work is ~120 kflop per row.
#include "common.cuh"
#include "kernels.h"

namespace serb {

constexpr int kMlpRows = 8;        // rows per CTA
constexpr int kMlpThreads = 320;


__global__ void __launch_bounds__(kMlpThreads) mlp_kernel(MlpParams p) {
    extern __shared__ double smem[];
    double* z = smem;                                   // [kMlpRows][n_in]
    double* h = z + kMlpRows * p.n_in;                  // [kMlpRows][n_hidden]
    double* o = h + kMlpRows * p.n_hidden;              // [kMlpRows][n_out]
    const long long row0 = static_cast<long long>(blockIdx.x) * kMlpRows;
    const int rows = static_cast<int>(min(static_cast<long long>(kMlpRows), p.n - row0));
    const int tid = threadIdx.x;

    for (int i = tid; i < rows * p.n_in; i += kMlpThreads) {
        const int r = i / p.n_in, c = i - r * p.n_in;
        const long long src = (row0 + r) * p.n_in + c;
        const double x = p.x64 ? p.x64[src] : static_cast<double>(p.x32[src]);
        z[i] = (x - p.mean[c]) / p.scale[c];
    }
    __syncthreads();
    for (int j = tid; j < p.n_hidden; j += kMlpThreads) {
        double acc[kMlpRows];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) acc[r] = 0.0;
        for (int i = 0; i < p.n_in; ++i) {
            const double w = p.w1[static_cast<long long>(i) * p.n_hidden + j];
#pragma unroll
            for (int r = 0; r < kMlpRows; ++r) acc[r] = fma(z[r * p.n_in + i], w, acc[r]);
        }
        const double b = p.b1[j];
#pragma unroll
        for (int r = 0; r < kMlpRows; ++r) h[r * p.n_hidden + j] = fmax(acc[r] + b, 0.0);
    }
    __syncthreads();
    for (int i = tid; i < rows * p.n_out; i += kMlpThreads) {
        const int r = i / p.n_out, c = i - r * p.n_out;
        double acc = 0.0;
        for (int j = 0; j < p.n_hidden; ++j) acc = fma(h[r * p.n_hidden + j], p.w2[j * p.n_out + c], acc);
        o[i] = acc + p.b2[c];
    }
    __syncthreads();
    if (tid < rows) {
        const int r = tid;
        double* orow = o + r * p.n_out;
        double* dst = p.proba + (row0 + r) * p.n_classes;
        if (p.out_activation == 0) {  // softmax
            double mx = orow[0];
            for (int c = 1; c < p.n_out; ++c) mx = fmax(mx, orow[c]);
            double sum = 0.0;
            for (int c = 0; c < p.n_out; ++c) { orow[c] = exp(orow[c] - mx); sum += orow[c]; }
            int best = 0;
            double bestv = -1.0;
            for (int c = 0; c < p.n_out; ++c) {
                const double v = orow[c] / sum;
                dst[c] = v;
                if (v > bestv) { bestv = v; best = c; }
            }
            p.label[row0 + r] = best;
        } else if (p.n_out == 1) {    // binary logistic: proba = [1 - p, p], label = p > 0.5
            const double v = 1.0 / (1.0 + exp(-orow[0]));
            dst[0] = 1.0 - v;
            dst[1] = v;
            p.label[row0 + r] = v > 0.5 ? 1 : 0;
        } else {                      // multilabel logistic (not produced by the reference's trainer)
            int best = 0;
            double bestv = -1.0;
            for (int c = 0; c < p.n_out; ++c) {
                const double v = 1.0 / (1.0 + exp(-orow[c]));
                dst[c] = v;
                if (v > bestv) { bestv = v; best = c; }
            }
            p.label[row0 + r] = best;
        }
    }
}

size_t mlp_smem_bytes(int n_in, int n_hidden, int n_out) {
    return sizeof(double) * static_cast<size_t>(kMlpRows) * (n_in + n_hidden + n_out);
}

cudaError_t configure_mlp(int n_in, int n_hidden, int n_out) {
    return cudaFuncSetAttribute(mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(mlp_smem_bytes(n_in, n_hidden, n_out)));
}

cudaError_t launch_mlp(const MlpParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    const int grid = static_cast<int>((p.n + kMlpRows - 1) / kMlpRows);
    mlp_kernel<<<grid, kMlpThreads, mlp_smem_bytes(p.n_in, p.n_hidden, p.n_out), stream>>>(p);
    return cudaGetLastError();
}

// -----------------------------------------------------------------------------------------
// PCM16 -> float32 with whole-buffer peak normalisation
// (ser/_internal/utils/audio_utils.py:28-60 for mono PCM16 input).
// -----------------------------------------------------------------------------------------
__global__ void pcm_absmax_kernel(const short* __restrict__ pcm, long long n, int* __restrict__ out_max) {
    int m = 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int v = pcm[i];
        m = max(m, v < 0 ? -v : v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out_max, m);
}

__global__ void pcm_scale_kernel(const short* __restrict__ pcm, long long n, const int* __restrict__ max_abs,
                                 float* __restrict__ out) {
    // x = pcm / 32768 (float32), then x / float(max|x|): the divisor is a Python float (float64
    // holding a float32 value) applied to a float32 array, i.e. a float32 division.
    const int m = *max_abs;
    const float peak = static_cast<float>(m) / 32768.0f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float x = static_cast<float>(pcm[i]) / 32768.0f;
        out[i] = (m == 0) ? 0.0f : __fdiv_rn(x, peak);
    }
}

// ---- segmented mean / std pooling over frame ranges -----------------------------------------
// Replaces ser/_internal/pool/stats_pool.py:15-43 `mean_std_pool` (and the mean-only pool of
// ser/_internal/repr/handcrafted.py:109-122): window w pools frames [lo[w], hi[w]).  One CTA per
// window, threads over the feature dimension (coalesced rows); each (window, feature) is summed
// in row order in float64, which is the order numpy reduces axis 0 of a C-contiguous matrix in,
// so mean and std (ddof = 0) come out bit-identical to numpy's.
__global__ void __launch_bounds__(128) pool_stats_kernel(const float* __restrict__ emb, int dim,
                                                         const int* __restrict__ lo, const int* __restrict__ hi,
                                                         int want_std, double* __restrict__ out) {
    const int w = blockIdx.x;
    const int a = lo[w], b = hi[w];
    const int width = want_std ? 2 * dim : dim;
    const double n = static_cast<double>(b - a);
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        double sum = 0.0;
        for (int r = a; r < b; ++r) sum += static_cast<double>(emb[static_cast<size_t>(r) * dim + d]);
        const double mean = sum / n;
        out[static_cast<size_t>(w) * width + d] = mean;
        if (want_std) {
            double ss = 0.0;
            for (int r = a; r < b; ++r) {
                const double x = static_cast<double>(emb[static_cast<size_t>(r) * dim + d]) - mean;
                ss = __dadd_rn(ss, __dmul_rn(x, x));     // numpy squares, then sums: no FMA contraction
            }
            out[static_cast<size_t>(w) * width + dim + d] = sqrt(ss / n);
        }
    }
}

// mean-only pooling in float32, as `embeddings[mask].mean(axis=0)` computes it on a float32 matrix
__global__ void __launch_bounds__(128) pool_mean_f32_kernel(const float* __restrict__ emb, int dim,
                                                            const int* __restrict__ lo, const int* __restrict__ hi,
                                                            double* __restrict__ out) {
    const int w = blockIdx.x;
    const int a = lo[w], b = hi[w];
    const float n = static_cast<float>(b - a);
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float sum = 0.0f;
        for (int r = a; r < b; ++r) sum = __fadd_rn(sum, emb[static_cast<size_t>(r) * dim + d]);
        out[static_cast<size_t>(w) * dim + d] = static_cast<double>(__fdiv_rn(sum, n));
    }
}

cudaError_t launch_pool_stats(const float* d_emb, int dim, const int* d_lo, const int* d_hi, long long n_windows,
                              int mode, double* d_out, cudaStream_t stream) {
    if (n_windows <= 0 || dim <= 0) return cudaSuccess;
    if (mode == 2) pool_mean_f32_kernel<<<static_cast<unsigned>(n_windows), 128, 0, stream>>>(d_emb, dim, d_lo, d_hi, d_out);
    else pool_stats_kernel<<<static_cast<unsigned>(n_windows), 128, 0, stream>>>(d_emb, dim, d_lo, d_hi, mode == 1 ? 1 : 0, d_out);
    return cudaGetLastError();
}

cudaError_t launch_prepare_pcm16(const short* d_pcm, long long n, int* d_scratch_max, float* d_out,
                                 cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(d_scratch_max, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    if (n <= 0) return cudaSuccess;
    const int threads = 256;
    long long want = (n + threads - 1) / threads;
    if (want > 148LL * 16) want = 148LL * 16;
    const int grid = static_cast<int>(want);
    pcm_absmax_kernel<<<grid, threads, 0, stream>>>(d_pcm, n, d_scratch_max);
    pcm_scale_kernel<<<grid, threads, 0, stream>>>(d_pcm, n, d_scratch_max, d_out);
    return cudaGetLastError();
}

}  // namespace serb
