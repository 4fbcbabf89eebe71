// tcgen05 decimator (ser_b200/csrc/decimate_mma.cu) against a float64 FIR and against the FFMA2
// kernel it replaces: accuracy per level of the recursion, bit-identity across batch positions,
// and time per level on a c2-shaped chunk.  Built by scripts/microbench/build_decimate_mma_test.sh:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I ser_b200/csrc \
//        scripts/microbench/decimate_mma_test.cu ser_b200/csrc/decimate_mma.cu ser_b200/csrc/cqt_kernels.cu \
//        ser_b200/csrc/cqt_tables.cpp -o scripts/microbench/decimate_mma_test
//   ./decimate_mma_test [n_clips] [length]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "cqt_tables.h"
#include "kernels.h"

using namespace serb;
#ifdef DM_TRACE
namespace serb { void* decimate_mma_trace_ptr(); }
#endif

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); std::exit(1); } } while (0)

int main(int argc, char** argv) {
    const int n_clips = argc > 1 ? std::atoi(argv[1]) : 1440;
    const int length = argc > 2 ? std::atoi(argv[2]) : 144000;
    std::vector<double> taps;
    decimation_taps(2, taps);
    if (static_cast<int>(taps.size()) != kDecTaps2) { std::printf("tap count %zu\n", taps.size()); return 1; }
    std::vector<double> t64(taps.size());
    std::vector<float> t32(taps.size());
    for (size_t i = 0; i < taps.size(); ++i) { t64[i] = taps[i] * std::sqrt(2.0); t32[i] = static_cast<float>(t64[i]); }
    CK(configure_cqt(t32.data(), t64.data()));
    CK(configure_decimate_mma());
    std::vector<unsigned char> table(decimate_mma_table_bytes());
    decimate_mma_table(t64.data(), table.data());
    void* d_table;
    CK(cudaMalloc(&d_table, table.size()));
    CK(cudaMemcpy(d_table, table.data(), table.size(), cudaMemcpyHostToDevice));
    int n_sms = 0;
    CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, 0));

    // ragged on purpose: every fourth clip is shorter, one is below the float64 threshold
    std::vector<TonClip> clips(n_clips);
    long long total0 = 0;
    int max_len0 = 0, max_length = 0, n_exact = 0;
    for (int c = 0; c < n_clips; ++c) {
        int len = length;
        if (c % 4 == 1) len = length - 1 - 977 * (c % 13);
        if (c == 2) len = 1500;
        if (len < 512) len = 512;
        TonClip t{};
        t.length = len;
        t.len0 = (len + 1) / 2;
        t.off0 = total0;
        t.hoff = total0 * 2;
        clips[c] = t;
        total0 += (t.len0 + 511) / 512 * 512;
        max_len0 = std::max(max_len0, t.len0);
        max_length = std::max(max_length, len);
        n_exact += len < kDecExactBelow;
    }
    std::vector<float> yharm(static_cast<size_t>(total0) * 2, 0.0f);
    std::mt19937 rng(7);
    std::normal_distribution<float> gauss(0.0f, 0.3f);
    for (int c = 0; c < n_clips; ++c) {
        // the first two long clips carry identical samples: their outputs must be bit-identical
        const int same_as = (c == 4 && clips[0].length == clips[4].length) ? 0 : -1;
        for (int i = 0; i < clips[c].length; ++i) {
            float v = gauss(rng) + 0.5f * std::sin(0.01f * i * (1 + c % 7));
            if (c % 5 == 3) v *= 1e-4f;                         // a quiet clip
            yharm[clips[c].hoff + i] = same_as >= 0 ? yharm[clips[same_as].hoff + i] : v;
        }
    }
    TonClip* d_clips;
    float *d_yharm, *d_yoct_a, *d_yoct_b;
    const size_t oct_floats = static_cast<size_t>(total0) * 2 + 64;
    CK(cudaMalloc(&d_clips, sizeof(TonClip) * n_clips));
    CK(cudaMalloc(&d_yharm, yharm.size() * 4));
    CK(cudaMalloc(&d_yoct_a, oct_floats * 4));
    CK(cudaMalloc(&d_yoct_b, oct_floats * 4));
    CK(cudaMemcpy(d_clips, clips.data(), sizeof(TonClip) * n_clips, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_yharm, yharm.data(), yharm.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_yoct_a, 0xff, oct_floats * 4));
    CK(cudaMemset(d_yoct_b, 0xff, oct_floats * 4));

    CqtParams p{};
    p.clips = d_clips;
    p.n_clips = n_clips;
    p.yharm = d_yharm;
    long long base = 0;
    for (int l = 0; l < kCqOctaves; ++l) { p.level_base[l] = base; base += total0 >> l; }
    p.early_factor = 2;
    p.max_len0 = max_len0;
    p.max_length = max_length;
    p.n_dec_exact = n_exact;
    p.dec_toeplitz = nullptr;       // A: the FFMA2 kernels
    p.n_sms = n_sms;
    CqtParams pa = p, pb = p;
    pa.yoct = d_yoct_a;
    pb.yoct = d_yoct_b;
    pb.dec_toeplitz = d_table;      // B: tcgen05

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    long long launches = 0;
    float ms_a = 0, ms_b = 0;
    const int reps = 20;
    for (int pass = 0; pass < 3; ++pass) {        // pass 0 warms the clocks up
        CK(cudaEventRecord(e0));
        for (int rep = 0; rep < reps; ++rep) CK(launch_decimations(pa, 0, &launches));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_a, e0, e1));
        CK(cudaEventRecord(e0));
        for (int rep = 0; rep < reps; ++rep) CK(launch_decimations(pb, 0, &launches));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms_b, e0, e1));
        std::printf("pass %d: FFMA2 chain %.3f ms   tcgen05 chain %.3f ms   (mean of %d)\n", pass, ms_a / reps, ms_b / reps, reps);
    }
    CK(cudaDeviceSynchronize());

#ifdef DM_TRACE
    {
        // one more level-0 launch, then the clock stamps of CTA 0
        std::vector<long long> tr(4096 * 16, 0);
        void* sym = serb::decimate_mma_trace_ptr();
        CK(cudaMemset(sym, 0, tr.size() * 8));
        CK(launch_decimate2_mma(pb, -1, max_length, d_table, n_sms, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(tr.data(), sym, tr.size() * 8, cudaMemcpyDeviceToHost));
        std::printf("CTA 0, clocks.  per group: [wait for the staged group | issue] ; producer warp 5 group 0: released at, stored at ; epilogue warp 1: accumulators full at, drained at\n");
        for (int n = 0; n < 10 && tr[n * 16] != 0; ++n) {
            const long long* t = &tr[n * 16];
            std::printf("%2d: start %7lld |", n, t[0] - tr[0]);
            for (int g = 0; g < 4; ++g) std::printf(" g%d wait %5lld issue %5lld |", g, t[3 * g + 1] - t[3 * g], t[3 * g + 2] - t[3 * g + 1]);
            std::printf(" loop top %7lld | prod g0 stored %7lld | epi %7lld .. %7lld\n", t[12] - tr[0], t[13] - tr[0], t[14] - tr[0], t[15] - tr[0]);
        }
    }
#endif
    std::vector<float> ya(oct_floats), yb(oct_floats);
    CK(cudaMemcpy(ya.data(), d_yoct_a, oct_floats * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(yb.data(), d_yoct_b, oct_floats * 4, cudaMemcpyDeviceToHost));

    // float64 recursion on a few clips: level l + 1 from the float32 level l each path produced
    // (per-level error), and the chain against a float64 chain with float32 stores (as the oracle)
    const int check[] = {0, 1, 2, 3, 5, n_clips - 1};
    int bad = 0;
    for (int level = 0; level < kCqOctaves; ++level) {
        double worst_a = 0, worst_b = 0, worst_ab = 0;
        double bias_a = 0, bias_b = 0, sq_a = 0, sq_b = 0;      // signed relative error where |out| > peak / 10
        long long n_big = 0;
        for (int ci : check) {
            if (ci >= n_clips) continue;
            const TonClip& c = clips[ci];
            int len_in = c.length;
            for (int l = 0; l < level; ++l) len_in = (len_in + 1) >> 1;      // length of the source of this level
            const int len_out = (len_in + 1) >> 1;
            auto src_of = [&](const std::vector<float>& y) -> const float* {
                return level == 0 ? yharm.data() + c.hoff : y.data() + p.level_base[level - 1] + (c.off0 >> (level - 1));
            };
            const float* sa = src_of(ya);
            const float* sb = src_of(yb);
            const float* oa = ya.data() + p.level_base[level] + (c.off0 >> level);
            const float* ob = yb.data() + p.level_base[level] + (c.off0 >> level);
            double peak = 0;
            std::vector<double> ra(len_out), rb(len_out);
            for (int m = 0; m < len_out; ++m) {
                double acc_a = 0, acc_b = 0;
                const int centre = 2 * m + (kDecTaps2 - 1) / 2;
                for (int k = std::max(0, centre - (len_in - 1)); k <= std::min(kDecTaps2 - 1, centre); ++k) {
                    acc_a += t64[k] * sa[centre - k];
                    acc_b += t64[k] * sb[centre - k];
                }
                ra[m] = acc_a;
                rb[m] = acc_b;
                peak = std::max(peak, std::fabs(acc_b));
            }
            if (peak == 0) peak = 1;
            for (int m = 0; m < len_out; ++m) {
                worst_a = std::max(worst_a, std::fabs(oa[m] - ra[m]) / peak);
                worst_b = std::max(worst_b, std::fabs(ob[m] - rb[m]) / peak);
                worst_ab = std::max(worst_ab, std::fabs(static_cast<double>(oa[m]) - ob[m]) / peak);
                if (!std::isfinite(ob[m])) ++bad;
                if (std::fabs(rb[m]) > 0.1 * peak && std::fabs(ra[m]) > 0.1 * peak) {
                    const double ea = (oa[m] - ra[m]) / ra[m], eb = (ob[m] - rb[m]) / rb[m];
                    bias_a += ea; bias_b += eb; sq_a += ea * ea; sq_b += eb * eb; ++n_big;
                }
            }
        }
        std::printf("level %d: max |out - f64 FIR of own input| / peak   FFMA2 %.3e   tcgen05 %.3e   | FFMA2 vs tcgen05 chain %.3e\n",
                    level, worst_a, worst_b, worst_ab);
        if (n_big) std::printf("         signed relative error on large outputs: FFMA2 mean %+.2e rms %.2e   tcgen05 mean %+.2e rms %.2e  (%lld outputs)\n",
                               bias_a / n_big, std::sqrt(sq_a / n_big), bias_b / n_big, std::sqrt(sq_b / n_big), n_big);
    }
    // bit-identity of equal signals at different batch positions (clips 0 and 4)
    if (n_clips > 4 && clips[0].length == clips[4].length) {
        long long diff = 0;
        for (int level = 0; level < kCqOctaves; ++level) {
            int len = clips[0].len0;
            for (int l = 0; l < level; ++l) len = (len + 1) >> 1;
            const float* x0 = yb.data() + p.level_base[level] + (clips[0].off0 >> level);
            const float* x4 = yb.data() + p.level_base[level] + (clips[4].off0 >> level);
            for (int m = 0; m < len; ++m) diff += std::memcmp(x0 + m, x4 + m, 4) != 0;
        }
        std::printf("batch-position bit differences (clip 0 vs clip 4, all levels): %lld\n", diff);
        bad += diff != 0;
    }
    // the float64 clip must be identical under both chains
    {
        const TonClip& c = clips[2];
        long long diff = 0;
        int len = c.len0;
        for (int level = 0; level < kCqOctaves; ++level) {
            const float* xa = ya.data() + p.level_base[level] + (c.off0 >> level);
            const float* xb = yb.data() + p.level_base[level] + (c.off0 >> level);
            for (int m = 0; m < len; ++m) diff += std::memcmp(xa + m, xb + m, 4) != 0;
            len = (len + 1) >> 1;
        }
        std::printf("short (float64) clip differences between the chains: %lld\n", diff);
        bad += diff != 0;
    }
    std::printf(bad ? "FAILED\n" : "OK\n");
    return bad ? 1 : 0;
}
