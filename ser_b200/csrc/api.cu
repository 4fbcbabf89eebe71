// C ABI of libser_b200.so (see include/ser_b200.h): context, chunked launch chains, staging.
#include "../../include/ser_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"
#include "cqt_tables.h"
#include "filterbanks.h"
#include "host_stage.h"
#include "kernels.h"

namespace {

using namespace serb;

thread_local std::string g_create_error;

struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (ptr) { cudaFree(ptr); ptr = nullptr; bytes = 0; }
        // exact-size growth (bytes is 0 here: the old block is gone, so want == need); sizes repeat per workload
        size_t want = std::max(need, bytes + bytes / 2);
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();          // the failed attempt must not surface at the next launch check
            want = need;
            e = cudaMalloc(&ptr, want);
        }
        if (e == cudaSuccess) bytes = want;
        else { cudaGetLastError(); ptr = nullptr; }
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

struct SrTables {
    int sample_rate = 0;
    DevBuf chroma_banks;   // [100][1025][12] float
    DevBuf mel_start, mel_count, mel_offset, mel_weights, mel_points;
    int mel_nnz = 0;
    int kmin = 0, kmax = 0, peak_cap = 0;
    // tonnetz chain (built on first use)
    bool cqt_ready = false;
    CqtPlan plan;
    DevBuf cq_rows, cq_vals, early_taps;
    DevBuf cq_sets;         // [100][7] CqSetBank: the rows mapped lane = column (cqtc_kernel), if the basis fits
    bool cq_sets_ok = false;
    int n_early_taps = 0;
};

struct Mlp {
    bool loaded = false;
    int n_in = 0, n_hidden = 0, n_out = 0, n_classes = 0, out_activation = 0;
    DevBuf mean, scale, w1, b1, w2, b2;
};

}  // namespace

constexpr int kProfKinds = 14;

struct serb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr, ev_done = nullptr, ev_chain = nullptr;
    bool chain_recorded = false;   // ev_chain marks the end of the last launch chain (on whatever stream it used)
    std::vector<cudaEvent_t> piece_events;
    std::mutex mu;
    std::string err;
    long long launches = 0;
    int chunk_cols = 1048576;   // STFT columns per launch chain (SERB_CHUNK_COLS): 21 KB of scratch per column, 22 GB at most;
                                // 262 144 cost 1.3 ms more per c2 step in launch tails (gpurun_out sweep)
    int ramp_start = 65536;     // first chunk of a host-buffer call (SERB_RAMP_START), then x ramp_factor_x10 / 10 per chunk
    int ramp_factor_x10 = 40;   // SERB_RAMP_FACTOR_X10 (profiles/r02_ramp_sweep_pcm16.txt)
    bool ramp_chunks = false;   // set by the host-buffer entries for the duration of one call
    // pageable caller memory goes through a ring of pinned slots filled by stage_threads host threads
    // (host_stage.h; SERB_STAGE_THREADS, 0 = hand pageable pointers to cudaMemcpyAsync as they are).  The
    // gather runs at a few times the chain's consumption rate, not at PCIe speed, so the ramp is gentler.
    StageRing stage_ring;
    int stage_threads = 4;
    int ramp_factor_pageable_x10 = 20;   // SERB_RAMP_FACTOR_PAGEABLE_X10
    bool ramp_pageable = false; // this call's sources are pageable (set with ramp_chunks)
    float last_ms = 0.f;
    bool timed = false;

    DevBuf edges, dct, tables, tile_clip;
    std::map<int, SrTables> sr_tables;
    Mlp mlp;

    // scratch, reused by every chunk so it stays L2 resident
    DevBuf spill, logmel, tile_mel, tile_lmax, tile_chroma, peaks, peak_count;
    DevBuf clips, short_clips, tuning, short_tuning, status;
    // host-entry staging
    DevBuf wave, out, proba, labels, x64, pcm, pcm_max, pcm_files, pcm_peaks;
    // tonnetz chain
    DevBuf hann_sq, cq_twiddles, dec_toeplitz;
    int n_sms = 0;
    bool dec_mma = true;        // factor-2 decimation on tcgen05 (SERB_DECIMATE=ffma keeps the FFMA2 kernel)
    DevBuf cspec, perc, frames, yharm, yoct, cqmag, cq_chroma, ton_part;
    bool keep_cqmag = false;    // serb_debug_tonnetz_stages: also write the 252-wide constant-Q magnitudes
    DevBuf long_idx, long_state;
    DevBuf ton_clips, ton_clips_a, ton_clips_b, ton_segs, ton_runs, ton_tuning, ton_tile_clip;
    bool cqt_shared = true;     // low octaves share the first FFT stage between frames (SERB_CQT=percolumn turns it off)
    int cqtc_max_hop = 32;     // cqtc_kernel: largest hop that shares the first FFT stage (at 64 the table costs as many
                                // first-stage transforms as the frames do); SERB_CQT_SHARED_MAXHOP
    bool cqt_cols = true;       // n_fft 1024 octaves multiply the rows lane = column (SERB_CQT=rows keeps the lane = row kernels)
    bool istft_fused = true;    // inverse STFT + overlap-add in one kernel (SERB_ISTFT=split keeps the two HBM-bound kernels)
    int harm_seg = 512;
    std::vector<int> last_tuning_rows;   // out_row per main clip, in clips-array order
    std::vector<int> last_short_rows;
    long long last_n_clips = 0;
    bool last_had_chroma = false;

    // optional per-kernel timing (serb_debug_set_profile): event pairs around every launch
    bool profile = false;
    struct ProfRec { int kind; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[kProfKinds] = {};
    long long prof_n[kProfKinds] = {};
};

namespace {

int fail(serb_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}
int fail_cuda(serb_ctx* ctx, cudaError_t e, const char* what) {
    return fail(ctx, SERB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define SERB_CUDA(ctx, call)                                          \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #call);    \
    } while (0)

// kinds: 0 stft, 1 tuning, 2 proj, 3 pool, 4 short, 5 mlp, 6 hpss_harm, 7 hpss_perc, 8 istft, 9 ola,
// 10 decimations, 11 constant-Q octaves, 12 tonnetz, 13 PCM16 preparation
constexpr const char* kProfNames[] = {"serb:stft", "serb:tuning", "serb:proj", "serb:pool", "serb:short", "serb:mlp",
                                      "serb:hpss_harm", "serb:hpss_perc", "serb:istft", "serb:ola", "serb:decimate",
                                      "serb:cqt", "serb:tonnetz", "serb:pcm_prepare"};
struct ProfScope {
    serb_ctx* ctx; int kind; cudaStream_t stream; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(serb_ctx* c, int k, cudaStream_t s) : ctx(c), kind(k), stream(s) {
        nvtxRangePushA(kProfNames[k]);   // one NVTX range per launch kind (a no-op without a tool attached)
        if (!ctx->profile) return;
        auto take = [&]() {
            cudaEvent_t ev = nullptr;
            if (!ctx->prof_pool.empty()) { ev = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
            else cudaEventCreate(&ev);
            return ev;
        };
        a = take(); b = take();
        cudaEventRecord(a, stream);
    }
    ~ProfScope() {
        nvtxRangePop();
        if (!a) return;
        cudaEventRecord(b, stream);
        ctx->prof_recs.push_back({kind, a, b});
    }
};

// Small per-call tables travel as pageable copies on purpose: the call returns once the copy engine has taken
// them, i.e. right behind the first waveform piece.  Asynchronous copies from a pinned arena were measured
// and dropped: the host then enqueues every later 32 MiB piece within a millisecond, the copy engine
// alternates between the two streams, and each small table waits for one more piece (chain start 8 ms late).
template <typename T>
int upload(serb_ctx* ctx, DevBuf& buf, const T* host, size_t count, cudaStream_t stream) {
    SERB_CUDA(ctx, buf.reserve(std::max<size_t>(count, 1) * sizeof(T)));
    if (count) SERB_CUDA(ctx, cudaMemcpyAsync(buf.ptr, host, count * sizeof(T), cudaMemcpyHostToDevice, stream));
    return SERB_OK;
}

int get_sr_tables(serb_ctx* ctx, int sr, SrTables** out) {
    auto it = ctx->sr_tables.find(sr);
    if (it != ctx->sr_tables.end()) { *out = &it->second; return SERB_OK; }
    // built in a local and moved into the cache only when every upload succeeded: a transient
    // cudaMalloc failure must not leave a half-initialised entry behind for the next call
    SrTables t;
    struct Guard {
        SrTables* t; bool keep = false;
        ~Guard() {
            if (keep) return;
            for (DevBuf* b : {&t->chroma_banks, &t->mel_start, &t->mel_count, &t->mel_offset, &t->mel_weights, &t->mel_points})
                b->release();
        }
    } guard{&t};
    t.sample_rate = sr;
    // mel (sparse CSR)
    std::vector<float> dense;
    mel_filterbank(sr, kNFft, dense);
    MelSparse ms;
    mel_sparse(dense, kNBins, ms);
    if (ms.weights.size() > 2560) return fail(ctx, SERB_ERR_UNSUPPORTED, "mel filterbank has more non-zeros than the kernel stages");
    t.mel_nnz = static_cast<int>(ms.weights.size());
    int rc;
    if ((rc = upload(ctx, t.mel_start, ms.start.data(), ms.start.size(), ctx->stream))) return rc;
    if ((rc = upload(ctx, t.mel_count, ms.count.data(), ms.count.size(), ctx->stream))) return rc;
    if ((rc = upload(ctx, t.mel_offset, ms.offset.data(), ms.offset.size(), ctx->stream))) return rc;
    if ((rc = upload(ctx, t.mel_weights, ms.weights.data(), ms.weights.size(), ctx->stream))) return rc;
    std::vector<double> pts;
    mel_points(sr, pts);
    if ((rc = upload(ctx, t.mel_points, pts.data(), pts.size(), ctx->stream))) return rc;
    // chroma banks, one per tuning-histogram bin, transposed to [bin][12]
    std::vector<float> banks(static_cast<size_t>(kNTunings) * kNBins * 12);
    std::vector<float> w;
    for (int i = 0; i < kNTunings; ++i) {
        chroma_filterbank(sr, kNFft, tuning_edge(i), w);
        float* dst = banks.data() + static_cast<size_t>(i) * kNBins * 12;
        for (int c = 0; c < 12; ++c)
            for (int k = 0; k < kNBins; ++k) dst[k * 12 + c] = w[static_cast<size_t>(c) * kNBins + k];
    }
    if ((rc = upload(ctx, t.chroma_banks, banks.data(), banks.size(), ctx->stream))) return rc;
    // piptrack band: fmin <= k * (1 / (n_fft * (1/sr))) < min(4000, sr/2)
    const double val = 1.0 / (static_cast<double>(kNFft) * (1.0 / static_cast<double>(sr)));
    const double fmax = std::min(4000.0, static_cast<double>(sr) / 2.0);
    int kmin = kNBins, kmax = 0;
    for (int k = 0; k < kNBins; ++k) {
        const double f = k * val;
        if (f >= 150.0 && f < fmax) { kmin = std::min(kmin, k); kmax = std::max(kmax, k + 1); }
    }
    if (kmin >= kmax) { kmin = 1; kmax = 1; }
    kmin = std::max(kmin, 1);
    kmax = std::min(kmax, kNBins - 1);
    t.kmin = kmin;
    t.kmax = kmax;
    t.peak_cap = std::max(1, (kmax - kmin + 1) / 2 + 1);
    cqt_plan(sr, t.plan);
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host vectors go out of scope
    guard.keep = true;
    SrTables& slot = ctx->sr_tables[sr];
    slot = t;
    *out = &slot;
    return SERB_OK;
}

struct Offsets { int dim, mfcc, chroma, mel, contrast, tonnetz; };
Offsets make_offsets(uint32_t flags) {
    Offsets o{0, -1, -1, -1, -1, -1};
    if (flags & SERB_FLAG_MFCC) { o.mfcc = o.dim; o.dim += 40; }
    if (flags & SERB_FLAG_CHROMA) { o.chroma = o.dim; o.dim += 12; }
    if (flags & SERB_FLAG_MEL) { o.mel = o.dim; o.dim += 128; }
    if (flags & SERB_FLAG_CONTRAST) { o.contrast = o.dim; o.dim += 7; }
    if (flags & SERB_FLAG_TONNETZ) { o.tonnetz = o.dim; o.dim += 6; }
    return o;
}

struct Chunk { int clip_lo, clip_hi, n_cols, n_tiles; long long max_end; };

// Validates the request and splits it into the main path (n_fft = 2048) and short clips.
int plan(serb_ctx* ctx, long long n_wave, const int64_t* starts, const int64_t* lengths, long long n_clips,
         int sr, uint32_t flags, std::vector<ClipDev>& main_clips, std::vector<ShortClip>& short_clips,
         std::vector<Chunk>& chunks) {
    // host entries stream the waveform in while the chain runs: small first chunks let compute start
    // after a few megabytes have landed, later chunks grow to the full size the kernels like
    long long total_cols = 0, consumed_cols = 0;     // STFT columns of the main path: all, and in finished chunks
    if (starts && lengths)
        for (long long i = 0; i < n_clips; ++i)
            if (lengths[i] >= kNFft) total_cols += 1 + lengths[i] / kHop;
    auto chunk_limit = [&](size_t index) -> int {
        long long limit = ctx->chunk_cols;
        if (ctx->ramp_chunks) {
            // H2D from pinned memory runs about three times faster than the chain consumes columns, so
            // each chunk may be three times the previous one without starving
            long long ramp = ctx->ramp_start;
            const int factor_x10 = ctx->ramp_pageable ? ctx->ramp_factor_pageable_x10 : ctx->ramp_factor_x10;
            for (size_t i = 0; i < std::min<size_t>(index, 12) && ramp < ctx->chunk_cols; ++i) ramp = ramp * factor_x10 / 10;
            limit = std::min(limit, ramp);
        }
        if (limit >= ctx->chunk_cols) {
            // full-size chunks share what is left evenly, so the last one is not a sliver whose twenty
            // launches run mostly empty
            const long long left = std::max<long long>(1, total_cols - consumed_cols);
            const long long n = (left + ctx->chunk_cols - 1) / ctx->chunk_cols;
            limit = std::min<long long>(limit, (left + n - 1) / n);
        }
        return static_cast<int>(limit);
    };
    if (sr <= 0) return fail(ctx, SERB_ERR_SAMPLE_RATE, "Sample rate must be a positive integer.");
    if (n_clips < 0 || (n_clips > 0 && (!starts || !lengths))) return fail(ctx, SERB_ERR_INVALID_ARG, "bad clip arrays");
    if (flags & ~SERB_FLAG_ALL) return fail(ctx, SERB_ERR_INVALID_ARG, "unknown feature flag bits");
    if ((flags & SERB_FLAG_CONTRAST) && !(6400.0 < 0.5 * sr))
        return fail(ctx, SERB_ERR_NYQUIST, "Frequency band exceeds Nyquist. Reduce either fmin or n_bands.");
    Chunk cur{0, 0, 0, 0, 0};
    for (long long i = 0; i < n_clips; ++i) {
        const long long s = starts[i], len = lengths[i];
        if (len <= 0) return fail(ctx, SERB_ERR_EMPTY, "Audio contains no samples.");
        if (s < 0 || s + len > n_wave || len > 0x7fffffffLL)
            return fail(ctx, SERB_ERR_INVALID_ARG, "clip " + std::to_string(i) + " lies outside the waveform buffer");
        if (len < kNFft) {
            short_clips.push_back(ShortClip{s, static_cast<int>(len), static_cast<int>(i)});
            continue;
        }
        ClipDev c{};
        c.start = s;
        c.length = static_cast<int>(len);
        c.n_cols = 1 + static_cast<int>(len / kHop);
        c.out_row = static_cast<int>(i);
        const int tiles = (c.n_cols + kColsPerTile - 1) / kColsPerTile;
        if (cur.clip_hi > cur.clip_lo && cur.n_cols + c.n_cols > chunk_limit(chunks.size())) {
            consumed_cols += cur.n_cols;
            chunks.push_back(cur);
            cur = Chunk{cur.clip_hi, cur.clip_hi, 0, 0, 0};
        }
        c.col_base = cur.n_cols;
        c.tile_base = cur.n_tiles;
        cur.n_cols += c.n_cols;
        cur.n_tiles += tiles;
        cur.max_end = std::max(cur.max_end, s + len);
        cur.clip_hi += 1;
        main_clips.push_back(c);
    }
    if (cur.clip_hi > cur.clip_lo) chunks.push_back(cur);
    return SERB_OK;
}

// ---- tonnetz chain -------------------------------------------------------------------------
int get_cqt_tables(serb_ctx* ctx, SrTables* tab) {
    if (tab->cqt_ready) return SERB_OK;
    const CqtPlan& plan = tab->plan;
    // 100 tuning-indexed sparse bases, built in parallel on the host (float64 FFTs)
    std::vector<CqRow> rows(static_cast<size_t>(kNTunings) * kCqtBins);
    std::vector<float> vals(static_cast<size_t>(kNTunings) * kCqtBins * kCqtRowCap * 2);
    std::vector<int> ok(kNTunings, 1);
    static_assert(sizeof(CqtSetBank) == sizeof(CqSetBank) && sizeof(CqtSet) == sizeof(CqSet), "host and device layouts of the row sets");
    bool any_1024 = false;      // octaves with an FFT size cqtc_kernel is built for (1024, 512)
    for (int i = 0; i < kCqtOctaves; ++i) any_1024 = any_1024 || plan.n_fft[i] == 1024 || plan.n_fft[i] == 512;
    std::vector<CqtSetBank> sets(any_1024 ? static_cast<size_t>(kNTunings) * kCqtOctaves : 0);
    std::vector<int> sets_ok(kNTunings, any_1024 ? 1 : 0);
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> workers;
    for (unsigned w = 0; w < hw; ++w)
        workers.emplace_back([&, w]() {
            CqtBank bank;
            for (int t = static_cast<int>(w); t < kNTunings; t += static_cast<int>(hw)) {
                ok[t] = cqt_bank(plan, t, bank) ? 1 : 0;
                for (int r = 0; r < kCqtBins; ++r) {
                    const CqtRow& src = bank.rows[r];
                    rows[static_cast<size_t>(t) * kCqtBins + r] = CqRow{src.start, src.count, src.scale, src.bin};
                }
                std::memcpy(vals.data() + static_cast<size_t>(t) * kCqtBins * kCqtRowCap * 2, bank.vals.data(),
                            bank.vals.size() * sizeof(float));
                if (any_1024) sets_ok[t] = cqt_set_banks(plan, t, sets.data() + static_cast<size_t>(t) * kCqtOctaves) ? 1 : 0;
            }
        });
    for (std::thread& th : workers) th.join();
    for (int t = 0; t < kNTunings; ++t)
        if (!ok[t]) return fail(ctx, SERB_ERR_UNSUPPORTED, "tonnetz: a constant-Q basis row is wider than the kernel stages");
    int rc;
    if ((rc = upload(ctx, tab->cq_rows, rows.data(), rows.size(), ctx->stream))) return rc;
    if ((rc = upload(ctx, tab->cq_vals, vals.data(), vals.size(), ctx->stream))) return rc;
    tab->cq_sets_ok = any_1024;
    for (int t = 0; t < kNTunings; ++t) tab->cq_sets_ok = tab->cq_sets_ok && sets_ok[t];
    if (tab->cq_sets_ok && (rc = upload(ctx, tab->cq_sets, sets.data(), sets.size(), ctx->stream))) return rc;
    std::vector<float> taps32;
    if (plan.early_factor > 2) {
        std::vector<double> taps;
        decimation_taps(plan.early_factor, taps);
        const double sc = std::sqrt(static_cast<double>(plan.early_factor));
        for (double t : taps) taps32.push_back(static_cast<float>(t * sc));
        tab->n_early_taps = static_cast<int>(taps32.size());
        if ((rc = upload(ctx, tab->early_taps, taps32.data(), taps32.size(), ctx->stream))) return rc;
    }
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tab->cqt_ready = true;
    return SERB_OK;
}

// clips of one chunk that take the multi-CTA tuning path (more than kTuneLongCols columns)
struct LongList {
    const int* d_idx = nullptr;
    int n = 0;
    int max_cols = 0;
};

struct TonChunk {
    int clip_lo, clip_hi;       // range in the request-wide TonClip array
    int seg_lo, n_segs;
    int run_lo, n_runs;         // (clip, first column) per kIstftRun columns: work items of istft_ola_kernel
    int n_cols, n_tiles, cq_rows, max_len0, max_cq_cols, n_parts, max_length, n_dec_exact;
    long long total0, max_end;
};

struct TonPlan {
    std::vector<TonClip> clips;
    std::vector<ClipDev> clips_b;   // the harmonic signals as STFT clips (tuning pass)
    std::vector<int2> segs;
    std::vector<int2> runs;
    std::vector<TonChunk> chunks;
};

TonChunk ton_open_chunk(const TonPlan& tp) {
    TonChunk c{};
    c.clip_lo = c.clip_hi = static_cast<int>(tp.clips.size());
    c.seg_lo = static_cast<int>(tp.segs.size());
    c.run_lo = static_cast<int>(tp.runs.size());
    return c;
}

// appends one clip (padded length plen >= 512, dsp.py:38-45) to the open chunk
int ton_add_clip(serb_ctx* ctx, const CqtPlan& plan, TonPlan& tp, TonChunk& cur, const ClipDev& a, long long plen) {
    const int fe = plan.early_factor;
    const long long len0 = (plen + fe - 1) / fe;
    TonClip c{};
    c.off0 = cur.total0;
    c.hoff = cur.total0 * fe;
    c.length = static_cast<int>(plen);
    c.n_cols = a.n_cols;
    c.col_base = a.col_base;
    c.tile_base = a.tile_base;
    c.len0 = static_cast<int>(len0);
    int cq = 0x7fffffff, ln = c.len0;
    for (int l = 0; l < kCqOctaves; ++l) {
        cq = std::min(cq, 1 + ln / (plan.hop0 >> l));
        ln = (ln + 1) >> 1;
    }
    c.cq_cols = cq;
    c.cq_base = cur.cq_rows;
    c.out_row = a.out_row;
    c.part_base = cur.n_parts;
    ClipDev b = a;
    b.start = c.hoff;
    b.length = c.length;
    const int local = static_cast<int>(tp.clips.size()) - cur.clip_lo;
    for (int t0 = 0; t0 < a.n_cols; t0 += ctx->harm_seg) tp.segs.push_back(make_int2(local, t0));
    for (int t0 = 0; t0 < a.n_cols; t0 += kIstftRun) tp.runs.push_back(make_int2(local, t0));
    tp.clips.push_back(c);
    tp.clips_b.push_back(b);
    // 512-sample granules keep every level of every clip 32-byte aligned (off0 >> 6 is a multiple of 8)
    cur.total0 += (len0 + 511) / 512 * 512;
    cur.n_cols = std::max(cur.n_cols, a.col_base + a.n_cols);
    cur.n_tiles = std::max(cur.n_tiles, a.tile_base + (a.n_cols + kColsPerTile - 1) / kColsPerTile);
    cur.n_segs = static_cast<int>(tp.segs.size()) - cur.seg_lo;
    cur.n_runs = static_cast<int>(tp.runs.size()) - cur.run_lo;
    cur.cq_rows += cq;
    cur.n_parts += (cq + kTonTile - 1) / kTonTile;
    cur.max_len0 = std::max(cur.max_len0, c.len0);
    cur.max_length = std::max(cur.max_length, c.length);
    cur.n_dec_exact += c.length < kDecExactBelow;
    cur.max_cq_cols = std::max(cur.max_cq_cols, cq);
    cur.max_end = std::max(cur.max_end, a.start + a.length);
    cur.clip_hi += 1;
    return SERB_OK;
}

int ton_reserve_and_upload(serb_ctx* ctx, SrTables* tab, const TonPlan& tp, cudaStream_t stream) {
    const int fe = tab->plan.early_factor;
    int max_cols = 0, max_tiles = 0, max_cq = 0, max_parts = 0;
    long long max_total0 = 0;
    for (const TonChunk& c : tp.chunks) {
        max_parts = std::max(max_parts, c.n_parts);
        max_cols = std::max(max_cols, c.n_cols);
        max_tiles = std::max(max_tiles, c.n_tiles);
        max_cq = std::max(max_cq, c.cq_rows);
        max_total0 = std::max(max_total0, c.total0);
    }
    const size_t col_f = static_cast<size_t>(max_cols) * kSpillStride;
    SERB_CUDA(ctx, ctx->spill.reserve(col_f * sizeof(float)));
    SERB_CUDA(ctx, ctx->cspec.reserve(col_f * sizeof(float2)));
    SERB_CUDA(ctx, ctx->perc.reserve(col_f * sizeof(float)));
    if (!ctx->istft_fused) SERB_CUDA(ctx, ctx->frames.reserve(static_cast<size_t>(max_cols) * kNFft * sizeof(float)));
    SERB_CUDA(ctx, ctx->yharm.reserve((static_cast<size_t>(max_total0) * fe + 64) * sizeof(float)));
    SERB_CUDA(ctx, ctx->yoct.reserve((static_cast<size_t>(max_total0) * 2 + 64) * sizeof(float)));
    if (ctx->keep_cqmag) SERB_CUDA(ctx, ctx->cqmag.reserve(static_cast<size_t>(std::max(max_cq, 1)) * kCqBins * sizeof(float)));
    SERB_CUDA(ctx, ctx->cq_chroma.reserve(static_cast<size_t>(std::max(max_cq, 1)) * kCqOctaves * 12 * sizeof(float)));
    SERB_CUDA(ctx, ctx->ton_part.reserve(static_cast<size_t>(std::max(max_parts, 1)) * 6 * sizeof(double)));
    SERB_CUDA(ctx, ctx->ton_tile_clip.reserve(static_cast<size_t>(std::max(max_tiles, 1)) * sizeof(int)));
    SERB_CUDA(ctx, ctx->peaks.reserve(static_cast<size_t>(max_cols) * tab->peak_cap * sizeof(float2)));
    SERB_CUDA(ctx, ctx->peak_count.reserve(static_cast<size_t>(max_cols) * sizeof(int)));
    SERB_CUDA(ctx, ctx->ton_tuning.reserve(std::max<size_t>(tp.clips.size(), 1) * sizeof(int)));
    int rc;
    if ((rc = upload(ctx, ctx->ton_clips, tp.clips.data(), tp.clips.size(), stream))) return rc;
    if ((rc = upload(ctx, ctx->ton_clips_b, tp.clips_b.data(), tp.clips_b.size(), stream))) return rc;
    if ((rc = upload(ctx, ctx->ton_segs, tp.segs.data(), tp.segs.size(), stream))) return rc;
    if ((rc = upload(ctx, ctx->ton_runs, tp.runs.data(), tp.runs.size(), stream))) return rc;
    return SERB_OK;
}

// librosa.effects.harmonic + librosa.feature.tonnetz for the clips of one chunk, written to
// out[row][off_tonnetz .. +6).  d_clips_a / d_tile_clip describe the chunk's clips inside the
// waveform; when stft_done the chunk's |X| (ctx->spill) and X (ctx->cspec) are already there.
int ton_run_chunk(serb_ctx* ctx, SrTables* tab, int sr, const Offsets& off, float* d_out, cudaStream_t stream,
                  const TonChunk& c, const float* d_wave, const ClipDev* d_clips_a, const int* d_tile_clip,
                  bool stft_done, const LongList& longs) {
    const CqtPlan& plan = tab->plan;
    const int nc = c.clip_hi - c.clip_lo;
    if (nc <= 0) return SERB_OK;
    const TonClip* d_clips = ctx->ton_clips.as<TonClip>() + c.clip_lo;
    const ClipDev* d_b = ctx->ton_clips_b.as<ClipDev>() + c.clip_lo;
    int* d_tuning = ctx->ton_tuning.as<int>() + c.clip_lo;
    StftParams sp{};
    sp.wave = d_wave;
    sp.clips = d_clips_a;
    sp.n_clips = nc;
    sp.tile_clip = d_tile_clip;
    sp.tables = ctx->tables.as<float2>();
    sp.spill = ctx->spill.as<float>();
    sp.cspill = ctx->cspec.as<float2>();
    sp.do_peaks = 0;
    sp.kmin = tab->kmin; sp.kmax = tab->kmax; sp.peak_cap = tab->peak_cap;
    sp.peaks = ctx->peaks.as<float2>();
    sp.peak_count = ctx->peak_count.as<int>();
    sp.sr_over_nfft_num = static_cast<double>(sr);
    sp.status = ctx->status.as<int>();
    if (!stft_done) {
        // 1. complex STFT of the clip
        { ProfScope ps(ctx, 0, stream); SERB_CUDA(ctx, launch_stft(sp, c.n_tiles, stream)); }
        ctx->launches += 1;
    }
    // 2. HPSS medians; the time median leaves the soft mask in ctx->perc
    HpssParams hp{};
    hp.clips = d_clips;
    hp.segs = ctx->ton_segs.as<int2>() + c.seg_lo;
    hp.seg_len = ctx->harm_seg;
    hp.one = 1.0f;
    hp.mag = ctx->spill.as<float>();
    hp.perc = ctx->perc.as<float>();
    hp.cspec = ctx->cspec.as<float2>();
    { ProfScope ps(ctx, 7, stream); SERB_CUDA(ctx, launch_hpss_perc(hp, c.n_cols, stream)); }
    { ProfScope ps(ctx, 6, stream); SERB_CUDA(ctx, launch_hpss_harm(hp, c.n_segs, stream)); }
    ctx->launches += 2;
    // 3. inverse STFT + overlap-add
    IstftParams ip{};
    ip.cspec = ctx->cspec.as<float2>();
    ip.mask = ctx->perc.as<float>();
    ip.tables = ctx->tables.as<float2>();
    ip.frames = ctx->frames.as<float>();
    OlaParams op{};
    op.clips = d_clips;
    op.tile_clip = d_tile_clip;
    op.frames = ip.frames;
    op.hann_sq = ctx->hann_sq.as<double>();
    op.wss4 = reinterpret_cast<const float*>(ctx->hann_sq.as<double>() + 2048);
    op.yharm = ctx->yharm.as<float>();
    if (ctx->istft_fused) {
        ProfScope ps(ctx, 8, stream);
        SERB_CUDA(ctx, launch_istft_ola(ip, op, ctx->ton_runs.as<int2>() + c.run_lo, c.n_runs, ctx->n_sms, stream));
        ctx->launches += 1;
    } else {
        { ProfScope ps(ctx, 8, stream); SERB_CUDA(ctx, launch_istft(ip, c.n_cols, stream)); }
        { ProfScope ps(ctx, 9, stream); SERB_CUDA(ctx, launch_ola(op, c.n_tiles, stream)); }
        ctx->launches += 2;
    }
    // 4. tuning of the harmonic signal (36 bins per octave)
    sp.wave = ctx->yharm.as<float>();
    sp.clips = d_b;
    sp.spill = nullptr;          // only the piptrack peaks of the harmonic signal are needed
    sp.cspill = nullptr;
    sp.do_peaks = 1;
    { ProfScope ps(ctx, 0, stream); SERB_CUDA(ctx, launch_stft(sp, c.n_tiles, stream)); }
    TuneParams tp{};
    tp.clips = d_b;
    tp.peaks = sp.peaks;
    tp.peak_count = sp.peak_count;
    tp.peak_cap = tab->peak_cap;
    tp.bins_per_octave = 36;
    tp.edges = ctx->edges.as<double>();
    tp.tuning_idx = d_tuning;
    tp.long_clips = longs.d_idx;
    tp.n_long = longs.n;
    tp.max_long_cols = longs.max_cols;
    tp.long_state = ctx->long_state.as<TuneLongState>();
    { ProfScope ps(ctx, 1, stream); SERB_CUDA(ctx, launch_tuning(tp, nc, stream, &ctx->launches)); }
    ctx->launches += 1;
    // 5. decimations, constant-Q, chroma, tonnetz
    CqtParams qp{};
    qp.clips = d_clips;
    qp.n_clips = nc;
    qp.tuning_idx = d_tuning;
    qp.yharm = ctx->yharm.as<float>();
    qp.yoct = ctx->yoct.as<float>();
    long long base = 0;
    for (int l = 0; l < kCqOctaves; ++l) { qp.level_base[l] = base; base += c.total0 >> l; }
    qp.early_factor = plan.early_factor;
    qp.hop0 = plan.hop0;
    for (int l = 0; l < kCqOctaves; ++l) qp.n_fft[l] = plan.n_fft[l];
    qp.early_taps = tab->early_taps.as<float>();
    qp.n_early_taps = tab->n_early_taps;
    qp.rows = tab->cq_rows.as<CqRow>();
    qp.vals = tab->cq_vals.as<float2>();
    qp.set_banks = (ctx->cqt_cols && tab->cq_sets_ok) ? tab->cq_sets.as<CqSetBank>() : nullptr;
    qp.twiddles = ctx->cq_twiddles.as<float2>();
    qp.cqmag = ctx->keep_cqmag ? ctx->cqmag.as<float>() : nullptr;
    qp.cq_chroma = ctx->cq_chroma.as<float>();
    qp.ton_part = ctx->ton_part.as<double>();
    qp.out = d_out;
    qp.dim = off.dim;
    qp.off_tonnetz = off.tonnetz;
    qp.max_len0 = c.max_len0;
    qp.max_cq_cols = c.max_cq_cols;
    qp.max_length = c.max_length;
    qp.n_dec_exact = c.n_dec_exact;
    qp.dec_toeplitz = ctx->dec_mma ? ctx->dec_toeplitz.ptr : nullptr;
    qp.n_sms = ctx->n_sms;
    qp.cqt_no_shared = ctx->cqt_shared ? 0 : 1;
    qp.cqtc_shared_max_hop = ctx->cqtc_max_hop;
    { ProfScope ps(ctx, 10, stream); SERB_CUDA(ctx, launch_decimations(qp, stream, &ctx->launches)); }
    { ProfScope ps(ctx, 11, stream); SERB_CUDA(ctx, launch_cqt_octaves(qp, stream, &ctx->launches)); }
    { ProfScope ps(ctx, 12, stream); SERB_CUDA(ctx, launch_tonnetz(qp, stream)); }
    ctx->launches += 2;
    return SERB_OK;
}

// Enqueues the whole launch chain on `stream`.  before_chunk(max_end) lets the host entry make
// the stream wait for the H2D piece that holds the chunk's last sample.
template <typename BeforeChunk>
int run_features(serb_ctx* ctx, const float* d_wave, long long n_wave, const int64_t* starts,
                 const int64_t* lengths, long long n_clips, int sr, uint32_t flags, float* d_out,
                 cudaStream_t stream, BeforeChunk&& before_chunk) {
    std::vector<ClipDev> main_clips;
    std::vector<ShortClip> short_clips;
    std::vector<Chunk> chunks;
    int rc = plan(ctx, n_wave, starts, lengths, n_clips, sr, flags, main_clips, short_clips, chunks);
    if (rc) return rc;
    ctx->last_n_clips = n_clips;
    ctx->last_had_chroma = (flags & SERB_FLAG_CHROMA) != 0;
    ctx->last_tuning_rows.clear();
    ctx->last_short_rows.clear();
    const Offsets off = make_offsets(flags);
    if (n_clips == 0 || off.dim == 0) return SERB_OK;
    if (off.contrast >= 0 && off.dim == 7 && main_clips.empty() && short_clips.empty()) return SERB_OK;

    SrTables* tab = nullptr;
    if ((rc = get_sr_tables(ctx, sr, &tab))) return rc;
    // the scratch buffers are per context, not per stream: a caller alternating streams must not
    // start this chain before the previous one (on another stream) has finished with them
    if (ctx->chain_recorded) SERB_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_chain, 0));
    const bool want_chroma = off.chroma >= 0;
    const bool want_mel = off.mel >= 0 || off.mfcc >= 0;
    const bool want_ton = off.tonnetz >= 0;

    // ---- tonnetz plan: one chunk per main chunk (sharing its STFT), then the short clips ----
    TonPlan tp;
    std::vector<ClipDev> short_a;       // short clips as STFT-2048 clips (harmonic() always uses n_fft = 2048)
    size_t n_main_ton_chunks = 0;
    if (want_ton) {
        const CqtPlan& cp = tab->plan;
        if (cp.status == 1) return fail(ctx, SERB_ERR_NYQUIST, cp.message);
        if (cp.status != 0) return fail(ctx, SERB_ERR_UNSUPPORTED, cp.message);
        if ((rc = get_cqt_tables(ctx, tab))) return rc;
        for (const Chunk& c : chunks) {
            TonChunk cur = ton_open_chunk(tp);
            for (int i = c.clip_lo; i < c.clip_hi; ++i)
                if ((rc = ton_add_clip(ctx, cp, tp, cur, main_clips[i], main_clips[i].length))) return rc;
            tp.chunks.push_back(cur);
        }
        n_main_ton_chunks = tp.chunks.size();
        TonChunk cur = ton_open_chunk(tp);
        int cols = 0, tiles = 0;
        for (const ShortClip& sc : short_clips) {
            const long long plen = std::max<long long>(sc.length, 512);
            ClipDev a{};
            a.start = sc.start;
            a.length = sc.length;
            a.n_cols = 1 + static_cast<int>(plen / kHop);
            a.out_row = sc.out_row;
            if (cur.clip_hi > cur.clip_lo && cols + a.n_cols > ctx->chunk_cols) {
                tp.chunks.push_back(cur);
                cur = ton_open_chunk(tp);
                cols = 0; tiles = 0;
            }
            a.col_base = cols;
            a.tile_base = tiles;
            cols += a.n_cols;
            tiles += (a.n_cols + kColsPerTile - 1) / kColsPerTile;
            short_a.push_back(a);
            if ((rc = ton_add_clip(ctx, cp, tp, cur, a, plen))) return rc;
        }
        if (cur.clip_hi > cur.clip_lo) tp.chunks.push_back(cur);
    }

    int max_cols = 0, max_tiles = 0, max_clips = 0;
    for (const Chunk& c : chunks) {
        max_cols = std::max(max_cols, c.n_cols);
        max_tiles = std::max(max_tiles, c.n_tiles);
        max_clips = std::max(max_clips, c.clip_hi - c.clip_lo);
    }
    if (!chunks.empty()) {
        SERB_CUDA(ctx, ctx->spill.reserve(static_cast<size_t>(max_cols) * kSpillStride * sizeof(float)));
        SERB_CUDA(ctx, ctx->tile_clip.reserve(static_cast<size_t>(max_tiles) * sizeof(int)));
        SERB_CUDA(ctx, ctx->logmel.reserve(static_cast<size_t>(max_tiles) * 128 * kColsPerTile * sizeof(float)));
        SERB_CUDA(ctx, ctx->tile_mel.reserve(static_cast<size_t>(max_tiles) * 128 * sizeof(float)));
        SERB_CUDA(ctx, ctx->tile_lmax.reserve(static_cast<size_t>(max_tiles) * sizeof(float)));
        SERB_CUDA(ctx, ctx->tile_chroma.reserve(static_cast<size_t>(max_tiles) * 12 * sizeof(float)));
        if (want_chroma || want_ton) {
            SERB_CUDA(ctx, ctx->peaks.reserve(static_cast<size_t>(max_cols) * tab->peak_cap * sizeof(float2)));
            SERB_CUDA(ctx, ctx->peak_count.reserve(static_cast<size_t>(max_cols) * sizeof(int)));
        }
        SERB_CUDA(ctx, ctx->tuning.reserve(main_clips.size() * sizeof(int)));
        if ((rc = upload(ctx, ctx->clips, main_clips.data(), main_clips.size(), stream))) return rc;
        for (const ClipDev& c : main_clips) ctx->last_tuning_rows.push_back(c.out_row);
    }
    // clips long enough for the multi-CTA tuning path, per chunk
    std::vector<LongList> longs(chunks.size());
    {
        std::vector<int> idx;
        std::vector<size_t> first(chunks.size(), 0);
        int max_n = 0;
        for (size_t ci = 0; ci < chunks.size(); ++ci) {
            first[ci] = idx.size();
            for (int i = chunks[ci].clip_lo; i < chunks[ci].clip_hi; ++i)
                if (main_clips[i].n_cols > kTuneLongCols) {
                    idx.push_back(i - chunks[ci].clip_lo);
                    longs[ci].n += 1;
                    longs[ci].max_cols = std::max(longs[ci].max_cols, main_clips[i].n_cols);
                }
            max_n = std::max(max_n, longs[ci].n);
        }
        if (!idx.empty() && (want_chroma || want_ton)) {
            if ((rc = upload(ctx, ctx->long_idx, idx.data(), idx.size(), stream))) return rc;
            SERB_CUDA(ctx, ctx->long_state.reserve(static_cast<size_t>(max_n) * sizeof(TuneLongState)));
            // prefixes and histograms start every call at zero (a few megabytes at most)
            SERB_CUDA(ctx, cudaMemsetAsync(ctx->long_state.ptr, 0, static_cast<size_t>(max_n) * sizeof(TuneLongState), stream));
            for (size_t ci = 0; ci < chunks.size(); ++ci) longs[ci].d_idx = ctx->long_idx.as<int>() + first[ci];
        } else {
            for (LongList& l : longs) l = LongList{};
        }
    }
    if (want_ton) {
        if ((rc = ton_reserve_and_upload(ctx, tab, tp, stream))) return rc;
        if ((rc = upload(ctx, ctx->ton_clips_a, short_a.data(), short_a.size(), stream))) return rc;
    }
    if (!short_clips.empty()) {
        if ((rc = upload(ctx, ctx->short_clips, short_clips.data(), short_clips.size(), stream))) return rc;
        SERB_CUDA(ctx, ctx->short_tuning.reserve(short_clips.size() * sizeof(int)));
    }
    SERB_CUDA(ctx, ctx->status.reserve(sizeof(int)));
    SERB_CUDA(ctx, cudaMemsetAsync(ctx->status.ptr, 0, sizeof(int), stream));

    if (ctx->timed) SERB_CUDA(ctx, cudaEventRecord(ctx->ev_start, stream));
    for (size_t ci = 0; ci < chunks.size(); ++ci) {
        const Chunk& c = chunks[ci];
        before_chunk(c.max_end);
        const ClipDev* d_clips = ctx->clips.as<ClipDev>() + c.clip_lo;
        const int nc = c.clip_hi - c.clip_lo;
        SERB_CUDA(ctx, launch_expand_tiles(d_clips, nc, ctx->tile_clip.as<int>(), stream));
        ctx->launches += 1;
        StftParams sp{};
        sp.wave = d_wave;
        sp.clips = d_clips;
        sp.n_clips = nc;
        sp.tile_clip = ctx->tile_clip.as<int>();
        sp.tables = ctx->tables.as<float2>();
        sp.spill = ctx->spill.as<float>();
        sp.cspill = want_ton ? ctx->cspec.as<float2>() : nullptr;
        sp.do_peaks = want_chroma ? 1 : 0;
        sp.kmin = tab->kmin;
        sp.kmax = tab->kmax;
        sp.peak_cap = tab->peak_cap;
        sp.peaks = ctx->peaks.as<float2>();
        sp.peak_count = ctx->peak_count.as<int>();
        sp.sr_over_nfft_num = static_cast<double>(sr);
        sp.status = ctx->status.as<int>();
        { ProfScope ps(ctx, 0, stream); SERB_CUDA(ctx, launch_stft(sp, c.n_tiles, stream)); }
        ctx->launches += 1;
        int* d_tuning = ctx->tuning.as<int>() + c.clip_lo;
        if (want_chroma) {
            TuneParams tu{};
            tu.clips = d_clips;
            tu.peaks = sp.peaks;
            tu.peak_count = sp.peak_count;
            tu.peak_cap = tab->peak_cap;
            tu.bins_per_octave = 12;
            tu.edges = ctx->edges.as<double>();
            tu.tuning_idx = d_tuning;
            tu.long_clips = longs[ci].d_idx;
            tu.n_long = longs[ci].n;
            tu.max_long_cols = longs[ci].max_cols;
            tu.long_state = ctx->long_state.as<TuneLongState>();
            { ProfScope ps(ctx, 1, stream); SERB_CUDA(ctx, launch_tuning(tu, nc, stream, &ctx->launches)); }
        }
        ProjParams pp{};
        pp.clips = d_clips;
        pp.n_clips = nc;
        pp.tile_clip = sp.tile_clip;
        pp.spill = sp.spill;
        pp.do_mel = want_mel ? 1 : 0;
        pp.mel_start = tab->mel_start.as<int>();
        pp.mel_count = tab->mel_count.as<int>();
        pp.mel_offset = tab->mel_offset.as<int>();
        pp.mel_weights = tab->mel_weights.as<float>();
        pp.mel_nnz = tab->mel_nnz;
        pp.logmel = ctx->logmel.as<float>();
        pp.tile_mel = ctx->tile_mel.as<float>();
        pp.tile_lmax = ctx->tile_lmax.as<float>();
        pp.do_chroma = want_chroma ? 1 : 0;
        pp.chroma_banks = tab->chroma_banks.as<float>();
        pp.tuning_idx = d_tuning;
        pp.tile_chroma = ctx->tile_chroma.as<float>();
        if (want_mel || want_chroma) {
            { ProfScope ps(ctx, 2, stream); SERB_CUDA(ctx, launch_proj(pp, c.n_tiles, stream)); }
            ctx->launches += 1;
        }
        PoolParams qp{};
        qp.clips = d_clips;
        qp.logmel = pp.logmel;
        qp.tile_mel = pp.tile_mel;
        qp.tile_lmax = pp.tile_lmax;
        qp.tile_chroma = pp.tile_chroma;
        qp.dct = ctx->dct.as<double>();
        qp.out = d_out;
        qp.dim = off.dim;
        qp.off_mfcc = off.mfcc;
        qp.off_chroma = off.chroma;
        qp.off_mel = off.mel;
        qp.off_contrast = off.contrast;
        { ProfScope ps(ctx, 3, stream); SERB_CUDA(ctx, launch_pool(qp, nc, stream)); }
        ctx->launches += 1;
        if (want_ton) {
            // same clips, same columns: the chunk's |X| and X feed the harmonic separation directly
            if ((rc = ton_run_chunk(ctx, tab, sr, off, d_out, stream, tp.chunks[ci], d_wave, d_clips,
                                    ctx->tile_clip.as<int>(), true, longs[ci]))) return rc;
        }
    }
    if (!short_clips.empty()) {
        long long max_end = 0;
        for (const ShortClip& s : short_clips) {
            max_end = std::max(max_end, s.start + s.length);
            ctx->last_short_rows.push_back(s.out_row);
        }
        before_chunk(max_end);
        ShortParams hp{};
        hp.wave = d_wave;
        hp.clips = ctx->short_clips.as<ShortClip>();
        hp.sample_rate = sr;
        hp.mel_points = tab->mel_points.as<double>();
        hp.edges = ctx->edges.as<double>();
        hp.dct = ctx->dct.as<double>();
        hp.out = d_out;
        hp.dim = off.dim;
        hp.off_mfcc = off.mfcc;
        hp.off_chroma = off.chroma;
        hp.off_mel = off.mel;
        hp.off_contrast = off.contrast;
        hp.tuning_idx = ctx->short_tuning.as<int>();
        hp.status = ctx->status.as<int>();
        { ProfScope ps(ctx, 4, stream); SERB_CUDA(ctx, launch_short(hp, static_cast<int>(short_clips.size()), stream)); }
        ctx->launches += 1;
        // tonnetz of the short clips: their own STFT-2048 pass (the short kernel uses n_fft = len)
        for (size_t ci = n_main_ton_chunks; ci < tp.chunks.size(); ++ci) {
            const TonChunk& c = tp.chunks[ci];
            const int first = c.clip_lo - static_cast<int>(main_clips.size());
            const ClipDev* d_a = ctx->ton_clips_a.as<ClipDev>() + first;
            SERB_CUDA(ctx, launch_expand_tiles(d_a, c.clip_hi - c.clip_lo, ctx->ton_tile_clip.as<int>(), stream));
            ctx->launches += 1;
            if ((rc = ton_run_chunk(ctx, tab, sr, off, d_out, stream, c, d_wave, d_a, ctx->ton_tile_clip.as<int>(), false,
                                    LongList{})))
                return rc;
        }
    }
    if (ctx->timed) SERB_CUDA(ctx, cudaEventRecord(ctx->ev_stop, stream));
    SERB_CUDA(ctx, cudaEventRecord(ctx->ev_chain, stream));
    ctx->chain_recorded = true;
    return SERB_OK;
}

int check_status(serb_ctx* ctx, cudaStream_t stream) {
    int status = 0;
    SERB_CUDA(ctx, cudaMemcpyAsync(&status, ctx->status.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(stream));
    if (status & 1) return fail(ctx, SERB_ERR_NOT_FINITE, "Audio buffer is not finite everywhere.");
    return SERB_OK;
}

// host waveform -> ctx->wave in pieces on the copy stream; returns a functor making `stream`
// wait for the piece holding sample (max_end - 1)
// Host waveform -> ctx->wave in pieces on the copy stream.  The copies are enqueued lazily, at the
// first chunk: every small table upload of the call has been issued by then, so none of them sits
// behind a gigabyte of waveform in the H2D copy engine's queue.  operator()(max_end) makes the
// compute stream wait for the piece holding sample (max_end - 1).
// whether this call's host sources go through the pinned ring (the first source decides; a wrong
// guess only costs speed, never correctness)
bool use_stage_ring(serb_ctx* ctx, const void* first_source) {
    if (ctx->stage_threads <= 0 || !first_source || !host_pointer_is_pageable(first_source)) return false;
    if (ctx->stage_ring.ensure(ctx->stage_threads) != cudaSuccess) {
        cudaGetLastError();      // no pinned memory to be had: plain copies
        return false;
    }
    return true;
}

struct PieceWaiter {
    serb_ctx* ctx;
    cudaStream_t stream;
    const float* h_wave;
    long long n_wave;
    bool staged = false;           // pageable source: pieces are gathered into pinned slots as the chunks ask for them
    long long piece = 8LL << 20;   // samples per piece (32 MiB)
    int n_pieces = 0;
    int enqueued = 0;
    int waited = -1;
    bool started = false;
    cudaError_t error = cudaSuccess;
    void start() {
        started = true;
        n_pieces = static_cast<int>((n_wave + piece - 1) / piece);
        // the previous call's kernels may still read ctx->wave
        if ((error = cudaEventRecord(ctx->ev_done, ctx->stream)) != cudaSuccess) return;
        error = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done, 0);
    }
    void enqueue_through(int last) {
        for (; enqueued <= last && error == cudaSuccess; ++enqueued) {
            const long long lo = enqueued * piece, hi = std::min(n_wave, lo + piece);
            if (staged) {
                StagedWriter writer{&ctx->stage_ring, ctx->copy_stream};
                if ((error = writer.add(ctx->wave.as<float>() + lo, h_wave + lo, (hi - lo) * sizeof(float))) != cudaSuccess) return;
                if ((error = writer.flush()) != cudaSuccess) return;
            } else if ((error = cudaMemcpyAsync(ctx->wave.as<float>() + lo, h_wave + lo, (hi - lo) * sizeof(float),
                                                cudaMemcpyHostToDevice, ctx->copy_stream)) != cudaSuccess) {
                return;
            }
            error = cudaEventRecord(ctx->piece_events[enqueued], ctx->copy_stream);
        }
    }
    // every enqueued copy is ordered before whatever follows on the compute stream (a call that planned
    // no chunk at all never waited for the prefetched piece, and the caller's buffer must not be in
    // flight when the entry returns)
    void finish() {
        if (enqueued > 0) cudaStreamWaitEvent(stream, ctx->piece_events[enqueued - 1], 0);
    }
    // before the host plans the chain: the first piece of a pinned source starts crossing PCIe now
    void prefetch() {
        if (staged || n_wave <= 0) return;
        if (!started) start();
        if (n_pieces > 0 && error == cudaSuccess) enqueue_through(0);
    }
    void operator()(long long max_end) {
        if (!started) start();
        if (n_pieces == 0 || error != cudaSuccess) return;
        int idx = static_cast<int>(std::min<long long>((std::max<long long>(max_end, 1) - 1) / piece, n_pieces - 1));
        // pinned sources: every copy is enqueued at once (they are asynchronous); pageable ones: up to the
        // piece this chunk needs, the host gathering the next pieces while the chunk's kernels run
        enqueue_through(staged ? idx : n_pieces - 1);
        if (error != cudaSuccess) return;
        if (idx > waited) {
            cudaStreamWaitEvent(stream, ctx->piece_events[idx], 0);
            waited = idx;
        }
    }
};

int ensure_piece_events(serb_ctx* ctx, int n_pieces) {
    while (static_cast<int>(ctx->piece_events.size()) < n_pieces) {
        cudaEvent_t ev;
        SERB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->piece_events.push_back(ev);
    }
    return SERB_OK;
}

int stage_wave(serb_ctx* ctx, const float* h_wave, long long n_wave, PieceWaiter& waiter) {
    SERB_CUDA(ctx, ctx->wave.reserve(std::max<long long>(n_wave, 1) * sizeof(float) + 64));
    waiter = PieceWaiter{ctx, ctx->stream, h_wave, n_wave};
    waiter.staged = n_wave > 0 && use_stage_ring(ctx, h_wave);
    return ensure_piece_events(ctx, static_cast<int>((n_wave + waiter.piece - 1) / waiter.piece));
}

// The same for a list of separately allocated clips (the training loader's shape: one array per
// file).  Clips are copied to 16-byte aligned offsets of ctx->wave in pieces of consecutive clips;
// a piece is enqueued when the first chunk that needs it is about to launch, so the host's copy
// work for chunk k + 1 (pageable memory: the driver stages it) overlaps the kernels of chunk k and
// the caller never builds a packed copy of its own.
struct ClipStager {
    serb_ctx* ctx;
    cudaStream_t stream;
    const float* const* clips;
    const int64_t* lengths;
    const std::vector<int64_t>* starts;
    long long n_clips;
    long long next_clip = 0;      // first clip not yet enqueued
    long long enqueued_end = 0;   // wave offset covered by the enqueued copies
    bool staged = false;          // pageable clips: gathered into pinned slots (host_stage.h)
    int n_pieces = 0;
    bool started = false;
    cudaError_t error = cudaSuccess;
    void operator()(long long max_end) {
        if (error != cudaSuccess) return;
        if (!started) {
            started = true;
            // the previous call's kernels may still read ctx->wave
            if ((error = cudaEventRecord(ctx->ev_done, ctx->stream)) != cudaSuccess) return;
            if ((error = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done, 0)) != cudaSuccess) return;
        }
        constexpr long long kPiece = 8LL << 20;   // samples per piece (32 MiB)
        while (enqueued_end < max_end && next_clip < n_clips) {
            long long in_piece = 0;
            StagedWriter writer{&ctx->stage_ring, ctx->copy_stream};
            while (next_clip < n_clips && in_piece < kPiece) {
                const long long len = lengths[next_clip];
                float* dst = ctx->wave.as<float>() + (*starts)[next_clip];
                if (staged) error = writer.add(dst, clips[next_clip], static_cast<size_t>(len) * sizeof(float));
                else error = cudaMemcpyAsync(dst, clips[next_clip], static_cast<size_t>(len) * sizeof(float),
                                             cudaMemcpyHostToDevice, ctx->copy_stream);
                if (error != cudaSuccess) return;
                enqueued_end = (*starts)[next_clip] + len;
                in_piece += len;
                ++next_clip;
            }
            if (staged && (error = writer.flush()) != cudaSuccess) return;
            if (n_pieces >= static_cast<int>(ctx->piece_events.size())) {
                cudaEvent_t ev;
                if ((error = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return;
                ctx->piece_events.push_back(ev);
            }
            if ((error = cudaEventRecord(ctx->piece_events[n_pieces], ctx->copy_stream)) != cudaSuccess) return;
            ++n_pieces;
        }
        if (n_pieces > 0) cudaStreamWaitEvent(stream, ctx->piece_events[n_pieces - 1], 0);
    }
};

int mlp_run(serb_ctx* ctx, const float* d_x32, const double* d_x64, long long n, double* d_proba,
            int* d_label, cudaStream_t stream) {
    if (!ctx->mlp.loaded) return fail(ctx, SERB_ERR_NO_MODEL, "no classifier loaded (call serb_mlp_load)");
    MlpParams p{};
    const Mlp& m = ctx->mlp;
    p.n_in = m.n_in; p.n_hidden = m.n_hidden; p.n_out = m.n_out; p.n_classes = m.n_classes;
    p.out_activation = m.out_activation;
    p.mean = m.mean.as<double>(); p.scale = m.scale.as<double>();
    p.w1 = m.w1.as<double>(); p.b1 = m.b1.as<double>(); p.w2 = m.w2.as<double>(); p.b2 = m.b2.as<double>();
    p.x32 = d_x32; p.x64 = d_x64; p.n = n; p.proba = d_proba; p.label = d_label;
    { ProfScope ps(ctx, 5, stream); SERB_CUDA(ctx, launch_mlp(p, stream)); }
    if (n > 0) ctx->launches += 1;
    return SERB_OK;
}

// PCM16 files -> ctx->pcm (int16, 16-byte aligned file starts) on the copy stream, then, on the
// compute stream, the segmented preparation kernels -> ctx->wave.  Runs of files that are
// contiguous in host memory (and whose sizes keep the alignment) travel as one copy.  A piece is
// enqueued when the first chunk that needs it is about to launch, so copies of later files overlap
// the kernels of earlier chunks.
struct PcmStager {
    serb_ctx* ctx;
    cudaStream_t stream;
    const int16_t* const* files;
    const std::vector<PcmFile>* layout;
    long long max_frames;
    int next_file = 0;            // first file whose copy is not enqueued yet
    int prepared = 0;             // first file not yet converted
    bool staged = false;          // pageable files: gathered into pinned slots (host_stage.h)
    int n_pieces = 0;
    bool started = false;
    cudaError_t error = cudaSuccess;
    void begin() {
        if (started) return;
        started = true;
        // the previous call's kernels may still read ctx->pcm / ctx->wave
        if ((error = cudaEventRecord(ctx->ev_done, ctx->stream)) != cudaSuccess) return;
        error = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_done, 0);
    }
    // one piece of about 32 MiB: whole files from next_file on, one event
    void enqueue_piece() {
        const std::vector<PcmFile>& lay = *layout;
        const int n_files = static_cast<int>(lay.size());
        constexpr long long kPiece = 16LL << 20;   // int16 samples per piece (32 MiB)
        long long in_piece = 0;
        StagedWriter writer{&ctx->stage_ring, ctx->copy_stream};
        while (next_file < n_files && in_piece < kPiece) {
            // extend a run while host and device layouts stay contiguous
            int run_hi = next_file + 1;
            long long run_samples = lay[next_file].frames * lay[next_file].channels;
            while (run_hi < n_files && in_piece + run_samples < kPiece &&
                   files[run_hi] == files[run_hi - 1] + lay[run_hi - 1].frames * lay[run_hi - 1].channels &&
                   lay[run_hi].pcm_off == lay[run_hi - 1].pcm_off + lay[run_hi - 1].frames * lay[run_hi - 1].channels) {
                run_samples += lay[run_hi].frames * lay[run_hi].channels;
                ++run_hi;
            }
            short* dst = ctx->pcm.as<short>() + lay[next_file].pcm_off;
            if (staged) error = writer.add(dst, files[next_file], static_cast<size_t>(run_samples) * sizeof(int16_t));
            else error = cudaMemcpyAsync(dst, files[next_file], static_cast<size_t>(run_samples) * sizeof(int16_t),
                                         cudaMemcpyHostToDevice, ctx->copy_stream);
            if (error != cudaSuccess) return;
            in_piece += run_samples;
            next_file = run_hi;
        }
        if (staged && (error = writer.flush()) != cudaSuccess) return;
        if (n_pieces >= static_cast<int>(ctx->piece_events.size())) {
            cudaEvent_t ev;
            if ((error = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return;
            ctx->piece_events.push_back(ev);
        }
        if ((error = cudaEventRecord(ctx->piece_events[n_pieces], ctx->copy_stream)) != cudaSuccess) return;
        ++n_pieces;
    }
    void finish() {       // see PieceWaiter::finish
        if (n_pieces > 0) cudaStreamWaitEvent(stream, ctx->piece_events[n_pieces - 1], 0);
    }
    // Called before the host plans the launch chain: the first piece (about the first chunk of the ramp)
    // crosses PCIe while the host builds its clip tables, instead of after.  Pinned sources only -- a
    // pageable piece is gathered by this very thread.
    void prefetch() {
        if (staged || layout->empty()) return;
        begin();
        if (error == cudaSuccess) enqueue_piece();
    }
    void operator()(long long max_end) {
        if (error != cudaSuccess) return;
        const std::vector<PcmFile>& lay = *layout;
        const int n_files = static_cast<int>(lay.size());
        begin();
        while (error == cudaSuccess && next_file < n_files && lay[next_file].wave_off < max_end) enqueue_piece();
        if (error != cudaSuccess) return;
        if (prepared < next_file) {
            if ((error = cudaStreamWaitEvent(stream, ctx->piece_events[n_pieces - 1], 0)) != cudaSuccess) return;
            ProfScope ps(ctx, 13, stream);
            error = launch_pcm_prepare_files(ctx->pcm.as<short>(), ctx->pcm_files.as<PcmFile>(), prepared, next_file,
                                             max_frames, ctx->pcm_peaks.as<int>(), ctx->wave.as<float>(), stream,
                                             &ctx->launches);
            prepared = next_file;
        }
    }
};

// shared body of serb_features_host_pcm16 / serb_infer_host_pcm16
int run_pcm16(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames, const int32_t* file_channels,
              int64_t n_files, const int64_t* clip_file, const int64_t* clip_starts, const int64_t* clip_lengths,
              int64_t n_clips, int32_t sample_rate, uint32_t flag_bits, bool infer, float* h_features, double* h_proba,
              int32_t* h_label_index) {
    if (n_files < 0 || n_clips < 0 || (n_files > 0 && (!h_files || !file_frames)) ||
        (n_clips > 0 && (!clip_file || !clip_starts || !clip_lengths)))
        return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
    if (n_files > 0x7fffffffLL) return fail(ctx, SERB_ERR_INVALID_ARG, "too many files");
    const int dim = serb_feature_dim(flag_bits);
    const Mlp& m = ctx->mlp;
    if (infer) {
        if (!m.loaded) return fail(ctx, SERB_ERR_NO_MODEL, "no classifier loaded (call serb_mlp_load)");
        if (n_clips > 0 && (!h_proba || !h_label_index)) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
        if (dim != m.n_in)
            return fail(ctx, SERB_ERR_INVALID_ARG,
                        "Feature vector size mismatch for loaded model. Expected " + std::to_string(m.n_in) +
                            ", got [" + std::to_string(dim) + "].");
    } else if (n_clips > 0 && !h_features) {
        return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
    }
    std::vector<PcmFile> layout(static_cast<size_t>(n_files));
    long long n_pcm = 0, n_wave = 0, max_frames = 0;
    for (int64_t f = 0; f < n_files; ++f) {
        const int ch = file_channels ? file_channels[f] : 1;
        if (file_frames[f] <= 0) return fail(ctx, SERB_ERR_EMPTY, "Audio file contains no samples.");
        if (!h_files[f] || ch < 1 || ch > 256) return fail(ctx, SERB_ERR_INVALID_ARG, "bad PCM16 file descriptor");
        layout[f] = PcmFile{n_pcm, n_wave, file_frames[f], ch, 0};
        n_pcm += (file_frames[f] * ch + 7) / 8 * 8;
        n_wave += (file_frames[f] + 3) / 4 * 4;
        max_frames = std::max<long long>(max_frames, file_frames[f]);
    }
    std::vector<int64_t> starts(static_cast<size_t>(n_clips));
    int64_t prev_file = 0;
    for (int64_t i = 0; i < n_clips; ++i) {
        const int64_t f = clip_file[i];
        if (f < prev_file || f >= n_files) return fail(ctx, SERB_ERR_INVALID_ARG, "clip_file must be non-decreasing and < n_files");
        prev_file = f;
        if (clip_lengths[i] <= 0) return fail(ctx, SERB_ERR_EMPTY, "Audio contains no samples.");
        if (clip_starts[i] < 0 || clip_starts[i] + clip_lengths[i] > file_frames[f])
            return fail(ctx, SERB_ERR_INVALID_ARG, "clip " + std::to_string(i) + " lies outside its file");
        starts[i] = layout[f].wave_off + clip_starts[i];
    }
    SERB_CUDA(ctx, ctx->pcm.reserve(std::max<long long>(n_pcm, 1) * sizeof(int16_t) + 64));
    SERB_CUDA(ctx, ctx->wave.reserve(std::max<long long>(n_wave, 1) * sizeof(float) + 64));
    SERB_CUDA(ctx, ctx->pcm_peaks.reserve(std::max<size_t>(layout.size(), 1) * sizeof(int)));
    SERB_CUDA(ctx, ctx->out.reserve(std::max<size_t>(static_cast<size_t>(n_clips) * dim, 1) * sizeof(float)));
    if (infer) {
        SERB_CUDA(ctx, ctx->proba.reserve(std::max<size_t>(static_cast<size_t>(n_clips) * m.n_classes, 1) * sizeof(double)));
        SERB_CUDA(ctx, ctx->labels.reserve(std::max<size_t>(static_cast<size_t>(n_clips), 1) * sizeof(int)));
    }
    int rc = upload(ctx, ctx->pcm_files, layout.data(), layout.size(), ctx->stream);
    if (rc) return rc;
    PcmStager stager{ctx, ctx->stream, h_files, &layout, max_frames};
    stager.staged = n_files > 0 && use_stage_ring(ctx, h_files[0]);
    ctx->ramp_pageable = stager.staged;
    ctx->ramp_chunks = true;
    if (n_clips > 0) stager.prefetch();
    rc = run_features(ctx, ctx->wave.as<float>(), n_wave, starts.data(), clip_lengths, n_clips, sample_rate, flag_bits,
                      ctx->out.as<float>(), ctx->stream, stager);
    ctx->ramp_chunks = false;
    ctx->ramp_pageable = false;
    stager.finish();
    if (!rc && stager.error != cudaSuccess) rc = fail_cuda(ctx, stager.error, "PCM16 staging");
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return rc; }
    if (n_clips == 0 || dim == 0) { SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return SERB_OK; }
    if (infer) {
        rc = mlp_run(ctx, ctx->out.as<float>(), nullptr, n_clips, ctx->proba.as<double>(), ctx->labels.as<int>(), ctx->stream);
        if (rc) return rc;
        SERB_CUDA(ctx, cudaMemcpyAsync(h_proba, ctx->proba.ptr, static_cast<size_t>(n_clips) * m.n_classes * sizeof(double),
                                       cudaMemcpyDeviceToHost, ctx->stream));
        SERB_CUDA(ctx, cudaMemcpyAsync(h_label_index, ctx->labels.ptr, static_cast<size_t>(n_clips) * sizeof(int),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (h_features)
        SERB_CUDA(ctx, cudaMemcpyAsync(h_features, ctx->out.ptr, static_cast<size_t>(n_clips) * dim * sizeof(float),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_status(ctx, ctx->stream);
}

}  // namespace

// =========================================================================================
extern "C" {

const char* serb_version(void) { return "ser_b200 0.1.0 (sm_100a)"; }

int serb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* serb_last_error(const serb_ctx* ctx) {
    if (!ctx) return g_create_error.c_str();
    // copied under the context lock into this thread's buffer: valid until this thread asks again
    thread_local std::string copy;
    std::lock_guard<std::mutex> lock(const_cast<serb_ctx*>(ctx)->mu);
    copy = ctx->err;
    return copy.c_str();
}

int serb_feature_dim(uint32_t flag_bits) { return make_offsets(flag_bits & SERB_FLAG_ALL).dim; }

int serb_ctx_create(int device_ordinal, serb_ctx** out_ctx) {
    if (!out_ctx) return fail(nullptr, SERB_ERR_INVALID_ARG, "out_ctx is NULL");
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(nullptr, SERB_ERR_NO_DEVICE, "no CUDA device available: ser_b200 has no CPU fallback");
    }
    if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, SERB_ERR_INVALID_ARG, "device ordinal out of range");
    if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) return fail_cuda(nullptr, e, "cudaSetDevice");
    serb_ctx* ctx = new serb_ctx();
    ctx->device = device_ordinal;
    if (const char* env = std::getenv("SERB_CHUNK_COLS")) {
        const int v = std::atoi(env);
        if (v >= 64) ctx->chunk_cols = std::min(v, 1 << 20);    // int element indices stay below 2^31 up to here
    }
    if (const char* env = std::getenv("SERB_HARM_SEG")) { const int v = std::atoi(env); if (v >= 16) ctx->harm_seg = v; }
    if (const char* env = std::getenv("SERB_RAMP_START")) { const int v = std::atoi(env); if (v >= 1024) ctx->ramp_start = v; }
    if (const char* env = std::getenv("SERB_RAMP_FACTOR_X10")) { const int v = std::atoi(env); if (v >= 11 && v <= 100) ctx->ramp_factor_x10 = v; }
    if (const char* env = std::getenv("SERB_RAMP_FACTOR_PAGEABLE_X10")) { const int v = std::atoi(env); if (v >= 11 && v <= 100) ctx->ramp_factor_pageable_x10 = v; }
    {
        // staging threads: the host's cores shared between the visible GPUs (one rank per GPU), 2 .. 8
        int n_gpus = 1;
        if (cudaGetDeviceCount(&n_gpus) != cudaSuccess || n_gpus < 1) { cudaGetLastError(); n_gpus = 1; }
        const int cores = static_cast<int>(std::thread::hardware_concurrency());
        ctx->stage_threads = std::max(2, std::min(8, cores / n_gpus));
    }
    if (const char* env = std::getenv("SERB_STAGE_THREADS")) { const int v = std::atoi(env); if (v >= 0 && v <= 64) ctx->stage_threads = v; }
    ctx->timed = true;
#define CREATE_CHECK(call)                                                                     \
    do { cudaError_t e2 = (call); if (e2 != cudaSuccess) { int rc2 = fail_cuda(nullptr, e2, #call); delete ctx; return rc2; } } while (0)
    CREATE_CHECK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CREATE_CHECK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CREATE_CHECK(cudaEventCreate(&ctx->ev_start));
    CREATE_CHECK(cudaEventCreate(&ctx->ev_stop));
    CREATE_CHECK(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
    CREATE_CHECK(cudaEventCreateWithFlags(&ctx->ev_chain, cudaEventDisableTiming));
    CREATE_CHECK(configure_stft());
    CREATE_CHECK(configure_proj());
    CREATE_CHECK(configure_short());
    CREATE_CHECK(configure_hpss());
    {
        std::vector<double> taps, hsq;
        decimation_taps(2, taps);
        if (static_cast<int>(taps.size()) != kDecTaps2) {
            delete ctx;
            return fail(nullptr, SERB_ERR_UNSUPPORTED, "factor-2 decimation filter length differs from the compiled kernel");
        }
        std::vector<float> taps32(taps.size());
        std::vector<double> taps64(taps.size());
        for (size_t i = 0; i < taps.size(); ++i) {
            taps64[i] = taps[i] * std::sqrt(2.0);
            taps32[i] = static_cast<float>(taps64[i]);
        }
        CREATE_CHECK(configure_cqt(taps32.data(), taps64.data()));
        CREATE_CHECK(configure_decimate_mma());
        std::vector<unsigned char> toeplitz(decimate_mma_table_bytes());
        decimate_mma_table(taps64.data(), toeplitz.data());
        CREATE_CHECK(ctx->dec_toeplitz.reserve(toeplitz.size()));
        CREATE_CHECK(cudaMemcpy(ctx->dec_toeplitz.ptr, toeplitz.data(), toeplitz.size(), cudaMemcpyHostToDevice));
        CREATE_CHECK(cudaDeviceGetAttribute(&ctx->n_sms, cudaDevAttrMultiProcessorCount, device_ordinal));
        if (const char* env = std::getenv("SERB_DECIMATE")) ctx->dec_mma = std::string(env) != "ffma";
        if (const char* env = std::getenv("SERB_ISTFT")) ctx->istft_fused = std::string(env) != "split";
        if (const char* env = std::getenv("SERB_CQT_SHARED_MAXHOP")) { const int v = std::atoi(env); if (v >= 0 && v <= 256) ctx->cqtc_max_hop = v; }
        if (const char* env = std::getenv("SERB_CQT")) {
            ctx->cqt_shared = std::string(env) != "percolumn";
            ctx->cqt_cols = std::string(env) != "rows" && std::string(env) != "percolumn";
        }
        hann_squared_2048(hsq);
        // behind the 2048 doubles: the overlap-add's window sum of squares where four frames
        // overlap, accumulated in frame order exactly as ola_sample does (float64 add, float32 store)
        std::vector<float> wss4(kHop);
        for (int r = 0; r < kHop; ++r) {
            float wss = 0.0f;
            for (int j = r + 3 * kHop; j >= r; j -= kHop) wss = static_cast<float>(static_cast<double>(wss) + hsq[j]);
            wss4[r] = wss;
        }
        CREATE_CHECK(ctx->hann_sq.reserve(hsq.size() * sizeof(double) + wss4.size() * sizeof(float)));
        CREATE_CHECK(cudaMemcpy(ctx->hann_sq.ptr, hsq.data(), hsq.size() * sizeof(double), cudaMemcpyHostToDevice));
        CREATE_CHECK(cudaMemcpy(static_cast<char*>(ctx->hann_sq.ptr) + hsq.size() * sizeof(double), wss4.data(),
                                wss4.size() * sizeof(float), cudaMemcpyHostToDevice));
        // constant-Q FFT twiddles: W_N^j = (cos, -sin) for N = 128..1024, then (cos, sin) 2 pi k / (2N)
        std::vector<float> tw;
        const double pi = 3.14159265358979323846;
        for (int n = 128; n <= 1024; n *= 2)
            for (int j = 0; j < n; ++j) {
                tw.push_back(static_cast<float>(std::cos(2.0 * pi * j / n)));
                tw.push_back(static_cast<float>(-std::sin(2.0 * pi * j / n)));
            }
        for (int n = 128; n <= 1024; n *= 2)
            for (int k = 0; k < n; ++k) {
                tw.push_back(static_cast<float>(std::cos(2.0 * pi * k / (2.0 * n))));
                tw.push_back(static_cast<float>(std::sin(2.0 * pi * k / (2.0 * n))));
            }
        CREATE_CHECK(ctx->cq_twiddles.reserve(tw.size() * sizeof(float)));
        CREATE_CHECK(cudaMemcpy(ctx->cq_twiddles.ptr, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    std::vector<double> edges(101), dct;
    for (int i = 0; i <= 100; ++i) edges[i] = tuning_edge(i);
    dct_matrix(dct);
    CREATE_CHECK(ctx->edges.reserve(edges.size() * sizeof(double)));
    CREATE_CHECK(cudaMemcpy(ctx->edges.ptr, edges.data(), edges.size() * sizeof(double), cudaMemcpyHostToDevice));
    {
        // FFT twiddles: W_1024^(k1 n2) as (cos, -sin) [32][32], then W_2048^k as (cos, sin) [1024]
        std::vector<float> tables(2 * 2048);
        const double pi = 3.14159265358979323846;
        for (int k1 = 0; k1 < 32; ++k1)
            for (int n2 = 0; n2 < 32; ++n2) {
                const double a = 2.0 * pi * static_cast<double>((k1 * n2) & 1023) / 1024.0;
                tables[2 * (k1 * 32 + n2)] = static_cast<float>(std::cos(a));
                tables[2 * (k1 * 32 + n2) + 1] = static_cast<float>(-std::sin(a));
            }
        for (int k = 0; k < 1024; ++k) {
            const double a = 2.0 * pi * static_cast<double>(k) / 2048.0;
            tables[2048 + 2 * k] = static_cast<float>(std::cos(a));
            tables[2048 + 2 * k + 1] = static_cast<float>(std::sin(a));
        }
        CREATE_CHECK(ctx->tables.reserve(tables.size() * sizeof(float)));
        CREATE_CHECK(cudaMemcpy(ctx->tables.ptr, tables.data(), tables.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    CREATE_CHECK(ctx->dct.reserve(dct.size() * sizeof(double)));
    CREATE_CHECK(cudaMemcpy(ctx->dct.ptr, dct.data(), dct.size() * sizeof(double), cudaMemcpyHostToDevice));
#undef CREATE_CHECK
    *out_ctx = ctx;
    return SERB_OK;
}

void serb_ctx_destroy(serb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy_stream);
    for (DevBuf* b : {&ctx->edges, &ctx->dct, &ctx->tables, &ctx->tile_clip, &ctx->spill, &ctx->logmel, &ctx->tile_mel, &ctx->tile_lmax,
                      &ctx->tile_chroma, &ctx->peaks, &ctx->peak_count, &ctx->clips, &ctx->short_clips,
                      &ctx->tuning, &ctx->short_tuning, &ctx->status, &ctx->wave, &ctx->out, &ctx->proba,
                      &ctx->labels, &ctx->x64, &ctx->pcm, &ctx->pcm_max, &ctx->pcm_files, &ctx->pcm_peaks, &ctx->mlp.mean, &ctx->mlp.scale,
                      &ctx->mlp.w1, &ctx->mlp.b1, &ctx->mlp.w2, &ctx->mlp.b2, &ctx->hann_sq, &ctx->cq_twiddles, &ctx->dec_toeplitz,
                      &ctx->cspec, &ctx->perc, &ctx->frames, &ctx->yharm, &ctx->yoct, &ctx->cqmag, &ctx->cq_chroma, &ctx->ton_part,
                      &ctx->long_idx, &ctx->long_state, &ctx->ton_clips, &ctx->ton_clips_a, &ctx->ton_clips_b, &ctx->ton_segs, &ctx->ton_runs, &ctx->ton_tuning,
                      &ctx->ton_tile_clip})
        b->release();
    for (auto& kv : ctx->sr_tables) {
        SrTables& t = kv.second;
        for (DevBuf* b : {&t.chroma_banks, &t.mel_start, &t.mel_count, &t.mel_offset, &t.mel_weights, &t.mel_points,
                          &t.cq_rows, &t.cq_vals, &t.cq_sets, &t.early_taps})
            b->release();
    }
    for (cudaEvent_t ev : ctx->piece_events) cudaEventDestroy(ev);
    ctx->stage_ring.destroy();
    cudaEventDestroy(ctx->ev_start);
    cudaEventDestroy(ctx->ev_stop);
    cudaEventDestroy(ctx->ev_done);
    cudaEventDestroy(ctx->ev_chain);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

int serb_features_device(serb_ctx* ctx, const float* d_wave, int64_t n_wave, const int64_t* starts,
                         const int64_t* lengths, int64_t n_clips, int32_t sample_rate, uint32_t flag_bits,
                         float* d_out, void* stream) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_clips > 0 && (!d_wave || !d_out)) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL device buffer");
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return run_features(ctx, d_wave, n_wave, starts, lengths, n_clips, sample_rate, flag_bits, d_out, s,
                        [](long long) {});
}

int serb_features_device_check(serb_ctx* ctx, void* stream) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->status.ptr) return SERB_OK;
    return check_status(ctx, stream ? static_cast<cudaStream_t>(stream) : ctx->stream);
}

int serb_features_host(serb_ctx* ctx, const float* h_wave, int64_t n_wave, const int64_t* starts,
                       const int64_t* lengths, int64_t n_clips, int32_t sample_rate, uint32_t flag_bits,
                       float* h_out) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_clips > 0 && (!h_wave || !h_out)) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
    const int dim = serb_feature_dim(flag_bits);
    PieceWaiter waiter{ctx, ctx->stream, nullptr, 0};
    int rc = stage_wave(ctx, h_wave, n_wave, waiter);
    if (rc) return rc;
    SERB_CUDA(ctx, ctx->out.reserve(std::max<size_t>(static_cast<size_t>(n_clips) * dim, 1) * sizeof(float)));
    ctx->ramp_pageable = waiter.staged;
    ctx->ramp_chunks = true;
    if (n_clips > 0) waiter.prefetch();
    rc = run_features(ctx, ctx->wave.as<float>(), n_wave, starts, lengths, n_clips, sample_rate, flag_bits,
                      ctx->out.as<float>(), ctx->stream, waiter);
    ctx->ramp_chunks = false;
    ctx->ramp_pageable = false;
    waiter.finish();
    if (!rc && waiter.error != cudaSuccess) rc = fail_cuda(ctx, waiter.error, "waveform staging");
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return rc; }
    if (n_clips > 0 && dim > 0)
        SERB_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->out.ptr, static_cast<size_t>(n_clips) * dim * sizeof(float),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_clips > 0 && dim > 0) return check_status(ctx, ctx->stream);
    return SERB_OK;
}

int serb_features_host_clips(serb_ctx* ctx, const float* const* h_clips, const int64_t* lengths, int64_t n_clips,
                             int32_t sample_rate, uint32_t flag_bits, float* h_out) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_clips < 0 || (n_clips > 0 && (!h_clips || !lengths || !h_out))) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
    std::vector<int64_t> starts(static_cast<size_t>(n_clips));
    long long n_wave = 0;
    for (int64_t i = 0; i < n_clips; ++i) {
        if (lengths[i] <= 0) return fail(ctx, SERB_ERR_EMPTY, "Audio contains no samples.");
        if (!h_clips[i]) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL clip pointer");
        starts[i] = n_wave;
        n_wave += (lengths[i] + 3) / 4 * 4;       // 16-byte aligned clip starts: every tile takes the TMA path
    }
    const int dim = serb_feature_dim(flag_bits);
    SERB_CUDA(ctx, ctx->wave.reserve(std::max<long long>(n_wave, 1) * sizeof(float) + 64));
    SERB_CUDA(ctx, ctx->out.reserve(std::max<size_t>(static_cast<size_t>(n_clips) * dim, 1) * sizeof(float)));
    ClipStager stager{ctx, ctx->stream, h_clips, lengths, &starts, n_clips};
    stager.staged = n_clips > 0 && use_stage_ring(ctx, h_clips[0]);
    ctx->ramp_pageable = stager.staged;
    ctx->ramp_chunks = true;
    int rc = run_features(ctx, ctx->wave.as<float>(), n_wave, starts.data(), lengths, n_clips, sample_rate, flag_bits,
                          ctx->out.as<float>(), ctx->stream, stager);
    ctx->ramp_chunks = false;
    ctx->ramp_pageable = false;
    if (!rc && stager.error != cudaSuccess) rc = fail_cuda(ctx, stager.error, "clip staging");
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return rc; }
    if (n_clips > 0 && dim > 0)
        SERB_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->out.ptr, static_cast<size_t>(n_clips) * dim * sizeof(float),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_clips > 0 && dim > 0) return check_status(ctx, ctx->stream);
    return SERB_OK;
}

int serb_mlp_load(serb_ctx* ctx, int32_t n_in, int32_t n_hidden, int32_t n_out, const double* mean,
                  const double* scale, const double* w1, const double* b1, const double* w2, const double* b2,
                  int32_t out_activation) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_in <= 0 || n_hidden <= 0 || n_out <= 0 || !mean || !scale || !w1 || !b1 || !w2 || !b2)
        return fail(ctx, SERB_ERR_INVALID_ARG, "bad classifier shapes");
    if (out_activation != SERB_OUT_SOFTMAX && out_activation != SERB_OUT_LOGISTIC)
        return fail(ctx, SERB_ERR_INVALID_ARG, "unsupported output activation");
    if (mlp_smem_bytes(n_in, n_hidden, n_out) > 200 * 1024)
        return fail(ctx, SERB_ERR_UNSUPPORTED, "classifier too wide for the fused kernel");
    Mlp& m = ctx->mlp;
    m.loaded = false;
    int rc;
    if ((rc = upload(ctx, m.mean, mean, n_in, ctx->stream))) return rc;
    if ((rc = upload(ctx, m.scale, scale, n_in, ctx->stream))) return rc;
    if ((rc = upload(ctx, m.w1, w1, static_cast<size_t>(n_in) * n_hidden, ctx->stream))) return rc;
    if ((rc = upload(ctx, m.b1, b1, n_hidden, ctx->stream))) return rc;
    if ((rc = upload(ctx, m.w2, w2, static_cast<size_t>(n_hidden) * n_out, ctx->stream))) return rc;
    if ((rc = upload(ctx, m.b2, b2, n_out, ctx->stream))) return rc;
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SERB_CUDA(ctx, configure_mlp(n_in, n_hidden, n_out));
    m.n_in = n_in; m.n_hidden = n_hidden; m.n_out = n_out; m.out_activation = out_activation;
    m.n_classes = (out_activation == SERB_OUT_LOGISTIC && n_out == 1) ? 2 : n_out;
    m.loaded = true;
    return SERB_OK;
}

int serb_mlp_n_classes(const serb_ctx* ctx) { return (ctx && ctx->mlp.loaded) ? ctx->mlp.n_classes : 0; }

int serb_mlp_predict_device(serb_ctx* ctx, const float* d_x, int64_t n, double* d_proba, int32_t* d_label_index,
                            void* stream) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n < 0 || (n > 0 && (!d_x || !d_proba || !d_label_index))) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL buffer");
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return mlp_run(ctx, d_x, nullptr, n, d_proba, d_label_index, s);
}

int serb_mlp_predict_host(serb_ctx* ctx, const double* h_x, int64_t n, double* h_proba, int32_t* h_label_index) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mlp.loaded) return fail(ctx, SERB_ERR_NO_MODEL, "no classifier loaded (call serb_mlp_load)");
    if (n < 0 || (n > 0 && (!h_x || !h_proba || !h_label_index))) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL buffer");
    if (n == 0) return SERB_OK;
    const Mlp& m = ctx->mlp;
    SERB_CUDA(ctx, ctx->x64.reserve(static_cast<size_t>(n) * m.n_in * sizeof(double)));
    SERB_CUDA(ctx, ctx->proba.reserve(static_cast<size_t>(n) * m.n_classes * sizeof(double)));
    SERB_CUDA(ctx, ctx->labels.reserve(static_cast<size_t>(n) * sizeof(int)));
    SERB_CUDA(ctx, cudaMemcpyAsync(ctx->x64.ptr, h_x, static_cast<size_t>(n) * m.n_in * sizeof(double),
                                   cudaMemcpyHostToDevice, ctx->stream));
    int rc = mlp_run(ctx, nullptr, ctx->x64.as<double>(), n, ctx->proba.as<double>(), ctx->labels.as<int>(), ctx->stream);
    if (rc) return rc;
    SERB_CUDA(ctx, cudaMemcpyAsync(h_proba, ctx->proba.ptr, static_cast<size_t>(n) * m.n_classes * sizeof(double),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(h_label_index, ctx->labels.ptr, static_cast<size_t>(n) * sizeof(int),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SERB_OK;
}

int serb_infer_host(serb_ctx* ctx, const float* h_wave, int64_t n_wave, const int64_t* starts,
                    const int64_t* lengths, int64_t n_clips, int32_t sample_rate, uint32_t flag_bits,
                    float* h_features, double* h_proba, int32_t* h_label_index) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mlp.loaded) return fail(ctx, SERB_ERR_NO_MODEL, "no classifier loaded (call serb_mlp_load)");
    if (n_clips > 0 && (!h_wave || !h_proba || !h_label_index)) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
    const int dim = serb_feature_dim(flag_bits);
    const Mlp& m = ctx->mlp;
    if (dim != m.n_in)
        return fail(ctx, SERB_ERR_INVALID_ARG,
                    "Feature vector size mismatch for loaded model. Expected " + std::to_string(m.n_in) +
                        ", got [" + std::to_string(dim) + "].");
    PieceWaiter waiter{ctx, ctx->stream, nullptr, 0};
    int rc = stage_wave(ctx, h_wave, n_wave, waiter);
    if (rc) return rc;
    SERB_CUDA(ctx, ctx->out.reserve(std::max<size_t>(static_cast<size_t>(n_clips) * dim, 1) * sizeof(float)));
    SERB_CUDA(ctx, ctx->proba.reserve(std::max<size_t>(static_cast<size_t>(n_clips) * m.n_classes, 1) * sizeof(double)));
    SERB_CUDA(ctx, ctx->labels.reserve(std::max<size_t>(static_cast<size_t>(n_clips), 1) * sizeof(int)));
    ctx->ramp_pageable = waiter.staged;
    ctx->ramp_chunks = true;
    if (n_clips > 0) waiter.prefetch();
    rc = run_features(ctx, ctx->wave.as<float>(), n_wave, starts, lengths, n_clips, sample_rate, flag_bits,
                      ctx->out.as<float>(), ctx->stream, waiter);
    ctx->ramp_chunks = false;
    ctx->ramp_pageable = false;
    waiter.finish();
    if (!rc && waiter.error != cudaSuccess) rc = fail_cuda(ctx, waiter.error, "waveform staging");
    if (rc) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); return rc; }
    if (n_clips == 0) return SERB_OK;
    rc = mlp_run(ctx, ctx->out.as<float>(), nullptr, n_clips, ctx->proba.as<double>(), ctx->labels.as<int>(), ctx->stream);
    if (rc) return rc;
    if (h_features)
        SERB_CUDA(ctx, cudaMemcpyAsync(h_features, ctx->out.ptr, static_cast<size_t>(n_clips) * dim * sizeof(float),
                                       cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(h_proba, ctx->proba.ptr, static_cast<size_t>(n_clips) * m.n_classes * sizeof(double),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(h_label_index, ctx->labels.ptr, static_cast<size_t>(n_clips) * sizeof(int),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_status(ctx, ctx->stream);
}

int serb_features_host_pcm16(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames,
                             const int32_t* file_channels, int64_t n_files, const int64_t* clip_file,
                             const int64_t* clip_starts, const int64_t* clip_lengths, int64_t n_clips,
                             int32_t sample_rate, uint32_t flag_bits, float* h_out) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_pcm16(ctx, h_files, file_frames, file_channels, n_files, clip_file, clip_starts, clip_lengths, n_clips,
                     sample_rate, flag_bits, false, h_out, nullptr, nullptr);
}

int serb_infer_host_pcm16(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames,
                          const int32_t* file_channels, int64_t n_files, const int64_t* clip_file,
                          const int64_t* clip_starts, const int64_t* clip_lengths, int64_t n_clips,
                          int32_t sample_rate, uint32_t flag_bits, float* h_features, double* h_proba,
                          int32_t* h_label_index) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    return run_pcm16(ctx, h_files, file_frames, file_channels, n_files, clip_file, clip_starts, clip_lengths, n_clips,
                     sample_rate, flag_bits, true, h_features, h_proba, h_label_index);
}

int serb_prepare_pcm16_files_host(serb_ctx* ctx, const int16_t* const* h_files, const int64_t* file_frames,
                                  const int32_t* file_channels, int64_t n_files, float* const* h_out) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_files < 0 || n_files > 0x7fffffffLL || (n_files > 0 && (!h_files || !file_frames || !h_out)))
        return fail(ctx, SERB_ERR_INVALID_ARG, "NULL host buffer");
    std::vector<PcmFile> layout(static_cast<size_t>(n_files));
    long long n_pcm = 0, n_wave = 0, max_frames = 0;
    for (int64_t f = 0; f < n_files; ++f) {
        const int ch = file_channels ? file_channels[f] : 1;
        if (file_frames[f] <= 0) return fail(ctx, SERB_ERR_EMPTY, "Audio file contains no samples.");
        if (!h_files[f] || !h_out[f] || ch < 1 || ch > 256) return fail(ctx, SERB_ERR_INVALID_ARG, "bad PCM16 file descriptor");
        layout[f] = PcmFile{n_pcm, n_wave, file_frames[f], ch, 0};
        n_pcm += (file_frames[f] * ch + 7) / 8 * 8;
        n_wave += (file_frames[f] + 3) / 4 * 4;
        max_frames = std::max<long long>(max_frames, file_frames[f]);
    }
    if (n_files == 0) return SERB_OK;
    SERB_CUDA(ctx, ctx->pcm.reserve(n_pcm * sizeof(int16_t) + 64));
    SERB_CUDA(ctx, ctx->wave.reserve(n_wave * sizeof(float) + 64));
    SERB_CUDA(ctx, ctx->pcm_peaks.reserve(layout.size() * sizeof(int)));
    int rc = upload(ctx, ctx->pcm_files, layout.data(), layout.size(), ctx->stream);
    if (rc) return rc;
    for (int64_t f = 0; f < n_files; ++f)
        SERB_CUDA(ctx, cudaMemcpyAsync(ctx->pcm.as<short>() + layout[f].pcm_off, h_files[f],
                                       static_cast<size_t>(layout[f].frames) * layout[f].channels * sizeof(int16_t),
                                       cudaMemcpyHostToDevice, ctx->stream));
    {
        ProfScope ps(ctx, 13, ctx->stream);
        SERB_CUDA(ctx, launch_pcm_prepare_files(ctx->pcm.as<short>(), ctx->pcm_files.as<PcmFile>(), 0, static_cast<int>(n_files),
                                                max_frames, ctx->pcm_peaks.as<int>(), ctx->wave.as<float>(), ctx->stream,
                                                &ctx->launches));
    }
    for (int64_t f = 0; f < n_files; ++f)
        SERB_CUDA(ctx, cudaMemcpyAsync(h_out[f], ctx->wave.as<float>() + layout[f].wave_off,
                                       static_cast<size_t>(layout[f].frames) * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SERB_OK;
}

int serb_pool_frames_host(serb_ctx* ctx, const float* h_embeddings, int64_t n_frames, int32_t dim,
                          const int32_t* h_lo, const int32_t* h_hi, int64_t n_windows, int32_t mode, double* h_out) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mode < 0 || mode > 2 || dim <= 0 || n_frames < 0 || n_windows < 0)
        return fail(ctx, SERB_ERR_INVALID_ARG, "bad pooling arguments");
    if (n_windows == 0) return SERB_OK;
    if (!h_embeddings || !h_lo || !h_hi || !h_out) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL buffer");
    for (int64_t w = 0; w < n_windows; ++w)
        if (h_lo[w] < 0 || h_hi[w] > n_frames || h_lo[w] >= h_hi[w])
            return fail(ctx, SERB_ERR_INVALID_ARG, "Pooling window does not overlap any encoded frames");
    const size_t width = static_cast<size_t>(mode == 1 ? 2 * dim : dim);
    DevBuf& emb = ctx->wave;      // staging buffers are free between calls
    DevBuf& idx = ctx->labels;
    DevBuf& out = ctx->proba;
    SERB_CUDA(ctx, emb.reserve(static_cast<size_t>(n_frames) * dim * sizeof(float) + 64));
    SERB_CUDA(ctx, idx.reserve(static_cast<size_t>(n_windows) * 2 * sizeof(int)));
    SERB_CUDA(ctx, out.reserve(static_cast<size_t>(n_windows) * width * sizeof(double)));
    SERB_CUDA(ctx, cudaMemcpyAsync(emb.ptr, h_embeddings, static_cast<size_t>(n_frames) * dim * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(idx.ptr, h_lo, static_cast<size_t>(n_windows) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(idx.as<int>() + n_windows, h_hi, static_cast<size_t>(n_windows) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    SERB_CUDA(ctx, launch_pool_stats(emb.as<float>(), dim, idx.as<int>(), idx.as<int>() + n_windows, n_windows, mode, out.as<double>(), ctx->stream));
    ctx->launches += 1;
    SERB_CUDA(ctx, cudaMemcpyAsync(h_out, out.ptr, static_cast<size_t>(n_windows) * width * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SERB_OK;
}

int serb_prepare_pcm16_device(serb_ctx* ctx, const int16_t* d_pcm, int64_t n, float* d_out, void* stream) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n < 0 || (n > 0 && (!d_pcm || !d_out))) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL buffer");
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    SERB_CUDA(ctx, ctx->pcm_max.reserve(sizeof(int)));
    SERB_CUDA(ctx, launch_prepare_pcm16(d_pcm, n, ctx->pcm_max.as<int>(), d_out, s));
    if (n > 0) ctx->launches += 2;
    return SERB_OK;
}

int serb_prepare_pcm16_host(serb_ctx* ctx, const int16_t* h_pcm, int64_t n, float* h_out) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n < 0 || (n > 0 && (!h_pcm || !h_out))) return fail(ctx, SERB_ERR_INVALID_ARG, "NULL buffer");
    if (n == 0) return SERB_OK;
    SERB_CUDA(ctx, ctx->pcm.reserve(static_cast<size_t>(n) * sizeof(int16_t)));
    SERB_CUDA(ctx, ctx->pcm_max.reserve(sizeof(int)));
    SERB_CUDA(ctx, ctx->wave.reserve(static_cast<size_t>(n) * sizeof(float) + 64));
    SERB_CUDA(ctx, cudaMemcpyAsync(ctx->pcm.ptr, h_pcm, static_cast<size_t>(n) * sizeof(int16_t),
                                   cudaMemcpyHostToDevice, ctx->stream));
    SERB_CUDA(ctx, launch_prepare_pcm16(ctx->pcm.as<short>(), n, ctx->pcm_max.as<int>(), ctx->wave.as<float>(), ctx->stream));
    ctx->launches += 2;
    SERB_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->wave.ptr, static_cast<size_t>(n) * sizeof(float),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SERB_OK;
}

// ---- introspection -----------------------------------------------------------------------
int serb_debug_filterbank(int32_t kind, int32_t sample_rate, int32_t n_fft, int32_t tuning_index, float* out) {
    if (!out || n_fft < 2) return SERB_ERR_INVALID_ARG;
    if (kind == 0) {
        if (sample_rate <= 0) return SERB_ERR_SAMPLE_RATE;
        std::vector<float> w;
        mel_filterbank(sample_rate, n_fft, w);
        std::memcpy(out, w.data(), w.size() * sizeof(float));
    } else if (kind == 1) {
        if (sample_rate <= 0) return SERB_ERR_SAMPLE_RATE;
        if (tuning_index < 0 || tuning_index >= kNTunings) return SERB_ERR_INVALID_ARG;
        std::vector<float> w;
        chroma_filterbank(sample_rate, n_fft, tuning_edge(tuning_index), w);
        std::memcpy(out, w.data(), w.size() * sizeof(float));
    } else if (kind == 2) {
        std::vector<double> d;
        dct_matrix(d);
        for (size_t i = 0; i < d.size(); ++i) out[i] = static_cast<float>(d[i]);
    } else if (kind == 3) {
        std::vector<double> w;
        hann_periodic(n_fft, w);
        for (size_t i = 0; i < w.size(); ++i) out[i] = static_cast<float>(w[i]);
    } else {
        return SERB_ERR_INVALID_ARG;
    }
    return SERB_OK;
}

int serb_debug_stft_host(serb_ctx* ctx, const float* h_wave, int64_t n, float* h_out, int64_t n_cols) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!h_wave || !h_out || n < kNFft || n > 0x7fffffffLL) return fail(ctx, SERB_ERR_INVALID_ARG, "need one clip of >= 2048 samples");
    ClipDev c{};
    c.start = 0; c.length = static_cast<int>(n); c.n_cols = 1 + static_cast<int>(n / kHop);
    if (n_cols != c.n_cols) return fail(ctx, SERB_ERR_INVALID_ARG, "n_cols must be 1 + n / 512");
    const int tiles = (c.n_cols + kColsPerTile - 1) / kColsPerTile;
    SERB_CUDA(ctx, ctx->wave.reserve(static_cast<size_t>(n) * sizeof(float) + 64));
    SERB_CUDA(ctx, ctx->spill.reserve(static_cast<size_t>(c.n_cols) * kSpillStride * sizeof(float)));
    SERB_CUDA(ctx, ctx->tile_clip.reserve(static_cast<size_t>(tiles) * sizeof(int)));
    SERB_CUDA(ctx, ctx->status.reserve(sizeof(int)));
    SERB_CUDA(ctx, cudaMemsetAsync(ctx->status.ptr, 0, sizeof(int), ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(ctx->wave.ptr, h_wave, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int rc = upload(ctx, ctx->clips, &c, 1, ctx->stream);
    if (rc) return rc;
    SERB_CUDA(ctx, launch_expand_tiles(ctx->clips.as<ClipDev>(), 1, ctx->tile_clip.as<int>(), ctx->stream));
    StftParams sp{};
    sp.wave = ctx->wave.as<float>();
    sp.clips = ctx->clips.as<ClipDev>();
    sp.n_clips = 1;
    sp.tile_clip = ctx->tile_clip.as<int>();
    sp.tables = ctx->tables.as<float2>();
    sp.spill = ctx->spill.as<float>();
    sp.status = ctx->status.as<int>();
    SERB_CUDA(ctx, launch_stft(sp, tiles, ctx->stream));
    ctx->launches += 2;
    SERB_CUDA(ctx, cudaMemcpy2DAsync(h_out, kNBins * sizeof(float), ctx->spill.ptr, kSpillStride * sizeof(float),
                                     kNBins * sizeof(float), c.n_cols, cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SERB_OK;
}

int serb_debug_last_tuning(serb_ctx* ctx, int32_t* h_out, int64_t n_clips) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!h_out || n_clips != ctx->last_n_clips) return fail(ctx, SERB_ERR_INVALID_ARG, "n_clips differs from the last features call");
    for (long long i = 0; i < n_clips; ++i) h_out[i] = -1;
    if (!ctx->last_had_chroma) return SERB_OK;
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int> tmp(ctx->last_tuning_rows.size());
    if (!tmp.empty()) {
        SERB_CUDA(ctx, cudaMemcpy(tmp.data(), ctx->tuning.ptr, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) h_out[ctx->last_tuning_rows[i]] = tmp[i];
    }
    tmp.resize(ctx->last_short_rows.size());
    if (!tmp.empty()) {
        SERB_CUDA(ctx, cudaMemcpy(tmp.data(), ctx->short_tuning.ptr, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tmp.size(); ++i) h_out[ctx->last_short_rows[i]] = tmp[i];
    }
    return SERB_OK;
}

int serb_debug_tonnetz_stages(serb_ctx* ctx, const float* h_wave, int64_t n, int32_t sample_rate, float* h_yharm,
                              int32_t* tuning_index, float* h_cqmag, int64_t cq_capacity_rows, int32_t* cq_cols,
                              float* h_tonnetz6) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!h_wave || n <= 0 || n > 0x7fffffffLL) return fail(ctx, SERB_ERR_INVALID_ARG, "need one non-empty clip");
    if (sample_rate <= 0) return fail(ctx, SERB_ERR_SAMPLE_RATE, "Sample rate must be a positive integer.");
    SrTables* tab = nullptr;
    int rc = get_sr_tables(ctx, sample_rate, &tab);
    if (rc) return rc;
    SERB_CUDA(ctx, ctx->wave.reserve(static_cast<size_t>(n) * sizeof(float) + 64));
    SERB_CUDA(ctx, ctx->out.reserve(6 * sizeof(float)));
    SERB_CUDA(ctx, ctx->status.reserve(sizeof(int)));
    SERB_CUDA(ctx, cudaMemsetAsync(ctx->status.ptr, 0, sizeof(int), ctx->stream));
    SERB_CUDA(ctx, cudaMemcpyAsync(ctx->wave.ptr, h_wave, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    const int64_t start = 0, length = n;
    ctx->keep_cqmag = h_cqmag != nullptr;
    rc = run_features(ctx, ctx->wave.as<float>(), n, &start, &length, 1, sample_rate, SERB_FLAG_TONNETZ,
                      ctx->out.as<float>(), ctx->stream, [](long long) {});
    ctx->keep_cqmag = false;
    if (rc) return rc;
    const long long plen = std::max<long long>(n, 512);
    int cq = 0x7fffffff;
    {
        long long ln = (plen + tab->plan.early_factor - 1) / tab->plan.early_factor;
        for (int l = 0; l < kCqOctaves; ++l) { cq = std::min<long long>(cq, 1 + ln / (tab->plan.hop0 >> l)); ln = (ln + 1) >> 1; }
    }
    if (cq_cols) *cq_cols = cq;
    if (h_yharm) SERB_CUDA(ctx, cudaMemcpyAsync(h_yharm, ctx->yharm.ptr, static_cast<size_t>(plen) * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    if (tuning_index) SERB_CUDA(ctx, cudaMemcpyAsync(tuning_index, ctx->ton_tuning.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_cqmag) {
        if (cq_capacity_rows < cq) return fail(ctx, SERB_ERR_INVALID_ARG, "cqmag capacity too small");
        SERB_CUDA(ctx, cudaMemcpyAsync(h_cqmag, ctx->cqmag.ptr, static_cast<size_t>(cq) * kCqBins * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (h_tonnetz6) SERB_CUDA(ctx, cudaMemcpyAsync(h_tonnetz6, ctx->out.ptr, 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    SERB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SERB_OK;
}

int serb_debug_cqt_plan(int32_t sample_rate, int32_t* out10) {
    if (!out10 || sample_rate <= 0) return SERB_ERR_INVALID_ARG;
    CqtPlan plan;
    cqt_plan(sample_rate, plan);
    out10[0] = plan.status;
    out10[1] = plan.early_factor;
    out10[2] = plan.hop0;
    for (int i = 0; i < kCqtOctaves; ++i) out10[3 + i] = plan.n_fft[i];
    return SERB_OK;
}

int serb_debug_cqt_set_basis(int32_t sample_rate, int32_t tuning_index, int32_t octave, float* out_basis, float* out_scale36) {
    if (sample_rate <= 0 || tuning_index < 0 || tuning_index >= kNTunings || octave < 0 || octave >= kCqtOctaves || !out_basis)
        return SERB_ERR_INVALID_ARG;
    CqtPlan plan;
    cqt_plan(sample_rate, plan);
    if (plan.status != 0) return SERB_ERR_UNSUPPORTED;
    std::vector<CqtSetBank> banks(kCqtOctaves);
    if (!cqt_set_banks(plan, tuning_index, banks.data())) return SERB_ERR_UNSUPPORTED;
    // the column-mapped layout expanded back to [36][1 + n_fft/2] complex64, exactly as cqtc_kernel reads it
    const CqtSetBank& sb = banks[octave];
    const int n_bins = 1 + plan.n_fft[octave] / 2;
    std::memset(out_basis, 0, static_cast<size_t>(kCqtBpo) * n_bins * 2 * sizeof(float));
    for (int s = 0; s < kCqtSets; ++s) {
        const CqtSet& set = sb.sets[s];
        const int rpad = s < 4 ? 4 : 2;
        for (int q = 0; q < set.nrows; ++q) {
            const int j = set.bin[q] % kCqtBpo;
            if (out_scale36) out_scale36[j] = set.scale[q];
            for (int b = 0; b < set.ulen; ++b) {
                const int k = sb.bin_lo + set.u0 + b;
                if (k < 0 || k >= n_bins) return SERB_ERR_UNSUPPORTED;
                const size_t at = static_cast<size_t>(set.off) + static_cast<size_t>(b) * rpad + q;
                out_basis[(static_cast<size_t>(j) * n_bins + k) * 2] = sb.vals[2 * at];
                out_basis[(static_cast<size_t>(j) * n_bins + k) * 2 + 1] = sb.vals[2 * at + 1];
            }
        }
    }
    return SERB_OK;
}

int serb_debug_cqt_basis(int32_t sample_rate, int32_t tuning_index, int32_t octave, float* out_basis, float* out_scale36) {
    if (sample_rate <= 0 || tuning_index < 0 || tuning_index >= kNTunings || octave < 0 || octave >= kCqtOctaves)
        return SERB_ERR_INVALID_ARG;
    CqtPlan plan;
    cqt_plan(sample_rate, plan);
    if (plan.status != 0) return SERB_ERR_UNSUPPORTED;
    if (out_basis) {
        std::vector<float> dense;
        cqt_basis_dense(plan, tuning_index, octave, dense);
        std::memcpy(out_basis, dense.data(), dense.size() * sizeof(float));
    }
    if (out_scale36) {
        CqtBank bank;
        if (!cqt_bank(plan, tuning_index, bank, false)) return SERB_ERR_UNSUPPORTED;
        for (int j = 0; j < kCqtBpo; ++j) out_scale36[j] = bank.rows[static_cast<size_t>(octave) * kCqtBpo + j].scale;
    }
    return SERB_OK;
}

int serb_debug_decimation_taps(int32_t factor, double* out, int32_t capacity) {
    if (factor < 2 || factor > 8) return SERB_ERR_INVALID_ARG;
    std::vector<double> taps;
    decimation_taps(factor, taps);
    if (out) {
        if (capacity < static_cast<int32_t>(taps.size())) return SERB_ERR_INVALID_ARG;
        std::memcpy(out, taps.data(), taps.size() * sizeof(double));
    }
    return static_cast<int>(taps.size());
}

int64_t serb_debug_launch_count(const serb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int serb_debug_fp32_peak(serb_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop{};
    SERB_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    SERB_CUDA(ctx, ctx->x64.reserve(static_cast<size_t>(blocks) * 256 * sizeof(float)));
    cudaEvent_t a, b;
    SERB_CUDA(ctx, cudaEventCreate(&a));
    SERB_CUDA(ctx, cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {           // first pass warms up; keep the best of the rest
        cudaEventRecord(a, ctx->stream);
        cudaError_t e = launch_fp32_peak(ctx->x64.as<float>(), blocks, iters, ctx->stream);
        cudaEventRecord(b, ctx->stream);
        if (e == cudaSuccess) e = cudaEventSynchronize(b);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
        if (e != cudaSuccess) { cudaEventDestroy(a); cudaEventDestroy(b); return fail_cuda(ctx, e, "fp32 peak probe"); }
        const double flops = 2.0 * 16.0 * iters * 256.0 * blocks;
        if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    ctx->launches += 5;
    *tflops = best;
    return SERB_OK;
}

int serb_debug_set_profile(serb_ctx* ctx, int32_t enabled) {
    if (!ctx) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    ctx->profile = enabled != 0;
    for (int i = 0; i < kProfKinds; ++i) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
    return SERB_OK;
}

int serb_debug_kernel_ms(serb_ctx* ctx, int32_t kind, double* total_ms, int64_t* n_launches) {
    if (!ctx || kind < 0 || kind >= kProfKinds || !total_ms || !n_launches) return SERB_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(ctx->mu);
    SERB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (auto& rec : ctx->prof_recs) {
        float ms = 0.f;
        SERB_CUDA(ctx, cudaEventSynchronize(rec.b));
        SERB_CUDA(ctx, cudaEventElapsedTime(&ms, rec.a, rec.b));
        ctx->prof_ms[rec.kind] += ms;
        ctx->prof_n[rec.kind] += 1;
        ctx->prof_pool.push_back(rec.a);
        ctx->prof_pool.push_back(rec.b);
    }
    ctx->prof_recs.clear();
    *total_ms = ctx->prof_ms[kind];
    *n_launches = ctx->prof_n[kind];
    return SERB_OK;
}

float serb_debug_last_compute_ms(serb_ctx* ctx) {
    if (!ctx) return -1.f;
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(ctx->ev_stop) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    if (cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    return ms;
}

}  // extern "C"
