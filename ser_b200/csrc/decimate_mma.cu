// Factor-2 decimation of the constant-Q recursion as a banded-Toeplitz GEMM on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Arithmetic to reproduce (oracle/shim/librosa/core.py:_soxr_hq_decimate, resample(scale=True)):
//   out[m] = sum_{k < K} h[k] x[2 m + (K - 1) / 2 - k],   K = 389, h = sqrt(2) x the HQ low-pass.
//
// GEMM view.  128 consecutive outputs of one signal (a "window") read 643 consecutive samples;
// written as D[128 x N] = T[128 x 656] . X[656 x N], T is the banded Toeplitz matrix of the taps
// and every column of X is the sample span of another window.  float32 accuracy comes from three
// bf16 terms per operand (x = x0 + x1 + x2, h = h0 + h1 + h2, 8 mantissa bits each) and the six
// products of total order <= 2 (x0 h0 | x0 h1, x1 h0 | x1 h1, x0 h2, x2 h0), accumulated in float32:
// the leading product in one TMEM accumulator, the five small ones in a second (their sum is 2^-8
// of the result, so the tensor core's truncating accumulate costs them nothing), added once in
// the epilogue.
//
// Layouts (both operands K-major, no swizzle: 8 x 16-byte core matrices, UMMA "INTERLEAVE"):
// * T depends on (2 row - column) only.  With the output ROWS REVERSED (row i' = output 127 - i')
//   core matrix (r', c) depends on u = 2 r' + c alone, so the whole 128 x 656 operand is 112 core
//   matrices (14 KB per bf16 term) addressed with SBO = 2 cores, LBO = 1 core: it stays resident in
//   shared memory for the life of the CTA and costs no traffic.
// * X: the N = 8 R columns are 8 "lanes" (core-matrix rows) x R consecutive windows.  Lane a holds
//   one contiguous span of one signal as 16-byte chunks of 8 samples, chunk q at 128 q + 16 a;
//   window rn of the lane starts 32 chunks after window rn - 1, so core matrix (rn, c) is chunk
//   32 rn + c: SBO = 32 cores, LBO = 1 core, and the overlapping windows share their samples
//   (6 bytes of shared memory per input sample instead of 15).
// A CTA (one per SM, persistent) loops over tiles of 8 lanes: eight producer warps load float32
// samples, split them into the three bf16 terms and store the chunks; one thread issues the
// 41 K-slabs x 6 products = 246 tcgen05.mma (M 128, N 8 R, K 16); four epilogue warps read the two
// accumulators (tcgen05.ld), add them and store float32 outputs, overlapped with the next tile
// through a second pair of accumulators.
//
// Measured on a B200 (scripts/microbench/decimate_mma_test, 1 440 clips of 144 000 samples, all seven
// stages): 1.46 ms against 2.57 ms for the FFMA2 kernel.  With the producers' and the epilogue's
// work compiled out (MMAs only) the same launches take 1.26 ms, so the kernel runs at 86 % of what
// its MMA sequence allows: 246 MMAs take ~14 000 clocks per tile (57 per MMA where the N = 96 rate
// would be 48), paced by the operand fetch -- a slab's six MMAs read three A and three B tiles,
// 21 KB, at ~62 bytes per clock (an order that does not keep a slab's products together is twice as
// slow) -- and the SM clock settles near 1.4-1.5 GHz under this load.  N = 128 would balance fetch
// and MMA rate but needs 216 KB for the samples alone.
#include <cuda_bf16.h>

#include <cmath>
#include <cstring>

#include "kernels.h"

namespace serb {

namespace {

constexpr int kDmR = 14;                               // windows per lane
constexpr int kDmN = 8 * kDmR;                         // MMA N
constexpr int kDmWin = 128;                            // outputs per window = MMA M
constexpr int kDmSlabs = 41;                           // K = 16 x 41 = 656 >= 643 + left padding
constexpr int kDmHalf = (kDecTaps2 - 1) / 2;           // 194
constexpr int kDmLeft = 200;                           // window sample 0 = x[2 m0 - kDmLeft] (multiple of 8)
constexpr int kDmKoff = 2 * (kDmWin - 1) + kDmHalf + kDmLeft;   // tap index = kDmKoff - 8 u - 2 a - b
constexpr int kDmCoresA = 2 * (kDmWin / 8 - 1) + 2 * kDmSlabs;  // 112 distinct core matrices of T
constexpr int kDmChunks = 32 * (kDmR - 1) + 2 * kDmSlabs;       // 16-byte chunks per lane
constexpr int kDmABytes = kDmCoresA * 128;
// Core matrix u of T holds taps kDmKoff - 8 u - [0, 21]: all zero before kDmAFirst and after kDmALast,
// so the three bf16 terms' tables overlap on their zero cores (term t starts kDmAStride bytes after
// term t - 1): 35 KB instead of 42, which is what lets N = 112 windows fit the shared memory.
constexpr int kDmAFirst = (kDmKoff - 21 - (kDecTaps2 - 1) + 7) / 8;     // first core with a non-zero tap
constexpr int kDmALast = kDmKoff / 8;                                   // last one
constexpr int kDmAOverlap = (kDmAFirst < kDmCoresA - 1 - kDmALast) ? kDmAFirst : kDmCoresA - 1 - kDmALast;
constexpr int kDmAStride = (kDmCoresA - kDmAOverlap) * 128;
constexpr int kDmATotal = 2 * kDmAStride + kDmABytes;
static_assert(kDmAFirst > 0 && kDmALast < kDmCoresA - 1 && kDmAOverlap > 0, "zero cores at both ends of the Toeplitz table");
constexpr int kDmBBytes = kDmChunks * 128;
constexpr int kDmSegOut = kDmWin * kDmR;               // outputs per lane segment
// K slabs [kDmCentral0, kDmCentral1) hold the centre tap of some row: row i (output i) has it at
// window sample 2 i + kDmLeft, i = 0 .. 127
constexpr int kDmCentral0 = kDmLeft / 16;
constexpr int kDmCentral1 = (2 * (kDmWin - 1) + kDmLeft) / 16 + 1;
static_assert(kDmSlabs - kDmCentral1 <= kDmCentral0, "the outer-slab loop pairs slab t with slab kDmSlabs - 1 - t");
static_assert(kDmLeft % 8 == 0 && kDmLeft >= kDmHalf, "left padding");
static_assert(kDmKoff < 16 * kDmSlabs, "the last tap must fall inside the last K slab");
static_assert(kDmHalf + kDmLeft - (kDecTaps2 - 1) >= 0, "the first tap must fall inside the first K slab");
static_assert(kDmN % 16 == 0 && kDmN <= 256, "tcgen05.mma M = 128 needs N % 16 == 0");
static_assert(4 * kDmN <= 512, "two pairs of accumulators must fit the 512 TMEM columns");

constexpr int kDmProducerWarps = 15;
constexpr int kDmThreads = 32 * (1 + 4 + kDmProducerWarps);
constexpr int kDmProducers = 32 * kDmProducerWarps;
constexpr size_t kDmSmem = kDmATotal + 3 * kDmBBytes + 128;
static_assert(kDmSmem + 1024 <= 227 * 1024, "decimate2_mma_kernel shared memory");

#ifdef DM_TRACE
// per-tile clock64 stamps of CTA 0 (scripts/microbench/decimate_mma_test -DDM_TRACE): [tile][8]
__device__ long long dm_trace[4096 * 16];
#define DM_STAMP(n, slot) do { if (blockIdx.x == 0 && (n) < 4096 && (threadIdx.x & 31) == 0) dm_trace[(n) * 16 + (slot)] = clock64(); } while (0)
#define DM_STAMP1(n, slot) do { if (blockIdx.x == 0 && (n) < 4096) dm_trace[(n) * 16 + (slot)] = clock64(); } while (0)
#else
#define DM_STAMP(n, slot) do { } while (0)
#define DM_STAMP1(n, slot) do { } while (0)
#endif

struct DmLaneInfo {
    const float* src;
    float* dst;
    int len_in, len_out, m0, active;
};

__device__ __forceinline__ int dm_level_length(int len0, int level) {
    int n = len0;
    for (int i = 0; i < level; ++i) n = (n + 1) >> 1;
    return n;
}

// lane a of tile `tile` = segment (8 tile + a) of the launch: clip (segment / segs_per_clip),
// outputs [m0, m0 + kDmSegOut) of that clip's destination level
__device__ __forceinline__ DmLaneInfo dm_lane_info(const CqtParams& p, int src_level, int segs_per_clip, int tile, int a) {
    DmLaneInfo li{};
    const int g = 8 * tile + a;
    const int c = g / segs_per_clip;
    if (c >= p.n_clips) return li;
    const TonClip clip = p.clips[c];
    if (clip.length < kDecExactBelow) return li;         // float64 path (decimate2_kernel)
    li.len_in = src_level < 0 ? clip.length : dm_level_length(clip.len0, src_level);
    li.len_out = (li.len_in + 1) >> 1;
    li.m0 = (g - c * segs_per_clip) * kDmSegOut;
    if (li.m0 >= li.len_out) return li;
    li.src = (src_level < 0 || (src_level == 0 && p.early_factor == 1)) ? p.yharm + clip.hoff
                                                                        : p.yoct + p.level_base[src_level] + (clip.off0 >> src_level);
    li.dst = p.yoct + p.level_base[src_level + 1] + (clip.off0 >> (src_level + 1));
    li.active = 1;
    return li;
}

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return static_cast<uint32_t>(__cvta_generic_to_shared(ptr)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// A phase that never completes is a protocol bug: trap instead of hanging the device.  The
// waiting roles that are not on the critical path back off between polls (kSleepNs), so that
// their polling does not take issue slots from the warp that feeds the tensor core.
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if (kSleepNs > 0) __nanosleep(kSleepNs);
        if ((spins & 4095u) == 4095u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 8000000000ll) __trap();          // seconds, not microseconds
        }
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], bf16 x bf16 -> f32
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 consecutive TMEM columns of this thread's lane (32 lanes x 32 bit per warp)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: start >> 4 in
// bits [0,14), leading (K) byte offset >> 4 in [16,30), stride (M/N) byte offset >> 4 in [32,46),
// version 1 in [46,48), layout type 0 in [61,64))
__device__ __forceinline__ uint64_t dm_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
           (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D f32, A and B bf16, both K-major
constexpr uint32_t kDmIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kDmN >> 3) << 17) |
                              (static_cast<uint32_t>(kDmWin >> 4) << 24);

// eight float32 samples -> the three bf16 terms, 16 bytes each (element b at byte 2 b)
__device__ __forceinline__ void dm_split8(const float4& lo, const float4& hi, uint4& s0, uint4& s1, uint4& s2) {
    const float x[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t w0[4], w1[4], w2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        const float2 f0 = __bfloat1622float2(h0);
        const float ra = x[2 * i] - f0.x, rb = x[2 * i + 1] - f0.y;          // exact
        const __nv_bfloat162 h1 = __floats2bfloat162_rn(ra, rb);
        const float2 f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(ra - f1.x, rb - f1.y);
        w0[i] = *reinterpret_cast<const uint32_t*>(&h0);
        w1[i] = *reinterpret_cast<const uint32_t*>(&h1);
        w2[i] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    s0 = make_uint4(w0[0], w0[1], w0[2], w0[3]);
    s1 = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    s2 = make_uint4(w2[0], w2[1], w2[2], w2[3]);
}

}  // namespace

// Chunk q of a lane is read by the K slabs s with (2 s) mod 32 == q mod 32 rounded down to even, i.e.
// by slabs s, s + 16, s + 32 only.  The 41 slabs therefore fall into kDmGroups groups that read
// disjoint quarters of the operand buffer (group of slab s = (s mod 16) / 4, group of chunk q =
// (q mod 32) / 8): the single buffer works as a four-stage ring -- while the tensor core runs the
// slabs of group g + 1 of tile n, the producers already write group g of tile n + 1.
constexpr int kDmGroups = 4;
constexpr int kDmBlk = (kDmChunks + 31) / 32;                       // 32-chunk blocks per lane
constexpr int kDmGroupTasks = 8 * 8 * kDmBlk;                       // (lane, chunk) tasks per group, tail included
constexpr int kDmGroupIters = (kDmGroupTasks + kDmProducers - 1) / kDmProducers;
__host__ __device__ constexpr int dm_slab_group(int s) { return (s % 16) / 4; }
// issue order of group 0 is slabs 0..3, 16..19, 32..35: slab 0 opens d_small, slab 16 must be the first central one
static_assert(kDmCentral0 > 3 && kDmCentral0 <= 16 && kDmCentral1 > 16, "first MMA into d_main");

__global__ void __launch_bounds__(kDmThreads, 1)
decimate2_mma_kernel(CqtParams p, int src_level, int segs_per_clip, int n_tiles, const uint4* __restrict__ toeplitz) {
    extern __shared__ __align__(128) unsigned char dm_smem[];
    unsigned char* sm_a = dm_smem;                               // three overlapped tables, kDmAStride apart
    unsigned char* sm_b = dm_smem + kDmATotal;                   // [3][kDmBBytes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(dm_smem + kDmATotal + 3 * kDmBBytes);
    // bars[0..3] group full, [4..7] group empty, [8..9] accumulators full, [10..11] accumulators
    // empty; then the TMEM base address
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    __shared__ DmLaneInfo epi_info[4][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < kDmATotal / 16; i += kDmThreads) reinterpret_cast<uint4*>(sm_a)[i] = toeplitz[i];
    if (tid == 0) {
        for (int g = 0; g < kDmGroups; ++g) {
            mbar_init(smem_u32(&bars[g]), kDmProducerWarps);
            mbar_init(smem_u32(&bars[4 + g]), 1);
        }
        mbar_init(smem_u32(&bars[8]), 1);
        mbar_init(smem_u32(&bars[9]), 1);
        mbar_init(smem_u32(&bars[10]), 4);
        mbar_init(smem_u32(&bars[11]), 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the Toeplitz operand, written with generic stores
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- MMA issue: the whole warp walks the tiles, one elected lane issues ----
        const uint32_t a_addr = smem_u32(sm_a), b_addr = smem_u32(sm_b);
        // descriptors of K slab 0; slab s is 256 bytes (two core matrices) further in both operands
        uint64_t ad[3], bd[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            ad[t] = dm_desc(a_addr + t * kDmAStride, 128, 256);
            bd[t] = dm_desc(b_addr + t * kDmBBytes, 128, 32 * 128);
        }
        // one elected lane does all the waiting and issuing; the other lanes park at the warp barrier
        if (elect_one()) {
            int n = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
                const int buf = n & 1, use = n >> 1;
                DM_STAMP1(n, 12);
                mbar_wait<0>(smem_u32(&bars[10 + buf]), (use & 1) ^ 1);      // accumulators drained by the epilogue
                const uint32_t d_main = tmem_base + buf * 2 * kDmN, d_small = d_main + kDmN;
                // The tensor core's float32 accumulate truncates, so every add at full magnitude costs
                // up to an ulp, always towards zero.  The leading product x0 h0 of the CENTRAL slabs --
                // where every row's main lobe lies -- goes to d_main; its outer slabs (side lobes: a few
                // per cent of the result) and the five small products of all slabs go to d_small.
                // Within a slab the six MMAs stay together: they share three A and three B tiles, and
                // the operand fetch, not the MMA rate, is what paces the pipe at N = 96.
#pragma unroll 1
                for (int g = 0; g < kDmGroups; ++g) {
                    DM_STAMP1(n, 3 * g);
                    mbar_wait<0>(smem_u32(&bars[g]), n & 1);                 // group g of this tile staged
                    DM_STAMP1(n, 3 * g + 1);
                    tc_fence_after();
                    // rolled on purpose: 246 unrolled MMAs with their descriptor arithmetic are 12 KB of
                    // straight-line code that the issuing warp's instruction cache loses to the producer
                    // and epilogue warps between tiles
#pragma unroll 1
                    for (int o = 0; o < 48; o += 16) {
#pragma unroll 1
                        for (int j = 0; j < 4; ++j) {
                            const int s = 4 * g + j + o;
                            if (s >= kDmSlabs) break;
                            const uint64_t off = 16ull * s;                  // (256 s) >> 4 in the start-address field
                            const bool central = s >= kDmCentral0 && s < kDmCentral1;
                            // the tile's first MMA into d_small is slab 0's, into d_main slab 16's (both group 0)
                            tc_mma(d_small, ad[1] + off, bd[1] + off, kDmIdesc, s != 0);
                            tc_mma(d_small, ad[2] + off, bd[0] + off, kDmIdesc, 1);
                            tc_mma(central ? d_main : d_small, ad[0] + off, bd[0] + off, kDmIdesc, s != 16);
                            tc_mma(d_small, ad[0] + off, bd[2] + off, kDmIdesc, 1);
                            tc_mma(d_small, ad[1] + off, bd[0] + off, kDmIdesc, 1);
                            tc_mma(d_small, ad[0] + off, bd[1] + off, kDmIdesc, 1);
                        }
                    }
                    tc_commit(smem_u32(&bars[4 + g]));                    // group g may be overwritten
                    if (g == kDmGroups - 1) tc_commit(smem_u32(&bars[8 + buf]));     // the accumulators are complete
                    DM_STAMP1(n, 3 * g + 2);
                }
            }
        }
        __syncwarp();
    } else if (warp <= 4) {
        // ---- epilogue: TMEM lane quarter (warp & 3), row i' = output 127 - i' of every window ----
        const int quarter = warp & 3;
        DmLaneInfo* info = epi_info[warp - 1];
        const int mrow = kDmWin - 1 - (32 * quarter + lane);
        int n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
            const int buf = n & 1, use = n >> 1;
            if (lane < 8) info[lane] = dm_lane_info(p, src_level, segs_per_clip, tile, lane);
            __syncwarp();
            // per lane of the tile: where this thread's row lands and how many outputs are left there
            float* dptr[8];
            int left[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                const DmLaneInfo li = info[a];
                dptr[a] = li.dst + li.m0 + mrow;
                left[a] = li.active ? li.len_out - li.m0 - mrow : 0;
            }
            mbar_wait<64>(smem_u32(&bars[8 + buf]), use & 1);
            if (warp == 1) DM_STAMP(n, 14);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * quarter) << 16) + buf * 2 * kDmN;
#pragma unroll 2
            for (int g = 0; g < kDmN / 16; ++g) {
                float v[16], w[16];
                tc_ld16(taddr + 16 * g, v);
                tc_ld16(taddr + kDmN + 16 * g, w);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int off = kDmWin * (2 * g + (j >> 3));          // window rn = 2 g + j / 8 of lane j % 8
                    if (off < left[j & 7]) dptr[j & 7][off] = v[j] + w[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[10 + buf]));
            if (warp == 1) DM_STAMP(n, 15);
        }
    } else {
        // ---- producers ----
        // Task t of group g = (lane a = t % 8, chunk q = 32 (t / 64) + 8 g + (t / 8) % 8); a thread owns
        // tasks pt + kDmProducers k.  A group is refilled (load float32, split, store) as soon as the
        // tensor core releases it and is needed again a whole tile later, so the loads' latency is
        // off the critical path and nothing is prefetched into registers.
        const int pt = tid - 160, a = pt & 7;
        unsigned char* dst0 = sm_b + 16 * a;
        int n = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
            const DmLaneInfo li = dm_lane_info(p, src_level, segs_per_clip, tile, a);
            const int sbase = 2 * li.m0 - kDmLeft;
            const bool aligned = (reinterpret_cast<uintptr_t>(li.src) & 31) == 0;
#pragma unroll 1
            for (int g = 0; g < kDmGroups; ++g) {
                mbar_wait<32>(smem_u32(&bars[4 + g]), (n & 1) ^ 1);
                if (li.active) {
                    float4 lo[kDmGroupIters], hi[kDmGroupIters];
#pragma unroll
                    for (int k = 0; k < kDmGroupIters; ++k) {
                        const int t = pt + kDmProducers * k;
                        const int q = 32 * (t >> 6) + 8 * g + ((t >> 3) & 7);
                        const int s = sbase + 8 * q;
                        if (t >= kDmGroupTasks || q >= kDmChunks) continue;
                        if (aligned && s >= 0 && s + 8 <= li.len_in) {
                            lo[k] = __ldg(reinterpret_cast<const float4*>(li.src + s));
                            hi[k] = __ldg(reinterpret_cast<const float4*>(li.src + s + 4));
                        } else {
                            float x[8];
#pragma unroll
                            for (int b = 0; b < 8; ++b) x[b] = (s + b >= 0 && s + b < li.len_in) ? li.src[s + b] : 0.0f;
                            lo[k] = make_float4(x[0], x[1], x[2], x[3]);
                            hi[k] = make_float4(x[4], x[5], x[6], x[7]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kDmGroupIters; ++k) {
                        const int t = pt + kDmProducers * k;
                        const int q = 32 * (t >> 6) + 8 * g + ((t >> 3) & 7);
                        if (t >= kDmGroupTasks || q >= kDmChunks) continue;
                        uint4 s0, s1, s2;
                        dm_split8(lo[k], hi[k], s0, s1, s2);
                        unsigned char* d = dst0 + 128 * q;
                        *reinterpret_cast<uint4*>(d) = s0;
                        *reinterpret_cast<uint4*>(d + kDmBBytes) = s1;
                        *reinterpret_cast<uint4*>(d + 2 * kDmBBytes) = s2;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars[g]));
                if (warp == 5 && g == 0) DM_STAMP(n, 13);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host side ---------------------------------------------------------------------------
namespace {

uint16_t bf16_rn(double v, double* back) {
    const float f = static_cast<float>(v);
    uint32_t u;
    std::memcpy(&u, &f, 4);
    const uint32_t rounded = u + 0x7fffu + ((u >> 16) & 1u);     // round to nearest even on the top 16 bits
    const uint16_t h = static_cast<uint16_t>(rounded >> 16);
    const uint32_t w = static_cast<uint32_t>(h) << 16;
    float g;
    std::memcpy(&g, &w, 4);
    *back = static_cast<double>(g);
    return h;
}

}  // namespace

#ifdef DM_TRACE
void* decimate_mma_trace_ptr() {
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, dm_trace);
    return sym;
}
#endif

size_t decimate_mma_table_bytes() { return kDmATotal; }

// taps (x sqrt 2, float64) -> three overlapped tables of [112 core matrices][8 rows][8 columns] bf16
void decimate_mma_table(const double* taps2_scaled, unsigned char* out) {
    std::memset(out, 0, kDmATotal);
    uint16_t* t = reinterpret_cast<uint16_t*>(out);
    for (int u = 0; u < kDmCoresA; ++u)
        for (int a = 0; a < 8; ++a)
            for (int b = 0; b < 8; ++b) {
                const int k = kDmKoff - 8 * u - 2 * a - b;
                const double h = (k >= 0 && k < kDecTaps2) ? taps2_scaled[k] : 0.0;
                double b0, b1, b2;
                const uint16_t h0 = bf16_rn(h, &b0);
                const uint16_t h1 = bf16_rn(h - b0, &b1);
                const uint16_t h2 = bf16_rn(h - b0 - b1, &b2);
                const size_t at = (static_cast<size_t>(u) * 8 + a) * 8 + b;
                // overlapped tables: a core outside [kDmAFirst, kDmALast] is zero in every term, and only
                // such cores share storage
                if (u >= kDmAFirst && u <= kDmALast) {
                    t[at] = h0;
                    t[kDmAStride / 2 + at] = h1;
                    t[2 * (kDmAStride / 2) + at] = h2;
                }
            }
}

cudaError_t configure_decimate_mma() {
    return cudaFuncSetAttribute(decimate2_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDmSmem));
}

// level src_level -> src_level + 1 (src_level -1: yharm -> level 0) for every clip of at least
// kDecExactBelow samples; max_len_in = longest source signal of the launch
cudaError_t launch_decimate2_mma(const CqtParams& p, int src_level, int max_len_in, const void* d_toeplitz, int n_sms,
                                 cudaStream_t stream) {
    const int max_out = (max_len_in + 1) >> 1;
    const int segs_per_clip = (max_out + kDmSegOut - 1) / kDmSegOut;
    const long long segs = static_cast<long long>(p.n_clips) * segs_per_clip;
    if (segs <= 0) return cudaSuccess;
    if (segs > 0x3fffffffll) return cudaErrorInvalidValue;
    const int n_tiles = static_cast<int>((segs + 7) / 8);
    const int grid = n_tiles < n_sms ? n_tiles : n_sms;
    decimate2_mma_kernel<<<grid, kDmThreads, kDmSmem, stream>>>(p, src_level, segs_per_clip, n_tiles,
                                                                static_cast<const uint4*>(d_toeplitz));
    return cudaGetLastError();
}

}  // namespace serb
