// tcgen05 (5th-generation tensor core) evaluation of the radiance/visibility MLP, and the fused
// coarse+fine render of a ray batch in ONE launch.  sm_100a only.
//
// Reference being replaced: VipNeRF.render_rays (src/models/VipNeRF01.py:74-171) = get_z_vals_coarse :173,
// PositionalEncoder :416, MLP.forward :509-596, volume_rendering :331, get_z_vals_fine/sample_pdf :205-262.
//
// One persistent CTA per SM (clusters of two CTAs working on one tcgen05.mma.cta_group::2), 384 threads:
//   warps 0-3  epilogue group 0  (owns tile slot 0: TMEM columns [0,256),   activation buffer 0)
//   warps 4-7  epilogue group 1  (owns tile slot 1: TMEM columns [256,512), activation buffer 1)
//   warp  8    weight producer   (one lane: tensor-map TMA copies of this CTA's half of the weight chunk images into
//                                 the even stages of the shared-memory ring; all stages in the single-CTA variant)
//   warp  9    MMA issuer        (leader CTA; one lane: tcgen05.mma M=256 N=256 K=16, bf16 x bf16 -> fp32 in TMEM)
//   warp  10   second weight producer (CTA pairs: the odd ring stages)
//   warp  11   ray warp          (fused kernel: alpha compositing + hierarchical re-sampling of the rays the slots'
//                                 tiles complete, asynchronously to the slots' next tiles)
// A "tile" is 128 consecutive sample points (rows).  A row's activations live in shared memory as bf16 in the
// canonical K-major SWIZZLE_128B layout (four 16 KiB k-blocks of 128 rows x 64 columns; the tile's point encoding
// sits in k-block 0 while M0 and the encoding part of the skip layer run); the accumulator of a layer lives in the
// slot's 256 TMEM columns.  Per step (step_desc): the MMA warp streams the step's weight chunks against the slot's
// activation buffer; when the last MMA retires (tcgen05.commit -> d_ready) the slot's epilogue group pulls the
// accumulator with tcgen05.ld (the bias was added by the tensor core), applies ReLU, rounds to bf16 and overwrites
// the buffer in place (the MMAs that read it have completed), then signals a_ready.  Two slots ping-pong so the
// tensor pipe works on one tile while the other tile's epilogue runs on the CUDA cores.
// The sample points, their sinusoidal encodings and the density / colour / visibility heads are done by the epilogue
// groups, alpha compositing and hierarchical re-sampling by the ray warp, so no per-sample intermediate other than
// the 5 (+V) network outputs per sample leaves the SM.
//
// BF16X3 mode (parity mode): every operand is split x = hi + lo (both bf16) and each product is evaluated as
// hi*hi + lo*hi + hi*lo in the fp32 accumulator (3 MMAs).  Slot 1's buffers hold the lo parts, so only one
// tile is in flight.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "kernels.h"
#include "layout.cuh"

namespace vipnerf {
namespace {

constexpr int kTile = 128;
constexpr int kNumThreads = 384;   // 12 warps: a 10-warp CTA has the same per-thread register budget (3 warps per sub-partition)
constexpr uint32_t kABytes = 65536;
constexpr uint32_t kKBlockBytes = 16384;
// Shared memory: two activation buffers (the point encoding of a tile lives in k-block 0 of its slot's buffer while
// M0 / the encoding part of M5 run), then the weight ring, then barriers and small per-ray tables.
// VIPNERF_ONES_4K (build-time fallback): a conventional 4 KiB all-ones operand instead of the 128-byte block read
// through a zero-stride descriptor; costs one ring stage.
#ifdef VIPNERF_ONES_4K
#define VIPNERF_PAIR_STAGES 10
#define VIPNERF_SINGLE_STAGES 5
#else
#define VIPNERF_PAIR_STAGES 11
#define VIPNERF_SINGLE_STAGES 5
#endif
#define VIPNERF_STR2(x) #x
#define VIPNERF_STR(x) VIPNERF_STR2(x)
constexpr int kPairStages = VIPNERF_PAIR_STAGES;      // x 8 KiB (each CTA of a pair holds half of a chunk's rows)
constexpr int kSingleStages = VIPNERF_SINGLE_STAGES;  // x 16 KiB
constexpr uint32_t kOffA = 0;                 // 2 x 64 KiB
constexpr uint32_t kOffW = 131072;            // weight ring (88 KiB used) + the ray warp's re-sampling scratch
constexpr uint32_t kRingBytes = 98304;
constexpr uint32_t kOffRayScratch = kOffW + 90112;   // 2 KiB (resample_scratch_floats(64, 128) = 384 floats)
// Small fp32 head parameters of BOTH passes' MLPs, staged once per CTA (they were __ldg loads from a ~28 KiB L1 that
// the per-tile global traffic keeps evicting: every miss is an L2 round trip inside the view epilogue's dependency
// chain): per pass views_output_linear.weight^T [128][4] (2 KiB) and pts_output_linear.weight [256] (1 KiB).
constexpr uint32_t kOffHead = kOffW + 92160;         // 2 x 3 KiB
constexpr uint32_t kHeadFloats = 768;                // per pass: [0,512) W_out^T, [512,768) w_sigma
constexpr uint32_t kOffTail = kOffW + kRingBytes;
constexpr uint32_t kOffBar = kOffTail;        // up to 32 mbarriers
constexpr uint32_t kOffRayDone = kOffTail + 256;  // [2 slots] u32: per-ray events the slot's ray warp has completed
constexpr uint32_t kOffTmemPtr = kOffTail + 264;
constexpr uint32_t kOffVb = kOffTail + 272;   // [2 slots][2 rays][128] fp32: view-direction part of M9 + bias
constexpr uint32_t kOffPev = kOffVb + 2048;   // [2 slots][2 rays][32]  fp32: view-direction encodings
#ifdef VIPNERF_ONES_4K
#error "VIPNERF_ONES_4K is no longer supported (its 4 KiB now hold the staged head parameters)"
#else
constexpr uint32_t kOffOnes = kOffPev + 512;  // 128 B of bf16 1.0: the A operand of the bias chunks
constexpr uint32_t kOnesBytes = 128;
constexpr uint32_t kOffHeadBias = kOffOnes + 128;   // [2 passes][8] fp32: views_output_linear.bias (4), pts_output_linear.bias
constexpr uint32_t kSmemBytes = kOffHeadBias + 64;
#endif
static_assert(kSmemBytes <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
static_assert(kPairStages * 8192 <= 90112 && kSingleStages * 16384 <= 90112, "weight ring does not fit");
static_assert(kOffHead + 2 * kHeadFloats * 4 <= kOffTail, "head parameters do not fit");

enum { kBarWFull = 0, kBarWEmpty = 12, kBarAReady = 24, kBarDReady = 26, kBarRayFull = 28 };

// tcgen05 instruction descriptor: D=F32, A=B=BF16, both K-major, M=128, N=n (cute::UMMA::InstrDescriptor bits:
// c_format[4,6)=1, a_format[7,10)=1, b_format[10,13)=1, n_dim[17,23)=N>>3, m_dim[24,29)=M>>4)
// `half`: A = B = F16 (format 0) instead of BF16 (format 1) - the same kind::f16 instruction at the same rate with an
// 11-bit instead of an 8-bit significand (VIPNERF_PRECISION_FP16).
constexpr uint32_t instr_desc(uint32_t n, uint32_t m = 128, bool half = false) {
  return (1u << 4) | ((half ? 0u : 1u) << 7) | ((half ? 0u : 1u) << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

static_assert(resample_scratch_floats(64, 128) * 4 <= 2048, "ray scratch");
constexpr long long kTimeoutCycles = 4000000000ll;  // ~2 s: a protocol bug traps instead of hanging the GPU

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("vipnerf: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x,
         (bar / 8) % 32, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kTimeoutCycles) mbar_timeout(bar, parity);
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int group) {  // named barrier over one epilogue group (128 threads)
  asm volatile("bar.sync %0, 128;" ::"r"(group + 1) : "memory");
}
// ---- thread-block-cluster (CTA pair) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {  // shared::cta -> shared::cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// Arrive on an mbarrier of (possibly) another CTA of the cluster.  Default semantics (release at CTA scope), as
// CUTLASS's ClusterBarrier::arrive does: a cluster-scope release costs ~1000 cycles per arrival (it fences and
// invalidates L1), which throttled the CTA-pair weight relay to one chunk per microsecond.  CTA scope suffices
// here: what the signal publishes (weight chunks, activation tiles) is consumed by the tensor core of the
// signalling CTA itself through the async proxy (ordered by fence.proxy.async / the TMA's complete_tx), never by
// threads of the other CTA.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Position in the weight ring (stage index and the phase parity of its barriers) plus a cycle counter the
// profiling builds report; every role keeps its own copy and advances it chunk by chunk.
struct RingState {
  uint32_t stage = 0, phase = 0, wait_cycles = 0;
  uint32_t ready = 0;   // MMA issuer (probe-ahead loops): the current stage's full barrier was already seen complete
};

// The MMA issue loop of one run of `n_chunks` consecutive weight chunks, hand-written in PTX so that the per-chunk
// cost is ~25 instructions (ptxas turns the equivalent C++ into ~135 with reconvergence barriers and R2UR moves).
// Chunk c multiplies A columns [32c, 32c+32) - k-block c>>1 (1024 descriptor units apart), 64-byte half c&1
// (4 units) - with the current ring stage: waits the stage's full barrier, issues two N x K=16 MMAs from the
// elected lane, commits the stage's empty barrier and advances the ring.
//   single CTA :  5 stages x 16 KiB (1024 units), tcgen05.mma.cta_group::1, M=128
//   CTA pair   : 11 stages x  8 KiB ( 512 units: each CTA holds half of the chunk's rows), cta_group::2, M=256,
//                commits multicast to both CTAs' barriers
// The *_split variants are BF16X3: every weight chunk is two ring stages (hi image, lo image); per chunk
// A_hi*W_hi + A_lo*W_hi (4 MMAs, commit) then A_hi*W_lo (2 MMAs, commit).
__device__ __forceinline__ void issue_chunks(RingState& rs, uint32_t d_tmem, uint64_t a_desc, uint64_t w_desc0,
        uint32_t bar_full0, uint32_t bar_empty0, uint32_t n_chunks, uint32_t first_acc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, pw, e, pacc, pt;\n"
      ".reg .b32 c, fb, eb, t, spins, c0, c1;\n"
      ".reg .b64 a, b, a1, b1, t64;\n"
      ".reg .b16 mc;\n"
      "mov.b16 mc, 3;\n"
      "mov.u32 c, 0;\n"
      "setp.ne.b32 pacc, %9, 0;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "CHUNK_LOOP:\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %6, t;\n"
      "add.u32 eb, %7, t;\n"
      "mov.u32 spins, 0;\n"
      "mov.u32 c0, %clock;\n"
      "CHUNK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [fb], %1;\n"
      "@pw bra CHUNK_READY;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra CHUNK_WAIT;\n"
      "CHUNK_READY:\n"
      "mov.u32 c1, %clock;\n"
      "sub.u32 c1, c1, c0;\n"
      "add.u32 %2, %2, c1;\n"
      "tcgen05.fence::after_thread_sync;\n"
      "mul.wide.u32 b, %0, 1024;\n"
      "add.s64 b, b, %5;\n"
      "shr.u32 t, c, 1;\n"
      "mul.wide.u32 a, t, 1024;\n"
      "and.b32 t, c, 1;\n"
      "mul.wide.u32 t64, t, 4;\n"
      "add.s64 a, a, t64;\n"
      "add.s64 a, a, %4;\n"
      "add.s64 a1, a, 2;\n"
      "add.s64 b1, b, 2;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%3], a, b, %10, pacc;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%3], a1, b1, %10, pt;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [eb];\n"
      "setp.eq.b32 pacc, 0, 0;\n"
      "add.u32 %0, %0, 1;\n"
      "setp.eq.u32 p, %0, " VIPNERF_STR(VIPNERF_SINGLE_STAGES) ";\n"
      "@p mov.u32 %0, 0;\n"
      "@p xor.b32 %1, %1, 1;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 p, c, %8;\n"
      "@p bra CHUNK_LOOP;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase), "+r"(rs.wait_cycles)
      : "r"(d_tmem), "l"(a_desc), "l"(w_desc0), "r"(bar_full0), "r"(bar_empty0), "r"(n_chunks), "r"(first_acc),
        "r"(idesc)
      : "memory");
}
// Probe-ahead: an mbarrier probe has ~150-200 cycles of latency even when the phase is already complete, and the
// loop is one dependent chain (probe -> branch -> MMAs -> commit -> next probe), which alone costs more than the 256
// tensor cycles of a chunk.  So the probe of the NEXT stage (non-blocking test_wait) is issued before the current
// chunk's MMAs and its result consumed one iteration later (carried across calls in rs.ready - the ring position is
// continuous across layers and slots); only a stage found incomplete falls back to the blocking try_wait loop.
// rs.wait_cycles counts only that real waiting.
__device__ __forceinline__ void issue_chunks_pair(RingState& rs, uint32_t d_tmem, uint64_t a_desc, uint64_t w_desc0,
        uint32_t bar_full0, uint32_t bar_empty0, uint32_t n_chunks, uint32_t first_acc, uint32_t idesc,
        uint32_t skip_wait = 0) {
  asm volatile(
      "{\n"
      ".reg .pred p, pw, e, pacc, pt, pskip;\n"
      ".reg .b32 c, fb, eb, t, spins, c0, c1, ns, nph, nfb;\n"
      ".reg .b64 a, b, a1, b1, t64;\n"
      ".reg .b16 mc;\n"
      "mov.b16 mc, 3;\n"
      "mov.u32 c, 0;\n"
      "setp.ne.b32 pacc, %10, 0;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "setp.ne.b32 pskip, %12, 0;\n"
      "setp.ne.b32 pw, %3, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "CHUNK_LOOP:\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %7, t;\n"
      "add.u32 eb, %8, t;\n"
      "@pskip bra CHUNK_READY;\n"
      "@pw bra CHUNK_READY;\n"
      "mov.u32 spins, 0;\n"
      "mov.u32 c0, %clock;\n"
      "CHUNK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [fb], %1;\n"
      "@pw bra CHUNK_WAITED;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra CHUNK_WAIT;\n"
      "CHUNK_WAITED:\n"
      "mov.u32 c1, %clock;\n"
      "sub.u32 c1, c1, c0;\n"
      "add.u32 %2, %2, c1;\n"
      "CHUNK_READY:\n"
      "tcgen05.fence::after_thread_sync;\n"
      "mul.wide.u32 b, %0, 512;\n"
      "add.s64 b, b, %6;\n"
      "shr.u32 t, c, 1;\n"
      "mul.wide.u32 a, t, 1024;\n"
      "and.b32 t, c, 1;\n"
      "mul.wide.u32 t64, t, 4;\n"
      "add.s64 a, a, t64;\n"
      "add.s64 a, a, %5;\n"
      "add.s64 a1, a, 2;\n"
      "add.s64 b1, b, 2;\n"
      "add.u32 ns, %0, 1;\n"
      "mov.u32 nph, %1;\n"
      "setp.eq.u32 p, ns, " VIPNERF_STR(VIPNERF_PAIR_STAGES) ";\n"
      "@p mov.u32 ns, 0;\n"
      "@p xor.b32 nph, nph, 1;\n"
      "shl.b32 t, ns, 3;\n"
      "add.u32 nfb, %7, t;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 pw, [nfb], nph;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%4], a, b, %11, pacc;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%4], a1, b1, %11, pt;\n"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [eb], mc;\n"
      "setp.eq.b32 pacc, 0, 0;\n"
      "mov.u32 %0, ns;\n"
      "mov.u32 %1, nph;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 p, c, %9;\n"
      "@p bra CHUNK_LOOP;\n"
      "selp.u32 %3, 1, 0, pw;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase), "+r"(rs.wait_cycles), "+r"(rs.ready)
      : "r"(d_tmem), "l"(a_desc), "l"(w_desc0), "r"(bar_full0), "r"(bar_empty0), "r"(n_chunks), "r"(first_acc),
        "r"(idesc), "r"(skip_wait)
      : "memory");
}
__device__ __forceinline__ void issue_chunks_split(RingState& rs, uint32_t d_tmem, uint64_t a_hi_desc, uint64_t a_lo_desc,
        uint64_t w_desc0, uint32_t bar_full0, uint32_t bar_empty0, uint32_t n_chunks, uint32_t first_acc,
        uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, pw, e, pacc, pt;\n"
      ".reg .b32 c, fb, eb, t, spins, part;\n"
      ".reg .b64 a, l, b, a1, l1, b1, t64, off;\n"
      ".reg .b16 mc;\n"
      "mov.b16 mc, 3;\n"
      "mov.u32 c, 0;\n"
      "setp.ne.b32 pacc, %9, 0;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "SCHUNK_LOOP:\n"
      "mov.u32 part, 0;\n"
      "shr.u32 t, c, 1;\n"
      "mul.wide.u32 off, t, 1024;\n"
      "and.b32 t, c, 1;\n"
      "mul.wide.u32 t64, t, 4;\n"
      "add.s64 off, off, t64;\n"
      "add.s64 a, off, %3;\n"
      "add.s64 l, off, %4;\n"
      "add.s64 a1, a, 2;\n"
      "add.s64 l1, l, 2;\n"
      "SPART_LOOP:\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %6, t;\n"
      "add.u32 eb, %7, t;\n"
      "mov.u32 spins, 0;\n"
      "SCHUNK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [fb], %1;\n"
      "@pw bra SCHUNK_READY;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra SCHUNK_WAIT;\n"
      "SCHUNK_READY:\n"
      "tcgen05.fence::after_thread_sync;\n"
      "mul.wide.u32 b, %0, 1024;\n"
      "add.s64 b, b, %5;\n"
      "add.s64 b1, b, 2;\n"
      "setp.eq.u32 p, part, 0;\n"
      "@!p bra SLO_IMAGE;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], a, b, %10, pacc;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], l, b, %10, pt;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1, %10, pt;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], l1, b1, %10, pt;\n"
      "bra SPART_DONE;\n"
      "SLO_IMAGE:\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], a, b, %10, pt;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], a1, b1, %10, pt;\n"
      "SPART_DONE:\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [eb];\n"
      "setp.eq.b32 pacc, 0, 0;\n"
      "add.u32 %0, %0, 1;\n"
      "setp.eq.u32 p, %0, " VIPNERF_STR(VIPNERF_SINGLE_STAGES) ";\n"
      "@p mov.u32 %0, 0;\n"
      "@p xor.b32 %1, %1, 1;\n"
      "add.u32 part, part, 1;\n"
      "setp.lt.u32 p, part, 2;\n"
      "@p bra SPART_LOOP;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 p, c, %8;\n"
      "@p bra SCHUNK_LOOP;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase)
      : "r"(d_tmem), "l"(a_hi_desc), "l"(a_lo_desc), "l"(w_desc0), "r"(bar_full0), "r"(bar_empty0), "r"(n_chunks),
        "r"(first_acc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void issue_chunks_split_pair(RingState& rs, uint32_t d_tmem, uint64_t a_hi_desc, uint64_t a_lo_desc,
        uint64_t w_desc0, uint32_t bar_full0, uint32_t bar_empty0, uint32_t n_chunks, uint32_t first_acc,
        uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, pw, e, pacc, pt;\n"
      ".reg .b32 c, fb, eb, t, spins, part;\n"
      ".reg .b64 a, l, b, a1, l1, b1, t64, off;\n"
      ".reg .b16 mc;\n"
      "mov.b16 mc, 3;\n"
      "mov.u32 c, 0;\n"
      "setp.ne.b32 pacc, %9, 0;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "SCHUNK_LOOP:\n"
      "mov.u32 part, 0;\n"
      "shr.u32 t, c, 1;\n"
      "mul.wide.u32 off, t, 1024;\n"
      "and.b32 t, c, 1;\n"
      "mul.wide.u32 t64, t, 4;\n"
      "add.s64 off, off, t64;\n"
      "add.s64 a, off, %3;\n"
      "add.s64 l, off, %4;\n"
      "add.s64 a1, a, 2;\n"
      "add.s64 l1, l, 2;\n"
      "SPART_LOOP:\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %6, t;\n"
      "add.u32 eb, %7, t;\n"
      "mov.u32 spins, 0;\n"
      "SCHUNK_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [fb], %1;\n"
      "@pw bra SCHUNK_READY;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra SCHUNK_WAIT;\n"
      "SCHUNK_READY:\n"
      "tcgen05.fence::after_thread_sync;\n"
      "mul.wide.u32 b, %0, 512;\n"
      "add.s64 b, b, %5;\n"
      "add.s64 b1, b, 2;\n"
      "setp.eq.u32 p, part, 0;\n"
      "@!p bra SLO_IMAGE;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%2], a, b, %10, pacc;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%2], l, b, %10, pt;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%2], a1, b1, %10, pt;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%2], l1, b1, %10, pt;\n"
      "bra SPART_DONE;\n"
      "SLO_IMAGE:\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%2], a, b, %10, pt;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%2], a1, b1, %10, pt;\n"
      "SPART_DONE:\n"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [eb], mc;\n"
      "setp.eq.b32 pacc, 0, 0;\n"
      "add.u32 %0, %0, 1;\n"
      "setp.eq.u32 p, %0, " VIPNERF_STR(VIPNERF_PAIR_STAGES) ";\n"
      "@p mov.u32 %0, 0;\n"
      "@p xor.b32 %1, %1, 1;\n"
      "add.u32 part, part, 1;\n"
      "setp.lt.u32 p, part, 2;\n"
      "@p bra SPART_LOOP;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 p, c, %8;\n"
      "@p bra SCHUNK_LOOP;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase)
      : "r"(d_tmem), "l"(a_hi_desc), "l"(a_lo_desc), "l"(w_desc0), "r"(bar_full0), "r"(bar_empty0), "r"(n_chunks),
        "r"(first_acc), "r"(idesc)
      : "memory");
}
// Weight producer loop of one run of `n` consecutive chunks (`bytes` each, `stride` bytes apart in the packed
// stream), hand-written in PTX for the same reason as issue_chunks: wait for the ring stage to be free, arm its full
// barrier with the byte count, start the bulk copy (TMA engine), advance.  Executed by ONE lane.  Single-CTA variant.
__device__ __forceinline__ void produce_chunks(RingState& rs, const uint8_t* src, uint32_t bytes, uint32_t stride,
                                               uint32_t n, uint32_t bar_full0, uint32_t bar_empty0, uint32_t w_smem0) {
  asm volatile(
      "{\n"
      ".reg .pred p, pw;\n"
      ".reg .b32 c, fb, eb, t, dst, spins, par, c0, c1;\n"
      ".reg .b64 src, st64;\n"
      "mov.u32 c, 0;\n"
      "mov.u64 src, %3;\n"
      "cvt.u64.u32 st64, %5;\n"
      "PROD_LOOP:\n"
      "xor.b32 par, %1, 1;\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %7, t;\n"
      "add.u32 eb, %8, t;\n"
      "mov.u32 spins, 0;\n"
      "mov.u32 c0, %clock;\n"
      "PROD_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [eb], par;\n"
      "@pw bra PROD_READY;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra PROD_WAIT;\n"
      "PROD_READY:\n"
      "mov.u32 c1, %clock;\n"
      "sub.u32 c1, c1, c0;\n"
      "add.u32 %2, %2, c1;\n"
      "mbarrier.arrive.expect_tx.shared::cta.b64 _, [fb], %4;\n"
      "mad.lo.u32 dst, %0, 16384, %9;\n"
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [dst], [src], %4, [fb];\n"
      "add.u64 src, src, st64;\n"
      "add.u32 %0, %0, 1;\n"
      "setp.eq.u32 p, %0, " VIPNERF_STR(VIPNERF_SINGLE_STAGES) ";\n"
      "@p mov.u32 %0, 0;\n"
      "@p xor.b32 %1, %1, 1;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 p, c, %6;\n"
      "@p bra PROD_LOOP;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase), "+r"(rs.wait_cycles)
      : "l"(src), "r"(bytes), "r"(stride), "r"(n), "r"(bar_full0), "r"(bar_empty0), "r"(w_smem0)
      : "memory");
}
// CTA-pair variant.  The weight stream is addressed through a 2-D tensor map (rows of 64 bytes; the box is this CTA's
// half of a chunk's rows) so that the copy can be a cp.async.bulk.tensor with .cta_group::2, whose complete_tx may
// signal an mbarrier of the PEER CTA: both CTAs' copies complete directly on the LEADER's w_full (`bar_full_cl0` =
// its shared::cluster address), which the leader arms with the byte count of both halves (`expect_bytes`; 0 in the
// peer CTA, which only copies).  No relay hop between the peer's copy and the MMA-issuing lane.
__device__ __forceinline__ void produce_chunks_pair(RingState& rs, const void* tmap, uint32_t row0, uint32_t row_stride,
                                                    uint32_t n, uint32_t expect_bytes, uint32_t bar_full0,
                                                    uint32_t bar_full_cl0, uint32_t bar_empty0, uint32_t w_smem0,
                                                    uint32_t stage_step) {
  // probe-ahead like issue_chunks_pair: the next stage's empty barrier is probed before this stage's copy is issued
  asm volatile(
      "{\n"
      ".reg .pred p, pw, lead;\n"
      ".reg .b32 c, par, fb, fbc, eb, t, dst, spins, c0, c1, row, zero, ns, nph, neb;\n"
      "mov.u32 c, 0;\n"
      "mov.u32 zero, 0;\n"
      "mov.u32 row, %5;\n"
      "setp.ne.u32 lead, %8, 0;\n"
      "setp.ne.b32 pw, %3, 0;\n"
      "PRODP_LOOP:\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %9, t;\n"
      "add.u32 fbc, %10, t;\n"
      "add.u32 eb, %11, t;\n"
      "@pw bra PRODP_READY;\n"
      "xor.b32 par, %1, 1;\n"
      "mov.u32 spins, 0;\n"
      "mov.u32 c0, %clock;\n"
      "PRODP_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [eb], par;\n"
      "@pw bra PRODP_WAITED;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra PRODP_WAIT;\n"
      "PRODP_WAITED:\n"
      "mov.u32 c1, %clock;\n"
      "sub.u32 c1, c1, c0;\n"
      "add.u32 %2, %2, c1;\n"
      "PRODP_READY:\n"
      "add.u32 ns, %0, %13;\n"
      "mov.u32 nph, %1;\n"
      "setp.ge.u32 p, ns, " VIPNERF_STR(VIPNERF_PAIR_STAGES) ";\n"
      "@p sub.u32 ns, ns, " VIPNERF_STR(VIPNERF_PAIR_STAGES) ";\n"
      "@p xor.b32 nph, nph, 1;\n"
      "shl.b32 t, ns, 3;\n"
      "add.u32 neb, %11, t;\n"
      "xor.b32 par, nph, 1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 pw, [neb], par;\n"
      "@lead mbarrier.arrive.expect_tx.shared::cta.b64 _, [fb], %8;\n"
      "mad.lo.u32 dst, %0, 8192, %12;\n"
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [dst], [%4, {zero, row}], [fbc];\n"
      "add.u32 row, row, %6;\n"
      "mov.u32 %0, ns;\n"
      "mov.u32 %1, nph;\n"
      "add.u32 c, c, 1;\n"
      "setp.lt.u32 p, c, %7;\n"
      "@p bra PRODP_LOOP;\n"
      "selp.u32 %3, 1, 0, pw;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase), "+r"(rs.wait_cycles), "+r"(rs.ready)
      : "l"(tmap), "r"(row0), "r"(row_stride), "r"(n), "r"(expect_bytes), "r"(bar_full0), "r"(bar_full_cl0),
        "r"(bar_empty0), "r"(w_smem0), "r"(stage_step)
      : "memory");
}
// The bias chunk of a layer: ONE accumulating MMA - A = the all-ones block, B = the second K=16 step of the chunk
// (its column 31 holds the bias) - then the stage-release commit.
__device__ __forceinline__ void issue_bias_chunk(RingState& rs, uint32_t d_tmem, uint64_t a_desc, uint64_t w_desc0,
        uint32_t bar_full0, uint32_t bar_empty0, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p, pw, e, pt;\n"
      ".reg .b32 fb, eb, t, spins;\n"
      ".reg .b64 b;\n"
      ".reg .b16 mc;\n"
      "mov.b16 mc, 3;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %5, t;\n"
      "add.u32 eb, %6, t;\n"
      "mov.u32 spins, 0;\n"
      "BIAS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [fb], %1;\n"
      "@pw bra BIAS_READY;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra BIAS_WAIT;\n"
      "BIAS_READY:\n"
      "tcgen05.fence::after_thread_sync;\n"
      "mul.wide.u32 b, %0, 1024;\n"
      "add.s64 b, b, %4;\n"
      "add.s64 b, b, 2;\n"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%2], %3, b, %7, pt;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [eb];\n"
      "add.u32 %0, %0, 1;\n"
      "setp.eq.u32 p, %0, " VIPNERF_STR(VIPNERF_SINGLE_STAGES) ";\n"
      "@p mov.u32 %0, 0;\n"
      "@p xor.b32 %1, %1, 1;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase)
      : "r"(d_tmem), "l"(a_desc), "l"(w_desc0), "r"(bar_full0), "r"(bar_empty0), "r"(idesc)
      : "memory");
}
template <bool kProbe>
__device__ __forceinline__ void issue_bias_chunk_pair(RingState& rs, uint32_t d_tmem, uint64_t a_desc, uint64_t w_desc0,
        uint32_t bar_full0, uint32_t bar_empty0, uint32_t idesc, uint32_t skip_wait = 0) {
  if (!kProbe) rs.ready = 0;
  asm volatile(
      "{\n"
      ".reg .pred p, pw, e, pt, pskip;\n"
      ".reg .b32 fb, eb, t, spins, ns, nph, nfb;\n"
      ".reg .b64 b;\n"
      ".reg .b16 mc;\n"
      "mov.b16 mc, 3;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "elect.sync _|e, 0xffffffff;\n"
      "shl.b32 t, %0, 3;\n"
      "add.u32 fb, %6, t;\n"
      "add.u32 eb, %7, t;\n"
      "mov.u32 spins, 0;\n"
      "setp.ne.b32 pskip, %9, 0;\n"
      "setp.ne.b32 pw, %2, 0;\n"
      "@pskip bra BIAS_READY;\n"
      "@pw bra BIAS_READY;\n"
      "BIAS_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 pw, [fb], %1;\n"
      "@pw bra BIAS_READY;\n"
      "add.u32 spins, spins, 1;\n"
      "setp.gt.u32 p, spins, 4000000;\n"
      "@p trap;\n"
      "bra BIAS_WAIT;\n"
      "BIAS_READY:\n"
      "tcgen05.fence::after_thread_sync;\n"
      "mul.wide.u32 b, %0, 512;\n"
      "add.s64 b, b, %5;\n"
      "add.s64 b, b, 2;\n"
      "add.u32 ns, %0, 1;\n"
      "mov.u32 nph, %1;\n"
      "setp.eq.u32 p, ns, " VIPNERF_STR(VIPNERF_PAIR_STAGES) ";\n"
      "@p mov.u32 ns, 0;\n"
      "@p xor.b32 nph, nph, 1;\n"
      "shl.b32 t, ns, 3;\n"
      "add.u32 nfb, %6, t;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 pw, [nfb], nph;\n"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%3], %4, b, %8, pt;\n"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [eb], mc;\n"
      "mov.u32 %0, ns;\n"
      "mov.u32 %1, nph;\n"
      "selp.u32 %2, 1, 0, pw;\n"
      "}\n"
      : "+r"(rs.stage), "+r"(rs.phase), "+r"(rs.ready)
      : "r"(d_tmem), "l"(a_desc), "l"(w_desc0), "r"(bar_full0), "r"(bar_empty0), "r"(idesc), "r"(skip_wait)
      : "memory");
  if (!kProbe) rs.ready = 0;
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n.reg .pred e;\nelect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect_pair(uint32_t bar) {  // arrives on `bar` of BOTH CTAs of the pair
  asm volatile(
      "{\n.reg .pred e;\n.reg .b16 mc;\nmov.b16 mc, 3;\nelect.sync _|e, 0xffffffff;\n"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], mc;\n}\n" ::"r"(bar)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30), SBO = 1024 B (8 rows x 128 B) in [32,46),
// version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// Descriptor of the all-ones A operand of the bias chunks (128 rows x K=16, SWIZZLE_NONE: 8-row x 16-byte core
// matrices).  Every element is 1.0, so ONE 128-byte core matrix serves all of them through zero leading / stride
// offsets; the VIPNERF_ONES_4K build lays the 16 x 2 core matrices out conventionally (LBO 128 B, SBO 256 B).
__device__ __forceinline__ uint64_t make_desc_ones(uint32_t smem_addr) {
#ifdef VIPNERF_ONES_4K
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
#else
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 46);
#endif
}
// K-major SWIZZLE_64B descriptor of a weight chunk: 64-byte rows, SBO = 512 B (8 rows), layout type 4.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
         (4ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
// two fp32 -> packed 16-bit operands of the tensor core: bf16, or (kHalf) fp16 saturating to +-65504 instead of inf
template <bool kHalf>
__device__ __forceinline__ uint32_t pack_op(float lo, float hi) {
  if (!kHalf) return pack_bf16(lo, hi);
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// residual of a packed bf16 pair: (lo, hi) - float(bf16(lo, hi)), packed to bf16 again
__device__ __forceinline__ uint32_t pack_bf16_residual(float lo, float hi, uint32_t packed) {
  const float rlo = lo - __uint_as_float(packed << 16);
  const float rhi = hi - __uint_as_float(packed & 0xFFFF0000u);
  return pack_bf16(rlo, rhi);
}
// kAccurate (BF16X3): IEEE division and libm expf; otherwise the MUFU approximations (rel. error ~1e-6, far below the
// bf16 noise of the logits)
template <bool kAccurate>
__device__ __forceinline__ float sigmoidf(float x) {
  return kAccurate ? 1.f / (1.f + expf(-x)) : __fdividef(1.f, 1.f + __expf(-x));
}
// Packed fp32x2 arithmetic (Blackwell FADD2 / FFMA2) and fused ReLU + bf16x2 conversion (F2FP.RELU)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <bool kRelu, bool kHalf>
__device__ __forceinline__ uint32_t cvt_op_x2(uint64_t v) {  // element 0 (low half of v) -> low 16 bits
  float lo, hi;
  unpack_f32x2(v, lo, hi);
  uint32_t r;
  if (kHalf) {
    if (kRelu) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  } else {
    if (kRelu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  }
  return r;
}

// ------------------------------------------------------------------------------------------ parameters
struct PassDesc {
  float* z;               // [R,S] sample depths: read, or (fused coarse pass) written by the encoding stage
  const uint8_t* packed;  // packed weights of this pass's MLP
  float* sigma;           // [R,S]   network outputs of the pass
  float* rgb;             // [R,S,3]
  float* vis;             // [R,S]
  float* vis2;            // [R,S,V] visibility of every sample from the secondary views (null when V = 0)
  int64_t n_points;       // R*S
  int S;
  int compute_z;          // 1: z = get_z_vals_coarse(near, far) is generated here
};

struct TcParams {
  RayPtrs rp;
  RenderFlags fl;
  PassDesc pass[2];
  PassOutPtrs out[2];
  int64_t n_rays;
  int64_t n_units;  // staged: tiles of pass 0;  fused: ray pairs
  int n_fine;
  int has_fine;
  unsigned long long* prof;  // optional cycle counters of CTA 0 (vipnerf_debug_set_profile_buffer), else null
  // CTA-pair kernels: tensor maps over each pass's chunk-image stream seen as [rows][32] bf16 (64-byte rows, no
  // swizzle - the images are stored pre-swizzled); [pass][0] box = 128 rows (half of a 256-row chunk), [pass][1] box =
  // 64 rows (half of an M9 chunk)
  alignas(64) CUtensorMap wmap[2][2];
  int debug_noring;   // timing experiment (VIPNERF_TC_DEBUG_NORING=1): MMAs do not wait for weights - results are garbage
};

// Work of one tile slot: item i -> (pass, tile).  `unit` indexes the work units of the launch (fused: ray pairs,
// staged: tiles); `cs` is the slot's index over the whole grid.  With CTA pairs the two CTAs of a cluster share a
// slot index and split its units alternately (rank 0: even positions, rank 1: odd), and BOTH get the same number of
// items - the tensor core works on both CTAs' tiles with every MMA - so a missing unit becomes a dummy item (tile
// index past the end: every row invalid, nothing stored).
template <bool kFused, bool kPair>
struct WorkList {
  int64_t first, stride, end_unit, dummy;
  int rank;
  int n_first_pass;  // fused: coarse tiles (= ray pairs) of this CTA's slot;  staged: tiles
  int n_items;
  __device__ WorkList(const TcParams& p, int cs, int n_cs, int cta_rank) {
    rank = cta_rank;
    dummy = p.n_units;
    int64_t cnt;
    if (kFused) {
      const int64_t per = (p.n_units + n_cs - 1) / n_cs;
      first = (int64_t)cs * per;
      stride = 1;
      cnt = p.n_units - first;
      cnt = cnt < 0 ? 0 : (cnt > per ? per : cnt);
    } else {
      first = cs;
      stride = n_cs;
      cnt = first < p.n_units ? (p.n_units - first + stride - 1) / stride : 0;
    }
    end_unit = first + cnt * stride;
    n_first_pass = (int)(kPair ? (cnt + 1) / 2 : cnt);
    n_items = n_first_pass * ((kFused && p.has_fine) ? 4 : 1);
  }
  __device__ __forceinline__ int pass_of(int i) const { return kFused && i >= n_first_pass ? 1 : 0; }
  __device__ __forceinline__ int64_t unit_of(int i) const {  // i-th unit of this CTA's slot (or the dummy)
    const int64_t u = first + (kPair ? 2 * (int64_t)i + rank : (int64_t)i) * stride;
    return u < end_unit ? u : dummy;
  }
  __device__ __forceinline__ int64_t tile_of(int i) const {
    if (!kFused) return unit_of(i);
    if (i < n_first_pass) return unit_of(i);
    const int j = i - n_first_pass;
    return 3 * unit_of(j / 3) + j % 3;
  }
};

// ------------------------------------------------------------------------------------------ epilogue pieces
// sin / cos of 2^k * x for k = 0..L-1 with ONE range reduction: t = x / (2 pi) is formed in double-float
// (hi + lo, ~48 bits), scaling by 2^k is exact, and frac(2^k t) is exact, so every octave sees an argument in
// [-0.5, 0.5] turns with ~1e-7 absolute error - independent of the frequency (the reference evaluates
// sin(fl(x * 2^k)) with a full-precision libm).  kAccurate: sincospif (~1 ulp); otherwise the MUFU
// approximations (abs. error ~5e-7, far below the bf16 rounding the value gets next).
template <bool kAccurate>
__device__ __forceinline__ void sincos_octave(float t_hi, float t_lo, int k, float& s, float& c) {
  const float scale = (float)(1 << k);
  const float a = t_hi * scale;
  const float r = (a - rintf(a)) + t_lo * scale;   // turns, |r| <= 0.5 (+ tiny)
  if (kAccurate) {
    sincospif(2.f * r, &s, &c);
  } else {
    const float ang = r * 6.2831854820251465f;
    s = __sinf(ang);
    c = __cosf(ang);
  }
}

// bf16(gamma(point)), 64 columns (63 + the constant 1), packed two per word: enc[0..31] (BF16X3: enc[32..63] = the
// residuals).  Column order of PositionalEncoder.encode (VipNeRF01.py:439-448): x(3), then per octave sin(3), cos(3).
// Column 63 is the constant 1 that carries the biases of M0 / M5 through the tensor core (layout.cuh).
template <bool kSplit3, bool kHalf>
__device__ __forceinline__ void compute_point_encoding(float x, float y, float z, uint32_t* enc) {
  float v[64];
  v[0] = x; v[1] = y; v[2] = z;
  const float p[3] = {x, y, z};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float t_hi = p[a] * 0.15915493667125702f;
    const float t_lo = fmaf(p[a], 6.4206382432985265e-09f, fmaf(p[a], 0.15915493667125702f, -t_hi));
#pragma unroll
    for (int k = 0; k < kLPts; ++k) sincos_octave<kSplit3>(t_hi, t_lo, k, v[3 + 6 * k + a], v[6 + 6 * k + a]);
  }
  v[63] = 1.f;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    enc[q] = pack_op<kHalf>(v[2 * q], v[2 * q + 1]);
    if (kSplit3) enc[32 + q] = pack_bf16_residual(v[2 * q], v[2 * q + 1], enc[q]);
  }
}
// Row `row` of k-block 0 of the slot's activation buffer <- the packed encoding (K-major SWIZZLE_128B).
template <bool kSplit3>
__device__ __forceinline__ void store_point_encoding(uint8_t* smem, int slot, int row, const uint32_t* enc) {
  const uint32_t hi_base = smem_u32(smem + kOffA + (kSplit3 ? 0 : slot) * kABytes) + row * 128;
  const uint32_t lo_base = smem_u32(smem + kOffA + kABytes) + row * 128;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    const uint32_t off = (uint32_t)((ch ^ (row & 7)) << 4);
    st_shared_v4(hi_base + off, enc[4 * ch], enc[4 * ch + 1], enc[4 * ch + 2], enc[4 * ch + 3]);
    if (kSplit3) st_shared_v4(lo_base + off, enc[32 + 4 * ch], enc[33 + 4 * ch], enc[34 + 4 * ch], enc[35 + 4 * ch]);
  }
}

// One trunk / feature layer epilogue for one row: accumulator (256 fp32 TMEM columns, bias already added by the
// tensor core) (-> ReLU) -> bf16 -> A buffer (in place).  kSigma additionally accumulates the density head on the
// fp32 activations.  The eight 32-column TMEM loads are software-pipelined: block cb+1 is in flight while block cb
// is processed (tcgen05.wait::ld waits for everything outstanding, so the next load is issued right after it).
template <bool kSplit3, bool kRelu, bool kSigma, bool kHalf>
__device__ __forceinline__ float layer_epilogue(uint8_t* smem, int slot, int row, uint32_t taddr,
                                                const float* w_sigma /* shared */) {
  // density head: four independent partial sums (packed pairs on the throughput path) keep the FMA chain short
  float sigma_acc[4] = {0.f, 0.f, 0.f, 0.f};
  uint64_t sigma_acc2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) sigma_acc2[i] = pack_f32x2(0.f, 0.f);
  const uint32_t hi_base = smem_u32(smem + kOffA + (kSplit3 ? 0 : slot) * kABytes) + row * 128;
  const uint32_t lo_base = smem_u32(smem + kOffA + kABytes) + row * 128;
  uint32_t v[2][32];
  tmem_ld32(taddr, v[0]);
#pragma unroll
  for (int cb = 0; cb < 8; ++cb) {
    float ws[32];
    if (kSigma) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = *(reinterpret_cast<const float4*>(w_sigma + cb * 32) + q);   // shared memory (kOffHead)
        ws[4 * q] = t.x; ws[4 * q + 1] = t.y; ws[4 * q + 2] = t.z; ws[4 * q + 3] = t.w;
      }
    }
    tmem_ld_wait();
    if (cb + 1 < 8) tmem_ld32(taddr + (cb + 1) * 32, v[(cb + 1) & 1]);
    const uint32_t kb_off = (uint32_t)(cb >> 1) * kKBlockBytes;
    if (!kSplit3) {
      // throughput path: ReLU fused into the bf16x2 conversion - one instruction per two elements
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int e = 8 * j + 2 * q;
          w[q] = cvt_op_x2<kRelu, kHalf>(pack_f32x2(__uint_as_float(v[cb & 1][e]), __uint_as_float(v[cb & 1][e + 1])));
          if (kSigma)   // sigma += relu(h) . w_sigma on the fp32 accumulator values (VipNeRF01.py:546-553)
            sigma_acc2[q] = ffma2(pack_f32x2(fmaxf(__uint_as_float(v[cb & 1][e]), 0.f), fmaxf(__uint_as_float(v[cb & 1][e + 1]), 0.f)),
                                  pack_f32x2(ws[e], ws[e + 1]), sigma_acc2[q]);
        }
        const int ch = (cb & 1) * 4 + j;
        st_shared_v4(hi_base + kb_off + (uint32_t)((ch ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
      }
      continue;
    }
    float h[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      h[j] = __uint_as_float(v[cb & 1][j]);
      if (kRelu) h[j] = fmaxf(h[j], 0.f);
    }
    if (kSigma) {
#pragma unroll
      for (int j = 0; j < 32; ++j) sigma_acc[j & 3] = fmaf(h[j], ws[j], sigma_acc[j & 3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = (cb & 1) * 4 + j;
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = pack_op<kHalf>(h[8 * j + 2 * q], h[8 * j + 2 * q + 1]);
      const uint32_t off = kb_off + (uint32_t)((ch ^ (row & 7)) << 4);
      st_shared_v4(hi_base + off, w[0], w[1], w[2], w[3]);
      if (kSplit3) {
        uint32_t r[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = pack_bf16_residual(h[8 * j + 2 * q], h[8 * j + 2 * q + 1], w[q]);
        st_shared_v4(lo_base + off, r[0], r[1], r[2], r[3]);
      }
    }
  }
  if (!kSplit3) {
    float lo, hi;
    unpack_f32x2(fadd2(fadd2(sigma_acc2[0], sigma_acc2[1]), fadd2(sigma_acc2[2], sigma_acc2[3])), lo, hi);
    return lo + hi;
  }
  return (sigma_acc[0] + sigma_acc[1]) + (sigma_acc[2] + sigma_acc[3]);
}

// M9 epilogue for one row: relu(acc + (bias + view-direction part)) . views_output_linear -> 4 logits.
// Four independent accumulator pairs (one per 32-column block, interleaved even/odd columns) keep the FFMA2
// dependency chains short; they are summed at the end.
__device__ __forceinline__ void view_epilogue(uint32_t taddr, const float* vb_row, const float* w_out /* shared */,
                                              float (&o)[4]) {
  uint64_t acc01[4], acc23[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc01[i] = acc23[i] = pack_f32x2(0.f, 0.f);
  uint32_t v[2][32];
  tmem_ld32(taddr, v[0]);
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float b[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = *reinterpret_cast<const float4*>(vb_row + cb * 32 + 4 * q);
      b[4 * q] = t.x; b[4 * q + 1] = t.y; b[4 * q + 2] = t.z; b[4 * q + 3] = t.w;
    }
    tmem_ld_wait();
    if (cb + 1 < 4) tmem_ld32(taddr + (cb + 1) * 32, v[(cb + 1) & 1]);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float h = fmaxf(__uint_as_float(v[cb & 1][j]) + b[j], 0.f);
      const float4 w = *(reinterpret_cast<const float4*>(w_out) + cb * 32 + j);
      const uint64_t hh = pack_f32x2(h, h);
      acc01[j & 3] = ffma2(hh, pack_f32x2(w.x, w.y), acc01[j & 3]);
      acc23[j & 3] = ffma2(hh, pack_f32x2(w.z, w.w), acc23[j & 3]);
    }
  }
  const uint64_t o01 = fadd2(fadd2(acc01[0], acc01[1]), fadd2(acc01[2], acc01[3]));
  const uint64_t o23 = fadd2(fadd2(acc23[0], acc23[1]), fadd2(acc23[2], acc23[3]));
  unpack_f32x2(o01, o[0], o[1]);
  unpack_f32x2(o23, o[2], o[3]);
}

// The MMA steps of one tile, in issue order.  The skip layer M5 is two steps: its h4 part (K = 256 over the whole
// activation buffer; weight chunks 2..9 of the layer) and, once that has retired and the epilogue group has put the
// encoding back into k-block 0, its encoding part (K = 64, accumulating; chunks 0..1, carries the bias through the
// constant-one column).  With V secondary views, V more steps follow M9: E_v = direction encoding of view v (32
// columns of k-block v/2, written by the epilogue group once M9 has retired) x the view-direction chunk, into TMEM
// columns [128,256) of the slot - M9's accumulator (columns [0,128)) stays in place and is re-read for every view.
// Every step ends with a d_ready commit and starts with an a_ready wait.
constexpr int kNumBaseSteps = 10;
struct StepDesc {
  int layer;            // matrix layer M0..M9 / kViewChunkLayer (layout.cuh)
  uint32_t first_chunk; // first weight chunk of the layer's stream used by the step
  uint32_t n_chunks;    // chunks multiplied against the activation buffer
  uint32_t accumulate;  // 0: the first MMA overwrites the accumulator
  bool bias_chunk;      // followed by the layer's bias chunk (all-ones A operand)
  uint32_t a_units;     // A descriptor offset (16-byte units) of the step's first 32 columns
  uint32_t d_col;       // TMEM column offset of the accumulator inside the slot
};
__device__ __forceinline__ StepDesc step_desc(int st) {
  if (st == 0) return {0, 0, 2, 0, false, 0, 0};
  if (st == 5) return {5, 2, 8, 0, false, 0, 0};
  if (st == 6) return {5, 0, 2, 1, false, 0, 0};
  if (st >= kNumBaseSteps) {
    const uint32_t v = (uint32_t)(st - kNumBaseSteps);
    return {kViewChunkLayer, 0, 1, 0, false, (v >> 1) * (kKBlockBytes >> 4) + (v & 1) * 4, 128};
  }
  const int l = st < 5 ? st : (st == 9 ? 9 : st - 1);   // 7 -> M6, 8 -> M7, 9 -> M9 (feature_linear folded in, layout.cuh)
  return {l, 0, 8, 0, layer_has_bias_chunk(l), 0, 0};
}

// tc_layer_byte_offset for a run-time layer index without the run-time loop (the producer lanes evaluated it per step)
__device__ __forceinline__ uint32_t tc_layer_offset_rt(int l) {
  switch (l) {
    case 0: return tc_layer_byte_offset(0);
    case 1: return tc_layer_byte_offset(1);
    case 2: return tc_layer_byte_offset(2);
    case 3: return tc_layer_byte_offset(3);
    case 4: return tc_layer_byte_offset(4);
    case 5: return tc_layer_byte_offset(5);
    case 6: return tc_layer_byte_offset(6);
    case 7: return tc_layer_byte_offset(7);
    case 9: return tc_layer_byte_offset(9);
    default: return tc_layer_byte_offset(kViewChunkLayer);
  }
}

// Visibility of this row's sample from each secondary view (VipNeRF01.py:218-226, :527-530): the same hidden layer
// with the unit vector from that camera to the SAMPLE.  Called by every thread of an epilogue group once M9 has
// retired, so the activation buffer is free: the V direction encodings (27 columns + zeros + a constant 1 facing the
// bias column, 32 per view) go into its first k-blocks; then per view one K=32 MMA step into TMEM columns [128,256)
// and an epilogue relu(M9 accumulator + E_v) . w_out[:, 3].  Kept out of line: the eval render without secondary
// views must not pay registers for it.
template <bool kSplit3, bool kPair, bool kHalf>
__device__ __noinline__ void secondary_views(uint8_t* smem, const TcParams& p, const PassDesc& ps, int slot, int row,
                                             int lane, int64_t pg, bool valid, int64_t ray, uint32_t taddr,
                                             uint32_t d_ready_bar, uint32_t a_ready_bar, uint32_t& d_parity) {
  auto arrive_a_ready = [&]() {
    __syncwarp();
    if (lane == 0) { if (kPair) mbar_arrive_cluster(a_ready_bar); else mbar_arrive(a_ready_bar); }
  };
  const float* small = reinterpret_cast<const float*>(ps.packed);
  const int V = p.fl.n_sec_views;
  const int64_t pc = valid ? pg : ps.n_points - 1;
  const float o3[3] = {p.rp.rays_o[3 * ray], p.rp.rays_o[3 * ray + 1], p.rp.rays_o[3 * ray + 2]};
  const float d3[3] = {p.rp.rays_d[3 * ray], p.rp.rays_d[3 * ray + 1], p.rp.rays_d[3 * ray + 2]};
  float zz = ps.z[pc];
  if (p.fl.ndc) zz = depth_from_ndc_secondary(zz, o3[2], d3[2]);
  for (int v = 0; v < V; ++v) {
    const float* c2 = p.rp.rays_o2 + (ray * V + v) * 3;
    const float o2[3] = {c2[0], c2[1], c2[2]};
    float dd[3], pe[32];
    secondary_view_dir(o3, d3, zz, o2, dd);
#pragma unroll
    for (int j = 0; j < 32; ++j) pe[j] = 0.f;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      pe[ax] = dd[ax];
#pragma unroll
      for (int k = 0; k < kLView; ++k) {
        const float ang = dd[ax] * (float)(1 << k);
        if (kSplit3) { pe[3 + 6 * k + ax] = sinf(ang); pe[6 + 6 * k + ax] = cosf(ang); }
        else { pe[3 + 6 * k + ax] = __sinf(ang); pe[6 + 6 * k + ax] = __cosf(ang); }
      }
    }
    pe[31] = 1.f;
    const uint32_t kb = (uint32_t)(v >> 1) * kKBlockBytes + row * 128;
    const uint32_t hi_base = smem_u32(smem + kOffA + (kSplit3 ? 0 : slot) * kABytes) + kb;
    const uint32_t lo_base = smem_u32(smem + kOffA + kABytes) + kb;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4], r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        w[e] = pack_op<kHalf>(pe[8 * q + 2 * e], pe[8 * q + 2 * e + 1]);
        r[e] = kSplit3 ? pack_bf16_residual(pe[8 * q + 2 * e], pe[8 * q + 2 * e + 1], w[e]) : 0u;
      }
      const uint32_t off = (uint32_t)((((v & 1) * 4 + q) ^ (row & 7)) << 4);
      st_shared_v4(hi_base + off, w[0], w[1], w[2], w[3]);
      if (kSplit3) st_shared_v4(lo_base + off, r[0], r[1], r[2], r[3]);
    }
  }
  fence_proxy_async();
  arrive_a_ready();
  const float* head = reinterpret_cast<const float*>(smem + kOffHead) + (&ps == &p.pass[1] ? kHeadFloats : 0);
  const float b3 = reinterpret_cast<const float*>(smem + kOffHeadBias)[(&ps == &p.pass[1] ? 8 : 0) + 3];
  for (int v = 0; v < V; ++v) {
    mbar_wait(d_ready_bar, d_parity);
    d_parity ^= 1;
    tc_fence_after();
    float acc = 0.f;
    uint32_t f[32], e[32];
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      tmem_ld32(taddr + cb * 32, f);
      tmem_ld32(taddr + 128 + cb * 32, e);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float h = fmaxf(__uint_as_float(f[j]) + __uint_as_float(e[j]), 0.f);
        acc = fmaf(h, head[4 * (cb * 32 + j) + 3], acc);
      }
    }
    tc_fence_before();
    if (v + 1 < V) arrive_a_ready();   // columns [128,256) are drained: the next view's step may overwrite them
    if (valid) ps.vis2[pg * V + v] = sigmoidf<kSplit3>(acc + b3);
  }
}

// ------------------------------------------------------------------------------------------ the kernel
template <bool kSplit3, bool kFused, bool kProf, bool kPair, bool kSec, bool kHalf>
__global__ void __launch_bounds__(kNumThreads, 1) k_render_tc(const __grid_constant__ TcParams p) {
  static_assert(!(kSplit3 && kHalf), "the hi/lo split is a bf16 mode");
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int kSlots = kSplit3 ? 1 : 2;
  constexpr int kStages = kPair ? kPairStages : kSingleStages;
  constexpr uint32_t kStageBytes = kPair ? 8192 : 16384;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0;   // 0 = leader of the CTA pair (issues the MMAs)
  const uint32_t bar0 = smem_u32(smem + kOffBar);
  auto bar = [&](int idx) { return bar0 + 8u * idx; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("vipnerf: dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    // w_full: one arrival (the producer's arrive.expect_tx; in pair mode it covers both CTAs' copies, which complete
    // on the leader's barrier); a_ready: one arrival per epilogue warp of every CTA feeding the MMA
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar(kBarWFull + s), 1);
      mbar_init(bar(kBarWEmpty + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(kBarAReady + s), kPair ? 8 : 4);
      mbar_init(bar(kBarDReady + s), 1);
      mbar_init(bar(kBarRayFull + s), 4);   // one arrival per warp of the slot's epilogue group
      reinterpret_cast<volatile uint32_t*>(smem + kOffRayDone)[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    float* head = reinterpret_cast<float*>(smem + kOffHead);
    for (int i = threadIdx.x; i < 2 * (int)kHeadFloats; i += kNumThreads) {
      const float* small = reinterpret_cast<const float*>(p.pass[i / kHeadFloats].packed);
      const int j = i % kHeadFloats;
      head[i] = j < 512 ? small[kOffWOut + j] : small[kOffWSigma + j - 512];
    }
    if (threadIdx.x < 16) {
      const float* small = reinterpret_cast<const float*>(p.pass[threadIdx.x >> 3].packed);
      const int j = threadIdx.x & 7;
      reinterpret_cast<float*>(smem + kOffHeadBias)[threadIdx.x] = j < 4 ? small[kOffBOut + j] : (j == 4 ? small[kOffBSigma] : 0.f);
    }
  }
  if (warp == 9) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kOffTmemPtr)),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem + kOffTmemPtr)),
                   "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int n_cta_groups = kPair ? gridDim.x / 2 : gridDim.x;       // clusters (pair mode) or CTAs
  const int cta_group_idx = kPair ? blockIdx.x / 2 : blockIdx.x;
  const int n_slots_total = n_cta_groups * kSlots;

  if (warp < 8) {
    // =================================================================== epilogue groups
    const int group = warp >> 2;
    if (group < kSlots) {
      const int slot = group;
      const int row = threadIdx.x & 127;
      const WorkList<kFused, kPair> work(p, cta_group_idx * kSlots + slot, n_slots_total, (int)cta_rank);
      // a_ready lives in the leader CTA: the peer's epilogue threads arrive on it through the cluster window
      const uint32_t a_ready_bar = kPair ? map_to_cta(bar(kBarAReady + slot), 0) : bar(kBarAReady + slot);
      // one arrival per warp: every lane has fenced its writes to the async proxy, the warp converges, lane 0 signals
      auto arrive_a_ready = [&]() {
        __syncwarp();
        if (lane == 0) { if (kPair) mbar_arrive_cluster(a_ready_bar); else mbar_arrive(a_ready_bar); }
      };
      const uint32_t taddr = tmem_base + (uint32_t)(slot * 256) + ((uint32_t)((warp & 3) * 32) << 16);
      float* vb = reinterpret_cast<float*>(smem + kOffVb) + slot * 256;
      float* pev = reinterpret_cast<float*>(smem + kOffPev) + slot * 64;
      const bool prof_on = kProf && p.prof != nullptr && blockIdx.x == 0 && row == 0;
      long long c_enc = 0, c_vb = 0, c_wait = 0, c_epi = 0, c_view = 0, c_hook = 0;
      const long long c_begin = kProf ? clock64() : 0;

      // the all-ones operand of the bias chunks; written (and fenced with the first encoding) by group 0, whose first
      // a_ready precedes every MMA of the CTA pair
      if (slot == 0) {
        uint32_t* ones = reinterpret_cast<uint32_t*>(smem + kOffOnes);
        for (int i = row; i < (int)(kOnesBytes / 4); i += 128) ones[i] = kHalf ? 0x3C003C00u : 0x3F803F80u;   // two bf16 (fp16) 1.0
      }
      // This row's packed point encoding: computed ahead of time (during M6 of the previous tile), stored into k-block 0
      // at the tile boundary and once more for the encoding part of the skip layer M5.
      uint32_t enc[kSplit3 ? 64 : 32];
      // Sample position of this thread's row of item `it` and its encoding -> enc
      // (VipNeRF01.py:105-107, :173-203, :439-448).
      auto encode_item = [&](int it) {
        const long long t0 = kProf ? clock64() : 0;
        const PassDesc& ps = p.pass[work.pass_of(it)];
        const int64_t pg = work.tile_of(it) * kTile + row;
        const bool valid = pg < ps.n_points;
        const int64_t pc = valid ? pg : ps.n_points - 1;
        const int64_t ray = pc / ps.S;
        float zv;
        if (ps.compute_z) {
          const int s = (int)(pc - ray * ps.S);
          zv = coarse_z_at(p.rp.near[ray], p.rp.far[ray], p.rp.t_vals, s, ps.S, p.fl.lindisp,
                           p.rp.t_rand ? p.rp.t_rand + ray * ps.S : nullptr);
          if (valid) ps.z[pg] = zv;
        } else {
          zv = ps.z[pc];
        }
        const float px = fadd(p.rp.pts_o[3 * ray + 0], fmul(p.rp.pts_d[3 * ray + 0], zv));
        const float py = fadd(p.rp.pts_o[3 * ray + 1], fmul(p.rp.pts_d[3 * ray + 1], zv));
        const float pz = fadd(p.rp.pts_o[3 * ray + 2], fmul(p.rp.pts_d[3 * ray + 2], zv));
        compute_point_encoding<kSplit3, kHalf>(px, py, pz, enc);
        if (kProf) c_enc += clock64() - t0;
      };
      // enc -> k-block 0 of the slot's (idle) activation buffer, visible to the tensor core, then a_ready
      auto publish_encoding = [&]() {
        store_point_encoding<kSplit3>(smem, slot, row, enc);
        fence_proxy_async();
        arrive_a_ready();
      };
      // View-direction columns of views_linears.0 (+ bias) for the (at most two) rays of item `it`, fp32:
      // vb[rs][c] = b[c] + sum_j W[c][256 + j] * gamma(view_dir[ray_first + rs])[j]   (VipNeRF01.py:576-579)
      auto view_bias_item = [&](int it) {
        const long long t0 = kProf ? clock64() : 0;
        const PassDesc& ps = p.pass[work.pass_of(it)];
        const float* small = reinterpret_cast<const float*>(ps.packed);
        const int64_t ray_first = min((work.tile_of(it) * kTile) / ps.S, p.n_rays - 1);
        if (row < 2 * kEncView) {
          const int rs = row / kEncView, j = row % kEncView;
          const int64_t r2 = min(ray_first + rs, p.n_rays - 1);
          float val;
          if (j < 3) {
            val = p.rp.view_dirs[3 * r2 + j];
          } else {
            const int k = (j - 3) / 6, r6 = (j - 3) % 6;
            const float a = p.rp.view_dirs[3 * r2 + r6 % 3] * (float)(1 << k);
            val = r6 < 3 ? sinf(a) : cosf(a);
          }
          pev[rs * 32 + j] = val;
        }
        float w[kEncView];
#pragma unroll
        for (int j = 0; j < kEncView; ++j) w[j] = __ldg(small + kOffWViewDir + j * 128 + row);
        float a0 = small[kOffBiasViewsFused + row], a1 = a0;
        group_sync(group);
#pragma unroll
        for (int j = 0; j < kEncView; ++j) {
          a0 = fmaf(w[j], pev[j], a0);
          a1 = fmaf(w[j], pev[32 + j], a1);
        }
        vb[row] = a0;
        vb[128 + row] = a1;
        group_sync(group);
        if (kProf) c_vb += clock64() - t0;
      };

      // Fine depths of the j-th ray pair of this slot exist once the ray warp has completed its j-th event (the coarse
      // tiles are the slot's first events, one per pair).
      uint32_t n_ray_events = 0;
      auto ray_done_count = [&]() {
        uint32_t v;
        asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(smem + kOffRayDone) + 4u * slot) : "memory");
        return v;
      };
      auto depths_ready = [&](int it2) {
        if (!kFused || work.pass_of(it2) == 0 || (p.debug_noring & 4)) return true;
        return ray_done_count() > (uint32_t)((it2 - work.n_first_pass) / 3);
      };
      auto wait_depths = [&](int it2) {
        if (depths_ready(it2)) return;
        const long long t0 = clock64();
        while (!depths_ready(it2)) {
          if (clock64() - t0 > kTimeoutCycles) { printf("vipnerf: ray warp timed out (block %d slot %d)\n", blockIdx.x, slot); __trap(); }
        }
      };
      uint32_t d_parity = 0;
      if (work.n_items > 0) {
        encode_item(0);
        publish_encoding();
      }
      for (int it = 0; it < work.n_items; ++it) {
        const int pi = work.pass_of(it);
        const PassDesc& ps = p.pass[pi];
        const int64_t tile = work.tile_of(it);
        const int64_t pg = tile * kTile + row;
        const bool valid = pg < ps.n_points;
        const int64_t ray = (valid ? pg : ps.n_points - 1) / ps.S;
        const int64_t ray_first = min((tile * kTile) / ps.S, p.n_rays - 1);
        const float* head = reinterpret_cast<const float*>(smem + kOffHead) + pi * kHeadFloats;
        const float* head_bias = reinterpret_cast<const float*>(smem + kOffHeadBias) + pi * 8;
        // the next item's depths exist unless it is the fine tile of the pair whose coarse tile is this item
        const bool has_next = it + 1 < work.n_items;
        // (checked again, without blocking, right before the prefetch)
        bool next_encoded = false;

        float sigma_lin = 0.f;
#pragma unroll 1
        for (int l = 0; l < 8; ++l) {
          if (l == 5) {
            // skip layer M5 = h4 part (K=256, issued on a_ready(M4's epilogue)) + encoding part (K=64, carries the bias):
            // when the h4 part has retired, k-block 0 is free to take the encoding back
            // The per-ray view-direction term of this tile's M9 is computed HERE, in the shadow of M5's 2048 tensor
            // cycles the thread would otherwise spend waiting (vb / pev are free: the previous tile's M9 epilogue is done).
            view_bias_item(it);
            const long long tw = kProf ? clock64() : 0;
            mbar_wait(bar(kBarDReady + slot), d_parity);
            d_parity ^= 1;
            if (kProf) c_wait += clock64() - tw;
            publish_encoding();
          }
          const long long t0 = kProf ? clock64() : 0;
          mbar_wait(bar(kBarDReady + slot), d_parity);
          d_parity ^= 1;
          tc_fence_after();
          const long long t1 = kProf ? clock64() : 0;
          c_wait += t1 - t0;
          if (l == 7) sigma_lin = layer_epilogue<kSplit3, true, true, kHalf>(smem, slot, row, taddr, head + 512);
          else layer_epilogue<kSplit3, true, false, kHalf>(smem, slot, row, taddr, nullptr);
          fence_proxy_async();
          tc_fence_before();
          arrive_a_ready();
          if (kProf) c_epi += clock64() - t1;
          if (l == 5) {
            // M5 has retired: enc is free for the next tile's encoding.  Use the time this slot's M6 spends on the tensor pipe.
            if (has_next && depths_ready(it + 1)) { encode_item(it + 1); next_encoded = true; }
          }
        }
        const long long t0 = kProf ? clock64() : 0;
        mbar_wait(bar(kBarDReady + slot), d_parity);
        d_parity ^= 1;
        tc_fence_after();
        const long long t1 = kProf ? clock64() : 0;
        c_wait += t1 - t0;
        float o[4];
        if (!(p.debug_noring & 8)) view_epilogue(taddr, vb + (int)(ray - ray_first) * 128, head, o);
        else { o[0] = o[1] = o[2] = o[3] = 0.f; }
        tc_fence_before();  // orders this tile's last tcgen05.ld before the next tile's first MMA into the slot
        if (valid) {
          ps.sigma[pg] = fmaxf(sigma_lin + head_bias[4], 0.f);  // :546-553 (eval: no noise)
          ps.rgb[3 * pg + 0] = sigmoidf<kSplit3>(o[0] + head_bias[0]);   // :585-594
          ps.rgb[3 * pg + 1] = sigmoidf<kSplit3>(o[1] + head_bias[1]);
          ps.rgb[3 * pg + 2] = sigmoidf<kSplit3>(o[2] + head_bias[2]);
          ps.vis[pg] = sigmoidf<kSplit3>(o[3] + head_bias[3]);
        }
        if (kSec)   // compile-time: the eval render without secondary views pays nothing for this path
          secondary_views<kSplit3, kPair, kHalf>(smem, p, ps, slot, row, lane, pg, valid, ray, taddr, bar(kBarDReady + slot),
                                          a_ready_bar, d_parity);
        if (kProf) c_view += clock64() - t1;
        if (kFused) {
          // ---- per-ray stages on the two rays a pair of tiles completes: handed to the slot's ray warp.  Event k may
          // only be signalled once the ray warp has consumed event k-1 (the mbarrier holds a single pending phase).
          const bool pair_done = pi == 0 || (tile % 3) == 2;
          if (pair_done && !(p.debug_noring & 4)) {
            const long long th = kProf ? clock64() : 0;
            while (ray_done_count() < n_ray_events) {}
            ++n_ray_events;
            __threadfence_block();  // this thread's network outputs (global) before the signal
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarRayFull + slot));
            if (kProf) c_hook += clock64() - th;
          }
        }
        if (has_next) {
          if (!next_encoded) { wait_depths(it + 1); encode_item(it + 1); }
          publish_encoding();   // the activation buffer is idle: M9 has retired
        }
      }
      tc_fence_before();
      if (prof_on) {
        unsigned long long* q = p.prof + slot * 16;
        q[0] = c_enc; q[1] = c_vb; q[2] = c_wait; q[3] = c_epi; q[4] = c_view; q[5] = c_hook;
        q[6] = (unsigned long long)work.n_items; q[7] = (unsigned long long)(clock64() - c_begin);
      }
    }
  } else if (warp == 8 || (kPair && warp == 10)) {
    // =================================================================== weight producers
    // Stream the weight chunks of every (tile, step, slot) in the order the MMA issuer consumes them (StepDesc).
    // One lane can start a copy only every ~330 cycles (barrier probe + TMA issue latencies), more than the 256 tensor
    // cycles of a chunk, so in pair mode TWO lanes share the ring: warp 8 takes the even ring items, warp 10 the odd.
    if (lane == 0 && p.debug_noring != 1) {
      const uint32_t prod_par = warp == 8 ? 0u : 1u, prod_step = kPair ? 2u : 1u;
      uint32_t g_item = 0;   // ring items issued so far (by both producers)
      WorkList<kFused, kPair> work0(p, cta_group_idx * kSlots + 0, n_slots_total, (int)cta_rank);
      WorkList<kFused, kPair> work1(p, cta_group_idx * kSlots + (kSlots - 1), n_slots_total, (int)cta_rank);
      const int n_max = max(work0.n_items, kSlots > 1 ? work1.n_items : 0);
      // pair mode: this CTA streams its half of every chunk's rows; both CTAs' copies complete on the leader's w_full
      const uint32_t full_cluster0 = kPair ? map_to_cta(bar(kBarWFull), 0) : 0;
      const long long c_prod_begin = kProf ? clock64() : 0;
      const int n_steps = kNumBaseSteps + (kSec ? p.fl.n_sec_views : 0);
      RingState ring;
      ring.stage = kPair ? prod_par : 0;
      constexpr uint32_t kImg = kSplit3 ? 2 : 1;   // ring items per chunk (BF16X3: hi image, lo image)
      for (int it = 0; it < n_max; ++it) {
        for (int st = 0; st < n_steps; ++st) {
          for (int s = 0; s < kSlots; ++s) {
            const WorkList<kFused, kPair>& w = s == 0 ? work0 : work1;
            if (it >= w.n_items) continue;
            const StepDesc sd = step_desc(st);
            const uint32_t chunk_bytes = layer_chunk_bytes(sd.layer);
            const uint32_t n_items = (sd.n_chunks + (sd.bias_chunk ? 1 : 0)) * kImg;
            const uint32_t byte0 = tc_layer_offset_rt(sd.layer) * kImg + sd.first_chunk * chunk_bytes * kImg;
            if (kPair) {
              // rows of 64 bytes: consecutive chunk images, this CTA takes rows [rank * n/2, +n/2) of each
              const uint32_t rows = (uint32_t)layer_n(sd.layer);
              const uint32_t i0 = ((g_item & 1u) == prod_par) ? 0u : 1u;   // this lane's first item of the run
              if (n_items > i0)
                produce_chunks_pair(ring, &p.wmap[w.pass_of(it)][sd.layer >= 9 ? 1 : 0],
                                    byte0 / 64 + cta_rank * (rows / 2) + i0 * rows, 2 * rows, (n_items - i0 + 1) / 2,
                                    cta_rank == 0 ? (p.debug_noring == 3 ? chunk_bytes / 8 : chunk_bytes) : 0,
                                    bar(kBarWFull), full_cluster0, bar(kBarWEmpty), smem_u32(smem + kOffW), prod_step);
              g_item += n_items;
            } else {
              produce_chunks(ring, p.pass[w.pass_of(it)].packed + kSmallBytes + byte0, chunk_bytes, chunk_bytes, n_items,
                             bar(kBarWFull), bar(kBarWEmpty), smem_u32(smem + kOffW));
            }
          }
        }
      }
      if (kProf && p.prof != nullptr && blockIdx.x < 2 && warp == 8) {
        p.prof[40 + 4 * blockIdx.x] = ring.wait_cycles;
        p.prof[41 + 4 * blockIdx.x] = (unsigned long long)(clock64() - c_prod_begin);
        p.prof[42 + 4 * blockIdx.x] = 0;
      }
    }
  } else if (warp == 9 && (!kPair || cta_rank == 0)) {
    // =================================================================== MMA issuer (warp 9 of the leader CTA)
    // The whole warp runs the loop; one elected lane issues.  Per (tile, step, slot): wait a_ready, run the PTX chunk
    // loop over the slot's activation buffer from k-block 0 (the encoding sits there for M0 and for the encoding part
    // of M5), add the bias chunk, commit d_ready.  In pair mode every MMA is M=256: rows 0-127 from this CTA's
    // buffer and TMEM, rows 128-255 from the peer's (same shared-memory / TMEM addresses), B rows split between them.
    WorkList<kFused, kPair> work0(p, cta_group_idx * kSlots + 0, n_slots_total, (int)cta_rank);
    WorkList<kFused, kPair> work1(p, cta_group_idx * kSlots + (kSlots - 1), n_slots_total, (int)cta_rank);
    const int n_max = max(work0.n_items, kSlots > 1 ? work1.n_items : 0);
    const int n_steps = kNumBaseSteps + (kSec ? p.fl.n_sec_views : 0);
    RingState ring;
    uint32_t a_parity0 = 0, a_parity1 = 0, n_issued = 0;
    const uint64_t a_desc0 = make_desc(smem_u32(smem + kOffA));
    const uint64_t w_desc0 = make_desc_sw64(smem_u32(smem + kOffW));
    const uint64_t ones_desc = make_desc_ones(smem_u32(smem + kOffOnes));
    constexpr uint64_t kAUnits = kABytes >> 4;
    static_assert((kKBlockBytes >> 4) == 1024 && (kStageBytes >> 4) == (kPair ? 512 : 1024), "issue_chunks* assume these");
    const uint32_t bar_full0 = bar(kBarWFull), bar_empty0 = bar(kBarWEmpty);
    long long c_wait_a = 0;
    const long long c_begin = kProf ? clock64() : 0;
    for (int it = 0; it < n_max; ++it) {
      for (int st = 0; st < n_steps; ++st) {
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          const WorkList<kFused, kPair>& w = s == 0 ? work0 : work1;
          if (it >= w.n_items) continue;
          {
            const long long t0 = kProf ? clock64() : 0;
            mbar_wait(bar(kBarAReady + s), s == 0 ? a_parity0 : a_parity1);
            if (kProf) c_wait_a += clock64() - t0;
          }
          if (s == 0) a_parity0 ^= 1; else a_parity1 ^= 1;
          tc_fence_after();
          const StepDesc sd = step_desc(st);
          const uint32_t idesc = instr_desc(layer_n(sd.layer), kPair ? 256 : 128, kHalf);
          const uint32_t d_tmem = tmem_base + (uint32_t)(s * 256) + sd.d_col;
          const uint64_t a_hi = a_desc0 + (kSplit3 ? 0 : s) * kAUnits + sd.a_units, a_lo = a_desc0 + kAUnits + sd.a_units;
          if (!kSplit3 && !kPair) issue_chunks(ring, d_tmem, a_hi, w_desc0, bar_full0, bar_empty0, sd.n_chunks, sd.accumulate, idesc);
          else if (!kSplit3) issue_chunks_pair(ring, d_tmem, a_hi, w_desc0, bar_full0, bar_empty0, sd.n_chunks, sd.accumulate, idesc, p.debug_noring == 1);
          else if (!kPair) issue_chunks_split(ring, d_tmem, a_hi, a_lo, w_desc0, bar_full0, bar_empty0, sd.n_chunks, sd.accumulate, idesc);
          else issue_chunks_split_pair(ring, d_tmem, a_hi, a_lo, w_desc0, bar_full0, bar_empty0, sd.n_chunks, sd.accumulate, idesc);
          n_issued += sd.n_chunks;
          if (sd.bias_chunk) {
            // all-ones A operand x the bias chunk (second K=16 step of the chunk); BF16X3: hi and lo images
            for (int part = 0; part < (kSplit3 ? 2 : 1); ++part) {
              if (kPair) issue_bias_chunk_pair<!kSplit3>(ring, d_tmem, ones_desc, w_desc0, bar_full0, bar_empty0, idesc, p.debug_noring == 1);
              else issue_bias_chunk(ring, d_tmem, ones_desc, w_desc0, bar_full0, bar_empty0, idesc);
            }
            n_issued += 1;
          }
          if (kPair) umma_commit_elect_pair(bar(kBarDReady + s)); else umma_commit_elect(bar(kBarDReady + s));
        }
      }
    }
    tc_fence_before();
    if (kProf && p.prof != nullptr && blockIdx.x == 0 && lane == 0) {
      p.prof[32] = c_wait_a; p.prof[33] = ring.wait_cycles; p.prof[34] = (unsigned long long)(clock64() - c_begin);
      p.prof[35] = n_issued;
    }
  }
  if (kFused && warp == 11 && !(p.debug_noring & 4)) {
    // =================================================================== ray warp
    // Alpha compositing (volume_rendering, VipNeRF01.py:331-384) of the two rays a coarse tile / a fine tile triple
    // completes and, after the coarse pass, their hierarchical re-sampling (get_z_vals_fine :205-216) - one warp,
    // shuffle scans, asynchronous to the tiles the epilogue groups go on with.  Serves both tile slots, in the order
    // their events become due.
    const WorkList<kFused, kPair> work0(p, cta_group_idx * kSlots + 0, n_slots_total, (int)cta_rank);
    const WorkList<kFused, kPair> work1(p, cta_group_idx * kSlots + (kSlots - 1), n_slots_total, (int)cta_rank);
    float* scratch = reinterpret_cast<float*>(smem + kOffRayScratch);
    // Events of a slot, in order: one per coarse tile (= ray pair), then one per completed fine tile triple.  The two
    // slots' events are served in the order they BECOME DUE (non-blocking probes of both barriers): a slot must not
    // wait for its sibling's tile to finish before its own pair is composited.
    uint32_t parity[2] = {0, 0}, n_done[2] = {0, 0};
    const uint32_t n_events[2] = {(uint32_t)(work0.n_first_pass * (p.has_fine ? 2 : 1)),
                                  kSlots > 1 ? (uint32_t)(work1.n_first_pass * (p.has_fine ? 2 : 1)) : 0u};
    const long long t_start = clock64();
    int slot = 0;
    while (n_done[0] < n_events[0] || n_done[1] < n_events[1]) {
      slot ^= (kSlots > 1 ? 1 : 0);
      if (n_done[slot] >= n_events[slot]) continue;
      // warp-uniform probe: lane 0 decides, then every lane acquires the (already completed) phase itself
      uint32_t due = lane == 0 ? (uint32_t)mbar_try_wait(bar(kBarRayFull + slot), parity[slot]) : 0u;
      due = __shfl_sync(0xffffffffu, due, 0);
      if (!due) {
        if (clock64() - t_start > 8 * kTimeoutCycles) mbar_timeout(bar(kBarRayFull + slot), parity[slot]);
        continue;
      }
      mbar_wait(bar(kBarRayFull + slot), parity[slot]);
      parity[slot] ^= 1;
      const WorkList<kFused, kPair>& work = slot == 0 ? work0 : work1;
      const int ev = (int)n_done[slot];
      const int pi = ev < work.n_first_pass ? 0 : 1;
      const int64_t pair = work.unit_of(pi == 0 ? ev : ev - work.n_first_pass);
      const PassDesc& ps = p.pass[pi];
      for (int rr = 0; rr < 2; ++rr) {
        const int64_t r = 2 * pair + rr;
        if (r >= p.n_rays) continue;
        RayConsts rc;
        rc.dnorm = vec3_norm(p.rp.pts_d[3 * r], p.rp.pts_d[3 * r + 1], p.rp.pts_d[3 * r + 2]);
        rc.oz = p.rp.rays_o[3 * r + 2];
        rc.dz = p.rp.rays_d[3 * r + 2];
        if (pi == 0) {
          float z_reg[2], w_reg[2];
          const int V = kSec ? p.fl.n_sec_views : 0;
          composite_ray<2>(lane, 64, ps.z + r * 64, ps.sigma + r * 64, ps.rgb + r * 192,
                           V > 0 ? ps.vis2 + r * 64 * V : nullptr, V, p.fl.ndc, p.fl.white_bkgd, rc, p.out[0], r, z_reg, w_reg);
          if (p.has_fine) {
            const float* u = p.rp.u_rand ? p.rp.u_rand + r * p.n_fine : p.rp.u_vals;
            resample_ray<2>(lane, 64, p.n_fine, z_reg, w_reg, u, p.rp.u_rand == nullptr, scratch,
                            p.pass[1].z + r * (64 + p.n_fine));
          }
        } else {
          float z_reg[6], w_reg[6];
          const int V = kSec ? p.fl.n_sec_views : 0;
          composite_ray<6>(lane, 192, ps.z + r * 192, ps.sigma + r * 192, ps.rgb + r * 576,
                           V > 0 ? ps.vis2 + r * 192 * V : nullptr, V, p.fl.ndc, p.fl.white_bkgd, rc, p.out[1], r, z_reg, w_reg);
        }
      }
      ++n_done[slot];
      __threadfence_block();   // z_fine (global) before the count the epilogue group acquires
      __syncwarp();
      if (lane == 0)
        asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(smem + kOffRayDone) + 4u * slot), "r"(n_done[slot]) : "memory");
    }
  }
  __syncthreads();
  if (kPair) cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 9) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
std::mutex g_attr_mutex;
unsigned long long* g_prof_buffer = nullptr;  // debug: set by vipnerf_debug_set_profile_buffer
int g_sm_count[64] = {0};
bool g_attr_set[64][64] = {{false}};
// CTA-pair (cta_group::2) kernels unless VIPNERF_TC_CTA_PAIRS=0 (read once)
const bool g_use_cta_pairs = []() { const char* e = getenv("VIPNERF_TC_CTA_PAIRS"); return e == nullptr || e[0] != '0'; }();

cudaError_t device_sm_count(int* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  if (dev < 64 && g_sm_count[dev] > 0) { *out = g_sm_count[dev]; return cudaSuccess; }
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  if (dev < 64) g_sm_count[dev] = n;
  *out = n;
  return cudaSuccess;
}

// Tensor maps over a packed buffer's chunk-image stream (TcParams::wmap).  cuTensorMapEncodeTiled is a host-only
// encoder (no device work); it is fetched from the driver through the runtime so that the library does not link
// libcuda.
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}
cudaError_t encode_weight_maps(const uint8_t* packed, bool split3, CUtensorMap (&maps)[2], int box_div = 1) {
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {32, (cuuint64_t)kTcBigBytes * (split3 ? 2 : 1) / 64};
  const cuuint64_t strides[1] = {64};
  const cuuint32_t elem_strides[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    const cuuint32_t box[2] = {32, (i == 0 ? 128u : 64u) / box_div};
    const CUresult r = encode(&maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                              const_cast<uint8_t*>(packed) + kSmallBytes, dims, strides, box, elem_strides,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  return cudaSuccess;
}

template <bool kSplit3, bool kFused, bool kProf, bool kPair, bool kSec, bool kHalf = false>
cudaError_t launch_variant(const TcParams& p, int64_t n_units, cudaStream_t s) {
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if ((e = device_sm_count(&sms)) != cudaSuccess) return e;
  auto kernel = k_render_tc<kSplit3, kFused, kProf, kPair, kSec, kHalf>;
  {
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    const int variant = (kSplit3 ? 2 : 0) + (kFused ? 1 : 0) + (kProf ? 4 : 0) + (kPair ? 8 : 0) + (kSec ? 16 : 0) + (kHalf ? 32 : 0);
    if (dev >= 64 || !g_attr_set[dev][variant]) {
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
      if (e != cudaSuccess) return e;
      if (dev < 64) g_attr_set[dev][variant] = true;
    }
  }
  constexpr int kSlots = kSplit3 ? 1 : 2;
  constexpr int kCtasPerGroup = kPair ? 2 : 1;
  // one slot (per CTA group) per work unit at most; a CTA pair serves two units per slot
  int64_t groups = (n_units + kSlots * kCtasPerGroup - 1) / (kSlots * kCtasPerGroup);
  if (groups > sms / kCtasPerGroup) groups = sms / kCtasPerGroup;
  if (groups < 1) return cudaSuccess;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * kCtasPerGroup));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtasPerGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = kPair ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <bool kSplit3, bool kFused, bool kHalf = false>
cudaError_t launch(TcParams& p, int64_t n_units, cudaStream_t s) {
  const bool pair = g_use_cta_pairs || kHalf;   // fp16 operands: CTA-pair kernels only
  { const char* e = getenv("VIPNERF_TC_DEBUG_NORING"); p.debug_noring = (e != nullptr && pair && !kSplit3) ? atoi(e) : 0; }
  if (pair) {
    for (int pi = 0; pi < 2; ++pi) {
      const cudaError_t e = encode_weight_maps(p.pass[pi].packed, kSplit3, p.wmap[pi], p.debug_noring == 3 ? 8 : 1);
      if (e != cudaSuccess) return e;
    }
  }
  if (kHalf) {   // fp16 operands: CTA pairs, no profiling variant
    return p.fl.n_sec_views > 0 ? launch_variant<false, kFused, false, true, true, true>(p, n_units, s)
                                : launch_variant<false, kFused, false, true, false, true>(p, n_units, s);
  }
  if (p.fl.n_sec_views > 0) {   // secondary views: separate instantiations (no profiling variant)
    return pair ? launch_variant<kSplit3, kFused, false, true, true>(p, n_units, s)
                : launch_variant<kSplit3, kFused, false, false, true>(p, n_units, s);
  }
  if (p.prof != nullptr) {
    return pair ? launch_variant<kSplit3, kFused, true, true, false>(p, n_units, s)
                : launch_variant<kSplit3, kFused, true, false, false>(p, n_units, s);
  }
  return pair ? launch_variant<kSplit3, kFused, false, true, false>(p, n_units, s)
              : launch_variant<kSplit3, kFused, false, false, false>(p, n_units, s);
}

}  // namespace

void set_tc_profile_buffer(void* dev_ptr) {
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  g_prof_buffer = static_cast<unsigned long long*>(dev_ptr);
}

cudaError_t launch_mlp_tc(int precision, const RayPtrs& rp, const RenderFlags& fl, int64_t n_rays, int S,
                          const float* z, const void* packed, float* sigma, float* rgb, float* vis, float* vis2,
                          cudaStream_t s) {
  TcParams p{};
  p.rp = rp;
  p.fl = fl;
  p.n_rays = n_rays;
  p.pass[0].z = const_cast<float*>(z);
  p.pass[0].packed = static_cast<const uint8_t*>(packed);
  p.pass[0].sigma = sigma;
  p.pass[0].rgb = rgb;
  p.pass[0].vis = vis;
  p.pass[0].vis2 = vis2;
  if (vis2 == nullptr) p.fl.n_sec_views = 0;
  p.pass[0].n_points = n_rays * S;
  p.pass[0].S = S;
  p.pass[0].compute_z = 0;
  p.pass[1] = p.pass[0];
  p.n_units = (p.pass[0].n_points + kTile - 1) / kTile;
  p.prof = g_prof_buffer;
  if (p.n_units == 0) return cudaSuccess;
  if (precision == VIPNERF_PRECISION_BF16X3) return launch<true, false>(p, p.n_units, s);
  if (precision == VIPNERF_PRECISION_FP16) return launch<false, false, true>(p, p.n_units, s);
  return launch<false, false>(p, p.n_units, s);
}

cudaError_t launch_render_fused_tc(int precision, const FusedArgs& a, cudaStream_t s) {
  if (a.n_rays == 0) return cudaSuccess;
  TcParams p{};
  p.rp = a.rp;
  p.fl = a.fl;
  p.n_rays = a.n_rays;
  p.n_fine = a.n_fine;
  p.has_fine = a.n_fine > 0;
  const int Sf = a.n_coarse + a.n_fine;
  for (int pi = 0; pi < 2; ++pi) {
    const PassOutPtrs& o = pi ? a.out_fine : a.out_coarse;
    PassDesc& d = p.pass[pi];
    d.S = pi ? Sf : a.n_coarse;
    d.n_points = a.n_rays * d.S;
    d.packed = static_cast<const uint8_t*>(pi ? a.packed_fine : a.packed_coarse);
    d.z = pi ? (o.z_vals ? o.z_vals : a.ws_z_fine) : (o.z_vals ? o.z_vals : a.ws_z_coarse);
    // each pass has its own workspace region: the ray warps read a pass's outputs while the next tiles are written
    d.sigma = o.raw_sigma ? o.raw_sigma : (pi ? a.ws_sigma : a.ws_sigma_c);
    d.rgb = o.raw_rgb ? o.raw_rgb : (pi ? a.ws_rgb : a.ws_rgb_c);
    d.vis = o.raw_visibility ? o.raw_visibility : (pi ? a.ws_vis : a.ws_vis_c);
    d.vis2 = a.fl.n_sec_views > 0 ? (o.raw_visibility2 ? o.raw_visibility2 : (pi ? a.ws_vis2 : a.ws_vis2_c)) : nullptr;
    d.compute_z = pi == 0;
    p.out[pi] = o;
    p.out[pi].z_vals = nullptr;  // depths are produced in place
  }
  if (!p.has_fine) p.pass[1] = p.pass[0];
  p.n_units = (a.n_rays + 1) / 2;
  p.prof = g_prof_buffer;
  if (precision == VIPNERF_PRECISION_BF16X3) return launch<true, true>(p, p.n_units, s);
  if (precision == VIPNERF_PRECISION_FP16) return launch<false, true, true>(p, p.n_units, s);
  return launch<false, true>(p, p.n_units, s);
}

}  // namespace vipnerf
