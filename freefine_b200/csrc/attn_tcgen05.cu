// (a) masked, KV-injected flash attention for sm_100a: tcgen05.mma + TMEM accumulators + TMA operand staging.
//
// Replaces the baddbmm -> softmax -> bmm triples (x3) + mask builders + blends of
// Attention_Modulator.Temporal_contextal_attention (src/utils/attention.py:1043-1091) and its _bg / _compose /
// style-align / plain variants; see include/freefine_b200.h for the per-(stream,head) plan semantics.
//
// One CTA = one 128-row query tile of one (stream, head).  Ten warps:
//   warps 0-7  softmax + epilogue as TWO warpgroups: warpgroup wg owns the 64-key K/V tiles of parity wg -- its own S/P
//              buffer, its own O accumulator in TMEM, its own running reference point.  Thread (w, lane) owns query row
//              32*(w%4)+lane == its TMEM lane (tcgen05.ld/st 32x32b).  Four softmax warps per SM sub-partition (two
//              CTAs per SM) keep the MUFU unit -- the bound of the d=40 layers -- fed while others wait on barriers /
//              TMEM.  The two partial softmaxes of a pass are merged at its end (one exchange of the reference points
//              through shared memory), the cross-pass accumulator sum_p w_p*roww_p(q)*O_p/l_p lives in shared memory.
//   warp  8    TMA producer (one elected lane): Q once, then a ring of K/V tiles (64 keys x 64 channels boxes,
//              SWIZZLE_128B, channels beyond head_dim zero-filled by the TMA bounds check), + TMEM alloc/dealloc
//   warp  9    MMA issuer (warp-uniform loop, one elected lane issues):
//                  S_par = Q K^T  (SS, M128 N64 K16 x DPAD/16, both operands K-major)
//                  O_par += P V   (TS: P from TMEM over the S columns it came from, V MN-major from the box)
// Per K/V tile t:  QK^T -> [s_full] -> softmax (row max over the allowed keys, 37% of the exp2 on the FMA pipe as a
// degree-3 polynomial, P packed for the tensor core and written over S) -> [p_full] -> PV(t) and QK(t+2) back to back
// -> [kv_empty, s_full].  The running reference point is only refreshed when the row max grows by more than 2^8 / 2^24
// (lazy rescale: O stays in TMEM, read-modify-written only then).  V is ALWAYS the staging of ff_kv_gather_cast, whose
// ones column makes P.V also produce the softmax denominator: no row-sum arithmetic in the softmax warps, numerator
// and denominator see the same rounded P.  Region masks are bit-vectors; `allowed(q,k)` is evaluated in registers on
// 32-bit words, nothing [S,S]-shaped exists anywhere.
#include <cuda.h>
#include <cuda_bf16.h>

#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "ff_common.cuh"

namespace {

constexpr int BM = 128;                 // query rows per CTA
constexpr int BN = 64;                  // keys per tile (S is double-buffered in TMEM: 2 x 64 columns)
constexpr int BOX_COLS = 64;            // channels per TMA box (128 bytes of bf16 = one swizzle-128B row)
constexpr int TILE_BYTES = BM * 128;    // Q box: 128 rows x 128 B = 16 KiB
constexpr int KV_BYTES = BN * 128;      // K / V box: 64 rows x 128 B = 8 KiB
constexpr int NUM_SOFTMAX_WARPS = 8;    // two warpgroups: warps 0-3 take the even K/V tiles, warps 4-7 the odd ones
constexpr int NUM_THREADS = 32 * (NUM_SOFTMAX_WARPS + 2);
// P operand of the PV contraction (template parameter HILO), fixed by the dtype of the staged V -- tcgen05.mma
// kind::f16 wants A and B in the SAME 16-bit format (an f16 A with a bf16 B raises an illegal-instruction fault):
//   V fp16  (HILO=false)  ONE fp16 P operand: 11 significant bits, half the PV tensor work and a third of the packing
//                         ALU work of
//   V bf16  (HILO=true)   a hi + lo pair of bf16 P operands (16 significant bits, two TS-MMAs per 16 keys).
// Lazy-rescale threshold in log2 units: p <= 2^24 keeps fp32 / bf16 relative precision; p <= 2^8 stays far inside
// the fp16 range.  Rescales are rare either way.
template <bool HILO> constexpr float rescale_threshold() { return HILO ? 24.f : 8.f; }
// exp2 of pair i (mod 8) of every 16-score chunk goes to the FMA pipe (degree-3 polynomial) instead of the MUFU unit
// when bit i of this pattern is set: the softmax of the d=40 layers is bound by the 16 ex2/clk/SM of the MUFU unit.
// 0x92 = 3 pairs of 8 (37.5%): measured best on the S=4096 d=40 launch (2.49 -> 2.37 ms; 25% and 50% are slower).
#ifndef FF_RING_DEFAULT
#define FF_RING_DEFAULT 6
#endif
#ifndef FF_POLY_PATTERN
#define FF_POLY_PATTERN 0x92
#endif
#ifndef FF_T32_DEFAULT
#define FF_T32_DEFAULT 0
#endif
#ifndef FF_SEP_P
#define FF_SEP_P 0   // experiment switch (parity-green, slower: 2.55 vs 2.35 ms -- profiles/r2b_attn_experiments.txt)
#endif
#ifndef FF_SELF_ISSUE
#define FF_SELF_ISSUE 0   // experiment switch (measured slower: 2.84 vs 2.35 ms, profiles/r2b_attn_experiments.txt)
#endif

// DPAD: head_dim padded to the K-step of Q K^T (a multiple of 16).  DPV: channels per head of the staged V = columns
// of the V tile / of O: head_dim real channels, a column of ONES at channel head_dim (its P.V column is the softmax
// denominator l = sum_k P[q,k], from exactly the rounded P that multiplies V), zeros above.
template <int DPAD, bool HILO> struct Cfg {
  static constexpr int DPV = DPAD == 16 ? 16 : (DPAD == 48 ? 48 : DPAD + 16);
  static constexpr int NKT = (DPAD + BOX_COLS - 1) / BOX_COLS;          // 64-channel boxes per operand tile
  static_assert((DPV + BOX_COLS - 1) / BOX_COLS == NKT, "V tile must span as many boxes as the K tile");
  static constexpr int NSTAGE = DPAD <= 48 ? 4 : (DPAD <= 80 ? 3 : 2);  // K/V ring depth (what fits beside the accumulator)
  static_assert(NSTAGE >= 2, "QK(t+1) is issued before PV(t): tile t+1 must fit beside tile t");
  // TMEM: S buffer 0 / 1 (= tile parity), then one O accumulator PER PARITY (each warpgroup runs its own online
  // softmax over its tiles; the two are merged at the end of a pass)
  // SEP (d <= 40, fp16 P, -DFF_SEP_P): the packed P tile (32 columns) gets its OWN TMEM region, shared by the two
  // warpgroups in turn, instead of overwriting the S columns it came from.  The S buffer of a tile is then free as soon as
  // its scores have been read, QK(t+2) is issued UNDER the exp sweep of tile t, and the warpgroup's chain no longer
  // contains  p_full -> PV(t), QK(t+2) -> s_full  (about a third of its period in the round-2 timelines).
  static constexpr bool SEP = FF_SEP_P && DPAD == 48 && !HILO && NSTAGE == 4;
  static constexpr int TMEM_S = 0, TMEM_P = 2 * BN, TMEM_O = 2 * BN + (SEP ? BN / 2 : 0);   // O of parity b at TMEM_O + b * DPV
  static constexpr int TMEM_USED = TMEM_O + 2 * DPV;
  static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;
  static constexpr int SMEM_Q = NKT * TILE_BYTES;
  static constexpr int SMEM_STAGE = 2 * NKT * KV_BYTES;                 // K tiles then V tiles
  // cross-pass accumulator sum_p w_p*roww_p*O_p/l_p: fp32 [128 rows][ACC_LD] in shared memory (TMEM is full)
  static constexpr int ACC_LD = DPAD <= 80 ? DPAD + 4 : DPAD;          // (padding against bank conflicts where it fits)
  static constexpr int SMEM_ACC = BM * ACC_LD * 4;
  static constexpr int SMEM_MX = 2 * BM * 4;                            // running reference points of the two warpgroups
  static constexpr int SMEM_BYTES = SMEM_Q + NSTAGE * SMEM_STAGE + SMEM_ACC + SMEM_MX + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static constexpr int MIN_CTAS = (SMEM_BYTES <= 113 * 1024 && TMEM_COLS == 256) ? 2 : 1;
  // SELF (d <= 40, fp16 P): there is no MMA-issuer warp.  When a softmax warpgroup has stored its P tile it meets on a
  // named barrier and its first warp issues PV(t) and QK(t+2) itself -- the p_full mbarrier round trip (arrive -> the
  // sleeping issuer warp resumes: ~330 cycles in the round-2 timelines, on the critical chain of every tile) is replaced
  // by a bar.sync among four warps, and nine warps instead of ten leave 112 instead of 96 registers per thread.
  static constexpr bool SELF = FF_SELF_ISSUE && !SEP && DPAD == 48 && !HILO && NSTAGE == 4;
  static constexpr int THREADS = SELF ? 32 * (NUM_SOFTMAX_WARPS + 1) : NUM_THREADS;
};

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
#ifdef FF_TRYWAIT_NOHINT
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
#else
  // the suspend-time hint lets the thread sleep in hardware until the phase completes instead of spinning
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
#endif
  return done;
}
// non-blocking probe of a phase (used to hide the ~90-cycle fast-path latency of try_wait behind useful work)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Slow path of every wait.  A waiting warp is woken many times before its phase completes (ncu source counters, round 2:
// ~7 wake-ups per wait -- NANOSLEEP.SYNCS returns on every event of the barrier), and with a clock64 time-out check in
// the loop each wake-up cost 12 issue slots: 22 % of ALL instructions the d = 40 launch executed were this loop, taken
// from the softmax warps that share the schedulers.  The loop is now try_wait + a counter; the loud time-out (a protocol
// bug must fail, not hang the GPU) is counted in wake-ups instead of cycles.
#ifndef FF_WAIT_FORM
#define FF_WAIT_FORM 1   // 0: clock64 time-out (round 1), 1: counted, try_wait with suspend hint, 2: counted, plain try_wait
#endif
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
#if FF_WAIT_FORM == 0
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug must fail loudly, not hang the GPU
      printf("ff_attn: mbarrier wait timed out (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
#else
  const long long t0 = clock64();
#pragma unroll 1
  for (uint32_t n = 1;; ++n) {
    if ((n & 15u) == 0 && clock64() - t0 > 4000000000LL) break;    // ~2 s (a sleeping try_wait returns after ~10 ms)
#if FF_WAIT_FORM == 1
    if (mbar_try(bar, parity)) return;
#else
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
#endif
  }
  printf("ff_attn: mbarrier wait timed out (block %d,%d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y,
         blockIdx.z, threadIdx.x, bar, parity);
  __trap();
#endif
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
// Optional busy-polling wait per role (FF_SPIN_MASK bit 0: softmax warps, bit 1: MMA issuer, bit 2: TMA producer);
// measured within noise of the hardware-sleep wait once the MMA issuer loop became warp-uniform, so off by default.
// Bounded: a protocol bug falls through to the loud time-out of mbar_wait_slow.
#ifndef FF_SPIN_MASK
#define FF_SPIN_MASK 0
#endif
template <int ROLE_BIT>
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
  if constexpr ((FF_SPIN_MASK >> ROLE_BIT) & 1) {
    for (uint32_t n = 0; !mbar_test(bar, parity); ++n) {
#ifdef FF_SPIN_SLEEP
      __nanosleep(FF_SPIN_SLEEP);
#endif
      if (n > (1u << 24)) { mbar_wait_slow(bar, parity); return; }
    }
  } else {
    mbar_wait(bar, parity);
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// Named barriers 1, 2 (0 is __syncthreads): all-softmax-warps synchronisation at the end of a pass.
// (immediate ids: with a register id ptxas reserves all 16 hardware barriers for the CTA)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  switch (id) {
    case 1: asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory"); break;
    case 2: asm volatile("bar.sync 2, %0;" ::"r"(count) : "memory"); break;
    case 3: asm volatile("bar.sync 3, %0;" ::"r"(count) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"r"(count) : "memory"); break;
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// one column (the softmax denominator produced by the ones column of V); waits for it
__device__ __forceinline__ float tmem_ld1_wait(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;"
               : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.ld writes its destination registers ASYNCHRONOUSLY; they are valid only after tcgen05.wait::ld.  To the
// compiler an inline-asm output is defined at the asm statement, so arithmetic on the loaded values could legally be
// scheduled ABOVE a separate wait statement.  These wait variants take the loaded registers as in/out operands, which
// pins every use behind the wait.
__device__ __forceinline__ void tmem_wait_ld16(float (&v)[16]) {
#ifdef FF_PLAIN_WAIT
  tmem_wait_ld();
  return;
#endif
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld32(float (&v)[32]) {
#ifdef FF_PLAIN_WAIT
  tmem_wait_ld();
  return;
#endif
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) {
#ifdef FF_KO_MUFU   // timing experiment only (wrong results): how much of the tile time is the MUFU unit?
  return x;
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}

// Shared-memory matrix descriptor, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor layout:
// start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type=2 (128B swizzle) [61,64)).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor kind::f16: D f32, A/B bf16, A K-major, B K-major (b_mn=0) or MN-major (b_mn=1), M=128.
// f16: A and B are fp16 (format 0) instead of bf16 (format 1) -- the two formats cannot be mixed in one MMA.
__host__ __device__ constexpr uint32_t make_idesc(int n, int b_mn, bool f16 = false) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

// Optional progress trace for debugging protocol stalls (ff_debug_set_trace): a host-mapped buffer the kernel writes
// (cta, role) -> (tile counter, site) into; null in production (one predictable branch per barrier wait).
__device__ uint32_t* g_trace = nullptr;
__device__ __forceinline__ void trace(uint32_t* tr, int role, uint32_t it, uint32_t site) {
  if (tr) {
    const uint32_t cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    volatile uint32_t* t = tr + (size_t)cta * 8 + role * 2;
    t[0] = it;
    t[1] = site;
    __threadfence_system();
  }
}

// Timeline probe (build with -DFF_TIMELINE, ff_debug_set_timeline): clock64 stamps of two CTAs, per role and tile.
#ifdef FF_TIMELINE
__device__ unsigned long long* g_timeline = nullptr;
#define FF_TL_TILES 64
#define FF_TL(role_, tile_, site_)                                                                      \
  do {                                                                                                  \
    if (tl && (tile_) < FF_TL_TILES) tl[(((size_t)tl_cta * 2 + (role_)) * FF_TL_TILES + (tile_)) * 8 + (site_)] = clock64(); \
  } while (0)
#else
#define FF_TL(role_, tile_, site_) do { } while (0)
#endif

struct KParams {
  const FFAttnHeadPlan* plan;
  const uint32_t* bitmasks;
  const int32_t* popc;
  void* out;
  int heads, head_dim, s_q, s_kv, mask_words, out_dtype, n_kv_streams;
  int n_streams, n_qtiles;   // work items of the persistent ring kernel: n_qtiles * heads * n_streams
  float scale_log2;   // scale * log2(e)
};

// ---------------------------------------------------------------------------------------------------------------
// Per-pass / per-tile decisions.  EVERY role (TMA producer, MMA issuer, softmax warps) evaluates these identically,
// so that skipped K/V tiles are skipped by all of them without any communication.
//
//   keybit(k)   = bit k of the pass's key mask (bit-vector mode)  or  k < T (PREFIX mode: the caller sorted the keys
//                 of that stream so that the T = popcount set keys come first -- softmax is permutation invariant)
//   rowflip(q)  = ROW_XOR & rowbit(q)
//   allowed(q,k)= keybit(k) ^ KEY_INVERT ^ rowflip(q)
// A 128-key tile is IN (all keybits 1), OUT (all 0) or MIX; a ragged last tile is always MIX.  For IN/OUT tiles the
// mask degenerates to ONE predicate per query row (no per-element work at all); the tile is skipped entirely when
// that predicate is false for every row of the query tile.  Only MIX tiles evaluate bits per element.
// ---------------------------------------------------------------------------------------------------------------
enum { TILE_ALL = 0, TILE_IN = 1, TILE_OUT = 2, TILE_MIX = 3 };

struct SegCtx {
  int kv;           // K/V stream of this segment (-1: none)
  int kmask;        // key mask id (-1: every key)
  int T;            // popcount of the key mask
  bool kinv, prefix;
};
struct PassCtx {
  SegCtx s0, s1;
  bool rowxor, active, two_seg;
  bool has_rf0, has_rf1;   // the query tile has (valid) rows with rowflip 0 / 1
};

__device__ __forceinline__ PassCtx make_ctx(const FFAttnPass& ps, const KParams& p, int q0) {
  PassCtx c;
  c.s0.kv = ps.kv_stream;
  c.s1.kv = ps.kv_stream2;
  c.two_seg = ps.kv_stream2 >= 0;
  c.s0.kmask = ps.key_mask;
  c.s1.kmask = ps.key_mask2;
  c.s0.kinv = (ps.flags & FF_PASS_KEY_INVERT) != 0;
  c.s1.kinv = (ps.flags & FF_PASS_KEY2_INVERT) != 0;
  c.s0.prefix = (ps.flags & FF_PASS_KEY_PREFIX) != 0;
  c.s1.prefix = (ps.flags & FF_PASS_KEY2_PREFIX) != 0;
  c.s0.T = ps.key_mask >= 0 ? __ldg(p.popc + ps.key_mask) : 0;
  c.s1.T = ps.key_mask2 >= 0 ? __ldg(p.popc + ps.key_mask2) : 0;
  c.rowxor = (ps.flags & FF_PASS_ROW_XOR) != 0 && ps.row_mask >= 0;
  c.active = true;
  c.has_rf0 = true;
  c.has_rf1 = false;
  if (ps.row_mask >= 0 && (c.rowxor || (ps.flags & FF_PASS_ROW_WEIGHT))) {
    uint32_t any1 = 0, any0 = 0;
    for (int w = 0; w < BM / 32; ++w) {
      const int base = q0 + 32 * w;
      if (base >= p.s_q) break;
      const uint32_t bits = __ldg(p.bitmasks + (size_t)ps.row_mask * p.mask_words + (base >> 5));
      const int rem = p.s_q - base;
      const uint32_t valid = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
      any1 |= bits & valid;
      any0 |= ~bits & valid;
    }
    if (ps.flags & FF_PASS_ROW_WEIGHT) c.active = any1 != 0;   // roww(q) = rowbit(q): nothing to add for this tile
    if (c.rowxor) {
      c.has_rf0 = any0 != 0;
      c.has_rf1 = any1 != 0;
    }
  }
  return c;
}

// quirk Q4: is the allowed set of rows with flip value f EMPTY (=> those rows attend uniformly to every key)?
__device__ __forceinline__ bool uniform_for(const PassCtx& c, const SegCtx& g, bool f, int s_kv) {
  if (c.two_seg || g.kmask < 0) return false;
  return (f ? s_kv - g.T : g.T) == 0;
}

__device__ __forceinline__ int tile_class(const SegCtx& g, int j, const KParams& p) {
  const int lo = j * BN;
  if (lo + BN > p.s_kv) return TILE_MIX;                    // ragged tile: invalid columns need per-element masking
  if (g.kmask < 0) return TILE_ALL;
  if (g.prefix) return lo + BN <= g.T ? TILE_IN : (lo >= g.T ? TILE_OUT : TILE_MIX);
  uint32_t a = 0xffffffffu, o = 0;
  for (int w = 0; w < BN / 32; ++w) {
    const uint32_t bits = __ldg(p.bitmasks + (size_t)g.kmask * p.mask_words + (lo >> 5) + w);
    a &= bits;
    o |= bits;
  }
  return a == 0xffffffffu ? TILE_IN : (o == 0 ? TILE_OUT : TILE_MIX);
}

// Maximal run of K/V tiles [j, hi) that share the class of tile j.  With sorted keys (PREFIX) or no key mask a
// segment decomposes into at most four runs (IN*, MIX?, OUT*, ragged?), so every per-tile decision -- class, skip,
// the per-row predicate -- is taken once per run instead of once per tile; in bit-vector mode a run is one tile.
__device__ __forceinline__ int run_end(const SegCtx& g, int j, int cls, const KParams& p) {
  if (cls == TILE_MIX || (g.kmask >= 0 && !g.prefix)) return j + 1;
  const int n_full = p.s_kv / BN;                                      // tiles without ragged columns
  if (cls == TILE_IN) { const int t = g.T / BN; return t < n_full ? t : n_full; }
  return n_full;                                                       // ALL or OUT: up to the ragged tile
}

// one predicate per row for IN/OUT/ALL tiles
__device__ __forceinline__ bool row_allowed(int cls, bool flip, bool uniform) {
  return uniform || cls == TILE_ALL || ((cls == TILE_IN) != flip);
}

__device__ __forceinline__ bool tile_skip(const PassCtx& c, const SegCtx& g, int cls, int s_kv) {
  if (cls == TILE_MIX || cls == TILE_ALL) return false;
  bool need = false;
  if (c.has_rf0) need = need || row_allowed(cls, g.kinv, uniform_for(c, g, g.kinv, s_kv));
  if (c.has_rf1) need = need || row_allowed(cls, !g.kinv, uniform_for(c, g, !g.kinv, s_kv));
  return !need;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 b = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&b);
}

// 16 scores -> p = exp2(s*sc + nb) as hi/lo bf16 pairs: hl[0..8) = hi (truncated upper halves, key 2i in the low
// half), hl[8..16) = lo = bf16(p - hi).  MASKED: bit i of `bits` gates key i.
template <bool MASKED>
__device__ __forceinline__ void softmax_chunk_hilo(const float* s, uint32_t* hl, float sc, float nb, uint32_t bits) {
  const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(nb, nb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), sc2, nb2);
    if (MASKED) {
      x.x = (bits >> (2 * i)) & 1u ? x.x : -INFINITY;
      x.y = (bits >> (2 * i + 1)) & 1u ? x.y : -INFINITY;
    }
    const float2 e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
    const uint32_t b0 = __float_as_uint(e.x), b1 = __float_as_uint(e.y);
    hl[i] = __byte_perm(b0, b1, 0x7632);
    const float2 r = __fadd2_rn(e, make_float2(-__uint_as_float(b0 & 0xffff0000u), -__uint_as_float(b1 & 0xffff0000u)));
    hl[8 + i] = pack_bf16x2(r.x, r.y);
  }
}

// 2^x for two values on the FMA pipe: x = n + f, n = round(x), f in [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial
// (max relative error 7.6e-5, below the half-ulp 2.4e-4 of the fp16 P operand it feeds), 2^n by an integer add into the
// exponent field.  x is clamped at -125 (result 2^-125: rounds to 0 in fp16, adds nothing measurable to the row sum).
__device__ __forceinline__ float2 poly_exp2_x2(float2 x) {
  const float MAGIC = 12582912.f;   // 1.5 * 2^23: the low mantissa bits of x + MAGIC hold round(x)
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(MAGIC, MAGIC));
  const float2 n = __fadd2_rn(t, make_float2(-MAGIC, -MAGIC));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 r = __ffma2_rn(make_float2(0.055205512791872025f, 0.055205512791872025f), f,
                        make_float2(0.24261389672756195f, 0.24261389672756195f));
  r = __ffma2_rn(r, f, make_float2(0.6932547688484192f, 0.6932547688484192f));
  r = __ffma2_rn(r, f, make_float2(0.9999276995658875f, 0.9999276995658875f));
  r.x = __int_as_float(__float_as_int(r.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(r.y) + (__float_as_int(t.y) << 23));
  return r;
}

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// 16 scores -> p = exp2(s*sc + nb), P as 8 packed fp16 pairs (key 2i in the low half; the row sum comes out of the
// tensor core: ones column of V).  MASKED: bit i of
// `bits` gates key i (MUFU only: ex2(-inf) is an exact 0).
template <bool MASKED>
__device__ __forceinline__ void softmax_chunk_f16(const float* s, uint32_t* pk, float sc, float nb, uint32_t bits) {
  const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(nb, nb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), sc2, nb2);
    float2 e;
    if (MASKED) {
      x.x = (bits >> (2 * i)) & 1u ? x.x : -INFINITY;
      x.y = (bits >> (2 * i + 1)) & 1u ? x.y : -INFINITY;
      e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
#ifdef FF_KO_EXP     // timing experiment only (wrong results): no exp2 at all, neither MUFU nor polynomial
    } else if (true) {
      e = x;
#endif
    } else if ((FF_POLY_PATTERN >> i) & 1) {
      e = poly_exp2_x2(x);
    } else {
      e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
    }
    pk[i] = pack_f16x2(e.x, e.y);
  }
}


// ---- lean issue blocks of the d <= 40 fp16-P issuer warp (FF_LEAN_ISSUER): ONE asm statement per tile, executed by the
// whole warp (the election happens inside); the per-K-step descriptor / TMEM-address arithmetic is done inside on PTX
// registers.  (With `if (leader) { mma; ...; commit; }` in C++ ptxas keeps every precomputed operand in vector registers
// and moves it to a uniform register with R2UR per tile: ~120 mostly serial instructions = ~500 cycles per tile, all of
// them inside the p_full -> s_full(t+2) window during which the owning softmax warpgroup idles.)
#ifndef FF_LEAN_ISSUER
#define FF_LEAN_ISSUER 1
#endif
__device__ __forceinline__ void lean_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
#ifdef FF_LEAN_SPIN   // experiment: poll instead of the hardware-sleep wait (wake-up latency vs issue slots)
  for (int n = 0; !done && n < FF_LEAN_SPIN; ++n) done = mbar_test(bar, parity);
#endif
  if (!done) mbar_wait_slow(bar, parity);
}
// S = Q K^T for DPAD = 48: three K-steps of 16 channels (32 B inside the 128-B row = 2 descriptor units), then the commit
__device__ __forceinline__ void lean_qk48(uint32_t sbuf, uint64_t qdesc, uint64_t kdesc, uint32_t idesc, uint32_t bar_s_) {
  asm volatile(
      "{\n\t.reg .pred e, pf, pt;\n\t.reg .b64 q1, q2, k1, k2;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, 0, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "add.u64 q1, %1, 2;\n\tadd.u64 q2, %1, 4;\n\tadd.u64 k1, %2, 2;\n\tadd.u64 k2, %2, 4;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q1, k1, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q2, k2, %3, pt;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n\t}"
      ::"r"(sbuf), "l"(qdesc), "l"(kdesc), "r"(idesc), "r"(bar_s_)
      : "memory");
}
// O (+)= P V: four K-steps of 16 keys (A = packed fp16 P in TMEM at columns +0, +8, +32, +40 of the S/P buffer, B = V tile
// MN-major, 2048 B = 128 descriptor units per K-step), commit kv_empty; then -- WITH_QK -- S = Q K^T of the same buffer's
// next tile right behind it in the pipe and its s_full commit, or -- last tiles of a parity -- a virtual s_full commit.
// timing experiments only (wrong results): -DFF_KO_PV1 / -DFF_KO_QK1 issue only the first K-step of P.V / Q.K^T
#ifdef FF_KO_PV1
#define FF_KO_PVSTEPS(x)
#else
#define FF_KO_PVSTEPS(x) x
#endif
#ifdef FF_KO_QK1
#define FF_KO_QKSTEPS(x)
#else
#define FF_KO_QKSTEPS(x) x
#endif
template <bool WITH_QK>
__device__ __forceinline__ void lean_pv_qk48(uint32_t obuf, uint32_t sp, uint64_t vdesc, uint32_t idesc_pv, uint32_t acc0,
                                             uint32_t bar_kve, uint64_t qdesc, uint64_t kdesc, uint32_t idesc_qk,
                                             uint32_t bar_s_) {
  if constexpr (WITH_QK) {
    asm volatile(
        "{\n\t.reg .pred e, p0, pf, pt;\n\t.reg .b64 v1, v2, v3, q1, q2, k1, k2;\n\t.reg .b32 a1, a2, a3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\tsetp.ne.b32 pf, 0, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 v1, %2, 128;\n\tadd.u64 v2, %2, 256;\n\tadd.u64 v3, %2, 384;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 32;\n\tadd.u32 a3, %1, 40;\n\t"
        "add.u64 q1, %6, 2;\n\tadd.u64 q2, %6, 4;\n\tadd.u64 k1, %7, 2;\n\tadd.u64 k2, %7, 4;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
        FF_KO_PVSTEPS("@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], v1, %3, pt;\n\t")
        FF_KO_PVSTEPS("@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], v2, %3, pt;\n\t")
        FF_KO_PVSTEPS("@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], v3, %3, pt;\n\t")
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %6, %7, %8, pf;\n\t"
        FF_KO_QKSTEPS("@e tcgen05.mma.cta_group::1.kind::f16 [%1], q1, k1, %8, pt;\n\t")
        FF_KO_QKSTEPS("@e tcgen05.mma.cta_group::1.kind::f16 [%1], q2, k2, %8, pt;\n\t")
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t}"
        ::"r"(obuf), "r"(sp), "l"(vdesc), "r"(idesc_pv), "r"(acc0), "r"(bar_kve), "l"(qdesc), "l"(kdesc), "r"(idesc_qk),
          "r"(bar_s_)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred e, p0, pt;\n\t.reg .b64 v1, v2, v3;\n\t.reg .b32 a1, a2, a3;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p0, %4, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 v1, %2, 128;\n\tadd.u64 v2, %2, 256;\n\tadd.u64 v3, %2, 384;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 32;\n\tadd.u32 a3, %1, 40;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], v1, %3, pt;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], v2, %3, pt;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], v3, %3, pt;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
        ::"r"(obuf), "r"(sp), "l"(vdesc), "r"(idesc_pv), "r"(acc0), "r"(bar_kve), "r"(bar_s_)
        : "memory");
  }
}

// SEP: O (+)= P V with P in its own region (K-step k at packed columns 8k), then the commits kv_empty and pv_done
__device__ __forceinline__ void lean_pv48_sep(uint32_t obuf, uint32_t pbase, uint64_t vdesc, uint32_t idesc_pv, uint32_t acc0,
                                              uint32_t bar_kve, uint32_t bar_o_) {
  asm volatile(
      "{\n\t.reg .pred e, p0, pt;\n\t.reg .b64 v1, v2, v3;\n\t.reg .b32 a1, a2, a3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "add.u64 v1, %2, 128;\n\tadd.u64 v2, %2, 256;\n\tadd.u64 v3, %2, 384;\n\t"
      "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], v1, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], v2, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], v3, %3, pt;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
      ::"r"(obuf), "r"(pbase), "l"(vdesc), "r"(idesc_pv), "r"(acc0), "r"(bar_kve), "r"(bar_o_)
      : "memory");
}

template <int DPAD, bool P_HILO>
__global__ void __launch_bounds__(Cfg<DPAD, P_HILO>::THREADS, Cfg<DPAD, P_HILO>::MIN_CTAS)
attn_masked_kv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const KParams p) {
  using C = Cfg<DPAD, P_HILO>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // SWIZZLE_128B atoms are 1024-B aligned
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + C::SMEM_Q;
  const uint32_t sACC = sKV + C::NSTAGE * C::SMEM_STAGE;                // fp32 [BM][ACC_LD] cross-pass accumulator
  const uint32_t sMX = sACC + C::SMEM_ACC;                              // fp32 [2][BM] reference points (pass merge)
  const uint32_t bar_base = sMX + C::SMEM_MX;
  const uint32_t bar_q = bar_base;
  // [2] each, indexed by tile parity = S buffer = softmax warpgroup: s_full / p_full.  There is deliberately NO
  // "PV done" barrier that softmax threads wait on only occasionally: an mbarrier wait is exact only if the waiter has
  // observed every earlier phase (parity aliasing otherwise -- a skipped-phase wait fell through on hardware while PV
  // was still accumulating).  "PV(t) has completed" is instead derived from s_full, which the warpgroup of that parity
  // observes phase by phase: QK(t+2) is issued after PV(t), and a commit arrives only when ALL earlier MMAs of the
  // issuing thread are done, so s_full of tile t+2 implies PV(t).
  const uint32_t bar_s = bar_base + 16, bar_p = bar_base + 32;
  const uint32_t bar_kv_full = bar_base + 48, bar_kv_empty = bar_base + 48 + 8 * C::NSTAGE;
  const uint32_t tmem_slot = bar_base + 48 + 16 * C::NSTAGE;            // u32 written by tcgen05.alloc
  // SEP: s_free[2] (the scores of a parity's tile have been read: 4 warp arrivals) and pv_done[2] (one completion per PV
  // of that parity: PV(t) and every earlier MMA have completed; pv_done of the OTHER parity is also "the P region is free").
  // A parity wait is exact only when the barrier is at most one completion behind what is waited for and cannot run
  // ahead of it: see the notes at the two wait sites.
  const uint32_t bar_f = bar_base + 192, bar_o = bar_base + 208;
  static_assert(48 + 16 * C::NSTAGE + 8 <= 192, "barrier block layout");
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* const acc_smem = reinterpret_cast<float*>(gen_base + (sACC - smem_base));
  float* const mx_smem = reinterpret_cast<float*>(gen_base + (sMX - smem_base));
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, head = blockIdx.y, stream = blockIdx.z;
  const FFAttnHeadPlan* plan = p.plan + (size_t)stream * p.heads + head;
  int n_pass = __ldg(&plan->n_pass);
  n_pass = n_pass < 0 ? 0 : (n_pass > FF_MAX_PASS ? FF_MAX_PASS : n_pass);
  const int n_kv_tiles = (p.s_kv + BN - 1) / BN;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_s + 8, 1);
    mbar_init(bar_p, NUM_SOFTMAX_WARPS / 2);    // one elected arrival per warp of the warpgroup that owns the parity
    mbar_init(bar_p + 8, NUM_SOFTMAX_WARPS / 2);
    mbar_init(bar_f, NUM_SOFTMAX_WARPS / 2);
    mbar_init(bar_f + 8, NUM_SOFTMAX_WARPS / 2);
    mbar_init(bar_o, 1);
    mbar_init(bar_o + 8, 1);
    for (int i = 0; i < C::NSTAGE; ++i) {
      mbar_init(bar_kv_full + 8 * i, 1);
      mbar_init(bar_kv_empty + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NUM_SOFTMAX_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
#ifdef FF_ENABLE_TRACE   // build with FF_TRACE=1 (csrc/build.py); off in the product build
  uint32_t* const tr = g_trace;
  const int trole = warp == NUM_SOFTMAX_WARPS ? 0 : (warp == NUM_SOFTMAX_WARPS + 1 ? 1 : (threadIdx.x == 0 ? 2 : (threadIdx.x == 32 * NUM_SOFTMAX_WARPS - 32 ? 3 : -1)));
#define FF_TRACE(it_, site_) do { if (tr && trole >= 0 && (warp < NUM_SOFTMAX_WARPS || lane == 0)) trace(tr, trole, (uint32_t)(it_), (site_)); } while (0)
#else
#define FF_TRACE(it_, site_) do { } while (0)
#endif
  FF_TRACE(0, 1);
#ifdef FF_TIMELINE
  const int tl_lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int tl_cta = tl_lin == 4000 ? 0 : 1;
  unsigned long long* const tl = ((tl_lin == 4000 || tl_lin == 4101) && lane == 0 &&
                                  (warp == 0 || warp == NUM_SOFTMAX_WARPS + 1)) ? g_timeline : nullptr;
#endif

  if (warp == NUM_SOFTMAX_WARPS) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(bar_q, C::NKT * TILE_BYTES);
      for (int kt = 0; kt < C::NKT; ++kt)
        tma_load_4d(sQ + kt * TILE_BYTES, &tm_q, kt * BOX_COLS, head, q0, stream, bar_q);
      int it = 0;
#pragma unroll 1
      for (int ip = 0; ip < n_pass; ++ip) {
        const FFAttnPass ps = plan->pass[ip];
        const PassCtx cx = make_ctx(ps, p, q0);
        if (!cx.active) continue;
#pragma unroll 1
        for (int seg = 0; seg < 2; ++seg) {
          const SegCtx sg = seg ? cx.s1 : cx.s0;
          if (sg.kv < 0) continue;
#pragma unroll 1
          for (int j0 = 0; j0 < n_kv_tiles;) {
            const int cls = tile_class(sg, j0, p), j1 = run_end(sg, j0, cls, p), jb = j0;
            j0 = j1;
            if (tile_skip(cx, sg, cls, p.s_kv)) continue;
#pragma unroll 1
            for (int j = jb; j < j1; ++j) {
            const int stage = it % C::NSTAGE, use = it / C::NSTAGE;
            FF_TRACE(it, 10);
            if (use > 0) mbar_wait_hot<2>(bar_kv_empty + 8 * stage, (use - 1) & 1);
            FF_TRACE(it, 11);
            const uint32_t full = bar_kv_full + 8 * stage;
            const uint32_t sK = sKV + stage * C::SMEM_STAGE, sV = sK + C::NKT * KV_BYTES;
#ifdef FF_KO_TMA     // timing experiment only (wrong results): no K/V traffic after the ring is primed
            if (it >= C::NSTAGE) { mbar_arrive(full); ++it; continue; }
#endif
            mbar_expect_tx(full, 2 * C::NKT * KV_BYTES);
            for (int kt = 0; kt < C::NKT; ++kt) {
              tma_load_4d(sK + kt * KV_BYTES, &tm_k, kt * BOX_COLS, head, j * BN, sg.kv, full);
              tma_load_4d(sV + kt * KV_BYTES, &tm_v, kt * BOX_COLS, head, j * BN, sg.kv, full);
            }
            ++it;
            }
          }
        }
      }
    }
  } else if (!C::SELF && warp == NUM_SOFTMAX_WARPS + 1) {
    // ===================================== MMA issuer =======================================
    // The WHOLE warp runs this warp-uniform loop, so that descriptors, barrier addresses and counters live in uniform
    // registers; only the elected lane executes tcgen05.mma / tcgen05.commit.  (Running the loop inside `if (lane == 0)`
    // made every MMA operand a divergent-context ELECT + R2UR.BROADCAST chain: ~150 serial instructions, ~1500 cycles
    // per tile -- the bound of the whole kernel, above the MUFU unit.)
    //
    // The issuer only needs tile COUNTS (which K/V tile sits in which ring stage is the producer's business), so the
    // pass / run structure is reduced to cumulative tile counts once and the hot loop is flat:
    //     wait p_full(t)  ->  PV(t), QK(t+2) back to back  ->  commits.
    // A warpgroup owns ONE S buffer: its next tile t+2 can only start after PV(t) has consumed P(t), so everything that
    // stands between its p_full arrival and s_full(t+2) is exposed latency; kv_full(t+2) is therefore waited for BEFORE
    // p_full(t).
    {
      const bool leader = elect_one();
      constexpr uint32_t idesc_qk = make_idesc(BN, 0);
      constexpr uint32_t idesc_pv = make_idesc(C::DPV, 1, !P_HILO);
      // descriptor address fields are (byte address >> 4) in the low 14 bits: offsets are plain additions
      const uint64_t qdesc0 = smem_desc_sw128(sQ, 16);
      const uint64_t kdesc0 = smem_desc_sw128(sKV, 16);
      const uint64_t vdesc0 = smem_desc_sw128(sKV + C::NKT * KV_BYTES, KV_BYTES);
      // cumulative tile counts per pass (same decisions as the other roles)
      int pe0 = 0, pe1 = 0, pe2 = 0, n_total = 0;     // tiles up to and including pass 0 / 1 / 2; all passes
#pragma unroll
      for (int ip = 0; ip < FF_MAX_PASS; ++ip) {
        if (ip < n_pass) {
          const FFAttnPass ps = plan->pass[ip];
          const PassCtx cx = make_ctx(ps, p, q0);
          if (cx.active) {
#pragma unroll 1
            for (int seg = 0; seg < 2; ++seg) {
              const SegCtx sg = seg ? cx.s1 : cx.s0;
              if (sg.kv < 0) continue;
#pragma unroll 1
              for (int j0 = 0; j0 < n_kv_tiles;) {
                const int cls = tile_class(sg, j0, p), j1 = run_end(sg, j0, cls, p), jb = j0;
                j0 = j1;
                if (!tile_skip(cx, sg, cls, p.s_kv)) n_total += j1 - jb;
              }
            }
          }
        }
        if (ip == 0) pe0 = n_total;
        if (ip == 1) pe1 = n_total;
        if (ip == 2) pe2 = n_total;
      }
      FF_TRACE(0, 20);
      mbar_wait(bar_q, 0);
      if constexpr (C::SEP) {
        // Tile t = 4g + u: stage u, parity u & 1.  Per tile, in order:  [A] QK(t+2) as soon as the scores of tile t have
        // been read (s_free) and K/V(t+2) has landed;  [B] PV(t) when P(t) is in the P region.  The warpgroups run about
        // half a period apart, so waiting for p_full(t) before looking at s_free(t+1) costs the other parity nothing.
        tc_fence_after();
        if (n_total > 0) { lean_wait(bar_kv_full, 0); tc_fence_after(); lean_qk48(tmem + C::TMEM_S, qdesc0, kdesc0, idesc_qk, bar_s); }
        if (n_total > 1) {
          lean_wait(bar_kv_full + 8, 0);
          tc_fence_after();
          lean_qk48(tmem + C::TMEM_S + BN, qdesc0, kdesc0 + (uint64_t)(C::SMEM_STAGE >> 4), idesc_qk, bar_s + 8);
        }
        uint32_t gph = 0;
        int t = 0;
#pragma unroll 1
        while (t < n_total) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (t < n_total) {
              const int pstart = t >= pe2 ? pe2 : (t >= pe1 ? pe1 : (t >= pe0 ? pe0 : 0));
              const uint32_t acc0 = (t - pstart < 2) ? 0u : 1u;      // the first tile of each parity in a pass starts O
              FF_TL(1, t, 0);
              if (t + 2 < n_total) {
                lean_wait(bar_kv_full + 8 * ((u + 2) & 3), u < 2 ? gph : gph ^ 1u);
                lean_wait(bar_f + 8 * (u & 1), (uint32_t)((u >> 1) & 1));
                FF_TL(1, t, 1);
                tc_fence_after();
                lean_qk48(tmem + C::TMEM_S + BN * (u & 1), qdesc0, kdesc0 + (uint64_t)((((u + 2) & 3) * C::SMEM_STAGE) >> 4),
                          idesc_qk, bar_s + 8 * (u & 1));
              }
              FF_TL(1, t, 2);
              lean_wait(bar_p, (uint32_t)(u & 1));
              FF_TL(1, t, 3);
              tc_fence_after();
              lean_pv48_sep(tmem + C::TMEM_O + C::DPV * (u & 1), tmem + C::TMEM_P, vdesc0 + (uint64_t)((u * C::SMEM_STAGE) >> 4),
                            idesc_pv, acc0, bar_kv_empty + 8 * u, bar_o + 8 * (u & 1));
              FF_TL(1, t, 4);
              ++t;
            }
          }
          gph ^= 1u;
        }
        __syncwarp();
      } else if constexpr (FF_LEAN_ISSUER && DPAD == 48 && !P_HILO && C::NSTAGE == 4) {
        // Lean flat loop, unrolled over one turn of the four-stage K/V ring: tile t = 4g + u sits in stage u, its S/P
        // buffer and accumulator have parity u & 1, p_full(t) has phase (u >> 1) & 1, and the K/V stage of tile t + 2 is
        // (u + 2) & 3 with phase g (u < 2) or g + 1 -- everything but g is a compile-time constant.  kv_full(t + 2) is
        // waited for BEFORE p_full(t) (no deadlock with four stages: its refill only needed PV(t - 2)).
        tc_fence_after();
        if (n_total > 0) { lean_wait(bar_kv_full, 0); tc_fence_after(); lean_qk48(tmem + C::TMEM_S, qdesc0, kdesc0, idesc_qk, bar_s); }
        if (n_total > 1) {
          lean_wait(bar_kv_full + 8, 0);
          tc_fence_after();
          lean_qk48(tmem + C::TMEM_S + BN, qdesc0, kdesc0 + (uint64_t)(C::SMEM_STAGE >> 4), idesc_qk, bar_s + 8);
        }
        uint32_t gph = 0;
        int t = 0;
#pragma unroll 1
        while (t < n_total) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (t < n_total) {
              const bool more = t + 2 < n_total;
              const int pstart = t >= pe2 ? pe2 : (t >= pe1 ? pe1 : (t >= pe0 ? pe0 : 0));
              const uint32_t acc0 = (t - pstart < 2) ? 0u : 1u;      // the first tile of each parity in a pass starts O
              const uint32_t sp = tmem + C::TMEM_S + BN * (u & 1);
              const uint32_t obuf = tmem + C::TMEM_O + C::DPV * (u & 1);
              const uint64_t vd = vdesc0 + (uint64_t)((u * C::SMEM_STAGE) >> 4);
              const uint64_t kd = kdesc0 + (uint64_t)((((u + 2) & 3) * C::SMEM_STAGE) >> 4);
              if (more) lean_wait(bar_kv_full + 8 * ((u + 2) & 3), u < 2 ? gph : gph ^ 1u);
              lean_wait(bar_p + 8 * (u & 1), (uint32_t)((u >> 1) & 1));
              tc_fence_after();
              if (more) lean_pv_qk48<true>(obuf, sp, vd, idesc_pv, acc0, bar_kv_empty + 8 * u, qdesc0, kd, idesc_qk, bar_s + 8 * (u & 1));
              else lean_pv_qk48<false>(obuf, sp, vd, idesc_pv, acc0, bar_kv_empty + 8 * u, 0, 0, 0, bar_s + 8 * (u & 1));
              ++t;
            }
          }
          gph ^= 1u;
        }
        __syncwarp();
      } else {
      int qk_t = 0;                                   // next tile whose S = Q K^T is to be issued
      uint32_t q_stage = 0, q_phase = 0, p_stage = 0;
      auto wait_kv = [&]() { mbar_wait_hot<1>(bar_kv_full + 8 * q_stage, q_phase); };
      auto issue_qk = [&]() {                         // leader only; kv_full(qk_t) has been observed by the warp
        const uint64_t kd = kdesc0 + ((q_stage * (uint32_t)C::SMEM_STAGE) >> 4);
        const uint32_t sbuf = tmem + C::TMEM_S + BN * (qk_t & 1);
#pragma unroll
        for (int ks = 0; ks < DPAD / 16; ++ks) {
          const uint32_t qoff = (ks >> 2) * TILE_BYTES + (ks & 3) * 32;   // 16 channels = 32 B inside the row
          const uint32_t koff = (ks >> 2) * KV_BYTES + (ks & 3) * 32;
          mma_ss(sbuf, qdesc0 + (qoff >> 4), kd + (koff >> 4), idesc_qk, ks > 0);
        }
        tc_commit(bar_s + 8 * (qk_t & 1));
      };
      auto advance_qk = [&]() {
        ++qk_t;
        if (++q_stage == (uint32_t)C::NSTAGE) { q_stage = 0; q_phase ^= 1u; }
      };
      // prologue: S of the first tile of each parity
#pragma unroll 1
      for (int k = 0; k < 2 && qk_t < n_total; ++k) {
        wait_kv();
        tc_fence_after();
        if (leader) issue_qk();
        __syncwarp();
        advance_qk();
      }
#pragma unroll 1
      for (int t = 0; t < n_total; ++t) {
        const bool more = qk_t < n_total;
        // everything that does not depend on P(t) is prepared BEFORE the wait: between the arrival of p_full(t) and
        // s_full(t+2) the owning warpgroup idles.
        // O_par (+)= P V : A = P in TMEM over the S columns of tile t, B = V tile, MN-major; 16 keys = 2048 B per K-step,
        // 64-channel groups KV_BYTES apart (LBO).  hi/lo bf16: K-step ks keeps hi in columns [16ks,16ks+8) and lo in
        // [16ks+8,16ks+16); fp16: 8 packed columns at 32*(ks/2) + 8*(ks%2).  The first tile of each parity in a pass
        // starts its accumulator.
        const int pstart = t >= pe2 ? pe2 : (t >= pe1 ? pe1 : (t >= pe0 ? pe0 : 0));
        const bool first_of_parity = t - pstart < 2;
        const uint64_t vd = vdesc0 + ((p_stage * (uint32_t)C::SMEM_STAGE) >> 4);
        const uint32_t pbase = tmem + C::TMEM_S + BN * (t & 1);
        const uint32_t obuf = tmem + C::TMEM_O + C::DPV * (t & 1);
        // is the K/V tile of QK(t+2) already there?  (normally yes: the ring runs ahead; never with a two-stage ring,
        // where tile t+2 reuses the stage PV(t) is about to release).  Warp-uniform answer.
        bool kv_ready = more && __all_sync(0xffffffffu, mbar_test(bar_kv_full + 8 * q_stage, q_phase));
        FF_TL(1, t, 0);
        mbar_wait_hot<1>(bar_p + 8 * (t & 1), (t >> 1) & 1);
        FF_TL(1, t, 3);
        if (more && !kv_ready) kv_ready = __all_sync(0xffffffffu, mbar_test(bar_kv_full + 8 * q_stage, q_phase));
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            const uint64_t vdesc = vd + ((ks * 2048) >> 4);
            if constexpr (P_HILO) {
              mma_ts(obuf, pbase + 16 * ks, vdesc, idesc_pv, (!first_of_parity || ks > 0) ? 1u : 0u);
              mma_ts(obuf, pbase + 16 * ks + 8, vdesc, idesc_pv, 1u);
            } else {
              mma_ts(obuf, pbase + 32 * (ks >> 1) + 8 * (ks & 1), vdesc, idesc_pv,
                     (!first_of_parity || ks > 0) ? 1u : 0u);
            }
          }
          tc_commit(bar_kv_empty + 8 * p_stage);
          // S of the same warpgroup's next tile right behind PV(t) in the pipe -- or, for the last tile of a parity, a
          // virtual s_full commit in its place, so that "s_full(t+2) => PV(t) has completed" holds there too
          if (kv_ready) issue_qk();
          else if (!more) tc_commit(bar_s + 8 * (t & 1));
        }
        __syncwarp();
        if (more && !kv_ready) {
          wait_kv();
          tc_fence_after();
          if (leader) issue_qk();
          __syncwarp();
        }
        if (more) advance_qk();
        FF_TL(1, t, 4);
        if (++p_stage == (uint32_t)C::NSTAGE) p_stage = 0;
      }
      __syncwarp();
      }
    }
  } else {
    // ===================================== softmax + epilogue ===============================
    // Warpgroup wg (warps 4*wg .. 4*wg+3) owns the K/V tiles of parity wg: S buffer wg, accumulator O_wg, its own running
    // reference point.  A thread = one query row (TMEM lane).  The two partial softmaxes are merged at the end of a pass.
    const int wq = warp & 3;        // TMEM lane quarter of this warp (hardware rule: warp w reaches lanes 32*(w%4)..+31)
    const int wg = warp >> 2;       // tile parity of this warpgroup
    const int rloc = 32 * wq + lane;
    const int row = q0 + rloc;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t tS = tlane + C::TMEM_S + BN * wg;                  // my S / P buffer
    const uint32_t tO = tlane + C::TMEM_O + C::DPV * wg;              // my O accumulator
    const uint32_t tOx = tlane + C::TMEM_O + C::DPV * (wg ^ 1);       // the other warpgroup's
    float* const acc_row = acc_smem + (size_t)rloc * C::ACC_LD;
    bool acc_started = false;       // has any pass been added to the accumulator yet (uniform across the CTA)
    int it = 0;                     // global tile counter (all roles count alike)
    // ---- SELF: the first warp of each warpgroup is the MMA issuer of the warpgroup's tiles (tile t sits in K/V stage
    // t % 4, its S/P buffer and accumulator have the warpgroup's parity)
    int n_total = 0;
    [[maybe_unused]] constexpr uint32_t s_idesc_qk = make_idesc(BN, 0);
    [[maybe_unused]] constexpr uint32_t s_idesc_pv = make_idesc(C::DPV, 1, !P_HILO);
    [[maybe_unused]] const uint64_t s_qdesc0 = smem_desc_sw128(sQ, 16);
    [[maybe_unused]] const uint64_t s_kdesc0 = smem_desc_sw128(sKV, 16);
    [[maybe_unused]] const uint64_t s_vdesc0 = smem_desc_sw128(sKV + C::NKT * KV_BYTES, KV_BYTES);
    if constexpr (C::SELF) {
      if (wq == 0) {
#pragma unroll 1
        for (int ip = 0; ip < n_pass; ++ip) {             // tiles of the whole CTA (same decisions as every role)
          const FFAttnPass ps = plan->pass[ip];
          const PassCtx cx = make_ctx(ps, p, q0);
          if (!cx.active) continue;
#pragma unroll 1
          for (int seg = 0; seg < 2; ++seg) {
            const SegCtx sg = seg ? cx.s1 : cx.s0;
            if (sg.kv < 0) continue;
#pragma unroll 1
            for (int j0 = 0; j0 < n_kv_tiles;) {
              const int cls = tile_class(sg, j0, p), j1 = run_end(sg, j0, cls, p), jb = j0;
              j0 = j1;
              if (!tile_skip(cx, sg, cls, p.s_kv)) n_total += j1 - jb;
            }
          }
        }
        n_total = __shfl_sync(0xffffffffu, n_total, 0);
        if (wg < n_total) {                                 // S of my first tile (tile wg, stage wg)
          mbar_wait(bar_q, 0);
          lean_wait(bar_kv_full + 8 * wg, 0);
          tc_fence_after();
          lean_qk48(tmem + C::TMEM_S + BN * wg, s_qdesc0, s_kdesc0 + (uint64_t)((wg * C::SMEM_STAGE) >> 4), s_idesc_qk,
                    bar_s + 8 * wg);
        }
      }
    }
#pragma unroll 1
    for (int ip = 0; ip < n_pass; ++ip) {
      const FFAttnPass ps = plan->pass[ip];
      const PassCtx cx = make_ctx(ps, p, q0);
      if (!cx.active) continue;
      // ---- per-row constants of this pass
      uint32_t rb = 0;
      if (ps.row_mask >= 0 && row < p.s_q)
        rb = (__ldg(p.bitmasks + (size_t)ps.row_mask * p.mask_words + (row >> 5)) >> (row & 31)) & 1u;
      const bool rowflip = cx.rowxor && rb;
      float m_used = -INFINITY;     // reference point of MY tiles (log2 units); -inf: nothing read so far
      int n_mine = 0, last_mine = 0;
      const int it_pass0 = it;
#pragma unroll 1
      for (int seg = 0; seg < 2; ++seg) {
        const SegCtx sg = seg ? cx.s1 : cx.s0;
        if (sg.kv < 0) continue;
        const bool flip = sg.kinv != rowflip;
        const bool uniform = uniform_for(cx, sg, flip, p.s_kv);   // quirk Q4 (per row: depends on its flip value)
        const float sc = uniform ? 0.f : p.scale_log2;
#pragma unroll 1
        for (int j0 = 0; j0 < n_kv_tiles;) {
          const int cls = tile_class(sg, j0, p), j1 = run_end(sg, j0, cls, p), jb = j0;
          j0 = j1;
          if (tile_skip(cx, sg, cls, p.s_kv)) continue;
          const bool row_ok_cls = row_allowed(cls, flip, uniform);   // per-row predicate of this whole run
          // my tiles of the run: those with global index parity wg
          int j = jb + (((it ^ wg) & 1) ? 1 : 0);
          int itj = it + (j - jb);
          it += j1 - jb;
#pragma unroll 1
          for (; j < j1; j += 2, itj += 2) {
            FF_TRACE(itj, 30);
            FF_TL(0, itj, 0);
            mbar_wait_hot<0>(bar_s + 8 * wg, (itj >> 1) & 1);        // also: PV(itj-2) has finished writing O_wg
            FF_TL(0, itj, 1);
            FF_TRACE(itj, 31);
            tc_fence_after();
            // ---- allowed-key bits of this row for MIX tiles (boundary / ragged): bit i <=> key j*BN + i
            uint32_t kb_lo = 0xffffffffu, kb_hi = 0xffffffffu;         // columns [0,32) / [32,64)
            if (cls == TILE_MIX) {
#pragma unroll
              for (int w = 0; w < 2; ++w) {
                const int kbase = j * BN + 32 * w;
                const int rem = p.s_kv - kbase;
                const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                uint32_t kb = 0xffffffffu;
                if (sg.kmask >= 0 && !uniform && rem > 0) {
                  if (sg.prefix) {
                    const int t = sg.T - kbase;
                    kb = t >= 32 ? 0xffffffffu : (t <= 0 ? 0u : ((1u << t) - 1u));
                  } else {
                    kb = __ldg(p.bitmasks + (size_t)sg.kmask * p.mask_words + (kbase >> 5));   // BN % 32 == 0
                  }
                  if (flip) kb = ~kb;
                }
                if (w == 0) kb_lo = kb & valid; else kb_hi = kb & valid;
              }
            }
            // ---- row max over the ALLOWED keys of the tile (the fp16 P operand has a narrow exponent range: the
            // reference point must not come from keys this row does not read).  Columns [0,32) are reduced first and
            // dropped (two CTAs x 320 threads leave ~100 registers per thread), then re-read for the exp sweep.
            float mt, sb[32];
            {
              float sa[32];
              tmem_ld32(tS, sa);
#ifdef FF_LD_BOTH     // both halves of the row in flight together: one tcgen05.ld round trip instead of two per tile
              tmem_ld32(tS + 32, sb);
#endif
              tmem_wait_ld32(sa);
              float m0, m1;
              if (cls != TILE_MIX) {
                m0 = fmaxf(sa[0], sa[1]);
                m1 = fmaxf(sa[2], sa[3]);
#pragma unroll
                for (int i = 4; i < 32; i += 4) {
                  m0 = fmaxf(m0, fmaxf(sa[i], sa[i + 1]));
                  m1 = fmaxf(m1, fmaxf(sa[i + 2], sa[i + 3]));
                }
              } else {
                m0 = m1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  m0 = fmaxf(m0, (kb_lo >> i) & 1u ? sa[i] : -INFINITY);
                  m1 = fmaxf(m1, (kb_lo >> (i + 1)) & 1u ? sa[i + 1] : -INFINITY);
                }
              }
              mt = fmaxf(m0, m1);
            }
            {
#ifdef FF_KO_LD2      // timing experiment only (wrong results): columns [32,64) are never read
#pragma unroll
              for (int i = 0; i < 32; ++i) sb[i] = mt + (float)i;
#else
#ifndef FF_LD_BOTH
              tmem_ld32(tS + 32, sb);
#endif
              tmem_wait_ld32(sb);
#endif
              float m0, m1;
              if (cls != TILE_MIX) {
                m0 = fmaxf(sb[0], sb[1]);
                m1 = fmaxf(sb[2], sb[3]);
#pragma unroll
                for (int i = 4; i < 32; i += 4) {
                  m0 = fmaxf(m0, fmaxf(sb[i], sb[i + 1]));
                  m1 = fmaxf(m1, fmaxf(sb[i + 2], sb[i + 3]));
                }
              } else {
                m0 = m1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  m0 = fmaxf(m0, (kb_hi >> i) & 1u ? sb[i] : -INFINITY);
                  m1 = fmaxf(m1, (kb_hi >> (i + 1)) & 1u ? sb[i + 1] : -INFINITY);
                }
              }
              mt = fmaxf(mt, fmaxf(m0, m1));
              if (cls != TILE_MIX && !row_ok_cls) mt = -INFINITY;
#ifdef FF_KO_MAX    // timing experiment only
              mt = 0.f;
#endif
            }
            FF_TL(0, itj, 2);
            const float mts = uniform ? 0.f : mt * p.scale_log2;     // (-inf: the row reads nothing from this tile)
            // ---- running reference point, lazy rescale of O_wg (TMEM read-modify-write only when the max grew a lot)
            float alpha = 1.f;
            bool grow = false;
            if (n_mine == 0) {
              m_used = mts;
            } else if (mts > m_used + rescale_threshold<P_HILO>()) {
              alpha = fast_exp2(m_used - mts);       // (m_used = -inf, nothing read so far: alpha = 0, O is 0 anyway)
              m_used = mts;
              grow = true;
            }
            const bool any_grow = __any_sync(0xffffffffu, grow);
            auto rescale_o = [&]() {
#pragma unroll
              for (int c = 0; c < C::DPV / 16; ++c) {      // (includes the denominator column)
                float o[16];
                uint32_t ob[16];
                tmem_ld16(tO + 16 * c, o);
                tmem_wait_ld16(o);
#pragma unroll
                for (int i = 0; i < 16; ++i) ob[i] = __float_as_uint(o[i] * alpha);
                tmem_st16(tO + 16 * c, ob);
              }
            };
            // (SEP: s_full(itj) no longer implies PV(itj-2) -- the rescale waits for pv_done below, before P is stored)
            if (!C::SEP && any_grow) rescale_o();
            FF_TL(0, itj, 3);
            // ---- p = 2^(s*scale*log2e - m), packed for the tensor core over the S columns of the same keys:
            // fp16: K-step ks -> packed columns 32*(ks/2) + 8*(ks%2) + [0,8); hi/lo bf16: hi at 16*ks + [0,8), lo at
            // 16*ks + [8,16).  Columns [32,64) first (still in registers), then [0,32) re-read.
            const float nb = (cls == TILE_MIX || row_ok_cls) ? -m_used : -INFINITY;   // -inf: p = 0 for the whole row
            uint32_t pk[(P_HILO || C::SEP) ? 32 : 16];     // SEP: [0,16) = keys 0..31, [16,32) = keys 32..63, stored at the end
#pragma unroll
            for (int hb = 1; hb >= 0; --hb) {
              float sa[32];
#ifdef FF_KO_REREAD   // timing experiment only (wrong results): columns [0,32) are not read a second time
              const float* sv = sb;
#else
              if (hb == 0) {
                tmem_ld32(tS, sa);
                tmem_wait_ld32(sa);
                if constexpr (C::SEP) {
                  // every score of the tile is in registers: the S buffer may take QK(itj+2) now
                  tc_fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(bar_f + 8 * wg);
                }
              }
              const float* sv = hb ? sb : sa;
#endif
              const uint32_t kbits = hb ? kb_hi : kb_lo;
              if (cls != TILE_MIX) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  if constexpr (P_HILO) softmax_chunk_hilo<false>(sv + 16 * jj, pk + 16 * jj, sc, nb, 0u);
                  else softmax_chunk_f16<false>(sv + 16 * jj, pk + (C::SEP ? 16 * hb : 0) + 8 * jj, sc, nb, 0u);
                }
              } else {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  const uint32_t bits = (kbits >> (16 * jj)) & 0xffffu;
                  if constexpr (P_HILO) softmax_chunk_hilo<true>(sv + 16 * jj, pk + 16 * jj, sc, nb, bits);
                  else softmax_chunk_f16<true>(sv + 16 * jj, pk + (C::SEP ? 16 * hb : 0) + 8 * jj, sc, nb, bits);
                }
              }
              if constexpr (P_HILO) {
                tmem_st16(tS + 32 * hb, *reinterpret_cast<const uint32_t(*)[16]>(pk));
                tmem_st16(tS + 32 * hb + 16, *reinterpret_cast<const uint32_t(*)[16]>(pk + 16));
              } else if constexpr (!C::SEP) {
                tmem_st16(tS + 32 * hb, *reinterpret_cast<const uint32_t(*)[16]>(pk));
              }
            }
            if constexpr (C::SEP) {
              // the P region is shared with the other warpgroup: PV(itj-1) must have consumed P(itj-1); in-order
              // completion then also covers PV(itj-2), the last writer of O_wg (rescale).  Exact: the previous completion
              // of that barrier, PV(itj-3), is implied by s_full(itj) (QK(itj) was issued after it), and the next one,
              // PV(itj+1), needs a P that is only stored after PV(itj), i.e. after the P this warpgroup is about to store.
              if (itj >= 1) {
                const uint32_t bo = bar_o + 8 * (wg ^ 1), po = (uint32_t)(((itj - 1) >> 1) & 1);
                if (!mbar_test(bo, po)) mbar_wait(bo, po);
              }
              tc_fence_after();
              if (any_grow) rescale_o();
              tmem_st16(tlane + C::TMEM_P, *reinterpret_cast<const uint32_t(*)[16]>(pk));
              tmem_st16(tlane + C::TMEM_P + 16, *reinterpret_cast<const uint32_t(*)[16]>(pk + 16));
            }
            FF_TL(0, itj, 4);
            tmem_wait_st();
            FF_TL(0, itj, 5);
            tc_fence_before();
            if constexpr (C::SELF) {
              named_bar_sync(3 + wg, 128);                 // every row of P(itj) has been stored
              if (wq == 0) {
                const int stg = itj & 3, nxt = (itj + 2) & 3;
                const bool more = itj + 2 < n_total;
                if (more) lean_wait(bar_kv_full + 8 * nxt, (uint32_t)(((itj + 2) >> 2) & 1));
                tc_fence_after();
                const uint32_t sp = tmem + C::TMEM_S + BN * wg;
                const uint32_t obuf = tmem + C::TMEM_O + C::DPV * wg;
                const uint64_t vd = s_vdesc0 + (uint64_t)((stg * C::SMEM_STAGE) >> 4);
                const uint64_t kd = s_kdesc0 + (uint64_t)((nxt * C::SMEM_STAGE) >> 4);
                const uint32_t acc0 = n_mine == 0 ? 0u : 1u;   // my first tile of the pass starts the accumulator
                if (more) lean_pv_qk48<true>(obuf, sp, vd, s_idesc_pv, acc0, bar_kv_empty + 8 * stg, s_qdesc0, kd, s_idesc_qk, bar_s + 8 * wg);
                else lean_pv_qk48<false>(obuf, sp, vd, s_idesc_pv, acc0, bar_kv_empty + 8 * stg, 0, 0, 0, bar_s + 8 * wg);
              }
            } else {
              __syncwarp();
              if (lane == 0) mbar_arrive(C::SEP ? bar_p : bar_p + 8 * wg);     // (SEP: one p_full barrier, phases in tile order)
            }
            FF_TL(0, itj, 6);
            FF_TRACE(itj, 34);
            ++n_mine;
            last_mine = itj;
          }
        }
      }
      if (it == it_pass0) continue;     // (defensive) no tile of this pass was processed: nothing to add (CTA-uniform)
      // ---- end of pass: merge the two partial softmaxes, acc += weight * roww / l * O.
      // (1) my last PV has landed: s_full(last_mine + 2) -- a real tile of the next pass or one of the two virtual commits
      FF_TRACE(it, 35);
      if constexpr (C::SEP) {
        // every PV of the pass has completed <=> the PV of its last tile L = it - 1 has (in-order completion).  Exact: the
        // previous completion of that barrier, PV(L-2), was implied by the last per-tile wait of either warpgroup (or by
        // the previous end of pass), and PV(L+2) belongs to the next pass, which starts after this merge.
        const int L = it - 1;
        mbar_wait(bar_o + 8 * (L & 1), (uint32_t)((L >> 1) & 1));
        tc_fence_after();
      } else if (n_mine > 0) mbar_wait_hot<0>(bar_s + 8 * wg, ((last_mine + 2) >> 1) & 1);
      FF_TRACE(it, 36);
      // (2) exchange the reference points (row-wise, through shared memory)
      mx_smem[wg * BM + rloc] = n_mine > 0 ? m_used : -INFINITY;
      tc_fence_before();
      named_bar_sync(1, 32 * NUM_SOFTMAX_WARPS);
      tc_fence_after();
      const float m_other = mx_smem[(wg ^ 1) * BM + rloc];
      const float m_all = fmaxf(m_used, m_other);          // (n_mine == 0 -> m_used = -inf by construction)
      // tcgen05.ld is warp-collective: TMEM reads are guarded by CTA-uniform tile counts only (an accumulator that no
      // PV of this pass wrote is not read); rows that read nothing (reference point -inf) get weight 0 per lane
      const bool wrote_mine = n_mine > 0, wrote_other = (it - it_pass0) - n_mine > 0;
      const float a_mine = (wrote_mine && m_used > -INFINITY) ? fast_exp2(m_used - m_all) : 0.f;
      const float a_other = (wrote_other && m_other > -INFINITY) ? fast_exp2(m_other - m_all) : 0.f;
      // (3) denominators from the ones column of each accumulator
      float l = 0.f;
      if (wrote_mine) l = a_mine * tmem_ld1_wait(tO + p.head_dim);
      if (wrote_other) l = fmaf(a_other, tmem_ld1_wait(tOx + p.head_dim), l);
      float coef = ps.weight;
      if (ps.flags & FF_PASS_ROW_WEIGHT) coef = rb ? coef : 0.f;
      coef = l > 0.f ? coef / l : 0.f;
      const float c_mine = coef * a_mine, c_other = coef * a_other;
      // (4) my share of the channels: 16-channel chunks of parity wg, both accumulators
#pragma unroll
      for (int c = 0; c < DPAD / 16; ++c) {
        if ((c & 1) != wg) continue;
        float r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = acc_started ? acc_row[16 * c + i] : 0.f;
        if (wrote_mine) {
          float o[16];
          tmem_ld16(tO + 16 * c, o);
          tmem_wait_ld16(o);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = fmaf(c_mine, o[i], r[i]);      // (c_mine = 0 for rows that read nothing)
        }
        if (wrote_other) {
          float o[16];
          tmem_ld16(tOx + 16 * c, o);
          tmem_wait_ld16(o);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = fmaf(c_other, o[i], r[i]);
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(acc_row + 16 * c + i) = make_float4(r[i], r[i + 1], r[i + 2], r[i + 3]);
      }
      acc_started = true;
      // (5) both accumulators have been read by everybody: the next pass may overwrite them
      tc_fence_before();
      named_bar_sync(2, 32 * NUM_SOFTMAX_WARPS);
      tc_fence_after();
    }
    // ---- write the row: out[stream, row, head*d : (head+1)*d]; each thread writes the chunks it accumulated
    {
      const bool row_ok = row < p.s_q;
      const size_t o_off = ((size_t)stream * p.s_q + (row_ok ? row : 0)) * ((size_t)p.heads * p.head_dim) +
                           (size_t)head * p.head_dim;
#pragma unroll
      for (int c = 0; c < DPAD / 16; ++c) {
        if ((c & 1) != wg) continue;
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = acc_started ? acc_row[16 * c + i] : 0.f;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (row_ok && 16 * c + 8 * g < p.head_dim) {   // head_dim % 8 == 0
            if (p.out_dtype == FF_DT_BF16) {
              uint4 v;
              v.x = pack_bf16x2(o[8 * g + 0], o[8 * g + 1]);
              v.y = pack_bf16x2(o[8 * g + 2], o[8 * g + 3]);
              v.z = pack_bf16x2(o[8 * g + 4], o[8 * g + 5]);
              v.w = pack_bf16x2(o[8 * g + 6], o[8 * g + 7]);
              *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + o_off + 16 * c + 8 * g) = v;
            } else {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + o_off + 16 * c + 8 * g);
              dst[0] = make_float4(o[8 * g + 0], o[8 * g + 1], o[8 * g + 2], o[8 * g + 3]);
              dst[1] = make_float4(o[8 * g + 4], o[8 * g + 5], o[8 * g + 6], o[8 * g + 7]);
            }
          }
        }
      }
    }
  }
  FF_TRACE(0, 90);
  // ---- teardown: every tcgen05 op of this CTA has completed (softmax threads waited on o_done of the last tile)
  tc_fence_before();
  __syncthreads();
  if (warp == NUM_SOFTMAX_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}

#include "attn_ring.cuh"        // ring-buffered kernel: the product kernel for fp16-P, 8 < head_dim <= 80
#include "attn_t32.cuh"         // 32-key-tile kernel for head_dim <= 40 (FF_ATTN_T32)

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [streams, S, heads, d] bf16 view of a dense [streams, S, heads*d] tensor; box = 64 channels x 1 head x 128 rows.
// Channels >= d of a box are out of bounds in dimension 0 and therefore zero-filled.
// For the staged V, d = v_head_stride (real channels + ones column + zero padding, all in bounds).
int make_map(CUtensorMap* map, const void* base, int streams, int S, int heads, int d, int box_rows,
             bool f16 = false) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return ff::fail(FF_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const cuuint64_t C = (cuuint64_t)heads * d;
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)S, (cuuint64_t)streams};
  cuuint64_t strides[3] = {(cuuint64_t)d * 2, C * 2, (cuuint64_t)S * C * 2};
  cuuint32_t box[4] = {BOX_COLS, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return ff::fail(FF_E_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) dims=[%d,%d,%d,%d]", (int)r, d, heads, S,
                    streams);
  return FF_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE attribute: configure once per (kernel, device).
template <typename K>
int ensure_smem(K kernel, int bytes, std::atomic<uint64_t>& done) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return FF_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return ff::fail(FF_E_CUDA, "cudaFuncSetAttribute(smem=%d): %s", bytes, cudaGetErrorString(e));
  done.fetch_or(bit, std::memory_order_release);
  return FF_OK;
}

template <int DPAD, bool HILO>
int launch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const KParams& kp, int n_streams,
           cudaStream_t st) {
  using C = Cfg<DPAD, HILO>;
  static std::atomic<uint64_t> configured{0};
  int rc = ensure_smem(attn_masked_kv_kernel<DPAD, HILO>, C::SMEM_BYTES, configured);
  if (rc != FF_OK) return rc;
  dim3 grid((kp.s_q + BM - 1) / BM, kp.heads, n_streams);
  attn_masked_kv_kernel<DPAD, HILO><<<grid, C::THREADS, C::SMEM_BYTES, st>>>(mq, mk, mv, kp);
  return ff::check_launch("ff_attn_masked_kv");
}

int sm_count() {                    // SMs of the current device (cached per device)
  static std::atomic<int> cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int n = cached[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cached[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

template <int DPAD, int NWG, int NBUF, bool TOKEN>
int launch_ring(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const KParams& kp, int n_streams,
                cudaStream_t st) {
  using C = RCfg<DPAD, NWG, NBUF>;
  static std::atomic<uint64_t> configured{0};
  int rc = ensure_smem(attn_ring_kernel<DPAD, NWG, NBUF, TOKEN>, C::SMEM_BYTES, configured);
  if (rc != FF_OK) return rc;
  // persistent grid: one CTA per SM (two for the layouts that fit twice), never more CTAs than work items
  const int n_items = kp.n_qtiles * kp.heads * n_streams;
  int ctas = sm_count() * C::MIN_CTAS;
  if (ctas > n_items) ctas = n_items;
  attn_ring_kernel<DPAD, NWG, NBUF, TOKEN><<<ctas, C::NUM_THREADS, C::SMEM_BYTES, st>>>(mq, mk, mv, kp);
  return ff::check_launch("ff_attn_masked_kv (ring kernel)");
}

int launch_t32(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const KParams& kp, int n_streams, cudaStream_t st) {
  static std::atomic<uint64_t> configured{0};
  int rc = ensure_smem(attn_t32_kernel, T32::SMEM_BYTES, configured);
  if (rc != FF_OK) return rc;
  dim3 grid((kp.s_q + BM - 1) / BM, kp.heads, n_streams);
  attn_t32_kernel<<<grid, NUM_THREADS, T32::SMEM_BYTES, st>>>(mq, mk, mv, kp);
  return ff::check_launch("ff_attn_masked_kv (32-key tiles)");
}

int t32_mode() {      // FF_ATTN_T32 in the environment, read once: 1 = 32-key-tile kernel for 8 < head_dim <= 40 (fp16 P)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FF_ATTN_T32");
    mode = e ? (atoi(e) != 0) : FF_T32_DEFAULT;
  }
  return mode;
}

// Kernel selection for the fp16-P path, 8 < head_dim <= 80 (FF_ATTN_RING in the environment, read once; experiments only):
//   0 = legacy kernel; ring <DPAD, warpgroups, S/P buffers, exp token> for d<=40 / d<=80:
//   1 = <48,1,3,0> (two CTAs per SM) / <80,2,3,0>    2 = <48,2,4,0> / <80,2,3,0>    3 = <48,3,5,0> / <80,2,3,0>
//   4 = <48,3,5,1> / <80,2,3,1>                      5 = <48,2,4,1> / <80,2,3,1>
//   6 = legacy kernel for d<=40 / <80,2,3,0> for d<=80 -- the default: what measured fastest per head dim in round 2.
int ring_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("FF_ATTN_RING");
    mode = e ? atoi(e) : FF_RING_DEFAULT;
    if (mode < 0 || mode > 7) mode = FF_RING_DEFAULT;
  }
  return mode;
}

}  // namespace

// Debug hook (not part of the product path): device-visible pointer (e.g. cudaHostAlloc'ed, mapped) that receives
// 8 uint32 per CTA: {producer it, site, mma it, site, softmax-row0 it, site, softmax-row96 it, site}; NULL disables.
// Debug hook: device buffer of 2 CTAs x 2 roles x 64 tiles x 8 clock64 stamps (library built with -DFF_TIMELINE).
extern "C" int ff_debug_set_timeline(void* device_ptr) {
#ifdef FF_TIMELINE
  cudaError_t e = cudaMemcpyToSymbol(g_timeline, &device_ptr, sizeof(void*));
  if (e != cudaSuccess) return ff::fail(FF_E_CUDA, "ff_debug_set_timeline: %s", cudaGetErrorString(e));
  return FF_OK;
#else
  (void)device_ptr;
  return ff::fail(FF_E_UNSUPPORTED, "ff_debug_set_timeline: library built without -DFF_TIMELINE");
#endif
}

extern "C" int ff_debug_set_trace(void* device_visible_ptr) {
#ifndef FF_ENABLE_TRACE
  if (device_visible_ptr) return ff::fail(FF_E_UNSUPPORTED, "ff_debug_set_trace: library built without FF_TRACE=1");
#endif
  cudaError_t e = cudaMemcpyToSymbol(g_trace, &device_visible_ptr, sizeof(void*));
  if (e != cudaSuccess) return ff::fail(FF_E_CUDA, "ff_debug_set_trace: %s", cudaGetErrorString(e));
  return FF_OK;
}

// Channels per head of the staged V: head_dim real channels, a ones column at channel head_dim, zeros above;
// equals Cfg<DPAD,*>::DPV of the instantiation ff_attn_masked_kv picks for this head_dim.
extern "C" int ff_attn_v_head_stride(int32_t head_dim) {
  if (head_dim <= 8) return Cfg<16, false>::DPV;
  if (head_dim <= 40) return Cfg<48, false>::DPV;
  if (head_dim <= 80) return Cfg<80, false>::DPV;
  return Cfg<160, false>::DPV;
}

extern "C" int ff_attn_masked_kv(const FFAttnArgs* a, void* stream) {
  FF_REQUIRE(a != nullptr, "ff_attn_masked_kv: null args");
  FF_REQUIRE(a->q && a->k && a->v && a->out && a->plan, "ff_attn_masked_kv: null pointer");
  FF_REQUIRE(a->n_streams > 0 && a->n_kv_streams > 0 && a->heads > 0 && a->s_q > 0 && a->s_kv > 0,
             "ff_attn_masked_kv: bad shape");
  FF_REQUIRE(a->head_dim >= 8 && a->head_dim % 8 == 0, "ff_attn_masked_kv: head_dim=%d must be a multiple of 8",
             a->head_dim);
  if (a->head_dim > 160) return ff::fail(FF_E_UNSUPPORTED, "ff_attn_masked_kv: head_dim=%d > 160", a->head_dim);
  FF_REQUIRE(a->heads <= 65535 && a->n_streams <= 65535, "ff_attn_masked_kv: heads / streams exceed grid limits");
  FF_REQUIRE(ff::aligned16(a->q) && ff::aligned16(a->k) && ff::aligned16(a->v) && ff::aligned16(a->out),
             "ff_attn_masked_kv: q/k/v/out must be 16-byte aligned");
  FF_REQUIRE(a->out_dtype == FF_DT_BF16 || a->out_dtype == FF_DT_F32, "ff_attn_masked_kv: bad out_dtype");
  FF_REQUIRE(a->v_dtype == FF_DT_BF16 || a->v_dtype == FF_DT_F16, "ff_attn_masked_kv: v_dtype must be bf16 or f16");
  FF_REQUIRE(a->scale > 0.f, "ff_attn_masked_kv: scale must be positive");
  if (a->n_masks > 0) {
    FF_REQUIRE(a->bitmasks && a->mask_popcount, "ff_attn_masked_kv: n_masks>0 but no bitmasks / popcounts");
    const int need = ((a->s_q > a->s_kv ? a->s_q : a->s_kv) + 31) / 32;
    FF_REQUIRE(a->mask_words >= need, "ff_attn_masked_kv: mask_words=%d < %d", a->mask_words, need);
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  {
    static std::atomic<int> cached_major[64];          // per device, queried once (this runs once per attention launch)
    major = cached_major[dev & 63].load(std::memory_order_relaxed);
    if (major == 0) {
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
      cached_major[dev & 63].store(major, std::memory_order_relaxed);
    }
  }
  if (major != 10) return ff::fail(FF_E_ARCH, "ff_attn_masked_kv: needs an sm_100 device (got sm_%d*)", major);

  alignas(64) CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_map(&mq, a->q, a->n_streams, a->s_q, a->heads, a->head_dim, BM)) != FF_OK) return rc;
  const bool v_f16 = a->v_dtype == FF_DT_F16;
  const bool use_t32 = v_f16 && a->head_dim > 8 && a->head_dim <= 40 && t32_mode() != 0;
  const int kv_rows = use_t32 ? TBN : BN;                  // rows of a K / V box = keys per tile of the kernel that runs
  if ((rc = make_map(&mk, a->k, a->n_kv_streams, a->s_kv, a->heads, a->head_dim, kv_rows)) != FF_OK) return rc;
  FF_REQUIRE(a->v_head_stride == ff_attn_v_head_stride(a->head_dim),
             "ff_attn_masked_kv: V must be staged by ff_kv_gather_cast (v_head_stride=%d, expected %d)",
             a->v_head_stride, ff_attn_v_head_stride(a->head_dim));
  if ((rc = make_map(&mv, a->v, a->n_kv_streams, a->s_kv, a->heads, a->v_head_stride, kv_rows, v_f16)) != FF_OK) return rc;

  KParams kp;
  kp.plan = a->plan;
  kp.bitmasks = a->bitmasks;
  kp.popc = a->mask_popcount;
  kp.out = a->out;
  kp.heads = a->heads;
  kp.head_dim = a->head_dim;
  kp.s_q = a->s_q;
  kp.s_kv = a->s_kv;
  kp.mask_words = a->mask_words;
  kp.out_dtype = a->out_dtype;
  kp.n_kv_streams = a->n_kv_streams;
  kp.n_streams = a->n_streams;
  kp.n_qtiles = (a->s_q + BM - 1) / BM;
  kp.scale_log2 = a->scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int d = a->head_dim;
  // thresholds leave room for the ones column at channel d inside DPV (see ff_attn_v_head_stride)
  if (v_f16) {
    if (d <= 8) return launch<16, false>(mq, mk, mv, kp, a->n_streams, st);
    if (use_t32) return launch_t32(mq, mk, mv, kp, a->n_streams, st);
    const int rm = ring_mode();
    if (rm == 1 && d <= 40) return launch_ring<48, 1, 3, false>(mq, mk, mv, kp, a->n_streams, st);
    if (rm == 2 && d <= 40) return launch_ring<48, 2, 4, false>(mq, mk, mv, kp, a->n_streams, st);
    if (rm == 3 && d <= 40) return launch_ring<48, 3, 5, false>(mq, mk, mv, kp, a->n_streams, st);
    if (rm == 7 && d <= 40) return launch_ring<48, 4, 5, false>(mq, mk, mv, kp, a->n_streams, st);
    if (rm == 4 && d <= 40) return launch_ring<48, 3, 5, true>(mq, mk, mv, kp, a->n_streams, st);
    if (rm == 5 && d <= 40) return launch_ring<48, 2, 4, true>(mq, mk, mv, kp, a->n_streams, st);
    if (d > 40 && d <= 80 && (rm == 4 || rm == 5)) return launch_ring<80, 2, 3, true>(mq, mk, mv, kp, a->n_streams, st);
    if (d > 40 && d <= 80 && rm != 0) return launch_ring<80, 2, 3, false>(mq, mk, mv, kp, a->n_streams, st);
    if (d <= 40) return launch<48, false>(mq, mk, mv, kp, a->n_streams, st);
    if (d <= 80) return launch<80, false>(mq, mk, mv, kp, a->n_streams, st);
    return launch<160, false>(mq, mk, mv, kp, a->n_streams, st);
  }
  if (d <= 8) return launch<16, true>(mq, mk, mv, kp, a->n_streams, st);
  if (d <= 40) return launch<48, true>(mq, mk, mv, kp, a->n_streams, st);
  if (d <= 80) return launch<80, true>(mq, mk, mv, kp, a->n_streams, st);
  return launch<160, true>(mq, mk, mv, kp, a->n_streams, st);
}
