// Ring-buffered variant of the masked KV-injection attention kernel (fp16-P path, head_dim 16 < d <= 80): the product
// kernel for the SD1.5 d = 40 and d = 80 layers.  Included by attn_tcgen05.cu inside its anonymous namespace (shares the
// PTX wrappers, the per-pass / per-tile decision functions and the exp2 sweep with the legacy kernel above, which keeps
// serving d <= 8, d > 80 and the hi+lo bf16 P operand).
//
// What the round-1 timelines and SASS showed about the legacy kernel (profiles/r1_attn_timeline.txt):
//   * a softmax warpgroup owns ONE S buffer, so between its p_full arrival and s_full of its next tile it idles for the
//     whole  p_full -> PV(t), QK(t+2) -> s_full  hand-shake (800-1000 cycles per tile);
//   * the single MMA-issuer warp executes ~120 mostly serial instructions per tile (R2UR chains that rebuild every
//     descriptor from loop counters living in vector registers): ~500 cycles per tile -- in a one-CTA-per-SM layout with a
//     single issuer (round-1 draft attn_onecta.cuh, first run in round 2: parity-green, 3.59 ms vs 2.36 ms) that issuer IS
//     the bound of the kernel;
//   * 96 registers cannot hold a 64-column score row: columns [0,32) are read from TMEM twice.
// This kernel therefore has
//   * a RING of NBUF S/P buffers in TMEM shared by NWG softmax warpgroups (tile `it` is processed by warpgroup it % NWG in
//     buffer it % NBUF, NBUF > NWG): QK^T runs up to NBUF tiles ahead of the softmax, a warpgroup normally finds its next
//     S ready when it has stored P (packed fp16, in place over the S columns);
//   * TWO issuer warps with lean loops unrolled over one turn of their shared-memory ring (one asm statement per tile)
//     -- QK^T issuer (waits k_full + "PV of the buffer's previous use done") and P.V issuer (waits v_full + p_full) -- so
//     no wait of one contraction delays the other, and two single-thread TMA producers (K ring, V ring);
//   * the whole 64-column row in registers (one TMEM read per score), P stored half row by half row;
//   * an explicit pv_done[buffer] barrier: the softmax warps observe "PV of my previous tile has completed" before a lazy
//     rescale touches O and before the end-of-pass merge;
//   * a tile loop specialised per run (mixed / not mixed) with incremental ring bookkeeping: 345 instead of 500
//     instructions per warp and tile (ncu source counters, round 2), and the s_full wait probed ahead of time;
//   * PERSISTENT CTAs walking a static list of work items with all rings running on across items (see the kernel).
// Waiting on s_full / pv_done without having observed every earlier phase of that barrier is exact here: the tensor pipe
// completes in issue order and both issuers walk the tiles in order (see the notes at the waits).
// Tried and dropped in round 2 (profiles/experiments/r2_attn_ring_separate_p_ring.cuh.txt): a separate ring of fp16 P
// slots so that an S buffer is released as soon as the row is in registers -- one more mbarrier wait and arrive per tile
// cost the softmax warps more (each ~250 cycles of latency) than the earlier S bought (2.53 -> 2.91 ms at d = 40).
// Instantiations <DPAD, NWG, NBUF>: d<=40: <48,3,5> (one CTA per SM, 16 warps, 128 registers: three softmax warps per SM
// sub-partition), <48,2,4>, <48,1,3> (two CTAs per SM, 256 TMEM columns each); d<=80: <80,2,3>
// (the double-buffered Q of 2 x 32 KB leaves room for K / V rings of three 16 KB stages each).
#pragma once
#include <type_traits>

template <int DPAD_, int NWG_, int NBUF_> struct RCfg {
  static constexpr int DPAD = DPAD_, NWG = NWG_, NBUF = NBUF_;
  static_assert(NBUF > NWG || NWG == 1, "a warpgroup must find its next S ready: more S/P buffers than warpgroups");
  static constexpr int DPV = Cfg<DPAD, false>::DPV;
  static constexpr int NKT = (DPAD + BOX_COLS - 1) / BOX_COLS;
  // TMEM: ring of S/P buffers (fp32 scores, overwritten in place by the packed fp16 probabilities), then one O
  // accumulator per warpgroup
  static constexpr int TMEM_S = 0, TMEM_O = NBUF * BN;                 // O of warpgroup g at TMEM_O + g * DPV
  static constexpr int TMEM_USED = TMEM_O + NWG * DPV;
  static_assert(TMEM_USED <= 512, "TMEM budget");
  static constexpr int TMEM_COLS = TMEM_USED <= 256 ? 256 : 512;
  static constexpr int MIN_CTAS = TMEM_COLS == 256 ? 2 : 1;
  static constexpr int NUM_WARPS = 4 * NWG + 4;                        // softmax warps, K producer, QK issuer, PV issuer, V producer
  static constexpr int NUM_THREADS = 32 * NUM_WARPS;
  // Register rebalancing (setmaxnreg) for the 20-warp layout: ptxas caps the kernel at 65536 / 640 -> 96 registers; the four
  // auxiliary warps (one warpgroup) release down to 32, which lets the 16 softmax warps grow to 112 -- the whole 64-column
  // row stays in registers.  (512 * 112 + 128 * 32 = 61440 = 640 * 96.)
  static constexpr bool REBALANCE = false;
  static constexpr int REG_AUX = 32, REG_SOFTMAX = 112;
  static constexpr int SMEM_Q = NKT * TILE_BYTES;
  static constexpr int SMEM_K = NKT * KV_BYTES;                        // one K stage == one V stage
  static constexpr int ACC_LD = DPAD + 4;
  static constexpr int SMEM_ACC = BM * ACC_LD * 4;
  static constexpr int SMEM_MX = NWG > 1 ? NWG * BM * 4 : 0;
  static constexpr int SMEM_FIXED = 2 * SMEM_Q + SMEM_ACC + SMEM_MX + 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int SMEM_BUDGET = MIN_CTAS == 2 ? 113 * 1024 : 227 * 1024;
  static constexpr int NST_FIT = (SMEM_BUDGET - SMEM_FIXED) / (2 * SMEM_K);
  // K ring depth == V ring depth: a multiple of NBUF (the issuer loops are unrolled over one turn of the ring, so every
  // stage / buffer index in them is a compile-time constant), NBUF..8
  static constexpr int NST = (NST_FIT >= 2 * NBUF && 2 * NBUF <= 8) ? 2 * NBUF : NBUF;
  static_assert(NST <= NST_FIT, "shared memory budget: the K and V rings must hold NBUF tiles each");
  static constexpr int SMEM_BYTES = SMEM_FIXED + 2 * NST * SMEM_K;
};

// Ring timeline probe (-DFF_TIMELINE): [2 CTAs][4 roles][64 tiles][8 sites] clock64 stamps; roles: 0/1 = warp 0 of softmax
// warpgroup 0/1, 2 = QK issuer, 3 = PV issuer.
#ifdef FF_TIMELINE
#define RT_TL(role_, tile_, site_)                                                                                       \
  do {                                                                                                                   \
    if (tl && (tile_) < FF_TL_TILES) tl[(((size_t)tl_cta * 4 + (role_)) * FF_TL_TILES + (tile_)) * 8 + (site_)] = clock64(); \
  } while (0)
#define RT_TLW(wg_, tile_, site_) do { if ((wg_) < 2) RT_TL(wg_, tile_, site_); } while (0)
#else
#define RT_TL(role_, tile_, site_) do { } while (0)
#define RT_TLW(wg_, tile_, site_) do { } while (0)
#endif

// Enumerates the K/V tiles a CTA processes, in the order every role counts them: passes -> segments -> runs -> tiles,
// skipped runs left out.  Each producer walks its own.
struct TileIter {
  const FFAttnHeadPlan* plan;
  int n_pass, q0, n_kv_tiles;
  int ip, seg, jn, j, j1;
  bool seg_live, valid;
  PassCtx cx;
  SegCtx sg;
  __device__ __forceinline__ void init(const FFAttnHeadPlan* pl, int np, int q0_, int nkt, const KParams& p) {
    plan = pl; n_pass = np; q0 = q0_; n_kv_tiles = nkt;
    ip = -1; seg = 1; jn = 0; j = 0; j1 = 0; seg_live = false; valid = true;
    seek(p);
  }
  __device__ __forceinline__ void seek(const KParams& p) {          // find the next non-skipped run
#pragma unroll 1
    while (true) {
      if (seg_live && jn < n_kv_tiles) {
        const int cls = tile_class(sg, jn, p), e = run_end(sg, jn, cls, p), jb = jn;
        jn = e;
        if (tile_skip(cx, sg, cls, p.s_kv)) continue;
        j = jb; j1 = e;
        return;
      }
      if (ip >= 0 && seg == 0) {
        seg = 1; sg = cx.s1; seg_live = cx.active && sg.kv >= 0; jn = 0;
        continue;
      }
      if (++ip >= n_pass) { valid = false; return; }
      const FFAttnPass ps = plan->pass[ip];
      cx = make_ctx(ps, p, q0);
      seg = 0; sg = cx.s0; seg_live = cx.active && sg.kv >= 0; jn = 0;
    }
  }
  __device__ __forceinline__ void next(const KParams& p) { if (++j >= j1) seek(p); }
};

// cumulative tile counts per pass (all roles take the same decisions)
__device__ __forceinline__ int count_tiles(const FFAttnHeadPlan* plan, int n_pass, int q0, int n_kv_tiles, const KParams& p,
                                           int& pe0, int& pe1, int& pe2) {
  int n_total = 0;
  pe0 = pe1 = pe2 = 0;
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
  return n_total;
}

#ifndef FF_RING_POLY
#define FF_RING_POLY FF_POLY_PATTERN
#endif

// 16 scores -> 8 packed fp16 pairs, exp2 of pair i on the FMA pipe when bit i of PATTERN is set (see poly_exp2_x2)
template <bool MASKED, int PATTERN>
__device__ __forceinline__ void ring_chunk_f16(const float* s, uint32_t* pk, float sc, float nb, uint32_t bits) {
  const float2 sc2 = make_float2(sc, sc), nb2 = make_float2(nb, nb);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), sc2, nb2);
    float2 e;
    if (MASKED) {
      x.x = (bits >> (2 * i)) & 1u ? x.x : -INFINITY;
      x.y = (bits >> (2 * i + 1)) & 1u ? x.y : -INFINITY;
      e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
    } else if ((PATTERN >> i) & 1) {
      e = poly_exp2_x2(x);
    } else {
      e = make_float2(fast_exp2(x.x), fast_exp2(x.y));
    }
    pk[i] = pack_f16x2(e.x, e.y);
  }
}

// ---- lean issue blocks for the issuer warps: ONE asm statement per tile, executed by the whole warp (the election
// happens inside), with a handful of operands -- the per-K-step descriptor / TMEM-address arithmetic is done inside on
// PTX registers.  (With `if (leader) { mma; mma; ...; commit; commit; }` in C++ ptxas keeps ~17 precomputed operands in
// vector registers and moves them to uniform registers with R2UR for every tile.)
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!done) mbar_wait_slow(bar, parity);
}

// S = Q K^T: DPAD/16 K-steps; 16 channels = 32 B inside the 128-B row (descriptor units of 16 B: +2 per K-step), the
// second 64-channel box of Q / K is TILE_BYTES / KV_BYTES further (DPAD = 80: K-step 4).  Then the two commits.
template <int DPAD>
__device__ __forceinline__ void issue_qk(uint32_t sbuf, uint64_t qdesc, uint64_t kdesc, uint32_t idesc, uint32_t bar_s_,
                                         uint32_t bar_ke_) {
  static_assert(DPAD == 48 || DPAD == 80, "issue_qk: 3 or 5 K-steps");
  if constexpr (DPAD == 48) {
    asm volatile(
        "{\n\t.reg .pred e, pf, pt;\n\t.reg .b64 q1, q2, k1, k2;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 pf, 0, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 q1, %1, 2;\n\tadd.u64 q2, %1, 4;\n\tadd.u64 k1, %2, 2;\n\tadd.u64 k2, %2, 4;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pf;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q1, k1, %3, pt;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q2, k2, %3, pt;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
        ::"r"(sbuf), "l"(qdesc), "l"(kdesc), "r"(idesc), "r"(bar_s_), "r"(bar_ke_)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred e, pf, pt;\n\t.reg .b64 q1, q2, q3, q4, k1, k2, k3, k4;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 pf, 0, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
        "add.u64 q1, %1, 2;\n\tadd.u64 q2, %1, 4;\n\tadd.u64 q3, %1, 6;\n\tadd.u64 q4, %1, %6;\n\t"
        "add.u64 k1, %2, 2;\n\tadd.u64 k2, %2, 4;\n\tadd.u64 k3, %2, 6;\n\tadd.u64 k4, %2, %7;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pf;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q1, k1, %3, pt;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q2, k2, %3, pt;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q3, k3, %3, pt;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], q4, k4, %3, pt;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%4];\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t}"
        ::"r"(sbuf), "l"(qdesc), "l"(kdesc), "r"(idesc), "r"(bar_s_), "r"(bar_ke_), "n"(TILE_BYTES >> 4), "n"(KV_BYTES >> 4)
        : "memory");
  }
}

// O (+)= P V: four K-steps of 16 keys; A = packed fp16 P in TMEM at columns +0, +8, +32, +40 of the S/P buffer, B = V tile
// MN-major, 2048 B (128 descriptor units) per K-step.  acc0 = 0 starts the accumulator.  Then the two commits.
__device__ __forceinline__ void issue_pv(uint32_t obuf, uint32_t pbase, uint64_t vdesc, uint32_t idesc, uint32_t acc0,
                                         uint32_t bar_ve_, uint32_t bar_d_) {
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
      ::"r"(obuf), "r"(pbase), "l"(vdesc), "r"(idesc), "r"(acc0), "r"(bar_ve_), "r"(bar_d_)
      : "memory");
}

// named barriers: 1, 2 = end-of-pass merge (all softmax warps); 3 + g = exp token of warpgroup g
__device__ __forceinline__ void token_wait(int g) {
  switch (g) {
    case 0: asm volatile("bar.sync 3, 256;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 4, 256;" ::: "memory"); break;
    default: asm volatile("bar.sync 5, 256;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void token_pass(int g) {
  switch (g) {
    case 0: asm volatile("bar.arrive 3, 256;" ::: "memory"); break;
    case 1: asm volatile("bar.arrive 4, 256;" ::: "memory"); break;
    default: asm volatile("bar.arrive 5, 256;" ::: "memory"); break;
  }
}

// TOKEN: the exp2 sweeps of the warpgroups are serialised by a token that travels in tile order (named barriers): the
// warps that share an SM sub-partition then never run their MUFU-bound phase at the same time.  (Measured slower than
// free-running warpgroups in round 2 -- the extra named-barrier round trip costs more than the overlap gains; kept as an
// experiment switch.)
//
// PERSISTENT: the grid holds one CTA per SM (two for the 256-column layouts); a CTA walks the work items
// w = blockIdx.x, blockIdx.x + gridDim.x, ... (item = one 128-row query tile of one (stream, head), query tile fastest so
// that the CTAs running at the same time share K/V in L2).  All rings and their phase counters run on across items: the
// producers and the QK^T issuer are already loading / multiplying the next item while the softmax warps finish the
// current one, so the per-item prologue (Q load, first K tiles, pipeline fill) and epilogue (last P.V, merge, store) --
// ~12 % of a one-CTA-per-SM launch in the round-2 timelines -- disappear behind the main loop.  Q is double-buffered.
template <int DPAD, int NWG, int NBUF, bool TOKEN>
__global__ void __launch_bounds__(RCfg<DPAD, NWG, NBUF>::NUM_THREADS, RCfg<DPAD, NWG, NBUF>::MIN_CTAS)
attn_ring_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const KParams p) {
  using C = RCfg<DPAD, NWG, NBUF>;
  constexpr int NST = C::NST, SW = 4 * NWG;                            // SW: number of softmax warps
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                                       // Q double buffer
  const uint32_t sK = sQ + 2 * C::SMEM_Q;                              // K ring
  const uint32_t sV = sK + NST * C::SMEM_K;                            // V ring
  const uint32_t sACC = sV + NST * C::SMEM_K;
  const uint32_t sMX = sACC + C::SMEM_ACC;
  const uint32_t bar_base = sMX + C::SMEM_MX;
  const uint32_t bar_qf = bar_base;                                    // [2] Q of item i has landed (buffer i & 1)
  const uint32_t bar_qe = bar_base + 16;                               // [2] every QK^T of item i has completed
  const uint32_t bar_s = bar_base + 32;                                // [NBUF] S(t) = Q K^T has landed
  const uint32_t bar_p = bar_s + 8 * NBUF;                             // [NBUF] P(t) has been stored (4 warp arrivals)
  const uint32_t bar_d = bar_p + 8 * NBUF;                             // [NBUF] PV(t) has completed
  const uint32_t bar_kf = bar_d + 8 * NBUF, bar_ke = bar_kf + 8 * NST; // K ring full / empty
  const uint32_t bar_vf = bar_ke + 8 * NST, bar_ve = bar_vf + 8 * NST; // V ring full / empty
  const uint32_t tmem_slot = bar_ve + 8 * NST;
  static_assert(32 + 24 * NBUF + 32 * NST + 8 <= 512, "barrier block");
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* const acc_smem = reinterpret_cast<float*>(gen_base + (sACC - smem_base));
  float* const mx_smem = reinterpret_cast<float*>(gen_base + (sMX - smem_base));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.s_kv + BN - 1) / BN;
  const int n_items = p.n_qtiles * p.heads * p.n_streams;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qf + 8 * i, 1);
      mbar_init(bar_qe + 8 * i, 1);
    }
    for (int i = 0; i < NBUF; ++i) {
      mbar_init(bar_s + 8 * i, 1);
      mbar_init(bar_p + 8 * i, 4);
      mbar_init(bar_d + 8 * i, 1);
    }
    for (int i = 0; i < NST; ++i) {
      mbar_init(bar_kf + 8 * i, 1);
      mbar_init(bar_ke + 8 * i, 1);
      mbar_init(bar_vf + 8 * i, 1);
      mbar_init(bar_ve + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == SW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // (warp-uniform by construction; the shuffle lets the compiler keep it in uniform registers)
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);
#ifdef FF_TIMELINE
  const int tl_cta = blockIdx.x == 5 ? 0 : 1;
  const int tl_role = warp == 0 ? 0 : (warp == 4 && NWG > 1 ? 1 : (warp == SW + 1 ? 2 : (warp == SW + 2 ? 3 : -1)));
  unsigned long long* const tl = ((blockIdx.x == 5 || blockIdx.x == 77) && lane == 0 && tl_role >= 0) ? g_timeline : nullptr;
#endif
  // work item -> (query tile, head, stream)
  auto item_coords = [&](int w, int& q0, int& head, int& stream) {
    const int qt = w % p.n_qtiles, hs = w / p.n_qtiles;
    q0 = qt * BM;
    head = hs % p.heads;
    stream = hs / p.heads;
  };
  auto item_plan = [&](int head, int stream, int& n_pass) {
    const FFAttnHeadPlan* plan = p.plan + (size_t)stream * p.heads + head;
    n_pass = __ldg(&plan->n_pass);
    n_pass = n_pass < 0 ? 0 : (n_pass > FF_MAX_PASS ? FF_MAX_PASS : n_pass);
    return plan;
  };

  if constexpr (C::REBALANCE) {
    if (warp >= SW) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::REG_AUX));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::REG_SOFTMAX));
  }
  if (warp == SW || warp == SW + 3) {
    // ===================================== TMA producers ====================================
    // Two single-thread producers (every mbarrier operation of a thread costs ~100-250 cycles of latency, and one thread
    // serving both rings cannot keep up with a tile every ~500 cycles): warp SW loads Q and the K ring, warp SW+3 the
    // V ring.  Each walks the tile sequence with its own iterator and only waits for its own ring's `empty` barriers.
    if (lane == 0) {
      const bool is_k = warp == SW;
      const CUtensorMap* tm = is_k ? &tm_k : &tm_v;
      const uint32_t ring = is_k ? sK : sV, bfull = is_k ? bar_kf : bar_vf, bempty = is_k ? bar_ke : bar_ve;
      int stage = 0, use = 0, li = 0;                                  // li: local item counter of this CTA
#pragma unroll 1
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++li) {
        int q0, head, stream, n_pass;
        item_coords(w, q0, head, stream);
        const FFAttnHeadPlan* plan = item_plan(head, stream, n_pass);
        if (is_k) {
          const int qb = li & 1;
          if (li >= 2) mbar_wait(bar_qe + 8 * qb, ((li >> 1) - 1) & 1);        // QK^T of item li-2 done with this buffer
          mbar_expect_tx(bar_qf + 8 * qb, C::NKT * TILE_BYTES);
          for (int kt = 0; kt < C::NKT; ++kt)
            tma_load_4d(sQ + qb * C::SMEM_Q + kt * TILE_BYTES, &tm_q, kt * BOX_COLS, head, q0, stream, bar_qf + 8 * qb);
        }
        TileIter ti;
        ti.init(plan, n_pass, q0, n_kv_tiles, p);
#pragma unroll 1
        while (ti.valid) {
          if (use > 0) mbar_wait(bempty + 8 * stage, (use - 1) & 1);
          const uint32_t full = bfull + 8 * stage, dst = ring + stage * C::SMEM_K;
          mbar_expect_tx(full, C::NKT * KV_BYTES);
          for (int kt = 0; kt < C::NKT; ++kt)
            tma_load_4d(dst + kt * KV_BYTES, tm, kt * BOX_COLS, head, ti.j * BN, ti.sg.kv, full);
          if (++stage == NST) { stage = 0; ++use; }
          ti.next(p);
        }
      }
    }
  } else if (warp == SW + 1) {
    // ===================================== QK^T issuer ======================================
    // S[b] = Q K(t)^T as soon as K(t) has landed and the P.V that read buffer b's previous contents has completed.
    // The issuer warps share their SM sub-partitions with the softmax warps and get an issue slot only every few cycles,
    // so what bounds them is their INSTRUCTION COUNT per tile (round-2 timelines: ~75 instructions = ~800 cycles per
    // tile with 3 softmax warps per sub-partition).  The tile loop is therefore unrolled over one turn of the K ring
    // (NST tiles, a multiple of NBUF) with the ring position `u` carried across work items: stage, S buffer, barrier
    // addresses and every descriptor are compile-time offsets inside a switch on u, phases flip once per turn.
    constexpr uint32_t idesc_qk = make_idesc(BN, 0);
    const uint64_t qdesc00 = smem_desc_sw128(sQ, 16);
    const uint64_t kdesc0 = smem_desc_sw128(sK, 16);
    uint32_t kph = 0, bph0 = 0;      // parity of this turn's k_full phases / of the buffer use at ring position 0
    int u = 0, tg = 0, li = 0;       // ring position, global tile counter, local item counter
#pragma unroll 1
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++li) {
      int q0, head, stream, n_pass, pe0, pe1, pe2;
      item_coords(w, q0, head, stream);
      const FFAttnHeadPlan* plan = item_plan(head, stream, n_pass);
      int left = __shfl_sync(0xffffffffu, count_tiles(plan, n_pass, q0, n_kv_tiles, p, pe0, pe1, pe2), 0);
      const int qb = li & 1;
      const uint64_t qdesc0 = qdesc00 + (uint64_t)((qb * C::SMEM_Q) >> 4);
      mbar_wait(bar_qf + 8 * qb, (li >> 1) & 1);
#pragma unroll 1
      while (left > 0) {
#pragma unroll
        for (int uu = 0; uu < NST; ++uu) {
          if (uu < u) continue;                    // resume the turn where the previous item stopped (u is uniform)
          if (left == 0) break;
          const int b = uu % NBUF;
          const uint32_t bph = bph0 ^ (uint32_t)((uu / NBUF) & 1);
          RT_TL(2, tg, 0);
          mbar_wait_lean(bar_kf + 8 * uu, kph);
          RT_TL(2, tg, 1);
          if (tg >= NBUF) mbar_wait_lean(bar_d + 8 * b, bph ^ 1u);
          RT_TL(2, tg, 2);
          tc_fence_after();
          issue_qk<DPAD>(tmem + C::TMEM_S + BN * b, qdesc0, kdesc0 + (uint64_t)((uu * C::SMEM_K) >> 4), idesc_qk,
                         bar_s + 8 * b, bar_ke + 8 * uu);
          __syncwarp();
          RT_TL(2, tg, 3);
          --left;
          ++tg;
          u = uu + 1;
        }
        if (u == NST) {
          u = 0;
          kph ^= 1u;
          bph0 ^= (uint32_t)((NST / NBUF) & 1);
        }
      }
      // every QK^T of this item has been issued: its Q buffer is free when they have completed
      if (elect_one()) tc_commit(bar_qe + 8 * qb);
      __syncwarp();
    }
  } else if (warp == SW + 2) {
    // ===================================== P.V issuer =======================================
    // O_g (+)= P(t) V(t): A = P in TMEM over the S columns of buffer b (8 packed fp16 columns per 16 keys at
    // 32*(k16/2) + 8*(k16%2)), B = V tile, MN-major (16 keys = 2048 B per K-step, 64-channel groups KV_BYTES apart).
    // Unrolled over one turn of the V ring like the QK^T issuer.
    constexpr uint32_t idesc_pv = make_idesc(C::DPV, 1, true);
    const uint64_t vdesc0 = smem_desc_sw128(sV, KV_BYTES);
    uint32_t vph = 0, bph0 = 0, g = 0;
    int u = 0, tg = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      int q0, head, stream, n_pass, pe0, pe1, pe2;
      item_coords(w, q0, head, stream);
      const FFAttnHeadPlan* plan = item_plan(head, stream, n_pass);
      const int n_total = __shfl_sync(0xffffffffu, count_tiles(plan, n_pass, q0, n_kv_tiles, p, pe0, pe1, pe2), 0);
      pe0 = __shfl_sync(0xffffffffu, pe0, 0);
      pe1 = __shfl_sync(0xffffffffu, pe1, 0);
      pe2 = __shfl_sync(0xffffffffu, pe2, 0);
      int t = 0, pi = 0, pend = pe0, pstart = 0;      // tile within the item; current pass = tiles [pstart, pend)
      g = 0;                                          // tile t of an item belongs to warpgroup t % NWG
#pragma unroll 1
      while (t < n_total) {
#pragma unroll
        for (int uu = 0; uu < NST; ++uu) {
          if (uu < u) continue;
          if (t >= n_total) break;
          const int b = uu % NBUF;
          const uint32_t bph = bph0 ^ (uint32_t)((uu / NBUF) & 1);
          while (t >= pend && pi < 3) {          // (rare) next pass; empty passes have pend == pstart
            pstart = pend;
            ++pi;
            pend = pi == 1 ? pe1 : (pi == 2 ? pe2 : n_total);
          }
          // the first tile of each warpgroup in a pass starts its accumulator
          const uint32_t acc0 = (t - pstart < NWG) ? 0u : 1u;
          RT_TL(3, tg, 0);
          mbar_wait_lean(bar_vf + 8 * uu, vph);
          RT_TL(3, tg, 1);
          mbar_wait_lean(bar_p + 8 * b, bph);
          RT_TL(3, tg, 2);
          tc_fence_after();
          issue_pv(tmem + C::TMEM_O + C::DPV * g, tmem + C::TMEM_S + BN * b, vdesc0 + (uint64_t)((uu * C::SMEM_K) >> 4),
                   idesc_pv, acc0, bar_ve + 8 * uu, bar_d + 8 * b);
          __syncwarp();
          RT_TL(3, tg, 3);
          if (++g == (uint32_t)NWG) g = 0;
          ++t;
          ++tg;
          u = uu + 1;
        }
        if (u == NST) {
          u = 0;
          vph ^= 1u;
          bph0 ^= (uint32_t)((NST / NBUF) & 1);
        }
      }
    }
  } else if (warp < SW) {
    // ===================================== softmax + epilogue ===============================
    // Warpgroup wg owns the tiles with (global index) % NWG == wg: S/P buffer index % NBUF, accumulator O_wg, its own
    // running reference point.  A thread = one query row (TMEM lane).
    const int wq = warp & 3;        // TMEM lane quarter of this warp (hardware rule: warp w reaches lanes 32*(w%4)..+31)
    const int wg = warp >> 2;
    const int rloc = 32 * wq + lane;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t tO = tlane + C::TMEM_O + C::DPV * wg;
    float* const acc_row = acc_smem + (size_t)rloc * C::ACC_LD;
    // Tile t of a work item (all roles count alike) is processed by warpgroup t % NWG -- the split of an item over the
    // warpgroups, and with it the summation order, does not depend on what the CTA processed before, so a batched launch
    // is bit-identical to the same edit run alone -- in S/P buffer (base + t) % NBUF, base = tiles of all earlier items.
    int base = 0;
    int my_b = 0, prev_b = -1;      // S/P buffer of my next tile; of my previous tile (its PV must have completed before
    uint32_t my_ph = 0, prev_ph = 0;   // O_wg is touched), with the parities of those buffer uses
    // An mbarrier wait costs the waiting warp ~250 cycles even when the phase completed long ago (round-2 timelines), so
    // s_full of my next tile is PROBED with the non-blocking test_wait before the exp sweep of the current one; the
    // blocking wait runs only when the probe failed.
    bool s_ready = false;
    if (TOKEN && NWG > 1 && wg == NWG - 1) token_pass(0);
#pragma unroll 1
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      int q0, head, stream, n_pass;
      item_coords(w, q0, head, stream);
      const FFAttnHeadPlan* plan = item_plan(head, stream, n_pass);
      const int row = q0 + rloc;
      bool acc_started = false;
      int it = 0;                   // tile counter within the item
      int my_next = wg;             // my next tile of this item
      my_b = (base + wg) % NBUF;
      my_ph = (uint32_t)(((base + wg) / NBUF) & 1);
      s_ready = false;              // (the probe of the previous item's last tile looked NWG tiles ahead: not my tile)
#pragma unroll 1
      for (int ip = 0; ip < n_pass; ++ip) {
        const FFAttnPass ps = plan->pass[ip];
        const PassCtx cx = make_ctx(ps, p, q0);
        if (!cx.active) continue;
        uint32_t rb = 0;
        if (ps.row_mask >= 0 && row < p.s_q)
          rb = (__ldg(p.bitmasks + (size_t)ps.row_mask * p.mask_words + (row >> 5)) >> (row & 31)) & 1u;
        const bool rowflip = cx.rowxor && rb;
        float m_used = -INFINITY;     // reference point of MY tiles (log2 units); -inf: nothing read so far
        int n_mine = 0;
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
            const int it_run = it;
            it += j1 - jb;
            // ---- one tile of mine: MIXED (boundary / ragged tile: per-element masks) or not is a property of the run,
            // so the hot loop of a run carries no mask logic at all
            auto tile = [&](auto mixed_tag, int j) {
              constexpr bool MIXED = decltype(mixed_tag)::value;
              const uint32_t tS = tlane + C::TMEM_S + BN * my_b;
              RT_TLW(wg, my_next, 0);
              // (exact without observing every phase of s_full[buffer]: the tensor pipe completes in issue order, so the
              // QK^T of the buffer's previous tile completed before the QK^T of my previous tile, whose S I have read;
              // the buffer's next QK^T needs the P.V of THIS tile)
              if (!s_ready) mbar_wait_lean(bar_s + 8 * my_b, my_ph);
              RT_TLW(wg, my_next, 1);
              tc_fence_after();
              // ---- the whole 64-column row: both loads in flight together
              float s0[32], s1[32];
              tmem_ld32(tS, s0);
              tmem_ld32(tS + 32, s1);
              // ---- allowed-key bits of this row for MIX tiles (computed under the load latency)
              uint32_t kb_lo = 0xffffffffu, kb_hi = 0xffffffffu;
              if constexpr (MIXED) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const int kbase = j * BN + 32 * h;
                  const int rem = p.s_kv - kbase;
                  const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
                  uint32_t kb = 0xffffffffu;
                  if (sg.kmask >= 0 && !uniform && rem > 0) {
                    if (sg.prefix) {
                      const int tt = sg.T - kbase;
                      kb = tt >= 32 ? 0xffffffffu : (tt <= 0 ? 0u : ((1u << tt) - 1u));
                    } else {
                      kb = __ldg(p.bitmasks + (size_t)sg.kmask * p.mask_words + (kbase >> 5));   // BN % 32 == 0
                    }
                    if (flip) kb = ~kb;
                  }
                  if (h == 0) kb_lo = kb & valid; else kb_hi = kb & valid;
                }
              }
              tmem_wait_ld32(s0);
              tmem_wait_ld32(s1);
              RT_TLW(wg, my_next, 2);
              // ---- row max over the ALLOWED keys of the tile (the fp16 P operand has a narrow exponent range: the
              // reference point must not come from keys this row does not read)
              float mt;
              {
                float m0, m1, m2, m3;
                if constexpr (!MIXED) {
                  m0 = fmaxf(s0[0], s0[1]);
                  m1 = fmaxf(s0[2], s0[3]);
                  m2 = fmaxf(s1[0], s1[1]);
                  m3 = fmaxf(s1[2], s1[3]);
#pragma unroll
                  for (int i = 4; i < 32; i += 4) {
                    m0 = fmaxf(m0, fmaxf(s0[i], s0[i + 1]));
                    m1 = fmaxf(m1, fmaxf(s0[i + 2], s0[i + 3]));
                    m2 = fmaxf(m2, fmaxf(s1[i], s1[i + 1]));
                    m3 = fmaxf(m3, fmaxf(s1[i + 2], s1[i + 3]));
                  }
                } else {
                  m0 = m1 = m2 = m3 = -INFINITY;
#pragma unroll
                  for (int i = 0; i < 32; i += 2) {
                    m0 = fmaxf(m0, (kb_lo >> i) & 1u ? s0[i] : -INFINITY);
                    m1 = fmaxf(m1, (kb_lo >> (i + 1)) & 1u ? s0[i + 1] : -INFINITY);
                    m2 = fmaxf(m2, (kb_hi >> i) & 1u ? s1[i] : -INFINITY);
                    m3 = fmaxf(m3, (kb_hi >> (i + 1)) & 1u ? s1[i + 1] : -INFINITY);
                  }
                }
                mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
              }
              // (a row that reads nothing from this run keeps its reference point: -inf never grows it)
              const float mts = (!MIXED && !row_ok_cls) ? -INFINITY : (uniform ? 0.f : mt * p.scale_log2);
              // ---- running reference point (lazy rescale: O_wg is read-modify-written only when the max grew a lot)
              float alpha = 1.f;
              bool grow = false;
              if (n_mine == 0) {
                m_used = mts;
              } else if (mts > m_used + rescale_threshold<false>()) {
                alpha = fast_exp2(m_used - mts);       // (m_used = -inf, nothing read so far: alpha = 0, O is 0 anyway)
                m_used = mts;
                grow = true;
              }
              RT_TLW(wg, my_next, 3);
              // probe: has the S of my next tile landed?  Consumed at the top of the next tile.
              int nb_ = my_b + NWG;
              uint32_t nph = my_ph;
              if (nb_ >= NBUF) { nb_ -= NBUF; nph ^= 1u; }
              s_ready = mbar_test(bar_s + 8 * nb_, nph);
              // ---- p = 2^(s*scale*log2e - m) as packed fp16 pairs over the S columns of the same keys: keys
              // [32*hb, 32*hb+32) -> 16 packed columns at 32*hb.  (A row that reads nothing from a non-mixed run gets p = 0
              // through a large negative offset: ex2 underflows to 0 on the MUFU unit and the FMA-pipe polynomial clamps
              // its argument; m_used may still be -inf there.)
              const float nb = (MIXED || row_ok_cls) ? -m_used : -60000.f;
              if (TOKEN && NWG > 1) token_wait(wg);
#pragma unroll
              for (int hb = 0; hb < 2; ++hb) {
                const float* sv = hb ? s1 : s0;
                uint32_t pk[16];
                if constexpr (!MIXED) {
#pragma unroll
                  for (int jj = 0; jj < 2; ++jj) ring_chunk_f16<false, FF_RING_POLY>(sv + 16 * jj, pk + 8 * jj, sc, nb, 0u);
                } else {
                  const uint32_t kbits = hb ? kb_hi : kb_lo;
#pragma unroll
                  for (int jj = 0; jj < 2; ++jj)
                    ring_chunk_f16<true, 0>(sv + 16 * jj, pk + 8 * jj, sc, nb, (kbits >> (16 * jj)) & 0xffffu);
                }
                // (the scores of this tile are in registers: P may overwrite the S columns half by half)
                tmem_st16(tS + 32 * hb, pk);
              }
              if (TOKEN && NWG > 1) token_pass((wg + 1) % NWG);
              RT_TLW(wg, my_next, 4);
              // ---- (rare) PV of my previous tile has completed: O_wg may be rescaled, and the accumulate of PV(this
              // tile) will see the rescaled values.  Exact although this warp does not observe every phase of
              // pv_done[prev_b]: the QK^T that filled that buffer for my previous tile waited for the phase before, and
              // the buffer's next use needs a P.V that the in-order issuer has not issued yet (it is still waiting for
              // THIS tile's P).
              if (__any_sync(0xffffffffu, grow)) {
                if (prev_b >= 0) mbar_wait(bar_d + 8 * prev_b, prev_ph);
                tc_fence_after();
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
              }
              tmem_wait_st();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_p + 8 * my_b);
              RT_TLW(wg, my_next, 5);
              ++n_mine;
              // next tile of mine: NWG tiles further (NWG < NBUF: the buffer index wraps at most once)
              prev_b = my_b;
              prev_ph = my_ph;
              my_next += NWG;
              my_b = nb_;
              my_ph = nph;
            };
            if (cls == TILE_MIX) {
#pragma unroll 1
              while (my_next < it) tile(std::true_type{}, jb + (my_next - it_run));
            } else {
#pragma unroll 1
              while (my_next < it) tile(std::false_type{}, 0);
            }
          }
        }
        if (it == it_pass0) continue;     // (defensive) no tile of this pass was processed: nothing to add (CTA-uniform)
        // ---- end of pass: acc += weight * roww / l * O (merging the partial softmaxes of the warpgroups)
        // (1) my last PV has landed
        if (n_mine > 0) mbar_wait(bar_d + 8 * prev_b, prev_ph);
        tc_fence_after();
        // (2) exchange the reference points (row-wise, through shared memory); a_g = 2^(m_g - m_all) per warpgroup g.
        // tcgen05.ld is warp-collective: TMEM reads are guarded by CTA-uniform tile counts only (an accumulator that no
        // PV of this pass wrote is not read); rows that read nothing (reference point -inf) get weight 0 per lane
        float cg_[NWG];                // coef * a_g / l
        bool wrote[NWG];
        if constexpr (NWG > 1) {
          mx_smem[wg * BM + rloc] = n_mine > 0 ? m_used : -INFINITY;
          tc_fence_before();
          named_bar_sync(1, 32 * SW);
          tc_fence_after();
          float mg[NWG], m_all = -INFINITY;
#pragma unroll
          for (int g = 0; g < NWG; ++g) {
            const int first = it_pass0 + ((g - it_pass0) % NWG + NWG) % NWG;     // first tile of warpgroup g in this pass
            wrote[g] = first < it;
            mg[g] = mx_smem[g * BM + rloc];
            m_all = fmaxf(m_all, mg[g]);
          }
          // (3) denominators from the ones column of each accumulator
          float l = 0.f;
#pragma unroll
          for (int g = 0; g < NWG; ++g) {
            const float a = (wrote[g] && mg[g] > -INFINITY) ? fast_exp2(mg[g] - m_all) : 0.f;
            cg_[g] = a;
            if (wrote[g]) l = fmaf(a, tmem_ld1_wait(tlane + C::TMEM_O + C::DPV * g + p.head_dim), l);
          }
          float coef = ps.weight;
          if (ps.flags & FF_PASS_ROW_WEIGHT) coef = rb ? coef : 0.f;
          coef = l > 0.f ? coef / l : 0.f;
#pragma unroll
          for (int g = 0; g < NWG; ++g) cg_[g] *= coef;
        } else {
          wrote[0] = n_mine > 0;
          const float l = (wrote[0] && m_used > -INFINITY) ? tmem_ld1_wait(tO + p.head_dim) : 0.f;
          float coef = ps.weight;
          if (ps.flags & FF_PASS_ROW_WEIGHT) coef = rb ? coef : 0.f;
          cg_[0] = l > 0.f ? coef / l : 0.f;
        }
        // (4) my share of the channels: the 16-channel chunks c with c % NWG == wg, from every accumulator
#pragma unroll
        for (int c = 0; c < DPAD / 16; ++c) {
          if (c % NWG != wg) continue;
          float r[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = acc_started ? acc_row[16 * c + i] : 0.f;
#pragma unroll
          for (int g = 0; g < NWG; ++g) {
            if (!wrote[g]) continue;
            float o[16];
            tmem_ld16(tlane + C::TMEM_O + C::DPV * g + 16 * c, o);
            tmem_wait_ld16(o);
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = fmaf(cg_[g], o[i], r[i]);      // (0 for rows that read nothing)
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(acc_row + 16 * c + i) = make_float4(r[i], r[i + 1], r[i + 2], r[i + 3]);
        }
        acc_started = true;
        // (5) NWG > 1: every accumulator has been read by everybody before the next pass may overwrite it.  (NWG == 1:
        // PV of the next pass needs p_full from all four warps, each of which arrives after its own reads.)
        tc_fence_before();
        if constexpr (NWG > 1) named_bar_sync(2, 32 * SW);
        else __syncwarp();
        tc_fence_after();
      }
      base += it;
      // ---- write the row: out[stream, row, head*d : (head+1)*d]; each thread writes the chunks it accumulated
      {
        const bool row_ok = row < p.s_q;
        const size_t o_off = ((size_t)stream * p.s_q + (row_ok ? row : 0)) * ((size_t)p.heads * p.head_dim) +
                             (size_t)head * p.head_dim;
#pragma unroll
        for (int c = 0; c < DPAD / 16; ++c) {
          if (c % NWG != wg) continue;
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
  }
  // ---- teardown: every tcgen05 op of this CTA has completed (the softmax warps observed pv_done of their last tiles,
  // and a P.V only runs after the QK^T that produced its S)
  tc_fence_before();
  __syncthreads();
  if (warp == SW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}
