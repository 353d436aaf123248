// EXPERIMENTAL (next-round work, NOT in the product build): one-CTA-per-SM variant of the d<=40, fp16-P attention kernel.
// Compiled only with -DFF_ONE_CTA (python -m freefine_b200.csrc.build --variant=onecta -DFF_ONE_CTA) and then selected by
// ff_attn_masked_kv for 8 < head_dim <= 40 with fp16-staged V; every other shape keeps attn_masked_kv_kernel.
// Written at the end of round 1 WITHOUT GPU time left: it compiles for sm_100a (no spills) but has never run.  First
// thing to do with it: FREEFINE_B200_LIB=freefine_b200/lib/libfreefine_b200_onecta.so python -m pytest
// tests/test_gpu_attention.py -m gpu -x, then profiles/attn_case.py.
//
// Why ("Where the attention kernel stands"): the product kernel retires a 128x64 tile per SM every ~740 cycles
// although the MUFU unit / issue ports would allow ~330: per tile a softmax warp exposes three tcgen05.ld round trips (96
// registers cannot hold a 64-column row) and the p_full -> PV(t), QK(t+2) -> s_full hand-shake (~1100 cycles), because a
// warpgroup owns ONE S buffer.  Here a CTA owns the whole SM:
//   * 512 TMEM columns: TWO S/P buffers per warpgroup (4 x 64) + one O accumulator per warpgroup (2 x 48) = 352;
//   * the issuer runs QK four tiles ahead: QK(t+4) goes out right behind PV(t) into the buffer PV(t) has just consumed,
//     so S(t+2) -- the warpgroup's next tile -- has been ready long before its softmax(t) arrives on p_full;
//   * ~200 registers per thread: the whole 64-column row is loaded once (two tcgen05.ld in flight together);
//   * "PV of my previous tile has completed" (needed before a lazy rescale touches O, and before the end-of-pass merge) is
//     an explicit per-warpgroup barrier o_done, waited for EVERY tile (an mbarrier wait is only exact when the waiter
//     observes every phase) but late -- after the exp sweep -- when it has normally completed long ago;
//   * K/V ring of 8 stages (tile t+4's K must be resident while tile t's V is still being read).
// Next steps once it runs: (1) issue the tcgen05.ld of S(n+1) right after the P(n) store when s_full(n+1) has already
// completed (mbar_test), so that the only exposed TMEM round trip per tile disappears too; (2) a d = 80 instantiation
// (4 x 64 + 2 x 96 = 448 columns; a 4-stage K/V ring is what fits beside Q and the cross-pass accumulator).
// Global tile index `it` (all roles count alike): warpgroup = it & 1, S/P buffer sb(it) = 2*(it&1) + ((it>>1)&1), and the
// phase parity of s_full[sb] / p_full[sb] for tile it is (it>>2)&1.
#pragma once

struct OC {
  static constexpr int DPAD = 48, DPV = 48, NSTAGE = 8;
  static constexpr int TMEM_S = 0, TMEM_O = 4 * BN, TMEM_COLS = 512;
  static constexpr int SMEM_Q = TILE_BYTES;
  static constexpr int SMEM_STAGE = 2 * KV_BYTES;                        // K tile then V tile
  static constexpr int ACC_LD = DPAD + 4;
  static constexpr int SMEM_ACC = BM * ACC_LD * 4;
  static constexpr int SMEM_MX = 2 * BM * 4;
  static constexpr int SMEM_BYTES = SMEM_Q + NSTAGE * SMEM_STAGE + SMEM_ACC + SMEM_MX + 1024 + 512;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(TMEM_O + 2 * DPV <= TMEM_COLS, "TMEM budget");
};

__device__ __forceinline__ int oc_sbuf(int it) { return 2 * (it & 1) + ((it >> 1) & 1); }

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_onecta_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const KParams p) {
  using C = OC;
  constexpr int DPAD = C::DPAD;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + C::SMEM_Q;
  const uint32_t sACC = sKV + C::NSTAGE * C::SMEM_STAGE;
  const uint32_t sMX = sACC + C::SMEM_ACC;
  const uint32_t bar_base = sMX + C::SMEM_MX;
  const uint32_t bar_q = bar_base;
  const uint32_t bar_s = bar_base + 8;             // [4] s_full per S/P buffer
  const uint32_t bar_p = bar_base + 40;            // [4] p_full per S/P buffer
  const uint32_t bar_o = bar_base + 72;            // [2] o_done per warpgroup: one phase per PV of that warpgroup
  const uint32_t bar_kv_full = bar_base + 88, bar_kv_empty = bar_base + 88 + 8 * C::NSTAGE;
  const uint32_t tmem_slot = bar_base + 88 + 16 * C::NSTAGE;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* const acc_smem = reinterpret_cast<float*>(gen_base + (sACC - smem_base));
  float* const mx_smem = reinterpret_cast<float*>(gen_base + (sMX - smem_base));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, head = blockIdx.y, stream = blockIdx.z;
  const FFAttnHeadPlan* plan = p.plan + (size_t)stream * p.heads + head;
  int n_pass = __ldg(&plan->n_pass);
  n_pass = n_pass < 0 ? 0 : (n_pass > FF_MAX_PASS ? FF_MAX_PASS : n_pass);
  const int n_kv_tiles = (p.s_kv + BN - 1) / BN;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_s + 8 * i, 1);
      mbar_init(bar_p + 8 * i, NUM_SOFTMAX_WARPS / 2);
    }
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

  if (warp == NUM_SOFTMAX_WARPS) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      mbar_expect_tx(bar_q, TILE_BYTES);
      tma_load_4d(sQ, &tm_q, 0, head, q0, stream, bar_q);
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
              if (use > 0) mbar_wait_hot<2>(bar_kv_empty + 8 * stage, (use - 1) & 1);
              const uint32_t full = bar_kv_full + 8 * stage;
              const uint32_t sK = sKV + stage * C::SMEM_STAGE, sV = sK + KV_BYTES;
              mbar_expect_tx(full, 2 * KV_BYTES);
              tma_load_4d(sK, &tm_k, 0, head, j * BN, sg.kv, full);
              tma_load_4d(sV, &tm_v, 0, head, j * BN, sg.kv, full);
              ++it;
            }
          }
        }
      }
    }
  } else if (warp == NUM_SOFTMAX_WARPS + 1) {
    // ===================================== MMA issuer =======================================
    const bool leader = elect_one();
    constexpr uint32_t idesc_qk = make_idesc(BN, 0);
    constexpr uint32_t idesc_pv = make_idesc(C::DPV, 1, true);
    const uint64_t qdesc0 = smem_desc_sw128(sQ, 16);
    const uint64_t kdesc0 = smem_desc_sw128(sKV, 16);
    const uint64_t vdesc0 = smem_desc_sw128(sKV + KV_BYTES, KV_BYTES);
    int pe0 = 0, pe1 = 0, pe2 = 0, n_total = 0;
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
    mbar_wait(bar_q, 0);
    int qk_t = 0;
    uint32_t q_stage = 0, q_phase = 0, p_stage = 0;
    auto wait_kv = [&]() { mbar_wait_hot<1>(bar_kv_full + 8 * q_stage, q_phase); };
    auto issue_qk = [&]() {                         // leader only; kv_full(qk_t) has been observed by the warp
      const uint64_t kd = kdesc0 + ((q_stage * (uint32_t)C::SMEM_STAGE) >> 4);
      const int sb = oc_sbuf(qk_t);
      const uint32_t sbuf = tmem + C::TMEM_S + BN * sb;
#pragma unroll
      for (int ks = 0; ks < DPAD / 16; ++ks)
        mma_ss(sbuf, qdesc0 + ((ks * 32) >> 4), kd + ((ks * 32) >> 4), idesc_qk, ks > 0);
      tc_commit(bar_s + 8 * sb);
    };
    auto advance_qk = [&]() {
      ++qk_t;
      if (++q_stage == (uint32_t)C::NSTAGE) { q_stage = 0; q_phase ^= 1u; }
    };
    // prologue: S of the first two tiles of each warpgroup
#pragma unroll 1
    for (int k = 0; k < 4 && qk_t < n_total; ++k) {
      wait_kv();
      tc_fence_after();
      if (leader) issue_qk();
      __syncwarp();
      advance_qk();
    }
#pragma unroll 1
    for (int t = 0; t < n_total; ++t) {
      const bool more = qk_t < n_total;             // (qk_t == t + 4 while there are tiles left)
      const int pstart = t >= pe2 ? pe2 : (t >= pe1 ? pe1 : (t >= pe0 ? pe0 : 0));
      const bool first_of_wg = t - pstart < 2;      // first tile of its warpgroup in this pass: starts the accumulator
      const uint64_t vd = vdesc0 + ((p_stage * (uint32_t)C::SMEM_STAGE) >> 4);
      const int sb = oc_sbuf(t);
      const uint32_t pbase = tmem + C::TMEM_S + BN * sb;
      const uint32_t obuf = tmem + C::TMEM_O + C::DPV * (t & 1);
      bool kv_ready = more && __all_sync(0xffffffffu, mbar_test(bar_kv_full + 8 * q_stage, q_phase));
      mbar_wait_hot<1>(bar_p + 8 * sb, (t >> 2) & 1);
      if (more && !kv_ready) kv_ready = __all_sync(0xffffffffu, mbar_test(bar_kv_full + 8 * q_stage, q_phase));
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks)
          mma_ts(obuf, pbase + 32 * (ks >> 1) + 8 * (ks & 1), vd + ((ks * 2048) >> 4), idesc_pv,
                 (!first_of_wg || ks > 0) ? 1u : 0u);
        tc_commit(bar_kv_empty + 8 * p_stage);
        tc_commit(bar_o + 8 * (t & 1));
        if (kv_ready) issue_qk();                   // QK(t+4) into the buffer PV(t) has just been queued to consume
      }
      __syncwarp();
      if (more && !kv_ready) {
        wait_kv();
        tc_fence_after();
        if (leader) issue_qk();
        __syncwarp();
      }
      if (more) advance_qk();
      if (++p_stage == (uint32_t)C::NSTAGE) p_stage = 0;
    }
    __syncwarp();
  } else {
    // ===================================== softmax + epilogue ===============================
    const int wq = warp & 3;
    const int wg = warp >> 2;
    const int rloc = 32 * wq + lane;
    const int row = q0 + rloc;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t tO = tlane + C::TMEM_O + C::DPV * wg;
    const uint32_t tOx = tlane + C::TMEM_O + C::DPV * (wg ^ 1);
    float* const acc_row = acc_smem + (size_t)rloc * C::ACC_LD;
    bool acc_started = false;
    int it = 0;                     // global tile counter (all roles count alike)
    int m_glob = 0;                 // my tiles processed so far, all passes (= PVs of my warpgroup requested so far)
    int o_seen = 0;                 // o_done phases of my warpgroup observed so far (m_glob - 1 or m_glob)
#pragma unroll 1
    for (int ip = 0; ip < n_pass; ++ip) {
      const FFAttnPass ps = plan->pass[ip];
      const PassCtx cx = make_ctx(ps, p, q0);
      if (!cx.active) continue;
      uint32_t rb = 0;
      if (ps.row_mask >= 0 && row < p.s_q)
        rb = (__ldg(p.bitmasks + (size_t)ps.row_mask * p.mask_words + (row >> 5)) >> (row & 31)) & 1u;
      const bool rowflip = cx.rowxor && rb;
      float m_used = -INFINITY;
      int n_mine = 0;
      const int it_pass0 = it;
#pragma unroll 1
      for (int seg = 0; seg < 2; ++seg) {
        const SegCtx sg = seg ? cx.s1 : cx.s0;
        if (sg.kv < 0) continue;
        const bool flip = sg.kinv != rowflip;
        const bool uniform = uniform_for(cx, sg, flip, p.s_kv);
        const float sc = uniform ? 0.f : p.scale_log2;
#pragma unroll 1
        for (int j0 = 0; j0 < n_kv_tiles;) {
          const int cls = tile_class(sg, j0, p), j1 = run_end(sg, j0, cls, p), jb = j0;
          j0 = j1;
          if (tile_skip(cx, sg, cls, p.s_kv)) continue;
          const bool row_ok_cls = row_allowed(cls, flip, uniform);
          int j = jb + (((it ^ wg) & 1) ? 1 : 0);
          int itj = it + (j - jb);
          it += j1 - jb;
#pragma unroll 1
          for (; j < j1; j += 2, itj += 2) {
            const int sb = oc_sbuf(itj);
            const uint32_t tS = tlane + C::TMEM_S + BN * sb;
            mbar_wait_hot<0>(bar_s + 8 * sb, (itj >> 2) & 1);
            tc_fence_after();
            // ---- the whole 64-column row: both loads in flight together
            float sa[32], sb2[32];
            tmem_ld32(tS, sa);
            tmem_ld32(tS + 32, sb2);
            // ---- allowed-key bits of this row for MIX tiles (computed under the load latency)
            uint32_t kb_lo = 0xffffffffu, kb_hi = 0xffffffffu;
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
                    kb = __ldg(p.bitmasks + (size_t)sg.kmask * p.mask_words + (kbase >> 5));
                  }
                  if (flip) kb = ~kb;
                }
                if (w == 0) kb_lo = kb & valid; else kb_hi = kb & valid;
              }
            }
            tmem_wait_ld32(sa);
            tmem_wait_ld32(sb2);
            // ---- row max over the ALLOWED keys of the tile
            float mt;
            {
              float m0, m1, m2, m3;
              if (cls != TILE_MIX) {
                m0 = fmaxf(sa[0], sa[1]);
                m1 = fmaxf(sa[2], sa[3]);
                m2 = fmaxf(sb2[0], sb2[1]);
                m3 = fmaxf(sb2[2], sb2[3]);
#pragma unroll
                for (int i = 4; i < 32; i += 4) {
                  m0 = fmaxf(m0, fmaxf(sa[i], sa[i + 1]));
                  m1 = fmaxf(m1, fmaxf(sa[i + 2], sa[i + 3]));
                  m2 = fmaxf(m2, fmaxf(sb2[i], sb2[i + 1]));
                  m3 = fmaxf(m3, fmaxf(sb2[i + 2], sb2[i + 3]));
                }
              } else {
                m0 = m1 = m2 = m3 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  m0 = fmaxf(m0, (kb_lo >> i) & 1u ? sa[i] : -INFINITY);
                  m1 = fmaxf(m1, (kb_lo >> (i + 1)) & 1u ? sa[i + 1] : -INFINITY);
                  m2 = fmaxf(m2, (kb_hi >> i) & 1u ? sb2[i] : -INFINITY);
                  m3 = fmaxf(m3, (kb_hi >> (i + 1)) & 1u ? sb2[i + 1] : -INFINITY);
                }
              }
              mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
              if (cls != TILE_MIX && !row_ok_cls) mt = -INFINITY;
            }
            const float mts = uniform ? 0.f : mt * p.scale_log2;
            // ---- running reference point (lazy rescale decision)
            float alpha = 1.f;
            bool grow = false;
            if (n_mine == 0) {
              m_used = mts;
            } else if (mts > m_used + rescale_threshold<false>()) {
              alpha = fast_exp2(m_used - mts);
              m_used = mts;
              grow = true;
            }
            // ---- p = 2^(s*scale*log2e - m) as packed fp16 pairs: keys [32*hb, 32*hb+32) -> pk[16*hb .. 16*hb+16)
            const float nb = (cls == TILE_MIX || row_ok_cls) ? -m_used : -INFINITY;
            uint32_t pk[32];
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              const float* sv = hb ? sb2 : sa;
              const uint32_t kbits = hb ? kb_hi : kb_lo;
              if (cls != TILE_MIX) {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) softmax_chunk_f16<false>(sv + 16 * jj, pk + 16 * hb + 8 * jj, sc, nb, 0u);
              } else {
#pragma unroll
                for (int jj = 0; jj < 2; ++jj)
                  softmax_chunk_f16<true>(sv + 16 * jj, pk + 16 * hb + 8 * jj, sc, nb, (kbits >> (16 * jj)) & 0xffffu);
              }
            }
            // ---- PV of my previous tile has completed (observed every tile, in order): O_wg may be rescaled, and the
            // accumulate of PV(this tile) will see the rescaled values
            if (o_seen < m_glob) {
              mbar_wait_hot<0>(bar_o + 8 * wg, (m_glob - 1) & 1);
              o_seen = m_glob;
              tc_fence_after();
            }
            if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
              for (int c = 0; c < C::DPV / 16; ++c) {
                float o[16];
                uint32_t ob[16];
                tmem_ld16(tO + 16 * c, o);
                tmem_wait_ld16(o);
#pragma unroll
                for (int i = 0; i < 16; ++i) ob[i] = __float_as_uint(o[i] * alpha);
                tmem_st16(tO + 16 * c, ob);
              }
            }
            tmem_st16(tS, *reinterpret_cast<const uint32_t(*)[16]>(pk));
            tmem_st16(tS + 32, *reinterpret_cast<const uint32_t(*)[16]>(pk + 16));
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p + 8 * sb);
            ++n_mine;
            ++m_glob;
          }
        }
      }
      if (it == it_pass0) continue;
      // ---- end of pass: (1) my last PV has landed
      if (o_seen < m_glob) {
        mbar_wait_hot<0>(bar_o + 8 * wg, (m_glob - 1) & 1);
        o_seen = m_glob;
      }
      // (2) exchange the reference points
      mx_smem[wg * BM + rloc] = n_mine > 0 ? m_used : -INFINITY;
      tc_fence_before();
      named_bar_sync(1, 32 * NUM_SOFTMAX_WARPS);
      tc_fence_after();
      const float m_other = mx_smem[(wg ^ 1) * BM + rloc];
      const float m_all = fmaxf(m_used, m_other);
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
          for (int i = 0; i < 16; ++i) r[i] = fmaf(c_mine, o[i], r[i]);
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
          if (row_ok && 16 * c + 8 * g < p.head_dim) {
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
  // ---- teardown: every tcgen05 op of this CTA has completed (the softmax warps observed o_done of their last tiles)
  tc_fence_before();
  __syncthreads();
  if (warp == NUM_SOFTMAX_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}
