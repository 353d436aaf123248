// 32-key-tile variant of the masked KV-injection attention kernel for head_dim <= 40, fp16 P (experiment switch
// FF_ATTN_T32=1 in the environment; included by attn_tcgen05.cu inside its anonymous namespace, shares its PTX wrappers,
// pass / tile decision logic and exp2 sweep).
//
// Why: a softmax warpgroup of the 64-key kernel is busy ~2000 cycles per tile and then idles ~950 in the hand-shake
// p_full -> PV(t), QK(t+2) -> s_full, because P overwrites the ONE S buffer the warpgroup owns (TMEM: 4 chains x (64 S + 48 O)
// = 448 of 512 columns; a second S buffer or a private P region per chain does not fit -- profiles/r2b_attn_experiments.txt).
// With 32-key tiles a chain needs 2 x 32 (two S buffers) + 16 (its own packed P) + 48 (O) = 128 columns: four chains fit
// exactly (2 CTAs x 256 columns), QK(t+4) is issued into the buffer tile t has just vacated while the warpgroup works on
// tile t+2, and nothing of the hand-shake is left on the chain.  The price is twice the per-tile bookkeeping per key.
//
// Roles as in the 64-key kernel: warps 0-7 softmax (two warpgroups, tile parity), warp 8 TMA producer + TMEM alloc, warp 9
// MMA issuer (warp-uniform lean loop, one asm block per contraction).  Global tile t: warpgroup t & 1, S buffer (t >> 1) & 1
// of that warpgroup, K/V stage t % 8.
#pragma once

constexpr int TBN = 32;                 // keys per tile

struct T32 {
  static constexpr int DPAD = 48, DPV = 48, NST = 8;
  static constexpr int KVB = TBN * 128;                                   // one K or V tile: 32 rows x 128 B
  static constexpr int TMEM_S = 0, TMEM_P = 4 * TBN, TMEM_O = 4 * TBN + 2 * (TBN / 2), TMEM_COLS = 256;
  static_assert(TMEM_O + 2 * DPV == 256, "TMEM budget: 4 S buffers, 2 P regions, 2 accumulators");
  static constexpr int SMEM_Q = TILE_BYTES, SMEM_STAGE = 2 * KVB;
  static constexpr int ACC_LD = DPAD + 4, SMEM_ACC = BM * ACC_LD * 4, SMEM_MX = 2 * BM * 4;
  static constexpr int SMEM_BYTES = SMEM_Q + NST * SMEM_STAGE + SMEM_ACC + SMEM_MX + 1024 + 512;
  static_assert(SMEM_BYTES <= 113 * 1024, "two CTAs per SM");
};

__device__ __forceinline__ int t32_class(const SegCtx& g, int j, const KParams& p) {
  const int lo = j * TBN;
  if (lo + TBN > p.s_kv) return TILE_MIX;
  if (g.kmask < 0) return TILE_ALL;
  if (g.prefix) return lo + TBN <= g.T ? TILE_IN : (lo >= g.T ? TILE_OUT : TILE_MIX);
  const uint32_t bits = __ldg(p.bitmasks + (size_t)g.kmask * p.mask_words + (lo >> 5));
  return bits == 0xffffffffu ? TILE_IN : (bits == 0 ? TILE_OUT : TILE_MIX);
}
__device__ __forceinline__ int t32_run_end(const SegCtx& g, int j, int cls, const KParams& p) {
  if (cls == TILE_MIX || (g.kmask >= 0 && !g.prefix)) return j + 1;
  const int n_full = p.s_kv / TBN;
  if (cls == TILE_IN) { const int t = g.T / TBN; return t < n_full ? t : n_full; }
  return n_full;
}

// O (+)= P V for a 32-key tile: two K-steps of 16 keys (A = packed fp16 P at columns +0, +8 of the P region, B = V tile
// MN-major, 2048 B = 128 descriptor units per K-step), then the commits kv_empty and pv_done
__device__ __forceinline__ void t32_pv(uint32_t obuf, uint32_t pbase, uint64_t vdesc, uint32_t idesc_pv, uint32_t acc0,
                                       uint32_t bar_kve, uint32_t bar_o_) {
  asm volatile(
      "{\n\t.reg .pred e, p0, pt;\n\t.reg .b64 v1;\n\t.reg .b32 a1;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p0, %4, 0;\n\tsetp.eq.b32 pt, 0, 0;\n\t"
      "add.u64 v1, %2, 128;\n\tadd.u32 a1, %1, 8;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], v1, %3, pt;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t}"
      ::"r"(obuf), "r"(pbase), "l"(vdesc), "r"(idesc_pv), "r"(acc0), "r"(bar_kve), "r"(bar_o_)
      : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
attn_t32_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const KParams p) {
  using C = T32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = smem_base + C::SMEM_Q;
  const uint32_t sACC = sKV + C::NST * C::SMEM_STAGE;
  const uint32_t sMX = sACC + C::SMEM_ACC;
  const uint32_t bar_base = sMX + C::SMEM_MX;
  const uint32_t bar_q = bar_base;
  const uint32_t bar_s = bar_base + 16;           // [2 warpgroups][2 buffers]: 16 + 8 * (2 * wg + k)
  const uint32_t bar_p = bar_base + 48;           // [2]: P of the warpgroup's tile is in its P region (4 warp arrivals)
  const uint32_t bar_o = bar_base + 64;           // [2]: PV of the warpgroup's tile (and every earlier MMA) has completed
  const uint32_t bar_kv_full = bar_base + 80, bar_kv_empty = bar_base + 80 + 8 * C::NST;
  const uint32_t tmem_slot = bar_base + 80 + 16 * C::NST;
  static_assert(80 + 16 * C::NST + 8 <= 512, "barrier block");
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  float* const acc_smem = reinterpret_cast<float*>(gen_base + (sACC - smem_base));
  float* const mx_smem = reinterpret_cast<float*>(gen_base + (sMX - smem_base));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM, head = blockIdx.y, stream = blockIdx.z;
  const FFAttnHeadPlan* plan = p.plan + (size_t)stream * p.heads + head;
  int n_pass = __ldg(&plan->n_pass);
  n_pass = n_pass < 0 ? 0 : (n_pass > FF_MAX_PASS ? FF_MAX_PASS : n_pass);
  const int n_kv_tiles = (p.s_kv + TBN - 1) / TBN;

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar_s + 8 * i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_p + 8 * i, NUM_SOFTMAX_WARPS / 2);
      mbar_init(bar_o + 8 * i, 1);
    }
    for (int i = 0; i < C::NST; ++i) {
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
            const int cls = t32_class(sg, j0, p), j1 = t32_run_end(sg, j0, cls, p), jb = j0;
            j0 = j1;
            if (tile_skip(cx, sg, cls, p.s_kv)) continue;
#pragma unroll 1
            for (int j = jb; j < j1; ++j, ++it) {
              const int stage = it % C::NST, use = it / C::NST;
              if (use > 0) mbar_wait(bar_kv_empty + 8 * stage, (use - 1) & 1);
              const uint32_t full = bar_kv_full + 8 * stage;
              const uint32_t sK = sKV + stage * C::SMEM_STAGE, sV = sK + C::KVB;
              mbar_expect_tx(full, 2 * C::KVB);
              tma_load_4d(sK, &tm_k, 0, head, j * TBN, sg.kv, full);
              tma_load_4d(sV, &tm_v, 0, head, j * TBN, sg.kv, full);
            }
          }
        }
      }
    }
  } else if (warp == NUM_SOFTMAX_WARPS + 1) {
    // ===================================== MMA issuer =======================================
    constexpr uint32_t idesc_qk = make_idesc(TBN, 0);
    constexpr uint32_t idesc_pv = make_idesc(C::DPV, 1, true);
    const uint64_t qdesc0 = smem_desc_sw128(sQ, 16);
    const uint64_t kdesc0 = smem_desc_sw128(sKV, 16);
    const uint64_t vdesc0 = smem_desc_sw128(sKV + C::KVB, C::KVB);
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
              const int cls = t32_class(sg, j0, p), j1 = t32_run_end(sg, j0, cls, p), jb = j0;
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
    tc_fence_after();
    // prologue: S of the first two tiles of each warpgroup (tiles 0..3: stage t, buffer (t >> 1) & 1 of warpgroup t & 1)
#pragma unroll
    for (int t0 = 0; t0 < 4; ++t0) {
      if (t0 < n_total) {
        lean_wait(bar_kv_full + 8 * t0, 0);
        tc_fence_after();
        lean_qk48(tmem + C::TMEM_S + TBN * (2 * (t0 & 1) + ((t0 >> 1) & 1)), qdesc0, kdesc0 + (uint64_t)((t0 * C::SMEM_STAGE) >> 4),
                  idesc_qk, bar_s + 8 * (2 * (t0 & 1) + ((t0 >> 1) & 1)));
      }
    }
    // tile t = 8 g + u: stage u, warpgroup u & 1, S buffer (u >> 1) & 1; p_full(t) has phase (u >> 1) & 1; the K/V stage of
    // tile t + 4 is (u + 4) & 7 with phase g (u < 4) or g + 1.  Per tile: PV(t), then QK(t + 4) into the buffer tile t vacated.
    uint32_t gph = 0;
    int t = 0;
#pragma unroll 1
    while (t < n_total) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (t < n_total) {
          const int pstart = t >= pe2 ? pe2 : (t >= pe1 ? pe1 : (t >= pe0 ? pe0 : 0));
          const uint32_t acc0 = (t - pstart < 2) ? 0u : 1u;        // the first tile of each warpgroup in a pass starts O
          const int wgi = u & 1, kb = (u >> 1) & 1;
          lean_wait(bar_p + 8 * wgi, (uint32_t)kb);
          tc_fence_after();
          t32_pv(tmem + C::TMEM_O + C::DPV * wgi, tmem + C::TMEM_P + (TBN / 2) * wgi, vdesc0 + (uint64_t)((u * C::SMEM_STAGE) >> 4),
                 idesc_pv, acc0, bar_kv_empty + 8 * u, bar_o + 8 * wgi);
          if (t + 4 < n_total) {
            lean_wait(bar_kv_full + 8 * ((u + 4) & 7), u < 4 ? gph : gph ^ 1u);
            tc_fence_after();
            lean_qk48(tmem + C::TMEM_S + TBN * (2 * wgi + kb), qdesc0, kdesc0 + (uint64_t)((((u + 4) & 7) * C::SMEM_STAGE) >> 4),
                      idesc_qk, bar_s + 8 * (2 * wgi + kb));
          }
          ++t;
        }
      }
      gph ^= 1u;
    }
    __syncwarp();
  } else {
    // ===================================== softmax + epilogue ===============================
    const int wq = warp & 3, wg = warp >> 2;
    const int rloc = 32 * wq + lane;
    const int row = q0 + rloc;
    const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t tP = tlane + C::TMEM_P + (TBN / 2) * wg;
    const uint32_t tO = tlane + C::TMEM_O + C::DPV * wg;
    const uint32_t tOx = tlane + C::TMEM_O + C::DPV * (wg ^ 1);
    float* const acc_row = acc_smem + (size_t)rloc * C::ACC_LD;
    bool acc_started = false;
    int it = 0;
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
      int n_mine = 0, last_mine = 0;
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
          const int cls = t32_class(sg, j0, p), j1 = t32_run_end(sg, j0, cls, p), jb = j0;
          j0 = j1;
          if (tile_skip(cx, sg, cls, p.s_kv)) continue;
          const bool row_ok_cls = row_allowed(cls, flip, uniform);
          int j = jb + (((it ^ wg) & 1) ? 1 : 0);
          int itj = it + (j - jb);
          it += j1 - jb;
#pragma unroll 1
          for (; j < j1; j += 2, itj += 2) {
            const int kbuf = (itj >> 1) & 1;
            const uint32_t tS = tlane + C::TMEM_S + TBN * (2 * wg + kbuf);
            mbar_wait(bar_s + 8 * (2 * wg + kbuf), (uint32_t)((itj >> 2) & 1));
            tc_fence_after();
            float s[32];
            tmem_ld32(tS, s);
            // allowed-key bits of this row for MIX tiles (boundary / ragged): bit i <=> key j * 32 + i
            uint32_t kbits = 0xffffffffu;
            if (cls == TILE_MIX) {
              const int kbase = j * TBN;
              const int rem = p.s_kv - kbase;
              const uint32_t valid = rem >= 32 ? 0xffffffffu : (rem <= 0 ? 0u : ((1u << rem) - 1u));
              uint32_t kb = 0xffffffffu;
              if (sg.kmask >= 0 && !uniform && rem > 0) {
                if (sg.prefix) {
                  const int tt = sg.T - kbase;
                  kb = tt >= 32 ? 0xffffffffu : (tt <= 0 ? 0u : ((1u << tt) - 1u));
                } else {
                  kb = __ldg(p.bitmasks + (size_t)sg.kmask * p.mask_words + (kbase >> 5));
                }
                if (flip) kb = ~kb;
              }
              kbits = kb & valid;
            }
            tmem_wait_ld32(s);
            // ---- row max over the ALLOWED keys of the tile
            float mt;
            {
              float m0, m1;
              if (cls != TILE_MIX) {
                m0 = fmaxf(s[0], s[1]);
                m1 = fmaxf(s[2], s[3]);
#pragma unroll
                for (int i = 4; i < 32; i += 4) {
                  m0 = fmaxf(m0, fmaxf(s[i], s[i + 1]));
                  m1 = fmaxf(m1, fmaxf(s[i + 2], s[i + 3]));
                }
              } else {
                m0 = m1 = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  m0 = fmaxf(m0, (kbits >> i) & 1u ? s[i] : -INFINITY);
                  m1 = fmaxf(m1, (kbits >> (i + 1)) & 1u ? s[i + 1] : -INFINITY);
                }
              }
              mt = fmaxf(m0, m1);
              if (cls != TILE_MIX && !row_ok_cls) mt = -INFINITY;
            }
            const float mts = uniform ? 0.f : mt * p.scale_log2;
            float alpha = 1.f;
            bool grow = false;
            if (n_mine == 0) {
              m_used = mts;
            } else if (mts > m_used + rescale_threshold<false>()) {
              alpha = fast_exp2(m_used - mts);
              m_used = mts;
              grow = true;
            }
            const bool any_grow = __any_sync(0xffffffffu, grow);
            const float nb = (cls == TILE_MIX || row_ok_cls) ? -m_used : -INFINITY;
            uint32_t pk[16];
            if (cls != TILE_MIX) {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) softmax_chunk_f16<false>(s + 16 * jj, pk + 8 * jj, sc, nb, 0u);
            } else {
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) softmax_chunk_f16<true>(s + 16 * jj, pk + 8 * jj, sc, nb, (kbits >> (16 * jj)) & 0xffffu);
            }
            // ---- my P region and my accumulator are free when PV of my previous tile has completed (normally long ago: it
            // was issued a whole tile time back).  Every completion of pv_done[wg] is observed in order, one phase ahead at
            // most: PV(itj-4) is implied by s_full(itj) -- QK(itj) was issued behind it -- and PV(itj) needs the P stored below.
            if (itj >= 2) {
              const uint32_t po = (uint32_t)(((itj - 2) >> 1) & 1);
              if (!mbar_test(bar_o + 8 * wg, po)) mbar_wait(bar_o + 8 * wg, po);
            }
            tc_fence_after();
            if (any_grow) {
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
            tmem_st16(tP, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p + 8 * wg);
            ++n_mine;
            last_mine = itj;
          }
        }
      }
      if (it == it_pass0) continue;
      // ---- end of pass.  (1) my last PV has landed (the other warpgroup waits for its own before the barrier below)
      if (n_mine > 0) mbar_wait(bar_o + 8 * wg, (uint32_t)((last_mine >> 1) & 1));
      tc_fence_after();
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
      float l = 0.f;
      if (wrote_mine) l = a_mine * tmem_ld1_wait(tO + p.head_dim);
      if (wrote_other) l = fmaf(a_other, tmem_ld1_wait(tOx + p.head_dim), l);
      float coef = ps.weight;
      if (ps.flags & FF_PASS_ROW_WEIGHT) coef = rb ? coef : 0.f;
      coef = l > 0.f ? coef / l : 0.f;
      const float c_mine = coef * a_mine, c_other = coef * a_other;
#pragma unroll
      for (int c = 0; c < C::DPAD / 16; ++c) {
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
      tc_fence_before();
      named_bar_sync(2, 32 * NUM_SOFTMAX_WARPS);
      tc_fence_after();
    }
    // ---- write the row
    {
      const bool row_ok = row < p.s_q;
      const size_t o_off = ((size_t)stream * p.s_q + (row_ok ? row : 0)) * ((size_t)p.heads * p.head_dim) +
                           (size_t)head * p.head_dim;
#pragma unroll
      for (int c = 0; c < C::DPAD / 16; ++c) {
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
  tc_fence_before();
  __syncthreads();
  if (warp == NUM_SOFTMAX_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)C::TMEM_COLS)
                 : "memory");
  }
}
