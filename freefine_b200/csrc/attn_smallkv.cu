// ff_attn_plain_smallkv: plain softmax(Q K^T * scale) V for SHORT key sequences (s_kv <= 128) -- the text cross-attention of
// every transformer block (77 keys) and the plain 8 x 8 self-attention (64 keys).
//
// Replaces the plain branch of ca_forward (src/utils/attention.py:395-404) for those layers.  Why a second attention kernel:
// the tcgen05 kernel (attn_tcgen05.cu) is built for thousands of keys -- per CTA it allocates TMEM, initialises a dozen
// mbarriers, prefetches tensor maps and runs a four-role pipeline, ~7 us of fixed latency that a 128-row x 77-key tile
// cannot amortise: the round-2 launch list shows 207 us for the 32-stream cross-attention at 64 x 64 (8192 CTAs in 28
// waves at 2 CTAs per SM), 5.3 % of a pair of UNet calls in total, for launches whose traffic (Q in, O out: 168 MB) takes
// 26 us at the HBM peak and whose 13 GFLOP are nothing.  This kernel is HBM-bound by design instead: K and V of one
// (stream, head) sit in shared memory for the life of the CTA, each of 4 warps owns 16 query rows per step, S = Q K^T and
// O = P V run on mma.sync.m16n8k16 with the whole score row in registers (one pass, no online rescale: every key is
// there), Q tiles stream in through a cp.async double buffer and O leaves as 16-byte rows staged through the Q slot.
// No TMEM / TMA / mbarriers: there is nothing to pipeline across key tiles, and the tensor work is 1 % of the launch.
// Numerics follow the main kernel: Q, K bf16 operands, fp32 scores, P and V as fp16 operands (|v| >= 65504 saturates),
// fp32 accumulation and normalisation.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ff_common.cuh"

namespace {

constexpr int SK_THREADS = 128;     // 4 warps
constexpr int SK_ROWS = 64;         // query rows per step of a CTA (16 per warp)
constexpr int SK_STEPS = 4;         // steps per CTA: 256 query rows share one K/V load

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// first product of an accumulator: C = 0 comes from RZ, the accumulator registers need no zeroing moves (60 per 16-row step)
__device__ __forceinline__ void mma_bf16_z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void mma_f16_z(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;      // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct SkCfg {
  static constexpr int DP = (D + 15) / 16 * 16;   // contraction length of Q K^T padded to the k16 of the mma (40 -> 48)
  static constexpr int LDS = DP + 8;               // shared-memory row stride (elements): 16-byte rows, conflict-free ldmatrix
  static constexpr int DV = D / 8;                 // 16-byte vectors per row / n8 tiles of O
};

// shared memory: K [NBLK*KT16*16][LDS] bf16 | V [NBLK*KT16*16][LDS] fp16 | Q [2][64][LDS] bf16
template <int D, int KT16, int NBLK>
constexpr size_t sk_smem_bytes() { return (size_t)(2 * KT16 * 16 * NBLK + 2 * SK_ROWS) * SkCfg<D>::LDS * 2; }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}

constexpr int sk_pow2_at_least(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : (v <= 16 ? 16 : 32)))); }

// All shared-memory addressing is "per-lane base computed once + compile-time offset" (the first version recomputed
// row * LDS + column per ldmatrix: 121 IMAD per 16-row step, ncu), the key-padding mask is applied to the boundary tile only.
// NBLK > 1: key sequences of up to NBLK * 128 keys (the plain 16 x 16 self-attention: 256 keys) as NBLK blocks of KT16 * 16
// keys under one online softmax (running maximum / sum, O rescaled between the blocks); all K / V still resident.
template <int D, int KT16, int NBLK, bool OUT_F32>
__global__ void __launch_bounds__(SK_THREADS, NBLK > 1 ? (D <= 80 ? 2 : 1) : (D <= 40 ? 6 : (D <= 80 ? 4 : 2)))
attn_smallkv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                    void* __restrict__ out, int s_q, int s_kv, int heads, float scale_log2e) {
  using Cfg = SkCfg<D>;
  constexpr int LDS = Cfg::LDS, DP = Cfg::DP, DV = Cfg::DV, KP = KT16 * 16, NT = KT16 * 2, KPT = KP * NBLK;
  constexpr int TPR = sk_pow2_at_least(DP / 8);          // threads side by side over the 16-byte vectors of a row (loads)
  constexpr int TPW = sk_pow2_at_least(DV);              // ... of an output row (stores, per warp)
  constexpr uint32_t SLOT = SK_ROWS * LDS * 2;           // bytes of one Q slot
  extern __shared__ __align__(16) unsigned char sk_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(sk_smem);
  __half* Vs = reinterpret_cast<__half*>(Ks + KPT * LDS);
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(Vs + KPT * LDS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y, stream = blockIdx.z;
  const int C = heads * D;
  const int row0 = blockIdx.x * (SK_ROWS * SK_STEPS);
  const __nv_bfloat16* qg = q + ((size_t)stream * s_q) * C + head * D;
  const __nv_bfloat16* kg = k + ((size_t)stream * s_kv) * C + head * D;
  const __nv_bfloat16* vg = v + ((size_t)stream * s_kv) * C + head * D;
  const int ld_r = tid / TPR, ld_c = tid % TPR;           // (powers of two: shifts)
  const bool ld_on = ld_c < DP / 8, ld_data = ld_c < DV;
  const uint32_t qs_u32 = smem_u32(Qs);

  // Q tile of step `st` -> slot st & 1 (rows past s_q and the padding columns D..DP are zero-filled).  FF_SK_WP (default):
  // every warp copies ITS OWN 16 rows, so the step loop needs no CTA-wide barrier (cp.async wait + __syncwarp) and the
  // four warps of a CTA drift apart instead of meeting twice per step.
#ifndef FF_SK_WP
#define FF_SK_WP 1
#endif
#if FF_SK_WP
  constexpr int TPQ = sk_pow2_at_least(DP / 8);            // lanes side by side over the vectors of a row (<= 32)
  const int wq_r = lane / TPQ, wq_c = lane % TPQ;
  auto load_q = [&](int st) {
    if (wq_c >= DP / 8) return;
    const int base = row0 + st * SK_ROWS + warp * 16;
    uint32_t dst = qs_u32 + (st & 1) * SLOT + ((warp * 16 + wq_r) * LDS + 8 * wq_c) * 2;
    const __nv_bfloat16* src = qg + (size_t)(base + wq_r) * C + 8 * wq_c;
#pragma unroll
    for (int r = 0; r < 16; r += 32 / TPQ) {
      const bool ok = wq_c < DV && base + wq_r + r < s_q;
      cp_async16(dst, ok ? static_cast<const void*>(src) : static_cast<const void*>(qg), ok);
      dst += (32 / TPQ) * LDS * 2;
      src += (size_t)(32 / TPQ) * C;
    }
  };
#else
  auto load_q = [&](int st) {
    if (!ld_on) return;
    const int base = row0 + st * SK_ROWS;
    uint32_t dst = qs_u32 + (st & 1) * SLOT + (ld_r * LDS + 8 * ld_c) * 2;
    const __nv_bfloat16* src = qg + (size_t)(base + ld_r) * C + 8 * ld_c;
#pragma unroll
    for (int r = 0; r < SK_ROWS; r += SK_THREADS / TPR) {
      const bool ok = ld_data && base + ld_r + r < s_q;
      cp_async16(dst, ok ? static_cast<const void*>(src) : static_cast<const void*>(qg), ok);
      dst += (SK_THREADS / TPR) * LDS * 2;
      src += (size_t)(SK_THREADS / TPR) * C;
    }
  };
#endif
  // K and V of this (stream, head): every 16-byte vector goes out as a cp.async at once (rows >= s_kv and the K padding
  // columns zero-filled), together with the first Q tile -- ONE exposed memory latency for the whole prologue (the first
  // version walked the rows with load -> convert -> store per trip: 64 dependent trips for 256 keys x d=160).  V lands as
  // bf16 and is converted to fp16 (saturating) in place by the thread that copied it, before the CTA-wide barrier.
  load_q(0);
  if (ld_on) {
    const uint32_t ks_u32 = smem_u32(Ks) + (ld_r * LDS + 8 * ld_c) * 2, vs_u32 = smem_u32(Vs) + (ld_r * LDS + 8 * ld_c) * 2;
    for (int r = ld_r; r < KPT; r += SK_THREADS / TPR) {
      const bool ok = ld_data && r < s_kv;
      const uint32_t off = (uint32_t)(r - ld_r) * LDS * 2;
      cp_async16(ks_u32 + off, ok ? static_cast<const void*>(kg + (size_t)r * C + 8 * ld_c) : static_cast<const void*>(kg), ok);
      cp_async16(vs_u32 + off, ok ? static_cast<const void*>(vg + (size_t)r * C + 8 * ld_c) : static_cast<const void*>(vg), ok);
    }
  }
  cp_async_commit();
  cp_async_wait<0>();
  if (ld_data) {
    for (int r = ld_r; r < min(KPT, s_kv); r += SK_THREADS / TPR) {
      uint4* pv = reinterpret_cast<uint4*>(Vs + r * LDS + 8 * ld_c);
      const uint4 vb = *pv;
      const uint32_t w[4] = {vb.x, vb.y, vb.z, vb.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = fminf(fmaxf(__uint_as_float(w[j] << 16), -65504.f), 65504.f);
        const float hi = fminf(fmaxf(__uint_as_float(w[j] & 0xffff0000u), -65504.f), 65504.f);
        o[j] = pack_f16(lo, hi);
      }
      *pv = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
#if FF_SK_WP
  __syncthreads();                   // K / V of the CTA are in place (the step loop itself has no CTA-wide barrier)
#endif
  // per-lane ldmatrix bases (bytes)
  const uint32_t q_lane = (uint32_t)(((warp * 16 + (lane & 15)) * LDS + (lane >> 4) * 8) * 2);
  const uint32_t k_base = smem_u32(Ks) + (uint32_t)((((lane >> 4) * 8 + (lane & 7)) * LDS + ((lane >> 3) & 1) * 8) * 2);
  const uint32_t v_base = smem_u32(Vs) + (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + (lane >> 4) * 8) * 2);
  const uint32_t v_base1 = smem_u32(Vs) + (uint32_t)(((((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + (DV - 1) * 8) * 2);
  const uint32_t o_lane = (uint32_t)(((warp * 16 + (lane >> 2)) * LDS + 2 * (lane & 3)) * 2);        // my accumulator position
  const int st_r = lane / TPW, st_c = lane % TPW;
  const uint32_t o_ld = (uint32_t)(((warp * 16 + st_r) * LDS + 8 * st_c) * 2);
  const int kq = 2 * (lane & 3);                           // my first key column inside an n8 tile

  const int n_steps = min(SK_STEPS, (s_q - row0 + SK_ROWS - 1) / SK_ROWS);
  for (int st = 0; st < n_steps; ++st) {
    if (st + 1 < n_steps) load_q(st + 1);
    cp_async_commit();
    cp_async_wait<1>();              // the tile of this step has landed (the one just issued may still fly)
#if FF_SK_WP
    __syncwarp();                    // ... for every lane's copies of my warp's rows
#else
    __syncthreads();                 // ... for every thread's copies; also covers the K / V stores before the first step
#endif
    const uint32_t slot = qs_u32 + (st & 1) * SLOT;

    float o[DV][4];
    if (NBLK > 1) {                  // (one key block: the first P V product writes the accumulators)
#pragma unroll
      for (int n = 0; n < DV; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    }
    float mr0 = -INFINITY, mr1 = -INFINITY, l0 = 0.f, l1 = 0.f;      // running row maxima; per-thread partial row sums
#pragma unroll 1                     // (NBLK = 2 unrolled let ptxas interleave the two blocks: 255 registers)
    for (int blk = 0; blk < NBLK; ++blk) {
      constexpr uint32_t BLK = KP * LDS * 2;                          // bytes of one key block in Ks / Vs
      // ---- S = Q K^T: NT n8 tiles of 16 x 8 scores, fp32
      float s[NT][4];
#pragma unroll
      for (int kb = 0; kb < DP / 16; ++kb) {
        uint32_t a[4];
        ldsm_x4(slot + q_lane + kb * 32, a);
#pragma unroll
        for (int t = 0; t < NT; t += 2) {
          uint32_t b0, b1, b2, b3;   // keys 8t..8t+7 (k 0-7, k 8-15), keys 8t+8..8t+15 (k 0-7, k 8-15)
          ldsm_x4(k_base + blk * BLK + (uint32_t)((t * 8 * LDS + kb * 16) * 2), b0, b1, b2, b3);
          if (kb == 0) {
            mma_bf16_z(s[t], a, b0, b1);
            mma_bf16_z(s[t + 1], a, b2, b3);
          } else {
            mma_bf16(s[t], a, b0, b1);
            mma_bf16(s[t + 1], a, b2, b3);
          }
        }
      }
      // ---- softmax over the block (rows lane/4 and lane/4 + 8; a row lives in the 4 lanes of a quad)
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (blk * KP + t * 8 + 8 > s_kv) {        // (warp-uniform) tile with padding keys
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (blk * KP + t * 8 + kq + (e & 1) >= s_kv) s[t][e] = -INFINITY;
        }
        m0 = fmaxf(m0, fmaxf(s[t][0], s[t][1]));
        m1 = fmaxf(m1, fmaxf(s[t][2], s[t][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      if (NBLK > 1 && blk > 0) {
        // online softmax: every block holds at least one real key (the host picks NBLK = ceil(s_kv / 128)), so the maxima
        // are finite from the first block on
        const float n0 = fmaxf(mr0, m0), n1 = fmaxf(mr1, m1);
        const float al0 = ex2((mr0 - n0) * scale_log2e), al1 = ex2((mr1 - n1) * scale_log2e);
        l0 *= al0;
        l1 *= al1;
#pragma unroll
        for (int n = 0; n < DV; ++n) {
          o[n][0] *= al0;
          o[n][1] *= al0;
          o[n][2] *= al1;
          o[n][3] *= al1;
        }
        mr0 = n0;
        mr1 = n1;
      } else {
        mr0 = m0;
        mr1 = m1;
      }
      const float nb0 = -mr0 * scale_log2e, nb1 = -mr1 * scale_log2e;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        s[t][0] = ex2(fmaf(s[t][0], scale_log2e, nb0));
        s[t][1] = ex2(fmaf(s[t][1], scale_log2e, nb0));
        s[t][2] = ex2(fmaf(s[t][2], scale_log2e, nb1));
        s[t][3] = ex2(fmaf(s[t][3], scale_log2e, nb1));
        l0 += s[t][0] + s[t][1];
        l1 += s[t][2] + s[t][3];
      }
      // ---- O += P V: P straight from the score registers (two n8 tiles = one k16 A fragment), fp16
#pragma unroll
      for (int j = 0; j < KT16; ++j) {
        const uint32_t a[4] = {pack_f16(s[2 * j][0], s[2 * j][1]), pack_f16(s[2 * j][2], s[2 * j][3]),
                               pack_f16(s[2 * j + 1][0], s[2 * j + 1][1]), pack_f16(s[2 * j + 1][2], s[2 * j + 1][3])};
#pragma unroll
        for (int n = 0; n + 1 < DV; n += 2) {
          uint32_t b[4];           // channels 8n..8n+7 (keys 0-7, 8-15 of the block), channels 8n+8..8n+15 (same)
          ldsm_x4_t(v_base + blk * BLK + (uint32_t)((j * 16 * LDS + n * 8) * 2), b);
          if (NBLK == 1 && j == 0) {
            mma_f16_z(o[n], a, b[0], b[1]);
            mma_f16_z(o[n + 1], a, b[2], b[3]);
          } else {
            mma_f16(o[n], a, b[0], b[1]);
            mma_f16(o[n + 1], a, b[2], b[3]);
          }
        }
        if (DV & 1) {
          uint32_t b[2];
          ldsm_x2_t(v_base1 + blk * BLK + (uint32_t)(j * 16 * LDS * 2), b);
          if (NBLK == 1 && j == 0) mma_f16_z(o[DV - 1], a, b[0], b[1]);
          else mma_f16(o[DV - 1], a, b[0], b[1]);
        }
      }
    }
    // row sums: the partials of the four lanes of a quad (their rescale factors were identical)
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float r0 = 1.f / l0, r1 = 1.f / l1;
    const int grow = row0 + st * SK_ROWS + warp * 16;      // first global row of my warp
    if (OUT_F32) {
      float* og = static_cast<float*>(out) + ((size_t)stream * s_q) * C + head * D;
#pragma unroll
      for (int n = 0; n < DV; ++n) {
        const int col = n * 8 + kq;
        const int ra = grow + (lane >> 2), rb = ra + 8;
        if (ra < s_q) *reinterpret_cast<float2*>(og + (size_t)ra * C + col) = make_float2(o[n][0] * r0, o[n][1] * r0);
        if (rb < s_q) *reinterpret_cast<float2*>(og + (size_t)rb * C + col) = make_float2(o[n][2] * r1, o[n][3] * r1);
      }
    } else {
      // stage the 16 x D bf16 tile through my rows of the Q slot (all my ldmatrix reads of it are done), leave as 16-byte rows
      __syncwarp();
      unsigned char* sb = sk_smem + (slot - smem_u32(sk_smem));
#pragma unroll
      for (int n = 0; n < DV; ++n) {
        *reinterpret_cast<__nv_bfloat162*>(sb + o_lane + n * 16) = __floats2bfloat162_rn(o[n][0] * r0, o[n][1] * r0);
        *reinterpret_cast<__nv_bfloat162*>(sb + o_lane + n * 16 + 8 * LDS * 2) = __floats2bfloat162_rn(o[n][2] * r1, o[n][3] * r1);
      }
      __syncwarp();
      if (st_c < DV) {
        __nv_bfloat16* og = static_cast<__nv_bfloat16*>(out) + ((size_t)stream * s_q + grow + st_r) * C + head * D + 8 * st_c;
#pragma unroll
        for (int r = 0; r < 16; r += 32 / TPW) {
          if (grow + st_r + r < s_q)
            *reinterpret_cast<uint4*>(og + (size_t)r * C) = *reinterpret_cast<const uint4*>(sb + o_ld + r * LDS * 2);
        }
      }
    }
#if FF_SK_WP
    __syncwarp();                    // my rows of the slot are free for the tile of step st + 2
#else
    __syncthreads();                 // the slot is free for the tile of step st + 2
#endif
  }
  cp_async_wait<0>();
}

template <int D, int KT16, int NBLK>
int sk_launch(const void* q, const void* k, const void* v, void* out, int n_streams, int heads, int s_q, int s_kv, float scale,
              int out_dtype, cudaStream_t st) {
  constexpr size_t smem = sk_smem_bytes<D, KT16, NBLK>();
  auto kf = attn_smallkv_kernel<D, KT16, NBLK, true>;
  auto kb = attn_smallkv_kernel<D, KT16, NBLK, false>;
  if (smem > 48 * 1024) {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
      cudaError_t e1 = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaError_t e2 = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e1 != cudaSuccess || e2 != cudaSuccess)
        return ff::fail(FF_E_CUDA, "ff_attn_plain_smallkv: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
      configured[dev & 63] = true;
    }
  }
  const dim3 grid((s_q + SK_ROWS * SK_STEPS - 1) / (SK_ROWS * SK_STEPS), heads, n_streams);
  const float sl = scale * 1.4426950408889634f;
  if (out_dtype == FF_DT_F32)
    kf<<<grid, SK_THREADS, smem, st>>>(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                        static_cast<const __nv_bfloat16*>(v), out, s_q, s_kv, heads, sl);
  else
    kb<<<grid, SK_THREADS, smem, st>>>(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                        static_cast<const __nv_bfloat16*>(v), out, s_q, s_kv, heads, sl);
  return ff::check_launch("ff_attn_plain_smallkv");
}

}  // namespace

extern "C" int ff_attn_plain_smallkv(const void* q, const void* k, const void* v, void* out, int32_t n_streams, int32_t heads,
                                     int32_t head_dim, int32_t s_q, int32_t s_kv, float scale, int32_t out_dtype, void* stream) {
  FF_REQUIRE(q && k && v && out, "ff_attn_plain_smallkv: null pointer");
  FF_REQUIRE(n_streams > 0 && heads > 0 && s_q > 0 && s_kv > 0, "ff_attn_plain_smallkv: bad shape");
  FF_REQUIRE(n_streams <= 65535 && heads <= 65535, "ff_attn_plain_smallkv: n_streams / heads must be <= 65535");
  FF_REQUIRE(s_kv <= (head_dim == 8 ? 128 : 256),
             "ff_attn_plain_smallkv: s_kv=%d must be <= 256 (128 for head_dim 8); longer key sequences: ff_attn_masked_kv", s_kv);
  FF_REQUIRE(head_dim == 8 || head_dim == 40 || head_dim == 80 || head_dim == 160,
             "ff_attn_plain_smallkv: head_dim=%d must be 40, 80, 160 (SD1.5) or 8 (the reference's golden vectors)", head_dim);
  FF_REQUIRE(out_dtype == FF_DT_BF16 || out_dtype == FF_DT_F32, "ff_attn_plain_smallkv: out_dtype must be bf16 or f32");
  FF_REQUIRE(scale > 0.f, "ff_attn_plain_smallkv: scale must be positive");
  FF_REQUIRE(ff::aligned16(q) && ff::aligned16(k) && ff::aligned16(v) && ff::aligned16(out), "ff_attn_plain_smallkv: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FF_SK(Dv)                                                                                                 \
  do {                                                                                                            \
    if (s_kv <= 80) return sk_launch<Dv, 5, 1>(q, k, v, out, n_streams, heads, s_q, s_kv, scale, out_dtype, st);  \
    if (s_kv <= 128) return sk_launch<Dv, 8, 1>(q, k, v, out, n_streams, heads, s_q, s_kv, scale, out_dtype, st); \
    return sk_launch<Dv, 8, 2>(q, k, v, out, n_streams, heads, s_q, s_kv, scale, out_dtype, st);                  \
  } while (0)
  if (head_dim == 8) {
    if (s_kv <= 80) return sk_launch<8, 5, 1>(q, k, v, out, n_streams, heads, s_q, s_kv, scale, out_dtype, st);
    return sk_launch<8, 8, 1>(q, k, v, out, n_streams, heads, s_q, s_kv, scale, out_dtype, st);
  }
  if (head_dim == 40) FF_SK(40);
  if (head_dim == 80) FF_SK(80);
  FF_SK(160);
#undef FF_SK
}
