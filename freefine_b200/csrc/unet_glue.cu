// Channels-last (NHWC) glue kernels of the UNet body between the library convolutions / GEMMs.
//
// The reference runs diffusers' UNet eagerly: per resnet half  conv -> (+bias) -> (+temb) -> GroupNorm (moments, apply)
// -> SiLU, per transformer block  permute copy -> LayerNorm x3 -> GEGLU (gelu, mul), per conv a NCHW<->NHWC conversion
// pair inside cuDNN.  ncu launch list of round 1: 29 % of a UNet call in non-vectorised at::elementwise_kernel, 8 % in
// layout conversions, 6 % in RowwiseMoments, 6.5 % in layer norm -- all of it HBM-bound byte shuffling.  These kernels
// keep every activation bf16 [N, H*W, C] (= torch channels_last = the token layout of the transformer blocks, so the
// permutes become views) and touch each tensor the minimum number of times:
//   ff_group_norm_nhwc     y = act(GroupNorm(x + add[n,c]))   2 reads (second one L2-resident) + 1 write, fp32 statistics,
//                          conv bias + time embedding folded in as the per-(n,c) addend, SiLU fused
//   ff_bias_residual_nhwc  out = h + bias[c] + res            conv bias + skip connection in one pass
//   ff_geglu               out = x * gelu(gate)               1 read of [M,2F], 1 write of [M,F]
//   ff_layer_norm          y = LayerNorm(x)                   1 read + 1 write, one warp per token row, row in registers
// All accesses are 128-bit; sums are combined in a fixed order (bit-reproducible run to run).
#include <cuda_bf16.h>

#include <atomic>
#include <cstdlib>

#include "ff_common.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);            // bf16 -> fp32 is a shift
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// SiLU of eight values.  t * sigmoid(t) = t / (1 + 2^(-t * log2 e)): the straightforward form costs TWO MUFU operations per
// element (EX2 + RCP); at 16 MUFU lanes per clock and SM that is 8 elements per clock -- 9 TB/s of algorithmic GroupNorm
// traffic, the same order as the HBM peak -- while taking every reciprocal on the FMA pipe instead (bit-trick seed,
// |rel err| <= 0.101, + three Newton steps as packed FFMA2 / FMUL2: 1.5e-7, fp32-exact for a bf16 result) makes the apply
// loop issue-bound (8 instead of 5 instructions per element).  Measured on B200 (profiles/r2c_gn_kernels.txt, GroupNorm +
// SiLU 32x320x64x64 / 32x1920x32x32): mode 0 = MUFU reciprocal everywhere 62.7 / 94.1 us, mode 2 = Newton everywhere 63.3 /
// 94.8 us, mode 3 = HALF of the pairs each way (both pipes busy) 59.8 / 87.0 us -- the product.  Mode 1
// (t * (0.5 + 0.5 * tanh.approx(t / 2)), one MUFU and 4 instructions: 54.2 / 76.3 us) is NOT used: its absolute error
// |t| * 2.4e-4 is a 70 % relative error at t = -8, where the other modes are exact to fp32 rounding.
#ifndef FF_SILU_MODE
#define FF_SILU_MODE 3
#endif
__device__ __forceinline__ void silu_pairs(float2 (&f)[4]) {
#if FF_SILU_MODE == 0
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = make_float2(__fdividef(f[i].x, 1.f + __expf(-f[i].x)), __fdividef(f[i].y, 1.f + __expf(-f[i].y)));
#elif FF_SILU_MODE == 1
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(0.5f * f[i].x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(0.5f * f[i].y));
    f[i] = __fmul2_rn(f[i], __ffma2_rn(make_float2(0.5f, 0.5f), th, make_float2(0.5f, 0.5f)));
  }
#else
#ifndef FF_SILU_NR
#define FF_SILU_NR 3
#endif
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = f[i];
    const float2 a = __fmul2_rn(t, make_float2(-1.4426950408889634f, -1.4426950408889634f));
    float2 e, r;
#if FF_SILU_MODE == 3
    if (i >= 2) {                                          // (compile-time after unrolling) half of the pairs: MUFU.RCP;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));      // rcp(1 + inf) = 0, no clamp needed here
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
      const float2 d = __fadd2_rn(e, make_float2(1.f, 1.f));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    } else
#endif
    {
      // Newton path: exponent argument clamped to 2^126 so that d = 1 + e stays finite (t < -87: silu(t) = -0 either way)
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(fminf(a.x, 126.f)));
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(fminf(a.y, 126.f)));
      const float2 d = __fadd2_rn(e, make_float2(1.f, 1.f));
      r = make_float2(__uint_as_float(0x7EF311C7u - __float_as_uint(d.x)), __uint_as_float(0x7EF311C7u - __float_as_uint(d.y)));
      const float2 nd = make_float2(-d.x, -d.y), two = make_float2(2.f, 2.f);
#pragma unroll
      for (int it = 0; it < FF_SILU_NR; ++it) r = __fmul2_rn(r, __ffma2_rn(nd, r, two));
    }
    f[i] = __fmul2_rn(t, r);
  }
#endif
}

__device__ __forceinline__ void silu8(float (&f)[8]) {
  float2 q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = make_float2(f[2 * i], f[2 * i + 1]);
  silu_pairs(q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = q[i].x;
    f[2 * i + 1] = q[i].y;
  }
}

// y = act(x * sc + sh) of one bf16x8 vector, pairwise: unpack (shift / mask), FFMA2, SiLU on the pairs, pack
template <bool SILU>
__device__ __forceinline__ uint4 norm_act8(const uint4& raw, const float2 (&sc)[4], const float2 (&sh)[4]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  float2 f[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    f[i] = __ffma2_rn(make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u)), sc[i], sh[i]);
  if (SILU) silu_pairs(f);
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(f[i].x, f[i].y);
    o[i] = *reinterpret_cast<const uint32_t*>(&b);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// thread-block cluster primitives (PTX; sm_90+): split barrier and a distributed-shared-memory load
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  cluster_arrive_release();
  cluster_wait_acquire();
}
__device__ __forceinline__ unsigned int cluster_nctarank() {
  unsigned int v;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(v));
  return v;
}
__device__ __forceinline__ float ld_dsmem_f32(const float* my_smem, unsigned int rank) {
  const uint32_t local = static_cast<uint32_t>(__cvta_generic_to_shared(my_smem));
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_CHUNKS = 64;     // pixel chunks per image (partial statistics per chunk), two-kernel form
constexpr int GN_MAX_CHUNKS_FUSED = 128;   // ... single-read form (the workspace is sized for this one)
#ifndef FF_GN_MLP
#define FF_GN_MLP 8
#endif
constexpr int GN_MLP = FF_GN_MLP;     // independent 16-byte loads in flight per thread (statistics and apply loops)
constexpr int GN_CHUNK_PX = 128;      // pixels per CTA, images of more than 1024 pixels
constexpr int GN_CHUNK_PX_SMALL = 64;  // ... of at most 1024 pixels

// Thread layout shared by the two GroupNorm kernels: `cols` = min(C/8, 256) threads side by side over the 16-byte
// channel vectors of a pixel, R = 256 / cols pixel rows in flight; a thread keeps ONE channel vector column at a time,
// so per-channel constants / accumulators live in registers.
struct GnLayout {
  int CV, cols, R, col, r;
  __device__ GnLayout(int C) {
    CV = C >> 3;
    cols = CV < GN_THREADS ? CV : GN_THREADS;
    R = GN_THREADS / cols;
    col = threadIdx.x % cols;
    r = threadIdx.x / cols;        // r >= R: idle thread (256 is not a multiple of cols)
  }
};

// bf16x8 -> four fp32 pairs (low half, high half of each word: channels 2i, 2i+1)
__device__ __forceinline__ void unpack8_pairs(const uint4& q, float2 (&f)[4]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) f[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}

// partial[n][chunk][g] = (sum, sum of squares) of v = x + add over the pixels of the chunk and the channels of group g.
// Round 2 (third session): the loop accumulates the RAW sums of x as packed pairs (FADD2 / FFMA2: 16 instructions per
// 16-byte vector with the unpack, was 40 with the per-element addend) and folds the per-(n, c) addend in afterwards
// (sum(x + a) = sum(x) + k a,  sum((x + a)^2) = sum(x^2) + 2 a sum(x) + k a^2,  k = pixels this thread walked); full batches
// of GN_MLP loads run without per-load predicates, the ragged end as one predicated batch.  ncu before: 89 executed
// instructions per vector, issue slots 49 % busy on an HBM-bound kernel.
// (5 CTAs per SM measured the same, 6 slower: 58.4 / 60.9 vs 58.1 us at 32x320x64x64)
#ifndef FF_GN_STATS_MINB
#define FF_GN_STATS_MINB 4
#endif
__global__ void __launch_bounds__(GN_THREADS, FF_GN_STATS_MINB)
gn_stats_nhwc_kernel(const uint4* __restrict__ x, const float* __restrict__ add_nc, long long add_ld, float2* __restrict__ partial,
                     int HW, int C, int G, int chunk_px, int n_chunks) {
  extern __shared__ float sm[];                 // [R][2][C] per-row-slot channel sums, then reduced into slot 0
  const GnLayout L(C);
  // CTAs are dispatched in ascending linear order: the statistics pass walks the tensor from its END -- the part the
  // producer (convolution / residual kernel) wrote last and most likely still holds in L2 -- towards its beginning, and
  // the apply pass then walks forwards, starting with what this pass read last.
  const int n = gridDim.y - 1 - blockIdx.y, chunk = gridDim.x - 1 - blockIdx.x;
  const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
  if (L.r < L.R) {
    for (int v = L.col; v < L.CV; v += L.cols) {
      float2 s2[4], q2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s2[i] = q2[i] = make_float2(0.f, 0.f);
      const uint4* px = x + ((size_t)n * HW + p0 + L.r) * L.CV + v;
      const size_t step = (size_t)L.R * L.CV;
      int p = p0 + L.r;
      const int mine = p < p1 ? (p1 - p + L.R - 1) / L.R : 0;     // pixels of this thread
#pragma unroll 1
      for (; p + (GN_MLP - 1) * L.R < p1; p += GN_MLP * L.R, px += GN_MLP * step) {
        uint4 raw[GN_MLP];                                 // independent loads first, then the arithmetic
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u) raw[u] = __ldg(px + u * step);
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u) {
          float2 f[4];
          unpack8_pairs(raw[u], f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            s2[i] = __fadd2_rn(s2[i], f[i]);
            q2[i] = __ffma2_rn(f[i], f[i], q2[i]);
          }
        }
      }
      if (p < p1) {                                        // ragged end: ONE predicated batch (zeros add nothing)
        uint4 raw[GN_MLP];
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u) raw[u] = (p + u * L.R < p1) ? __ldg(px + u * step) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u) {
          float2 f[4];
          unpack8_pairs(raw[u], f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            s2[i] = __fadd2_rn(s2[i], f[i]);
            q2[i] = __ffma2_rn(f[i], f[i], q2[i]);
          }
        }
      }
      float* dst = sm + (size_t)L.r * 2 * C + 8 * v;
      const float k = (float)mine;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a0 = add_nc ? __ldg(add_nc + (size_t)n * add_ld + 8 * v + 2 * i) : 0.f;
        const float a1 = add_nc ? __ldg(add_nc + (size_t)n * add_ld + 8 * v + 2 * i + 1) : 0.f;
        dst[2 * i] = fmaf(k, a0, s2[i].x);
        dst[2 * i + 1] = fmaf(k, a1, s2[i].y);
        dst[C + 2 * i] = fmaf(a0, fmaf(k, a0, 2.f * s2[i].x), q2[i].x);
        dst[C + 2 * i + 1] = fmaf(a1, fmaf(k, a1, 2.f * s2[i].y), q2[i].y);
      }
    }
  }
  __syncthreads();
  // one warp per group: lanes stride over the (row slot, channel) pairs, butterfly reduction -- a fixed order, so the
  // result is bit-reproducible
  const int cpg = C / G, cnt = L.R * cpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < G; g += GN_THREADS / 32) {
    float S = 0.f, SS = 0.f;
    for (int i = lane; i < cnt; i += 32) {
      const int rr = i / cpg, c = i - rr * cpg;
      const float* src = sm + (size_t)rr * 2 * C + g * cpg + c;
      S += src[0];
      SS += src[C];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      S += __shfl_xor_sync(0xffffffffu, S, o);
      SS += __shfl_xor_sync(0xffffffffu, SS, o);
    }
    if (lane == 0) partial[((size_t)n * n_chunks + chunk) * G + g] = make_float2(S, SS);
  }
}

// 4 CTAs per SM (64 registers, 24 bytes of spill) beat 3 (78 registers: 84 vs 58 us at 32x320x64x64) and 5 (66 us)
#ifndef FF_GN_APPLY_MINB
#define FF_GN_APPLY_MINB 4
#endif
template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS, FF_GN_APPLY_MINB)
gn_apply_nhwc_kernel(const uint4* __restrict__ x, const float* __restrict__ add_nc, long long add_ld,
                     const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                     const float2* __restrict__ partial, uint4* __restrict__ y, int HW, int C, int G, int chunk_px,
                     int n_chunks, float eps) {
  extern __shared__ float sm[];                 // [G] mean, [G] rstd
  float* mean = sm;
  float* rstd = sm + G;
  const GnLayout L(C);
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int cpg = C / G;
  const float inv_cnt = 1.f / ((float)cpg * (float)HW);
  {
    // one warp per group, one chunk per lane (n_chunks <= 64), butterfly reduction: the loads are independent (one
    // latency instead of n_chunks) and every CTA of image n derives bit-identical statistics
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int g = warp; g < G; g += GN_THREADS / 32) {
      float S = 0.f, SS = 0.f;
      for (int ch = lane; ch < n_chunks; ch += 32) {
        const float2 t = __ldg(partial + ((size_t)n * n_chunks + ch) * G + g);
        S += t.x;
        SS += t.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
      }
      if (lane == 0) {
        const float m = S * inv_cnt;
        const float var = fmaxf(SS * inv_cnt - m * m, 0.f);
        mean[g] = m;
        rstd[g] = 1.f / sqrtf(var + eps);
      }
    }
  }
  __syncthreads();
  if (L.r >= L.R) return;
  const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
  for (int v = L.col; v < L.CV; v += L.cols) {
    float2 sc[4], sh[4];                                   // (channel pairs: the loop below is packed FFMA2)
    {
      int g = (8 * v) / cpg, rem = 8 * v - g * cpg;       // one division per vector; the group advances with the channel
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (rem == cpg) {
          rem = 0;
          ++g;
        }
        ++rem;
        const float a = add_nc ? __ldg(add_nc + (size_t)n * add_ld + 8 * v + j) : 0.f;
        const float scj = rstd[g] * __bfloat162float(gamma[8 * v + j]);
        const float shj = fmaf(a - mean[g], scj, __bfloat162float(beta[8 * v + j]));     // y = (x + a - mean) * rstd * gamma + beta
        if (j & 1) {
          sc[j >> 1].y = scj;
          sh[j >> 1].y = shj;
        } else {
          sc[j >> 1].x = scj;
          sh[j >> 1].x = shj;
        }
      }
    }
    const size_t off = ((size_t)n * HW + p0 + L.r) * L.CV + v;
    const uint4* px = x + off;
    uint4* py = y + off;
    const size_t step = (size_t)L.R * L.CV;
    int p = p0 + L.r;
#pragma unroll 1
    for (; p + (GN_MLP - 1) * L.R < p1; p += GN_MLP * L.R, px += GN_MLP * step, py += GN_MLP * step) {
      uint4 raw[GN_MLP];                                   // (independent loads first: see the statistics kernel)
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) raw[u] = __ldg(px + u * step);
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) py[u * step] = norm_act8<SILU>(raw[u], sc, sh);
    }
    if (p < p1) {                                          // ragged end: one predicated batch
      uint4 raw[GN_MLP];
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) raw[u] = (p + u * L.R < p1) ? __ldg(px + u * step) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) {
        if (p + u * L.R < p1) py[u * step] = norm_act8<SILU>(raw[u], sc, sh);
      }
    }
  }
}

// ---- register-resident GroupNorm for small images (round 2) ------------------------------------------------------------
// The 8 x 8 / 16 x 16 / 32 x 32 levels of the UNet are HALF of the GroupNorm launches of a call, and the two-kernel form
// is latency-bound there: its CTAs split an image by pixels only, so a 32 x 1280 x 8 x 8 tensor (5 MB) is walked by 32 CTAs
// whose threads each chase 64 dependent loads (ncu launch list of round 2: 35-50 us per kernel for tensors that take
// 2-5 us at the HBM peak; 60 % of all GroupNorm time of a UNet call).  Groups are independent, so this kernel splits by
// CHANNELS instead: one CTA per (image, bundle of B groups whose channels fill whole 16-byte vectors), all HW pixels of
// those channels in registers (VPT vectors per thread, all loads in flight at once), statistics and apply in one pass --
// x is read exactly once, there is no workspace traffic, and a 32 x 1280 x 16 x 16 launch has 1024 CTAs instead of 128.
//
// Cluster form (CL, experiment switch FF_GN_CLUSTER=1 -- see gn_small_plan): images too large for one CTA's registers (64 x 64 and up, and the wide 32 x 32 concatenations) are
// split by PIXELS over a thread-block cluster of S <= 8 CTAs on top of the channel split; each CTA reduces its part, the
// per-group sums are exchanged through distributed shared memory between two cluster barriers (every CTA adds the S
// partials in rank order: identical statistics everywhere, bit-reproducible), and the apply runs from registers as before.
// Still one read of x and no workspace; the only cross-CTA traffic is 2 * B floats per CTA.
template <bool SILU, int VPT, bool CL>
__global__ void __launch_bounds__(GN_THREADS, VPT >= 32 ? 1 : (VPT >= 12 ? 2 : 1))
gn_small_nhwc_kernel(const uint4* __restrict__ x, const float* __restrict__ add_nc, long long add_ld,
                     const __nv_bfloat16* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y, int HW, int C, int G, int B, int P, float eps) {
  __shared__ float sm[GN_THREADS / 32 * 8 * 2 + 16];   // [warp][group][S, SS] partials, then [B] mean, [B] rstd
  __shared__ float xch[16];                            // CL: this CTA's [group][S, SS], read by the cluster peers
  const int cpg = C / G, bch = B * cpg, BV = bch >> 3, CV = C >> 3;
  const int R = GN_THREADS / BV, col = threadIdx.x % BV, r = threadIdx.x / BV;
  const bool live = r < R;
  const int n = blockIdx.y, bundle = blockIdx.x;
  const int pb = CL ? (int)blockIdx.z * P : 0;              // my pixel range [pb, pe)
  const int pe = CL ? min(HW, pb + P) : HW;
  const int c0 = bundle * bch + 8 * col;        // my eight channels
  float* stat = sm + GN_THREADS / 32 * 8 * 2;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = add_nc ? __ldg(add_nc + (size_t)n * add_ld + c0 + j) : 0.f;
  uint4 raw[VPT];
  const uint4* px = x + ((size_t)n * HW) * CV + bundle * BV + col;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int p = pb + r + k * R;
    raw[k] = (live && p < pe) ? __ldg(px + (size_t)p * CV) : make_uint4(0u, 0u, 0u, 0u);
  }
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    if (live && pb + r + k * R < pe) {
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = f[j] + a[j];
        s[j] += t;
        ss[j] = fmaf(t, t, ss[j]);
      }
    }
  }
  // ---- per-group sums of the bundle: each thread folds its eight channels into (at most B <= 8) group partials, a warp
  // butterfly and one shared-memory hop combine them -- fixed order, bit-reproducible, ~100 instructions per thread
  // (a first version reduced [R][2][bch] channel sums from shared memory with one warp per group: with B = 1 that was ONE
  // warp walking 2040 values while seven idled)
  int gid[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) gid[j] = (8 * col + j) / cpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    if (g < B) {                                            // (CTA-uniform)
      float S = 0.f, SS = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        S += gid[j] == g ? s[j] : 0.f;
        SS += gid[j] == g ? ss[j] : 0.f;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
      }
      if (lane == 0) {
        sm[(warp * 8 + g) * 2] = S;
        sm[(warp * 8 + g) * 2 + 1] = SS;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < B) {
    float S = 0.f, SS = 0.f;
#pragma unroll
    for (int w = 0; w < GN_THREADS / 32; ++w) {
      S += sm[(w * 8 + threadIdx.x) * 2];
      SS += sm[(w * 8 + threadIdx.x) * 2 + 1];
    }
    if (CL) {
      xch[2 * threadIdx.x] = S;
      xch[2 * threadIdx.x + 1] = SS;
    } else {
      const float inv_cnt = 1.f / ((float)cpg * (float)HW);
      const float m = S * inv_cnt;
      const float var = fmaxf(SS * inv_cnt - m * m, 0.f);
      stat[threadIdx.x] = m;
      stat[B + threadIdx.x] = 1.f / sqrtf(var + eps);
    }
  }
  if (CL) {
    cluster_sync_all();                                     // partials of every CTA of the cluster are in place
    if (threadIdx.x < B) {
      float S = 0.f, SS = 0.f;
      const unsigned int nr = cluster_nctarank();
      for (unsigned int rk = 0; rk < nr; ++rk) {            // rank order: the same sum in every CTA
        S += ld_dsmem_f32(&xch[2 * threadIdx.x], rk);
        SS += ld_dsmem_f32(&xch[2 * threadIdx.x + 1], rk);
      }
      const float inv_cnt = 1.f / ((float)cpg * (float)HW);
      const float m = S * inv_cnt;
      const float var = fmaxf(SS * inv_cnt - m * m, 0.f);
      stat[threadIdx.x] = m;
      stat[B + threadIdx.x] = 1.f / sqrtf(var + eps);
    }
    cluster_arrive_release();                               // my reads of the peers are done (waited for at the exit)
  }
  __syncthreads();
  if (live) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = gid[j];
      sc[j] = stat[B + g] * __bfloat162float(gamma[c0 + j]);
      sh[j] = fmaf(a[j] - stat[g], sc[j], __bfloat162float(beta[c0 + j]));     // y = (x + a - mean) * rstd * gamma + beta
    }
    uint4* py = y + ((size_t)n * HW) * CV + bundle * BV + col;
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
      const int p = pb + r + k * R;
      if (p < pe) {
        float f[8];
        unpack8(raw[k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
        if (SILU) silu8(f);
        py[(size_t)p * CV] = pack8(f);
      }
    }
  }
  if (CL) cluster_wait_acquire();                           // nobody leaves while a peer may still read its xch[]
}

struct GnSmallArgs {
  const uint4* x;
  const float* add_nc;
  long long add_ld;
  const __nv_bfloat16 *gamma, *beta;
  uint4* y;
  int HW, C, G, B, P;
  float eps;
};

// grid.z = cluster size (the cluster spans z only); CL launches go through cudaLaunchKernelEx with the cluster attribute
template <int VPT, bool CL>
cudaError_t gn_small_launch(const GnSmallArgs& a, dim3 grid, bool silu, cudaStream_t st) {
  auto kern = silu ? gn_small_nhwc_kernel<true, VPT, CL> : gn_small_nhwc_kernel<false, VPT, CL>;
  if (!CL) {
    kern<<<grid, GN_THREADS, 0, st>>>(a.x, a.add_nc, a.add_ld, a.gamma, a.beta, a.y, a.HW, a.C, a.G, a.B, a.P, a.eps);
    return cudaSuccess;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(GN_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = grid.z;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, a.x, a.add_nc, a.add_ld, a.gamma, a.beta, a.y, a.HW, a.C, a.G, a.B, a.P, a.eps);
}

// groups per bundle B (the smallest B dividing G whose B * cpg channels are whole 16-byte vectors), vectors per thread
// and cluster size S (1 = plain launch); 0 = the shape does not fit the register-resident form
int gn_small_plan(int HW, int C, int G, int* B_out, int* S_out, int* P_out) {
  static const bool off = [] { const char* e = getenv("FF_GN_SMALL"); return e && atoi(e) == 0; }();
  // FF_GN_CLUSTER=1: cluster form for shapes beyond 24 vectors per thread -- an EXPERIMENT switch: measured on B200 it is
  // parity-green but slower than the two-kernel form at every such shape (77 vs 65 us at 32x320x64x64, 226 vs 157 us at
  // 32x960x64x64, 117 vs 85 us at 32x1920x32x32: with 2 CTAs of 128 registers per SM the load, reduce and apply phases
  // run in lock-step and nothing overlaps them, profiles/r2c_gn_kernels.txt);  FF_GN_CLUSTER_VPT: vectors per thread the
  // cluster split aims for (tuning knob)
  static const bool cl_on = [] { const char* e = getenv("FF_GN_CLUSTER"); return e && atoi(e) != 0; }();
  static const int cl_vpt = [] { const char* e = getenv("FF_GN_CLUSTER_VPT"); const int v = e ? atoi(e) : 0; return v >= 2 && v <= 32 ? v : 24; }();
  if (off) return 0;
  const int cpg = C / G;
  int B = 0;
  for (int b = 1; b <= 8; b <<= 1)
    if (G % b == 0 && (b * cpg) % 8 == 0) { B = b; break; }
  if (!B) return 0;
  const int BV = B * cpg / 8;
  if (BV > GN_THREADS) return 0;
  const int R = GN_THREADS / BV;
  auto bucket = [](int need) { return need <= 2 ? 2 : (need <= 4 ? 4 : (need <= 6 ? 6 : (need <= 12 ? 12 : (need <= 24 ? 24 : 32)))); };
  int S = 1, P = HW, need = (HW + R - 1) / R;
  if (need > 24) {
    if (!cl_on) return 0;
    int best = 0;
    for (int sz = 2; sz <= 8; sz <<= 1) {
      const int p = (HW + sz - 1) / sz, nd = (p + R - 1) / R;
      if ((sz - 1) * p >= HW) continue;                      // (an empty last CTA: pointless split)
      if (nd <= 32) {
        best = sz;
        if (nd <= cl_vpt) break;
      }
    }
    if (!best) return 0;
    S = best;
    P = (HW + S - 1) / S;
    need = (P + R - 1) / R;
  }
  *B_out = B;
  *S_out = S;
  *P_out = P;
  return bucket(need);
}

// ---- single-read GroupNorm (round 2) -----------------------------------------------------------------------------------
// The two-kernel form above moves three passes of real traffic (statistics read, apply read, write): 0.33 of the HBM peak
// on ALGORITHMIC bytes (x once + y once).  Here ONE kernel keeps its chunk of the image in shared memory between the
// statistics and the apply phase: x is read from HBM exactly once.  The CTAs of an image meet at a per-image barrier in
// global memory (partial sums published, counter incremented, everybody spins until the counter reaches n_chunks).
// Forward progress: a CTA takes its LOGICAL index from an atomic ticket when it starts, images are contiguous ticket
// ranges, so every CTA of an earlier image is already running (or done) and can finish without waiting for anybody else;
// of the image that is only partly started at most n_chunks - 1 <= 63 CTAs wait, fewer than the 148 that are resident even
// at one CTA per SM, so a slot for the next ticket always frees up.  The counters live in the caller's workspace, which
// must be ZERO before the first call and is left zeroed by every call (self-cleaning), see include/freefine_b200.h.
struct GnCtrl {                       // head of the workspace
  unsigned int ticket, pad[3];
};

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS)
gn_fused_nhwc_kernel(const uint4* __restrict__ x, const float* __restrict__ add_nc, long long add_ld,
                     const __nv_bfloat16* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y, unsigned int* __restrict__ ctrl,
                     float2* __restrict__ partial, int N, int HW, int C, int G, int chunk_px, int n_chunks, float eps) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const GnLayout L(C);
  uint4* tile = reinterpret_cast<uint4*>(gsm);                                    // [chunk_px][CV]
  float* red = reinterpret_cast<float*>(gsm + (size_t)chunk_px * L.CV * sizeof(uint4));   // [R][2][C]
  float* mean = red + (size_t)L.R * 2 * C;                                        // [G]
  float* rstd = mean + G;                                                         // [G]
  __shared__ unsigned int s_id;
  __shared__ __align__(8) unsigned long long s_bar;
  unsigned int* done = ctrl + 4;        // [N] CTAs of image n that have published their partial sums
  unsigned int* left = done + N;        // [N] CTAs of image n that have read the partial sums
  if (threadIdx.x == 0) {
    s_id = atomicAdd(&ctrl[0], 1u);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned int id = s_id;
  const int n = (int)(id / (unsigned)n_chunks), chunk = (int)(id % (unsigned)n_chunks);
  const int p0 = chunk * chunk_px, p1 = min(HW, p0 + chunk_px);
  const int cpg = C / G;
  // ---- phase 1: HBM -> shared memory.  A chunk of an NHWC image is one contiguous run of bytes: a single thread issues it
  // as bulk asynchronous copies (cp.async.bulk, completion on an mbarrier), so the whole 80-120 KB chunk is in flight at
  // once -- with per-thread 16-byte loads a CTA keeps ~16 KB in flight and the phase is latency-bound (measured: the
  // single-read kernel was SLOWER than the two-kernel form, 96 vs 76 us at 32x320x64x64).
  {
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
    if (threadIdx.x == 0) {
      const uint32_t bytes = (uint32_t)(p1 - p0) * (uint32_t)C * 2u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      const char* src = reinterpret_cast<const char*>(x + ((size_t)n * HW + p0) * L.CV);
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile);
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t sz = bytes - off < 32768u ? bytes - off : 32768u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + off), "l"(src + off), "r"(sz), "r"(bar) : "memory");
      }
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar) : "memory");
    }
  }
  // per-thread channel-column sums over the chunk, from shared memory
  if (L.r < L.R) {
    for (int v = L.col; v < L.CV; v += L.cols) {
      float a[8], sacc[8], ssacc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] = add_nc ? __ldg(add_nc + (size_t)n * add_ld + 8 * v + j) : 0.f;
        sacc[j] = ssacc[j] = 0.f;
      }
      const uint4* pt = tile + (size_t)L.r * L.CV + v;
      const size_t step = (size_t)L.R * L.CV;
#pragma unroll 4
      for (int p = p0 + L.r; p < p1; p += L.R, pt += step) {
        float f[8];
        unpack8(*pt, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = f[j] + a[j];
          sacc[j] += t;
          ssacc[j] = fmaf(t, t, ssacc[j]);
        }
      }
      float* dst = red + (size_t)L.r * 2 * C + 8 * v;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dst[j] = sacc[j];
        dst[C + j] = ssacc[j];
      }
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int cnt = L.R * cpg;
    for (int g = warp; g < G; g += GN_THREADS / 32) {
      float S = 0.f, SS = 0.f;
      for (int i = lane; i < cnt; i += 32) {
        const int rr = i / cpg, c = i - rr * cpg;
        const float* src = red + (size_t)rr * 2 * C + g * cpg + c;
        S += src[0];
        SS += src[C];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
      }
      if (lane == 0) partial[((size_t)n * n_chunks + chunk) * G + g] = make_float2(S, SS);
    }
  }
  // ---- per-image barrier: publish, then wait for the other chunks of this image
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&done[n], 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + n) : "memory");
      if (seen < (unsigned)n_chunks) __nanosleep(64);
    } while (seen < (unsigned)n_chunks);
  }
  __syncthreads();
  // ---- phase 2: statistics of the whole image (every CTA of image n derives bit-identical values)
  {
    const float inv_cnt = 1.f / ((float)cpg * (float)HW);
    for (int g = warp; g < G; g += GN_THREADS / 32) {
      float S = 0.f, SS = 0.f;
      for (int ch = lane; ch < n_chunks; ch += 32) {
        const float2 t = __ldcg(partial + ((size_t)n * n_chunks + ch) * G + g);      // written by other SMs: bypass L1
        S += t.x;
        SS += t.y;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        SS += __shfl_xor_sync(0xffffffffu, SS, o);
      }
      if (lane == 0) {
        const float m = S * inv_cnt;
        const float var = fmaxf(SS * inv_cnt - m * m, 0.f);
        mean[g] = m;
        rstd[g] = 1.f / sqrtf(var + eps);
      }
    }
  }
  __syncthreads();
  // the partial sums of this image have been read by this CTA: the last reader re-arms the image's counters, the CTA
  // with the last ticket re-arms the ticket (every ticket has been handed out by then)
  if (threadIdx.x == 0) {
    if (atomicAdd(&left[n], 1u) == (unsigned)n_chunks - 1u) {
      done[n] = 0u;
      left[n] = 0u;
    }
    if (id == (unsigned)N * (unsigned)n_chunks - 1u) ctrl[0] = 0u;
  }
  // ---- phase 3: apply from shared memory
  if (L.r >= L.R) return;
  for (int v = L.col; v < L.CV; v += L.cols) {
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * v + j, g = c / cpg;
      const float a = add_nc ? __ldg(add_nc + (size_t)n * add_ld + c) : 0.f;
      sc[j] = rstd[g] * __bfloat162float(gamma[c]);
      sh[j] = fmaf(a - mean[g], sc[j], __bfloat162float(beta[c]));     // y = (x + a - mean) * rstd * gamma + beta
    }
    const uint4* pt = tile + (size_t)L.r * L.CV + v;
    uint4* py = y + ((size_t)n * HW + p0 + L.r) * L.CV + v;
    const size_t step = (size_t)L.R * L.CV;
#pragma unroll 4
    for (int p = p0 + L.r; p < p1; p += L.R, pt += step, py += step) {
      float f[8];
      unpack8(*pt, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sh[j]);
      if (SILU) silu8(f);
      *py = pack8(f);
    }
  }
}

// out[m, c] = h[m, c] + bias[c] + res[m, c]   (bias / res optional), CV = C/8 vectors per row
__global__ void __launch_bounds__(256)
bias_residual_kernel(const uint4* __restrict__ h, const __nv_bfloat16* __restrict__ bias,
                     const uint4* __restrict__ res, uint4* __restrict__ out, long long total, int CV) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(__ldg(h + i), f);
    if (bias) {
      const int v = (int)(i % CV);
      float b[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + v), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += b[j];
    }
    if (res) {
      float r[8];
      unpack8(__ldg(res + i), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += r[j];
    }
    out[i] = pack8(f);
  }
}

// out[m, f] = x * gelu(gate), x = h[m, f], gate = h[m, F + f]; gelu = erf form, with the intermediate bf16 rounding of
// gelu(gate) of the eager pair of kernels (F.gelu -> bf16 tensor, then a bf16 multiply).
// gelu(g) = g * Phi(g), Phi from erfc(z) = t*(a1 + t*(a2 + t*(a3 + t*(a4 + t*a5)))) * exp(-z^2), t = 1/(1 + p z), z =
// |g|/sqrt(2)  (Abramowitz & Stegun 7.1.26, absolute error 1.5e-7 -- four orders below the bf16 rounding of the result):
// 2 MUFU ops + ~12 FMA-pipe ops per element instead of libdevice erff's ~30, which made the kernel ALU-bound.
// Round 2 (third session): the scalar form of this (one element at a time, x unpacked to fp32) executed ~26 instructions
// per element and the kernel was ISSUE-bound (335 M elements x 26 / 32 lanes / 592 schedulers = 186 us of issue slots for
// a launch that took 197 us; the HBM floor is 154 us).  gelu_erf_x_bf16x2 evaluates a PAIR of gates with packed FFMA2 / FMUL2 (0.5 folded into the polynomial,
// gelu = g / 2 + |g| * (1/2 - erfc/2): no select), rounds the pair to bf16x2 and multiplies it with the still-packed x
// pair by HMUL2.BF16 -- exact product, ONE rounding: what the eager bf16 multiply does -- so x is never
// unpacked: ~12 instructions per element.  The launch is then bound by PIPE throughput, not issue slots: per element 10.5
// fp32 FMA-pipe operations (128 lanes per clock and SM; FFMA2 saves issue slots, not lanes) + 2 MUFU operations (16 lanes)
// = 95 / 149 us for 335 M elements against an HBM floor of 154 us; a Newton reciprocal (seed + 3 steps, see silu8) moves
// one MUFU operation to six FMA ones.  Measured at 131072 x 1280 with 0 / 1 / 2 pairs in four on the FMA pipe: 178.2 / 188.1 /
// 189.6 us -- every FMA-pipe operation added costs more than the MUFU operation it replaces, so all reciprocals stay on
// the MUFU pipe (FF_GEGLU_NR_PAIRS = 0).
// pairs per 16-byte vector (of 4) whose reciprocal runs on the FMA pipe -- see the balance below
#ifndef FF_GEGLU_NR_PAIRS
#define FF_GEGLU_NR_PAIRS 0
#endif
template <bool NEWTON>
__device__ __forceinline__ uint32_t gelu_erf_x_bf16x2(uint32_t gate, uint32_t xw) {
  const float2 g = make_float2(__uint_as_float(gate << 16), __uint_as_float(gate & 0xffff0000u));
  const float2 ag = make_float2(fabsf(g.x), fabsf(g.y));
  const float2 den = __ffma2_rn(ag, make_float2(0.23164189045f, 0.23164189045f), make_float2(1.f, 1.f));   // 1 + p |g| / sqrt 2
  float2 t;
  if (NEWTON) {
    t = make_float2(__uint_as_float(0x7EF311C7u - __float_as_uint(den.x)), __uint_as_float(0x7EF311C7u - __float_as_uint(den.y)));
    const float2 nd = make_float2(-den.x, -den.y), two = make_float2(2.f, 2.f);
#pragma unroll
    for (int it = 0; it < 3; ++it) t = __fmul2_rn(t, __ffma2_rn(nd, t, two));
  } else {
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  }
  // p = erfc polynomial / 2 (Abramowitz & Stegun 7.1.26 coefficients halved)
  float2 p = __ffma2_rn(t, make_float2(0.5307027145f, 0.5307027145f), make_float2(-0.7265760135f, -0.7265760135f));
  p = __ffma2_rn(p, t, make_float2(0.7107068705f, 0.7107068705f));
  p = __ffma2_rn(p, t, make_float2(-0.142248368f, -0.142248368f));
  p = __ffma2_rn(p, t, make_float2(0.127414796f, 0.127414796f));
  const float2 arg = __fmul2_rn(__fmul2_rn(g, g), make_float2(-0.7213475204444817f, -0.7213475204444817f));   // -z^2 log2 e
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
  const float2 h = __fmul2_rn(__fmul2_rn(p, t), e);                                       // erfc(|g| / sqrt 2) / 2
  // g Phi(g) = relu(g) - |g| h  (g >= 0: g (1 - h); g < 0: g h): the relu is an ALU-pipe FMNMX, ONE FMA-pipe operation left
  const float2 ge = __ffma2_rn(make_float2(-ag.x, -ag.y), h, make_float2(fmaxf(g.x, 0.f), fmaxf(g.y, 0.f)));
  const __nv_bfloat162 gb = __floats2bfloat162_rn(ge.x, ge.y);                            // the bf16 tensor F.gelu returns
  const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&xw), gb);   // bf16 * bf16, one rounding
  return *reinterpret_cast<const uint32_t*>(&r);
}

__device__ __forceinline__ uint4 geglu8(const uint4& qx, const uint4& qg) {
  return make_uint4(gelu_erf_x_bf16x2<false>(qg.x, qx.x), gelu_erf_x_bf16x2<FF_GEGLU_NR_PAIRS >= 1>(qg.y, qx.y),
                    gelu_erf_x_bf16x2<false>(qg.z, qx.z), gelu_erf_x_bf16x2<FF_GEGLU_NR_PAIRS >= 2>(qg.w, qx.w));
}

// One warp per row of h, lanes over its 16-byte vectors, GEGLU_U vectors of x and of the gate in flight per lane: the
// flat-index kernel below spends ~100 of its ~200 instructions per trip on 64-bit (row, vector) arithmetic and selects
// (cuobjdump), here a vector costs one add.  Rows of F = 1280 / 2560 / 5120 channels are 160 / 320 / 640 vectors: whole
// trips of 32 x 5 lanes-vectors, no idle lanes.
constexpr int GEGLU_U = 5;
__global__ void __launch_bounds__(256)
geglu_rows_kernel(const uint4* __restrict__ h, uint4* __restrict__ out, long long M, int FV) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += nwarps) {
    const uint4* row = h + m * 2 * FV + lane;
    uint4* orow = out + m * FV + lane;
    for (int v0 = 0; v0 < FV; v0 += 32 * GEGLU_U) {
      uint4 qx[GEGLU_U], qg[GEGLU_U];
#pragma unroll
      for (int u = 0; u < GEGLU_U; ++u) {
        if (v0 + 32 * u + lane < FV) {
          qx[u] = __ldg(row + v0 + 32 * u);
          qg[u] = __ldg(row + FV + v0 + 32 * u);
        }
      }
#pragma unroll
      for (int u = 0; u < GEGLU_U; ++u)
        if (v0 + 32 * u + lane < FV) orow[v0 + 32 * u] = geglu8(qx[u], qg[u]);
    }
  }
}

// (dm, dv) = (2 * stride) div / mod FV from the host: the (row, vector) pair advances without a 64-bit division per trip
__global__ void __launch_bounds__(256)
geglu_kernel(const uint4* __restrict__ h, uint4* __restrict__ out, long long total, int FV, long long dm, int dv) {
  // two output vectors per trip: four independent 16-byte loads in flight per thread before the erf arithmetic
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  long long m = i / FV, m2 = (i + stride) / FV;
  int v = (int)(i - m * FV), v2 = (int)(i + stride - m2 * FV);
  for (; i < total; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < total;
    const uint4* row = h + m * 2 * FV;
    const uint4* row2 = h + (two ? m2 : m) * 2 * FV;
    const int w2 = two ? v2 : v;
    const uint4 qx = __ldg(row + v), qg = __ldg(row + FV + v);
    const uint4 qx2 = __ldg(row2 + w2), qg2 = __ldg(row2 + FV + w2);
    out[i] = geglu8(qx, qg);
    if (two) out[i2] = geglu8(qx2, qg2);
    m += dm; v += dv;
    if (v >= FV) { v -= FV; ++m; }
    m2 += dm; v2 += dv;
    if (v2 >= FV) { v2 -= FV; ++m2; }
  }
}

// One warp per ROWS consecutive rows of C = 8*CV channels (CV <= 32*VPL): all row loads are issued before the first
// reduction (memory-level parallelism: a single 640-byte row per warp left the kernel latency-bound), each row stays in
// registers for its mean, centred variance and output.
template <int VPL, int ROWS>
__global__ void __launch_bounds__(256)
layer_norm_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                  const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y, long long M, int CV, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (row0 >= M) return;
  uint4 raw[ROWS][VPL];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = lane + 32 * k;
      raw[r][k] = (row0 + r < M && v < CV) ? __ldg(x + (row0 + r) * CV + v) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const float inv_c = 1.f / (float)(8 * CV);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (row0 + r >= M) break;                    // warp-uniform
    float f[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      unpack8(raw[r][k], f[k]);                  // (vectors beyond CV are zeros: they add nothing to the sum)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[k][j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (lane + 32 * k < CV) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[k][j] - mean;
          ss = fmaf(d, d, ss);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = 1.f / sqrtf(ss * inv_c + eps);
    uint4* py = y + (row0 + r) * CV;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int v = lane + 32 * k;
      if (v < CV) {
        float ga[8], be[8], o[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + v), ga);
        unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + v), be);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((f[k][j] - mean) * rstd, ga[j], be[j]);
        py[v] = pack8(o);
      }
    }
  }
}

// Nearest 2x up-sampling and channel concatenation of dense NHWC tensors (Upsample2D / the up-block skip connections of the
// UNet, diffusers blocks walked by override_forward, src/utils/attention.py:13-223).  ATen's NHWC kernels for these run at
// 0.4 TB/s (upsample_nearest2d_nhwc_out_frame: 0.5 ms for 32x640x32x32 -> 64x64) and 1.8 TB/s (CatArrayBatchedCopy) in the
// round-2 launch list: 3.7 % of a pair of UNet calls for pure copies.  One 16-byte vector per thread and trip, two trips in
// flight; every warp reads and writes whole contiguous channel rows.
__global__ void __launch_bounds__(256)
upsample2x_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, unsigned total, unsigned W, unsigned CV) {
  // 32-bit index arithmetic (the host checks 4 * total < 2^31): 64-bit divisions would make this copy ALU-bound
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned row = 2u * W * CV;                             // one output pixel row, in vectors
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const unsigned i2 = i + stride;
    const bool two = i2 < total;
    const uint4 a = __ldg(x + i);
    const uint4 b = two ? __ldg(x + i2) : a;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !two) break;
      const unsigned j = k ? i2 : i;
      const uint4 val = k ? b : a;
      const unsigned p = j / CV, v = j - p * CV;
      const unsigned nh = p / W, w = p - nh * W;                // nh = n * H + h
      uint4* o = y + (size_t)(2u * nh) * row + (2u * w) * CV + v;
      o[0] = val;
      o[CV] = val;
      o[row] = val;
      o[row + CV] = val;
    }
  }
}

__global__ void __launch_bounds__(256)
concat_nhwc_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, unsigned total, unsigned CVa,
                   unsigned CVb) {
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned CVt = CVa + CVb;
  auto src = [&](unsigned i) -> const uint4* {
    const unsigned m = i / CVt, v = i - m * CVt;
    return v < CVa ? a + (size_t)m * CVa + v : b + (size_t)m * CVb + (v - CVa);
  };
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const unsigned i2 = i + stride;
    const bool two = i2 < total;
    const uint4 va = __ldg(src(i));
    const uint4 vb = two ? __ldg(src(i2)) : va;
    out[i] = va;
    if (two) out[i2] = vb;
  }
}

// LayerNorm for C = 40 * LPR channels (320 / 640 / 1280: every transformer width of SD1.5): LPR = 8 / 16 / 32 lanes share a
// row, FIVE 16-byte vectors per lane -- no idle lanes (the generic kernel above gives a 320-channel row to a whole warp:
// 40 vectors over 32 lanes = two rounds with 24 lanes idle in the second, and a five-step butterfly where three suffice).
// Round 2 (third session): the row stays PACKED (20 registers) and is unpacked again in each of the three passes (sum,
// centred variance, output) instead of living as 40 fp32 registers -- 80 registers, three CTAs per SM instead of two, i.e.
// 61 KB of loads in flight per SM instead of 41 KB (the kernel is latency-bound: ~10 instructions per element); the
// arithmetic runs on channel pairs (FADD2 / FFMA2).
#ifndef FF_LN5_MINB
#define FF_LN5_MINB 3
#endif
__device__ __forceinline__ uint4 ld_nc_volatile(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// optimisation barrier: the four words are "modified", so a later unpack is recomputed from them instead of being kept
__device__ __forceinline__ void keep_packed(uint4& q) { asm volatile("" : "+r"(q.x), "+r"(q.y), "+r"(q.z), "+r"(q.w)); }

template <int LPR>
__global__ void __launch_bounds__(256, FF_LN5_MINB)
layer_norm5_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                   const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y, long long M, float eps) {
  constexpr int RPW = 32 / LPR, CV = 5 * LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const uint4* g4 = reinterpret_cast<const uint4*>(gamma) + sub;
  const uint4* b4 = reinterpret_cast<const uint4*>(beta) + sub;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)(8 * CV);
  // FF_LN5_GROUPS row groups per trip, their loads all issued before the first reduction: 2 groups (10 x 16 bytes in flight
  // per lane) measured SLOWER than 1 (35.9 vs 29.9 us at 131072 x 320: 140 bytes of spills at 80 registers), so 1 it is
#ifndef FF_LN5_GROUPS
#define FF_LN5_GROUPS 1
#endif
  constexpr int NG = FF_LN5_GROUPS;
  for (long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < M; row0 += NG * nwarps * RPW) {
    uint4 raw[NG][5];
    bool valid[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const long long row = row0 + g * nwarps * RPW + rsel;
      valid[g] = row < M;
#pragma unroll
      for (int k = 0; k < 5; ++k) raw[g][k] = valid[g] ? __ldg(x + row * CV + sub + LPR * k) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const long long row = row0 + g * nwarps * RPW + rsel;
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float2 f[4];
        unpack8_pairs(raw[g][k], f);
#pragma unroll
        for (int i = 0; i < 4; ++i) s2 = __fadd2_rn(s2, f[i]);
      }
      float s = s2.x + s2.y;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      const float2 nm = make_float2(-mean, -mean);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float2 f[4];
        keep_packed(raw[g][k]);          // (without it the compiler keeps the 40 unpacked values of the first pass alive)
        unpack8_pairs(raw[g][k], f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 d = __fadd2_rn(f[i], nm);
          q2 = __ffma2_rn(d, d, q2);
        }
      }
      float ss = q2.x + q2.y;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = 1.f / sqrtf(ss * inv_c + eps);
      if (valid[g]) {
        const float2 r2 = make_float2(rstd, rstd);
        uint4* py = y + row * CV + sub;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          float2 f[4], ga[4], be[4];
          keep_packed(raw[g][k]);
          // gamma / beta come from L1 every row (volatile loads: hoisted out of the row loop they cost 40 registers packed,
          // 80 unpacked -- the occupancy this version is about)
          unpack8_pairs(ld_nc_volatile(g4 + LPR * k), ga);
          unpack8_pairs(ld_nc_volatile(b4 + LPR * k), be);
          unpack8_pairs(raw[g][k], f);
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 o = __ffma2_rn(__fmul2_rn(__fadd2_rn(f[i], nm), r2), ga[i], be[i]);
            const __nv_bfloat162 b = __floats2bfloat162_rn(o.x, o.y);
            w[i] = *reinterpret_cast<const uint32_t*>(&b);
          }
          py[LPR * k] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
}

// Pixels per CTA (at most GN_MAX_CHUNKS chunks per image).  Measured with profiles/gn_case.py (graph replays, us, 32 images;
// profiles/r1b_gn_chunk_sweep.txt):        chunk_px   320x64^2   960x64^2   640x32^2   1280x16^2
//                                               256       76.4      197.1       83.9      125.5
//                                               128       76.0      203.1       49.3       67.9
//                                                64       96.3      209.5       43.1       38.9
//                                                32     (=64: chunk cap)        45.1       25.9
// Large images want ~128 pixels per CTA (per-CTA prologue: statistics butterfly + per-channel coefficients), small ones
// want enough CTAs to fill 148 SMs.  Sizing the chunks for exactly ONE wave (148 x 4 CTAs) was 10-25 % slower than 1.7
// waves of 128-pixel CTAs: the load phase of one CTA then no longer overlaps the reduction / write-back of another.
// The product rule only uses the two sizes the whole parity suite has run with (64 and 128); FF_GN_CHUNK_PX overrides.
int gn_chunks(int HW, int N, int* chunk_px) {
  (void)N;
  static const int forced = [] {
    const char* e = getenv("FF_GN_CHUNK_PX");
    const int v = e ? atoi(e) : 0;
    return (v >= 16 && v <= 1024) ? v : 0;
  }();
  const int px = forced ? forced : (HW <= 1024 ? GN_CHUNK_PX_SMALL : GN_CHUNK_PX);
  int n_chunks = (HW + px - 1) / px;
  if (n_chunks > GN_MAX_CHUNKS) n_chunks = GN_MAX_CHUNKS;
  if (n_chunks < 1) n_chunks = 1;
  *chunk_px = (HW + n_chunks - 1) / n_chunks;
  return (HW + *chunk_px - 1) / *chunk_px;
}

int grid_for(long long total) {
  long long grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  return (int)(grid < 1 ? 1 : grid);
}

}  // namespace

// workspace = [GnCtrl | done[N] | left[N]] rounded up to 256 bytes, then the partial sums
int64_t gn_ctrl_bytes(int N) { return ((int64_t)sizeof(GnCtrl) + 8LL * N + 255) / 256 * 256; }

// Chunking of the single-read kernel: the chunk (chunk_px pixels x C channels, bf16) must fit in shared memory next to the
// reduction scratch; ~100 KB tiles keep two CTAs per SM where the image allows it.  Returns 0 when no chunking fits (then
// the two-kernel form runs).
int gn_fused_chunks(int HW, int C, int* chunk_px, size_t* smem) {
  const int CV = C / 8, cols = CV < GN_THREADS ? CV : GN_THREADS, R = GN_THREADS / cols;
  const size_t scratch = (size_t)R * 2 * C * sizeof(float) + 2 * 64 * sizeof(float) + 64;   // red + mean/rstd (G <= 64 here)
  static const int tile_kb = [] { const char* e = getenv("FF_GN_TILE_KB"); const int v = e ? atoi(e) : 0; return v >= 8 && v <= 180 ? v : 48; }();
  int px = (int)(((size_t)tile_kb * 1024) / ((size_t)C * 2));
  const int cap = HW <= 1024 ? GN_CHUNK_PX_SMALL : GN_CHUNK_PX;
  if (px > cap) px = cap;
  if (px < 1) px = 1;
  int n_chunks = (HW + px - 1) / px;
  if (n_chunks > GN_MAX_CHUNKS_FUSED) n_chunks = GN_MAX_CHUNKS_FUSED;
  px = (HW + n_chunks - 1) / n_chunks;
  n_chunks = (HW + px - 1) / px;
  const size_t need = (size_t)px * C * 2 + scratch;
  if (need > 200 * 1024) return 0;
  *chunk_px = px;
  *smem = need;
  return n_chunks;
}

extern "C" int64_t ff_group_norm_ws_bytes(int32_t N, int32_t G) {
  if (N <= 0 || G <= 0) return 0;
  return gn_ctrl_bytes(N) + (int64_t)N * GN_MAX_CHUNKS_FUSED * G * (int64_t)sizeof(float2);
}

extern "C" int ff_group_norm_nhwc(const void* x, const float* add_nc, int64_t add_ld, const void* gamma, const void* beta,
                                  void* y, void* workspace, int32_t N, int32_t HW, int32_t C, int32_t G, float eps,
                                  int32_t silu, void* stream) {
  FF_REQUIRE(x && gamma && beta && y && workspace, "ff_group_norm_nhwc: null pointer");
  FF_REQUIRE(N > 0 && HW > 0 && C > 0 && G > 0, "ff_group_norm_nhwc: bad shape");
  FF_REQUIRE(!add_nc || add_ld >= C, "ff_group_norm_nhwc: add_ld=%lld must be >= C=%d", (long long)add_ld, C);
  FF_REQUIRE(N <= 65535, "ff_group_norm_nhwc: N must be <= 65535");
  FF_REQUIRE(C % 8 == 0 && C % G == 0, "ff_group_norm_nhwc: C must be a multiple of 8 and of G (C=%d G=%d)", C, G);
  FF_REQUIRE(ff::aligned16(x) && ff::aligned16(y) && ff::aligned16(workspace),
             "ff_group_norm_nhwc: x / y / workspace must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int* ctrl = static_cast<unsigned int*>(workspace);
  float2* partial = reinterpret_cast<float2*>(static_cast<char*>(workspace) + gn_ctrl_bytes(N));
  // The single-read kernel is an EXPERIMENT switch (FF_GN_SINGLE_READ=1): measured on B200 in round 2 it only wins on
  // small images (32x1280x16x16: 23.6 vs 39.0 us); at the sizes that matter it is slower than the two-kernel form (83 vs 77
  // us at 32x320x64x64, 333 vs 206 us at 32x960x64x64) -- a CTA parked at the per-image barrier holds its shared memory, and
  // load, reduce and store phases of the CTAs on an SM do not overlap enough (profiles/r2_gn_single_read.txt).
  static const bool two_kernel = [] { const char* e = getenv("FF_GN_SINGLE_READ"); return !(e && atoi(e) != 0); }();
  {
    int B = 0, S = 1, P = HW;
    const int vpt = gn_small_plan(HW, C, G, &B, &S, &P);
    if (vpt > 0) {
      const GnSmallArgs a{static_cast<const uint4*>(x), add_nc, (long long)add_ld, static_cast<const __nv_bfloat16*>(gamma),
                          static_cast<const __nv_bfloat16*>(beta), static_cast<uint4*>(y), HW, C, G, B, P, eps};
      const dim3 grid(G / B, N, S);
      cudaError_t e;
      if (S > 1) {
        if (vpt <= 12) e = gn_small_launch<12, true>(a, grid, silu != 0, st);
        else if (vpt == 24) e = gn_small_launch<24, true>(a, grid, silu != 0, st);
        else e = gn_small_launch<32, true>(a, grid, silu != 0, st);
      } else {
        if (vpt == 2) e = gn_small_launch<2, false>(a, grid, silu != 0, st);
        else if (vpt == 4) e = gn_small_launch<4, false>(a, grid, silu != 0, st);
        else if (vpt == 6) e = gn_small_launch<6, false>(a, grid, silu != 0, st);
        else if (vpt == 12) e = gn_small_launch<12, false>(a, grid, silu != 0, st);
        else e = gn_small_launch<24, false>(a, grid, silu != 0, st);
      }
      if (e != cudaSuccess) return ff::fail(FF_E_CUDA, "ff_group_norm_nhwc (register-resident): %s", cudaGetErrorString(e));
      return ff::check_launch("ff_group_norm_nhwc (register-resident)");
    }
  }
  {
    int px = 0;
    size_t smem = 0;
    const int nc = (two_kernel || G > 64) ? 0 : gn_fused_chunks(HW, C, &px, &smem);
    if (nc > 0) {
      // single-read form: x is read from HBM once (see gn_fused_nhwc_kernel)
      static std::atomic<uint64_t> configured{0};
      int dev = 0;
      cudaGetDevice(&dev);
      const uint64_t bit = 1ull << (dev & 63);
      if (!(configured.load(std::memory_order_acquire) & bit)) {
        cudaError_t e1 = cudaFuncSetAttribute(gn_fused_nhwc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
        cudaError_t e2 = cudaFuncSetAttribute(gn_fused_nhwc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024);
        if (e1 != cudaSuccess || e2 != cudaSuccess)
          return ff::fail(FF_E_CUDA, "ff_group_norm_nhwc: cudaFuncSetAttribute: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        configured.fetch_or(bit, std::memory_order_release);
      }
      const unsigned int grid1 = (unsigned)N * (unsigned)nc;
      if (silu)
        gn_fused_nhwc_kernel<true><<<grid1, GN_THREADS, smem, st>>>(
            static_cast<const uint4*>(x), add_nc, add_ld, static_cast<const __nv_bfloat16*>(gamma),
            static_cast<const __nv_bfloat16*>(beta), static_cast<uint4*>(y), ctrl, partial, N, HW, C, G, px, nc, eps);
      else
        gn_fused_nhwc_kernel<false><<<grid1, GN_THREADS, smem, st>>>(
            static_cast<const uint4*>(x), add_nc, add_ld, static_cast<const __nv_bfloat16*>(gamma),
            static_cast<const __nv_bfloat16*>(beta), static_cast<uint4*>(y), ctrl, partial, N, HW, C, G, px, nc, eps);
      return ff::check_launch("ff_group_norm_nhwc (single-read)");
    }
  }
  int chunk_px = 0;
  const int n_chunks = gn_chunks(HW, N, &chunk_px);
  const int CV = C / 8, cols = CV < GN_THREADS ? CV : GN_THREADS, R = GN_THREADS / cols;
  const size_t smem_stats = (size_t)R * 2 * C * sizeof(float);
  FF_REQUIRE(smem_stats <= 48 * 1024, "ff_group_norm_nhwc: C=%d too large", C);
  const dim3 grid(n_chunks, N);
  gn_stats_nhwc_kernel<<<grid, GN_THREADS, smem_stats, st>>>(static_cast<const uint4*>(x), add_nc, add_ld, partial, HW, C, G, chunk_px,
                                                             n_chunks);
  int rc = ff::check_launch("ff_group_norm_nhwc (statistics)");
  if (rc != FF_OK) return rc;
  const size_t smem_apply = (size_t)2 * G * sizeof(float);
  if (silu)
    gn_apply_nhwc_kernel<true><<<grid, GN_THREADS, smem_apply, st>>>(
        static_cast<const uint4*>(x), add_nc, add_ld, static_cast<const __nv_bfloat16*>(gamma),
        static_cast<const __nv_bfloat16*>(beta), partial, static_cast<uint4*>(y), HW, C, G,
        chunk_px, n_chunks, eps);
  else
    gn_apply_nhwc_kernel<false><<<grid, GN_THREADS, smem_apply, st>>>(
        static_cast<const uint4*>(x), add_nc, add_ld, static_cast<const __nv_bfloat16*>(gamma),
        static_cast<const __nv_bfloat16*>(beta), partial, static_cast<uint4*>(y), HW, C, G,
        chunk_px, n_chunks, eps);
  return ff::check_launch("ff_group_norm_nhwc (apply)");
}

extern "C" int ff_bias_residual_nhwc(const void* h, const void* bias, const void* res, void* out, int64_t M, int32_t C,
                                     void* stream) {
  FF_REQUIRE(h && out, "ff_bias_residual_nhwc: null pointer");
  FF_REQUIRE(M > 0 && C > 0 && C % 8 == 0, "ff_bias_residual_nhwc: C must be a positive multiple of 8");
  FF_REQUIRE(ff::aligned16(h) && ff::aligned16(out) && ff::aligned16(bias) && ff::aligned16(res),
             "ff_bias_residual_nhwc: pointers must be 16-byte aligned");
  const long long total = (long long)M * (C / 8);
  bias_residual_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(h), static_cast<const __nv_bfloat16*>(bias), static_cast<const uint4*>(res),
      static_cast<uint4*>(out), total, C / 8);
  return ff::check_launch("ff_bias_residual_nhwc");
}

extern "C" int ff_geglu(const void* h, void* out, int64_t M, int32_t F, void* stream) {
  FF_REQUIRE(h && out, "ff_geglu: null pointer");
  FF_REQUIRE(M > 0 && F > 0 && F % 8 == 0, "ff_geglu: F must be a positive multiple of 8");
  FF_REQUIRE(ff::aligned16(h) && ff::aligned16(out), "ff_geglu: pointers must be 16-byte aligned");
  const long long total = (long long)M * (F / 8);
  // (a persistent grid of 148 x {4, 6, 8, 12} CTAs measured the same or slower than the 148 x 16 cap: 192 / 230 / 192 / 191
  // vs 190 us at 131072 x 1280 -- the launch sits at 5.3 TB/s of its two-reads-one-write traffic either way)
  const int FV = F / 8;
  static const bool rows_off = [] { const char* e = getenv("FF_GEGLU_ROWS"); return e && atoi(e) == 0; }();
  if (FV >= 32 && !rows_off) {            // one warp per row (FF_GEGLU_ROWS=0: the flat-index kernel, A/B switch)
    long long blocks = (M + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    geglu_rows_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(h), static_cast<uint4*>(out), M, FV);
    return ff::check_launch("ff_geglu");
  }
  const int grid = grid_for(total);
  const long long step2 = 2LL * grid * 256;
  geglu_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(h), static_cast<uint4*>(out), total, FV, step2 / FV, (int)(step2 % FV));
  return ff::check_launch("ff_geglu");
}

extern "C" int ff_layer_norm(const void* x, const void* gamma, const void* beta, void* y, int64_t M, int32_t C,
                             float eps, void* stream) {
  FF_REQUIRE(x && gamma && beta && y, "ff_layer_norm: null pointer");
  FF_REQUIRE(M > 0 && C > 0 && C % 8 == 0, "ff_layer_norm: C must be a positive multiple of 8");
  FF_REQUIRE(C <= 8 * 32 * 8, "ff_layer_norm: C must be <= 2048");
  FF_REQUIRE(ff::aligned16(x) && ff::aligned16(y) && ff::aligned16(gamma) && ff::aligned16(beta),
             "ff_layer_norm: pointers must be 16-byte aligned");
  const int CV = C / 8, vpl = (CV + 31) / 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint4* xp = static_cast<const uint4*>(x);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(gamma);
  const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(beta);
  uint4* yp = static_cast<uint4*>(y);
  const int rows_per_warp = vpl <= 1 ? 8 : (vpl <= 2 ? 4 : (vpl <= 3 ? 2 : 1));
  const long long blocks = (M + 8LL * rows_per_warp - 1) / (8LL * rows_per_warp);
  FF_REQUIRE(blocks <= 2147483647LL, "ff_layer_norm: too many rows");
  if (CV == 40 || CV == 80 || CV == 160) {                    // C = 320 / 640 / 1280: sub-warp rows, five vectors per lane
    const int lpr = CV / 5, rpw = 32 / lpr;
    long long nb = (M + 8LL * rpw * 4 - 1) / (8LL * rpw * 4);   // about four row groups per warp ...
    // ... but at most what is resident at once (3 CTAs per SM): a persistent grid, no partial last wave.  FF_LN5_CTAS: knob.
    static const int cap = [] { const char* e = getenv("FF_LN5_CTAS"); const int v = e ? atoi(e) : 0; return v >= 1 ? v : 148 * FF_LN5_MINB; }();
    if (nb > cap) nb = cap;
    if (lpr == 8) layer_norm5_kernel<8><<<(int)nb, 256, 0, st>>>(xp, gp, bp, yp, M, eps);
    else if (lpr == 16) layer_norm5_kernel<16><<<(int)nb, 256, 0, st>>>(xp, gp, bp, yp, M, eps);
    else layer_norm5_kernel<32><<<(int)nb, 256, 0, st>>>(xp, gp, bp, yp, M, eps);
    return ff::check_launch("ff_layer_norm");
  }
  if (vpl <= 1) layer_norm_kernel<1, 8><<<(int)blocks, 256, 0, st>>>(xp, gp, bp, yp, M, CV, eps);
  else if (vpl <= 2) layer_norm_kernel<2, 4><<<(int)blocks, 256, 0, st>>>(xp, gp, bp, yp, M, CV, eps);
  else if (vpl <= 3) layer_norm_kernel<3, 2><<<(int)blocks, 256, 0, st>>>(xp, gp, bp, yp, M, CV, eps);
  else if (vpl <= 5) layer_norm_kernel<5, 1><<<(int)blocks, 256, 0, st>>>(xp, gp, bp, yp, M, CV, eps);
  else layer_norm_kernel<8, 1><<<(int)blocks, 256, 0, st>>>(xp, gp, bp, yp, M, CV, eps);
  return ff::check_launch("ff_layer_norm");
}

extern "C" int ff_upsample2x_nhwc(const void* x, void* y, int32_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  FF_REQUIRE(x && y, "ff_upsample2x_nhwc: null pointer");
  FF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "ff_upsample2x_nhwc: bad shape (C must be a multiple of 8)");
  FF_REQUIRE(ff::aligned16(x) && ff::aligned16(y), "ff_upsample2x_nhwc: pointers must be 16-byte aligned");
  const long long total = (long long)N * H * W * (C / 8);
  FF_REQUIRE(4 * total < 2147483647LL, "ff_upsample2x_nhwc: tensor too large for 32-bit vector indices");
  upsample2x_nhwc_kernel<<<grid_for((total + 1) / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), (unsigned)total, (unsigned)W, (unsigned)(C / 8));
  return ff::check_launch("ff_upsample2x_nhwc");
}

extern "C" int ff_concat_nhwc(const void* a, const void* b, void* out, int64_t M, int32_t Ca, int32_t Cb, void* stream) {
  FF_REQUIRE(a && b && out, "ff_concat_nhwc: null pointer");
  FF_REQUIRE(M > 0 && Ca > 0 && Cb > 0 && Ca % 8 == 0 && Cb % 8 == 0, "ff_concat_nhwc: bad shape (channel counts must be multiples of 8)");
  FF_REQUIRE(ff::aligned16(a) && ff::aligned16(b) && ff::aligned16(out), "ff_concat_nhwc: pointers must be 16-byte aligned");
  const long long total = (long long)M * ((Ca + Cb) / 8);
  FF_REQUIRE(total < 2147483647LL, "ff_concat_nhwc: tensor too large for 32-bit vector indices");
  concat_nhwc_kernel<<<grid_for((total + 1) / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), (unsigned)total, (unsigned)(Ca / 8),
      (unsigned)(Cb / 8));
  return ff::check_launch("ff_concat_nhwc");
}
