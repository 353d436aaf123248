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

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_CHUNKS = 64;     // pixel chunks per image (partial statistics per chunk), two-kernel form
constexpr int GN_MAX_CHUNKS_FUSED = 128;   // ... single-read form (the workspace is sized for this one)
#ifndef FF_GN_MLP
#define FF_GN_MLP 4
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

// partial[n][chunk][g] = (sum, sum of squares) of v = x + add over the pixels of the chunk and the channels of group g
__global__ void __launch_bounds__(GN_THREADS)
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
      float a[8], s[8], ss[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] = add_nc ? __ldg(add_nc + (size_t)n * add_ld + 8 * v + j) : 0.f;
        s[j] = ss[j] = 0.f;
      }
      const uint4* px = x + ((size_t)n * HW + p0 + L.r) * L.CV + v;
      const size_t step = (size_t)L.R * L.CV;
      // batches of GN_MLP independent loads, THEN the arithmetic: with a plain `#pragma unroll 4` ptxas kept two loads in
      // flight per thread (each load sat next to its 32 dependent operations) and the pass ran at 2.5 TB/s
#pragma unroll 1
      for (int p = p0 + L.r; p < p1; p += GN_MLP * L.R, px += GN_MLP * step) {
        uint4 raw[GN_MLP];
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u)
          raw[u] = (p + u * L.R < p1) ? __ldg(px + u * step) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u) {
          if (p + u * L.R < p1) {
            float f[8];
            unpack8(raw[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t = f[j] + a[j];
              s[j] += t;
              ss[j] = fmaf(t, t, ss[j]);
            }
          }
        }
      }
      float* dst = sm + (size_t)L.r * 2 * C + 8 * v;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dst[j] = s[j];
        dst[C + j] = ss[j];
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

template <bool SILU>
__global__ void __launch_bounds__(GN_THREADS)
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
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * v + j, g = c / cpg;
      const float a = add_nc ? __ldg(add_nc + (size_t)n * add_ld + c) : 0.f;
      sc[j] = rstd[g] * __bfloat162float(gamma[c]);
      sh[j] = fmaf(a - mean[g], sc[j], __bfloat162float(beta[c]));     // y = (x + a - mean) * rstd * gamma + beta
    }
    const size_t off = ((size_t)n * HW + p0 + L.r) * L.CV + v;
    const uint4* px = x + off;
    uint4* py = y + off;
    const size_t step = (size_t)L.R * L.CV;
#pragma unroll 1
    for (int p = p0 + L.r; p < p1; p += GN_MLP * L.R, px += GN_MLP * step, py += GN_MLP * step) {
      uint4 raw[GN_MLP];                                   // (independent loads first: see the statistics kernel)
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u)
        raw[u] = (p + u * L.R < p1) ? __ldg(px + u * step) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) {
        if (p + u * L.R < p1) {
          float f[8];
          unpack8(raw[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float t = fmaf(f[j], sc[j], sh[j]);
            if (SILU) t = __fdividef(t, 1.f + __expf(-t));     // 2 MUFU ops; the IEEE division made this kernel ALU-bound
            f[j] = t;
          }
          py[u * step] = pack8(f);
        }
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
template <bool SILU, int VPT>
__global__ void __launch_bounds__(GN_THREADS, VPT >= 24 ? 2 : 1)
gn_small_nhwc_kernel(const uint4* __restrict__ x, const float* __restrict__ add_nc, long long add_ld,
                     const __nv_bfloat16* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y, int HW, int C, int G, int B, float eps) {
  __shared__ float sm[GN_THREADS / 32 * 8 * 2 + 16];   // [warp][group][S, SS] partials, then [B] mean, [B] rstd
  const int cpg = C / G, bch = B * cpg, BV = bch >> 3, CV = C >> 3;
  const int R = GN_THREADS / BV, col = threadIdx.x % BV, r = threadIdx.x / BV;
  const bool live = r < R;
  const int n = blockIdx.y, bundle = blockIdx.x;
  const int c0 = bundle * bch + 8 * col;        // my eight channels
  float* stat = sm + GN_THREADS / 32 * 8 * 2;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = add_nc ? __ldg(add_nc + (size_t)n * add_ld + c0 + j) : 0.f;
  uint4 raw[VPT];
  const uint4* px = x + ((size_t)n * HW) * CV + bundle * BV + col;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    const int p = r + k * R;
    raw[k] = (live && p < HW) ? __ldg(px + (size_t)p * CV) : make_uint4(0u, 0u, 0u, 0u);
  }
  float s[8], ss[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
#pragma unroll
  for (int k = 0; k < VPT; ++k) {
    if (live && r + k * R < HW) {
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
    const float inv_cnt = 1.f / ((float)cpg * (float)HW);
    const float m = S * inv_cnt;
    const float var = fmaxf(SS * inv_cnt - m * m, 0.f);
    stat[threadIdx.x] = m;
    stat[B + threadIdx.x] = 1.f / sqrtf(var + eps);
  }
  __syncthreads();
  if (!live) return;
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
    const int p = r + k * R;
    if (p < HW) {
      float f[8];
      unpack8(raw[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(f[j], sc[j], sh[j]);
        if (SILU) t = __fdividef(t, 1.f + __expf(-t));
        f[j] = t;
      }
      py[(size_t)p * CV] = pack8(f);
    }
  }
}

// groups per bundle B (the smallest B dividing G whose B * cpg channels are whole 16-byte vectors), vectors per thread;
// 0 = the shape does not fit the register-resident form
int gn_small_plan(int HW, int C, int G, int* B_out, size_t* smem) {
  static const bool off = [] { const char* e = getenv("FF_GN_SMALL"); return e && atoi(e) == 0; }();
  if (off) return 0;
  const int cpg = C / G;
  int B = 0;
  for (int b = 1; b <= 8; b <<= 1)
    if (G % b == 0 && (b * cpg) % 8 == 0) { B = b; break; }
  if (!B) return 0;
  const int BV = B * cpg / 8;
  if (BV > GN_THREADS) return 0;
  const int R = GN_THREADS / BV, need = (HW + R - 1) / R;
  if (need > 24) return 0;
  *B_out = B;
  *smem = 0;                                                 // (static shared memory only)
  return need <= 2 ? 2 : (need <= 4 ? 4 : (need <= 6 ? 6 : (need <= 12 ? 12 : 24)));
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
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(f[j], sc[j], sh[j]);
        if (SILU) t = __fdividef(t, 1.f + __expf(-t));
        f[j] = t;
      }
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
__device__ __forceinline__ float gelu_erf(float g) {
  const float z = fabsf(g) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float half_erfc = 0.5f * p * t * e;                                      // 0.5 * erfc(|g|/sqrt 2)
  return g * (g >= 0.f ? 1.f - half_erfc : half_erfc);
}

__global__ void __launch_bounds__(256)
geglu_kernel(const uint4* __restrict__ h, uint4* __restrict__ out, long long total, int FV) {
  // two output vectors per trip: four independent 16-byte loads in flight per thread before the erf arithmetic
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const long long i2 = i + stride;
    const bool two = i2 < total;
    const long long m = i / FV, m2 = two ? i2 / FV : m;
    const int v = (int)(i - m * FV), v2 = two ? (int)(i2 - m2 * FV) : v;
    const uint4* row = h + m * 2 * FV;
    const uint4* row2 = h + m2 * 2 * FV;
    const uint4 qx = __ldg(row + v), qg = __ldg(row + FV + v);
    const uint4 qx2 = __ldg(row2 + v2), qg2 = __ldg(row2 + FV + v2);
    float x[8], g[8];
    unpack8(qx, x);
    unpack8(qg, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= bf16_round(gelu_erf(g[j]));
    out[i] = pack8(x);
    if (two) {
      unpack8(qx2, x);
      unpack8(qg2, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] *= bf16_round(gelu_erf(g[j]));
      out[i2] = pack8(x);
    }
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
// gamma / beta stay packed in registers across the rows a warp walks.
template <int LPR>
__global__ void __launch_bounds__(256, 2)
layer_norm5_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                   const __nv_bfloat16* __restrict__ beta, uint4* __restrict__ y, long long M, float eps) {
  constexpr int RPW = 32 / LPR, CV = 5 * LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  uint4 gq[5], bq[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    gq[k] = __ldg(reinterpret_cast<const uint4*>(gamma) + sub + LPR * k);
    bq[k] = __ldg(reinterpret_cast<const uint4*>(beta) + sub + LPR * k);
  }
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)(8 * CV);
  for (long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < M; row0 += nwarps * RPW) {
    const long long row = row0 + rsel;
    const bool valid = row < M;
    uint4 raw[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) raw[k] = valid ? __ldg(x + row * CV + sub + LPR * k) : make_uint4(0u, 0u, 0u, 0u);
    float f[5][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unpack8(raw[k], f[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[k][j];
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[k][j] - mean;
        ss = fmaf(d, d, ss);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = 1.f / sqrtf(ss * inv_c + eps);
    if (valid) {
      uint4* py = y + row * CV + sub;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        float ga[8], be[8], o[8];
        unpack8(gq[k], ga);
        unpack8(bq[k], be);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((f[k][j] - mean) * rstd, ga[j], be[j]);
        py[LPR * k] = pack8(o);
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
    int B = 0;
    size_t smem = 0;
    const int vpt = gn_small_plan(HW, C, G, &B, &smem);
    if (vpt > 0) {
      const dim3 grid(G / B, N);
      const uint4* xp = static_cast<const uint4*>(x);
      const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(gamma);
      const __nv_bfloat16* bp = static_cast<const __nv_bfloat16*>(beta);
      uint4* yp = static_cast<uint4*>(y);
#define FF_GN_SMALL_LAUNCH(V)                                                                                            \
  do {                                                                                                                   \
    if (silu) gn_small_nhwc_kernel<true, V><<<grid, GN_THREADS, smem, st>>>(xp, add_nc, add_ld, gp, bp, yp, HW, C, G, B, eps);    \
    else gn_small_nhwc_kernel<false, V><<<grid, GN_THREADS, smem, st>>>(xp, add_nc, add_ld, gp, bp, yp, HW, C, G, B, eps);        \
  } while (0)
      if (vpt == 2) FF_GN_SMALL_LAUNCH(2);
      else if (vpt == 4) FF_GN_SMALL_LAUNCH(4);
      else if (vpt == 6) FF_GN_SMALL_LAUNCH(6);
      else if (vpt == 12) FF_GN_SMALL_LAUNCH(12);
      else FF_GN_SMALL_LAUNCH(24);
#undef FF_GN_SMALL_LAUNCH
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
  geglu_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(h), static_cast<uint4*>(out), total, F / 8);
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
    long long nb = (M + 8LL * rpw * 4 - 1) / (8LL * rpw * 4);   // about four row groups per warp
    if (nb > 148 * 16) nb = 148 * 16;
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
