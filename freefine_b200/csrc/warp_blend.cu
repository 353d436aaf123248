// (b) fused affine warp + resample + mask-guided blend.
//
// Reference arithmetic: wrapAffine_tensor (src/utils/geo_utils.py:304-341) = F.affine_grid + F.grid_sample
// (padding_mode='zeros', align_corners=False) for the image (bilinear) and for the mask (nearest), then the
// np.where(mask, warped, background) blend of re_edit_2d (src/utils/vis_utils.py:252-256,272), in ONE pass:
// the warped image, the warped mask and the blended result never make a round trip through HBM.
//
// Mapping: one CTA produces a 64x16 output tile for a chunk of channels.  The affine image of that tile is a
// parallelogram; its bounding box in the source is staged in shared memory with 16-byte cp.async copies through a
// ring of 2-4 channel-pair buffer sets (as many as fit 88 KB: up to 3 later pairs are in flight while one is resampled), so each
// source texel is read from HBM/L2 once per tile instead of up to 4 times and the 4 bilinear taps are LDS.  The
// background tile is read and the result written as 128-bit vectors.  If the bounding box does not fit (strong
// minification) or the layout is not 16-byte friendly, the taps go straight to global memory -- same arithmetic,
// same result.  Coordinates follow ATen's operation order in fp32 with explicit round-to-nearest ops (no FMA
// contraction):  x_n = (2j+1)/dW - 1;  g = x_n*t00 + y_n*t01 + t02;  ix = ((g+1)*W - 1)/2;  nearest = rint.
#include <cuda_bf16.h>

#include "ff_common.cuh"

namespace {

constexpr int TILE_X = 64, TILE_Y = 16, PX = 4;           // 4 consecutive output pixels per thread
constexpr int THREADS = (TILE_X / PX) * TILE_Y;           // 256
constexpr int CH_PER_ITER = 2;                            // channels per pipeline stage
constexpr int SMEM_BYTES = 88 * 1024;                     // staging ring: 88 KB -> 2 CTAs / SM
constexpr int MAX_SETS = 4;                               // ring depth: as many channel-pair sets as fit (2..4)

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<__nv_bfloat16> { using type = uint2; };

struct Coord { float ix, iy; };

__device__ __forceinline__ Coord src_coord(int x, int y, int dW, int dH, int W, int H, const float* t) {
  const float xn = __fsub_rn(__fdiv_rn((float)(2 * x + 1), (float)dW), 1.f);
  const float yn = __fsub_rn(__fdiv_rn((float)(2 * y + 1), (float)dH), 1.f);
  const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xn, t[0]), __fmul_rn(yn, t[1])), t[2]);
  const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xn, t[3]), __fmul_rn(yn, t[4])), t[5]);
  Coord c;
  c.ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)W), 1.f), 2.f);
  c.iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)H), 1.f), 2.f);
  return c;
}

// Taps of one output pixel, computed ONCE per pixel and reused by every channel: offset of the north-west tap inside
// the fetch domain (the staged box or the whole image, row pitch `pitch`), which of the 4 taps fall inside the domain
// (outside = the zero padding of grid_sample), and the 4 bilinear weights in ATen's operation order
// (nw = (x1-ix)*(y1-iy), ne = (ix-x0)*(y1-iy), sw = (x1-ix)*(iy-y0), se = (ix-x0)*(iy-y0)).  mode 1 (nearest): one tap.
struct Taps {
  int off;
  uint32_t valid;   // bit0 nw, bit1 ne, bit2 sw, bit3 se
  float w[4];
};

__device__ __forceinline__ Taps make_taps(Coord c, int mode, int dx0, int dy0, int dw, int dh, int pitch) {
  Taps t;
  if (mode == 1) {
    const int u = __float2int_rn(c.ix) - dx0, v = __float2int_rn(c.iy) - dy0;
    t.off = v * pitch + u;
    t.valid = (u >= 0 && u < dw && v >= 0 && v < dh) ? 1u : 0u;
    t.w[0] = 1.f; t.w[1] = t.w[2] = t.w[3] = 0.f;
    return t;
  }
  const float x0 = floorf(c.ix), y0 = floorf(c.iy);
  const float x1 = __fadd_rn(x0, 1.f), y1 = __fadd_rn(y0, 1.f);
  const float wx1 = __fsub_rn(x1, c.ix), wx0 = __fsub_rn(c.ix, x0);
  const float wy1 = __fsub_rn(y1, c.iy), wy0 = __fsub_rn(c.iy, y0);
  t.w[0] = __fmul_rn(wx1, wy1);
  t.w[1] = __fmul_rn(wx0, wy1);
  t.w[2] = __fmul_rn(wx1, wy0);
  t.w[3] = __fmul_rn(wx0, wy0);
  const int u = (int)x0 - dx0, v = (int)y0 - dy0;
  t.off = v * pitch + u;
  const bool ux0 = u >= 0 && u < dw, ux1 = u + 1 >= 0 && u + 1 < dw;
  const bool vy0 = v >= 0 && v < dh, vy1 = v + 1 >= 0 && v + 1 < dh;
  t.valid = (ux0 && vy0 ? 1u : 0u) | (ux1 && vy0 ? 2u : 0u) | (ux0 && vy1 ? 4u : 0u) | (ux1 && vy1 ? 8u : 0u);
  return t;
}

// out = nw*w_nw + ne*w_ne + sw*w_sw + se*w_se, products rounded separately, summed left to right (ATen's order)
template <typename T, typename P>
__device__ __forceinline__ float apply_taps(P base, const Taps& t, int pitch, int mode) {
  const float nw = (t.valid & 1u) ? to_f<T>(base[t.off]) : 0.f;
  if (mode == 1) return nw;
  const float ne = (t.valid & 2u) ? to_f<T>(base[t.off + 1]) : 0.f;
  const float sw = (t.valid & 4u) ? to_f<T>(base[t.off + pitch]) : 0.f;
  const float se = (t.valid & 8u) ? to_f<T>(base[t.off + pitch + 1]) : 0.f;
  float o = __fmul_rn(nw, t.w[0]);
  o = __fadd_rn(o, __fmul_rn(ne, t.w[1]));
  o = __fadd_rn(o, __fmul_rn(sw, t.w[2]));
  o = __fadd_rn(o, __fmul_rn(se, t.w[3]));
  return o;
}

// Staged form of the taps: one absolute offset per tap inside the channel's staging buffer; taps that fall outside the
// box point at a zeroed slot behind the box (`zero_off`), so the per-channel work is 4 LDS + 4 FMUL + 3 FADD with no
// predicates.  Same products, same summation order as apply_taps.
struct STaps {
  int o[4];
  float w[4];
};
__device__ __forceinline__ STaps to_staged(const Taps& t, int pitch, int zero_off, int mode) {
  STaps r;
  r.o[0] = (t.valid & 1u) ? t.off : zero_off;
  r.o[1] = (mode == 0 && (t.valid & 2u)) ? t.off + 1 : zero_off;
  r.o[2] = (mode == 0 && (t.valid & 4u)) ? t.off + pitch : zero_off;
  r.o[3] = (mode == 0 && (t.valid & 8u)) ? t.off + pitch + 1 : zero_off;
#pragma unroll
  for (int i = 0; i < 4; ++i) r.w[i] = t.w[i];
  return r;
}
template <typename T>
__device__ __forceinline__ float apply_staged(const T* s, const STaps& t, int mode) {
  const float nw = to_f<T>(s[t.o[0]]);
  if (mode == 1) return nw;
  float o = __fmul_rn(nw, t.w[0]);
  o = __fadd_rn(o, __fmul_rn(to_f<T>(s[t.o[1]]), t.w[1]));
  o = __fadd_rn(o, __fmul_rn(to_f<T>(s[t.o[2]]), t.w[2]));
  o = __fadd_rn(o, __fmul_rn(to_f<T>(s[t.o[3]]), t.w[3]));
  return o;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(THREADS, 2)
warp_affine_blend_kernel(const T* __restrict__ src, const float* __restrict__ theta,
                         const uint8_t* __restrict__ mask_src, const T* __restrict__ bg, T* __restrict__ out,
                         uint8_t* __restrict__ mask_out, int C, int H, int W, int dH, int dW, int mode,
                         int ch_per_cta, int tiles_x, int vec_ok) {
  extern __shared__ __align__(16) uint8_t stage_raw[];
  __shared__ float th[6];
  constexpr int VEC = 16 / (int)sizeof(T);                 // elements per 16-byte copy
  const int n = blockIdx.z;
  const int tile = blockIdx.x;
  const int tx0 = (tile % tiles_x) * TILE_X, ty0 = (tile / tiles_x) * TILE_Y;
  const int c_begin = blockIdx.y * ch_per_cta;
  const int c_end = min(C, c_begin + ch_per_cta);
  if (threadIdx.x < 6) th[threadIdx.x] = theta[n * 6 + threadIdx.x];
  __syncthreads();

  const int lx = (threadIdx.x % (TILE_X / PX)) * PX, ly = threadIdx.x / (TILE_X / PX);
  const int ox = tx0 + lx, oy = ty0 + ly;
  const bool active = oy < dH && ox < dW;

  // per-thread source coordinates of its 4 pixels (shared by every channel) and the warped mask
  Coord cd[PX];
  bool keep[PX];          // true -> take the warped source, false -> background
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    cd[j] = src_coord(ox + j, oy, dW, dH, W, H, th);
    keep[j] = true;
  }
  if (mask_src != nullptr) {
    const uint8_t* mp = mask_src + (size_t)n * H * W;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      const bool in = oy < dH && (ox + j) < dW;
      const int xi = __float2int_rn(cd[j].ix), yi = __float2int_rn(cd[j].iy);     // nearest: rint(coordinate)
      const uint8_t mv = (in && xi >= 0 && xi < W && yi >= 0 && yi < H) ? __ldg(mp + (size_t)yi * W + xi) : 0;
      keep[j] = mv != 0;
      if (in && mask_out != nullptr && blockIdx.y == 0) mask_out[((size_t)n * dH + oy) * dW + ox + j] = keep[j];
    }
  }

  // bounding box of the tile's pre-image (affine map => extreme values at the 4 corners), +1 texel for bilinear
  const int xe = min(tx0 + TILE_X, dW) - 1, ye = min(ty0 + TILE_Y, dH) - 1;
  const Coord k0 = src_coord(tx0, ty0, dW, dH, W, H, th), k1 = src_coord(xe, ty0, dW, dH, W, H, th);
  const Coord k2 = src_coord(tx0, ye, dW, dH, W, H, th), k3 = src_coord(xe, ye, dW, dH, W, H, th);
  int bx0 = (int)floorf(fminf(fminf(k0.ix, k1.ix), fminf(k2.ix, k3.ix))) - 1;
  int bx1 = (int)floorf(fmaxf(fmaxf(k0.ix, k1.ix), fmaxf(k2.ix, k3.ix))) + 2;
  int by0 = (int)floorf(fminf(fminf(k0.iy, k1.iy), fminf(k2.iy, k3.iy))) - 1;
  int by1 = (int)floorf(fmaxf(fmaxf(k0.iy, k1.iy), fmaxf(k2.iy, k3.iy))) + 2;
  bx0 = (max(bx0, 0) / VEC) * VEC;              // 16-byte aligned start for the vector copies
  by0 = max(by0, 0);
  bx1 = min(bx1, W - 1);
  by1 = min(by1, H - 1);
  int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
  bw = ((bw + VEC - 1) / VEC) * VEC;
  if (bx0 + bw > W) bw = W - bx0;               // ragged right edge (W % VEC != 0): scalar staging
  const bool empty_box = (bx1 < bx0) || (by1 < by0);
  // ring of n_sets buffer sets (one set = CH_PER_ITER boxes); the deeper the ring, the more copies are in flight
  const int box_elems = ((bw * bh + VEC - 1) / VEC) * VEC;
  const int stage_elems = box_elems + VEC;                  // + a zeroed slot: the target of out-of-box taps
  const long long set_bytes = (long long)CH_PER_ITER * stage_elems * (long long)sizeof(T);
  const int n_sets = empty_box ? 0 : (int)min((long long)MAX_SETS, (long long)SMEM_BYTES / max(set_bytes, 1LL));
  const bool staged = n_sets >= 2;
  const bool vec_stage = vec_ok && (bw % VEC == 0) && (W % VEC == 0);

  // per-pixel taps in the fetch domain of this CTA: the staged box (row pitch bw) or the whole image (row pitch W)
  Taps tp[PX];
  STaps sp4[PX];
#pragma unroll
  for (int j = 0; j < PX; ++j) {
    tp[j] = staged ? make_taps(cd[j], mode, bx0, by0, bw, bh, bw) : make_taps(cd[j], mode, 0, 0, W, H, W);
    sp4[j] = to_staged(tp[j], bw, box_elems, mode);
  }

  auto stage_ptr = [&](int set, int cc) { return reinterpret_cast<T*>(stage_raw) + (size_t)(set * CH_PER_ITER + cc) * stage_elems; };
  // Copy assignments of this thread are the same for every channel: the (row, 16-byte column) -> (source offset,
  // staging offset) index math is done ONCE here, the per-channel staging loop is then address adds + cp.async only
  // (the generic div/mod loop it replaces was 20% of all issued instructions).
  constexpr int KMAX = 4;                                   // covers boxes up to 4*256 vectors (a whole 64x64 fp32 channel)
  int cp_src[KMAX], cp_dst[KMAX];
  const int bwv = bw / VEC;
  const bool fast_stage = staged && vec_stage && bwv * bh <= KMAX * THREADS;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int i = threadIdx.x + k * THREADS;
    cp_src[k] = -1;
    cp_dst[k] = 0;
    if (fast_stage && i < bwv * bh) {
      const int v = i / bwv, u = (i - v * bwv) * VEC;
      cp_src[k] = (by0 + v) * W + bx0 + u;
      cp_dst[k] = v * bw + u;
    }
  }
  const size_t plane = (size_t)H * W;
  // asynchronous staging of channels [c0, c0+nc) into buffer set `set` (one cp.async group)
  auto issue = [&](int c0, int set) {
    const int nc = min(CH_PER_ITER, c_end - c0);
    for (int cc = 0; cc < nc; ++cc) {
      const T* p = src + ((size_t)n * C + c0 + cc) * plane;
      T* s = stage_ptr(set, cc);
      if (fast_stage) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (cp_src[k] >= 0) cp_async16(s + cp_dst[k], p + cp_src[k]);
      } else if (vec_stage) {
        for (int i = threadIdx.x; i < bwv * bh; i += THREADS) {
          const int v = i / bwv, u = (i - v * bwv) * VEC;
          cp_async16(s + v * bw + u, p + (size_t)(by0 + v) * W + bx0 + u);
        }
      } else {
        for (int i = threadIdx.x; i < bw * bh; i += THREADS) {
          const int v = i / bw, u = i - v * bw;
          s[i] = __ldg(p + (size_t)(by0 + v) * W + bx0 + u);
        }
      }
    }
    cp_async_commit();
  };

  if (staged) {                                             // zero slots (never touched by the copies)
    for (int i = threadIdx.x; i < n_sets * CH_PER_ITER * VEC; i += THREADS)
      stage_ptr(0, 0)[(size_t)(i / VEC) * stage_elems + box_elems + (i % VEC)] = from_f<T>(0.f);
  }
  using V = typename Vec4<T>::type;
  const size_t oplane = (size_t)dH * dW;
  const size_t opix = (size_t)oy * dW + ox;                 // this thread's pixel offset inside an output plane

  // ---- fast path (CTA-uniform): bilinear, masked blend, staged box with precomputed copy lists, whole 4-pixel
  // vectors, an even number of channels.  Same arithmetic as the general loop below with every run-time switch removed
  // and all addressing reduced to pointer increments (the general loop issues ~3.5x more instructions per pixel, and
  // this kernel is issue-bound at its 25% occupancy).
  if (fast_stage && vec_ok && mode == 0 && mask_src != nullptr && ((c_end - c_begin) % CH_PER_ITER) == 0) {
    const int npair = (c_end - c_begin) / CH_PER_ITER;
    const T* gsrc = src + ((size_t)n * C + c_begin) * plane;          // next channel pair to stage
    T* const sbase = reinterpret_cast<T*>(stage_raw);
    const size_t set_elems = (size_t)CH_PER_ITER * stage_elems;
    auto stage_pair = [&](int set) {
      T* s0 = sbase + set * set_elems;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
        if (cp_src[k] >= 0) {
          cp_async16(s0 + cp_dst[k], gsrc + cp_src[k]);
          cp_async16(s0 + stage_elems + cp_dst[k], gsrc + plane + cp_src[k]);
        }
      cp_async_commit();
      gsrc += CH_PER_ITER * plane;
    };
    int n_staged = 0;
    for (int k = 0; k < n_sets - 1; ++k) {
      if (n_staged < npair) { stage_pair(k); ++n_staged; }
      else cp_async_commit();
    }
    const T* bgp = bg + ((size_t)n * C + c_begin) * oplane + opix;
    T* outp = out + ((size_t)n * C + c_begin) * oplane + opix;
    int set = 0, fill = n_sets - 1;
    for (int i = 0; i < npair; ++i) {
      T b0[PX], b1[PX];
      if (active) {
        *reinterpret_cast<V*>(b0) = __ldg(reinterpret_cast<const V*>(bgp));
        *reinterpret_cast<V*>(b1) = __ldg(reinterpret_cast<const V*>(bgp + oplane));
      }
      if (n_staged < npair) { stage_pair(fill); ++n_staged; }
      else cp_async_commit();
      if (n_sets == 2) cp_async_wait<1>();
      else if (n_sets == 3) cp_async_wait<2>();
      else cp_async_wait<3>();
      __syncthreads();
      if (active) {
        const T* s0 = sbase + set * set_elems;
        const T* s1 = s0 + stage_elems;
        T o0[PX], o1[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
          const float r0 = apply_staged<T>(s0, sp4[j], 0), r1 = apply_staged<T>(s1, sp4[j], 0);
          o0[j] = keep[j] ? from_f<T>(r0) : b0[j];
          o1[j] = keep[j] ? from_f<T>(r1) : b1[j];
        }
        *reinterpret_cast<V*>(outp) = *reinterpret_cast<V*>(o0);
        *reinterpret_cast<V*>(outp + oplane) = *reinterpret_cast<V*>(o1);
      }
      __syncthreads();
      bgp += CH_PER_ITER * oplane;
      outp += CH_PER_ITER * oplane;
      set = set + 1 == n_sets ? 0 : set + 1;
      fill = fill + 1 == n_sets ? 0 : fill + 1;
    }
    return;
  }

  // prologue: n_sets-1 channel pairs in flight (empty groups keep the group count uniform at the tail)
  if (staged)
    for (int k = 0; k < n_sets - 1; ++k) {
      if (c_begin + k * CH_PER_ITER < c_end) issue(c_begin + k * CH_PER_ITER, k);
      else cp_async_commit();
    }
  int it = 0;
  const bool vec_io = vec_ok && ox + PX <= dW;
  for (int c0 = c_begin; c0 < c_end; c0 += CH_PER_ITER, ++it) {
    const int nc = min(CH_PER_ITER, c_end - c0);
    const size_t obase0 = ((size_t)n * C + c0) * oplane + opix;
    // background vectors first: these loads are in flight while we wait for the staged copies
    T b[CH_PER_ITER][PX];
    if (active && mask_src != nullptr) {
#pragma unroll
      for (int cc = 0; cc < CH_PER_ITER; ++cc) {
        if (cc < nc) {
          const size_t obase = obase0 + cc * oplane;
          if (vec_io) *reinterpret_cast<V*>(b[cc]) = __ldg(reinterpret_cast<const V*>(bg + obase));
          else
            for (int j = 0; j < PX && ox + j < dW; ++j) b[cc][j] = bg[obase + j];
        }
      }
    }
    const int cur = staged ? it % n_sets : 0;
    if (staged) {
      // refill the set consumed in the previous iteration with the pair n_sets-1 ahead, then wait for the current one
      const int ahead = c0 + (n_sets - 1) * CH_PER_ITER;
      if (ahead < c_end) issue(ahead, (it + n_sets - 1) % n_sets);
      else cp_async_commit();
      if (n_sets == 2) cp_async_wait<1>();
      else if (n_sets == 3) cp_async_wait<2>();
      else cp_async_wait<3>();
      __syncthreads();                              // copies of set `cur` are visible to every thread
    }
    if (active) {
#pragma unroll
      for (int cc = 0; cc < CH_PER_ITER; ++cc) {
        if (cc < nc) {
          const size_t obase = obase0 + cc * oplane;
          float r[PX];
          if (empty_box) {
#pragma unroll
            for (int j = 0; j < PX; ++j) r[j] = 0.f;
          } else if (staged) {
            const T* sp = stage_ptr(cur, cc);
#pragma unroll
            for (int j = 0; j < PX; ++j) r[j] = apply_staged<T>(sp, sp4[j], mode);
          } else {
            const T* gp = src + ((size_t)n * C + c0 + cc) * plane;
#pragma unroll
            for (int j = 0; j < PX; ++j) r[j] = apply_taps<T>(gp, tp[j], W, mode);
          }
          if (vec_io) {
            T o[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) o[j] = keep[j] ? from_f<T>(r[j]) : b[cc][j];
            *reinterpret_cast<V*>(out + obase) = *reinterpret_cast<V*>(o);
          } else {
            for (int j = 0; j < PX && ox + j < dW; ++j) out[obase + j] = keep[j] ? from_f<T>(r[j]) : b[cc][j];
          }
        }
      }
    }
    if (staged) __syncthreads();                    // set `cur` may be refilled by the next iteration's prefetch
  }
}

}  // namespace

extern "C" int ff_warp_affine_blend(const void* src, const float* theta, const uint8_t* mask_src, const void* bg,
                                    void* out, uint8_t* mask_out, int32_t N, int32_t C, int32_t H, int32_t W,
                                    int32_t dH, int32_t dW, int32_t mode, int32_t dtype, void* stream) {
  FF_REQUIRE(src && theta && out, "ff_warp_affine_blend: null pointer");
  FF_REQUIRE(mask_src == nullptr || bg != nullptr, "ff_warp_affine_blend: mask given without a background");
  FF_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && dH > 0 && dW > 0, "ff_warp_affine_blend: bad shape");
  FF_REQUIRE(N <= 65535, "ff_warp_affine_blend: N=%d > 65535 (fold N into C for identical transforms)", N);
  FF_REQUIRE(mode == 0 || mode == 1, "ff_warp_affine_blend: mode must be 0 (bilinear) or 1 (nearest)");
  FF_REQUIRE(dtype == FF_DT_F32 || dtype == FF_DT_BF16, "ff_warp_affine_blend: dtype must be f32 or bf16");
  const int tiles_x = (dW + TILE_X - 1) / TILE_X, tiles_y = (dH + TILE_Y - 1) / TILE_Y;
  // channel chunks: enough CTAs to cover the 148 SMs (2 resident CTAs each) a few times, but as few as possible so
  // that the per-tile coordinate / mask work is amortised over many channels and the copy pipeline gets long
  const long long tiles = (long long)tiles_x * tiles_y * N;
  int chunks = (int)((148LL * 2 * 4 + tiles - 1) / tiles);
  if (chunks < 1) chunks = 1;
  if (chunks > (C + CH_PER_ITER - 1) / CH_PER_ITER) chunks = (C + CH_PER_ITER - 1) / CH_PER_ITER;
  int ch_per_cta = (C + chunks - 1) / chunks;
  ch_per_cta = ((ch_per_cta + CH_PER_ITER - 1) / CH_PER_ITER) * CH_PER_ITER;
  chunks = (C + ch_per_cta - 1) / ch_per_cta;
  FF_REQUIRE(chunks <= 65535, "ff_warp_affine_blend: too many channel chunks");
  const size_t es = dtype == FF_DT_F32 ? 4 : 2;
  const int vec_ok = (dW % 4 == 0) && ff::aligned16(src) && ff::aligned16(out) && (!bg || ff::aligned16(bg)) &&
                     ((size_t)H * W * es % 16 == 0) && ((size_t)dH * dW * es % 16 == 0);
  dim3 grid(tiles_x * tiles_y, chunks, N);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static bool configured = false;
  if (!configured) {
    cudaError_t e1 = cudaFuncSetAttribute(warp_affine_blend_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SMEM_BYTES);
    cudaError_t e2 = cudaFuncSetAttribute(warp_affine_blend_kernel<__nv_bfloat16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e1 != cudaSuccess || e2 != cudaSuccess)
      return ff::fail(FF_E_CUDA, "ff_warp_affine_blend: cudaFuncSetAttribute: %s",
                      cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    configured = true;
  }
  if (dtype == FF_DT_F32)
    warp_affine_blend_kernel<float><<<grid, THREADS, SMEM_BYTES, st>>>(
        static_cast<const float*>(src), theta, mask_src, static_cast<const float*>(bg), static_cast<float*>(out),
        mask_out, C, H, W, dH, dW, mode, ch_per_cta, tiles_x, vec_ok);
  else
    warp_affine_blend_kernel<__nv_bfloat16><<<grid, THREADS, SMEM_BYTES, st>>>(
        static_cast<const __nv_bfloat16*>(src), theta, mask_src, static_cast<const __nv_bfloat16*>(bg),
        static_cast<__nv_bfloat16*>(out), mask_out, C, H, W, dH, dW, mode, ch_per_cta, tiles_x, vec_ok);
  return ff::check_launch("ff_warp_affine_blend");
}
