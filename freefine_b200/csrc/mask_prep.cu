// Mask preparation of an edit in ONE launch: dilate_mask + prepare_various_mask.
//
// Reference: FreeFinePipeline.prepare_various_mask (src/demo/model.py:1432-1512) with dilate_mask (:927-934 =
// cv2.dilate(mask, ones(k, k)), anchor k//2, k = 15 for the moved object, 30 for the vacated region), prepare_tensor_mask
// (:1622-1639, binarisation `> 0`) and the two nearest down-samplings to the latent grid (:1505-1511).  Integer work:
// bit-exact, INCLUDING the uint8 wrap-around of the reference's tensor algebra (quirk Q1): `cons - ori` is 255 where the
// original object is not part of the constraint area, and `1 - 255` is 2, so completion / local-variance masks take the
// values {0, 1, 2}.  The reference does this with cv2 on the CPU plus ~15 eager uint8 tensor ops; round 1 of this repo with
// F.max_pool2d over the whole image (at::max_pool_forward_nchw: 1.7 ms x 2 per batch) plus a dozen eager ops.
//
// One thread per LATENT pixel.  It writes the r x r block of full-resolution outputs it covers (fg / shifted / ori, the
// controller's attention masks) and evaluates the dilations only where they are consumed -- at the sampled pixel
// (i*r, j*r) of the nearest down-sampling -- by scanning the k x k window directly (binary masks: dilation = OR).
#include "ff_common.cuh"

namespace {

__device__ __forceinline__ uint8_t bin(const uint8_t* m, int idx) { return m ? (uint8_t)(__ldg(m + idx) != 0) : (uint8_t)0; }

// OR over the k x k window of cv2.dilate anchored at k/2: rows [y - k/2, y + k - 1 - k/2], same for columns; outside = 0
__device__ __forceinline__ uint8_t window_or(const uint8_t* __restrict__ m, int H, int W, int y, int x, int k) {
  const int a = k >> 1;
  const int y0 = max(0, y - a), y1 = min(H - 1, y + k - 1 - a);
  const int x0 = max(0, x - a), x1 = min(W - 1, x + k - 1 - a);
  uint32_t acc = 0;
  for (int yy = y0; yy <= y1; ++yy) {
    const uint8_t* row = m + (size_t)yy * W;
    int xx = x0;
    for (; xx <= x1 && (xx & 3); ++xx) acc |= __ldg(row + xx);
    for (; xx + 3 <= x1; xx += 4) acc |= __ldg(reinterpret_cast<const uint32_t*>(row + xx));
    for (; xx <= x1; ++xx) acc |= __ldg(row + xx);
    if (acc) return 1;
  }
  return acc != 0;
}

__global__ void __launch_bounds__(256)
mask_prep_kernel(const uint8_t* __restrict__ shifted, const uint8_t* __restrict__ ori, const uint8_t* __restrict__ draw,
                 const uint8_t* __restrict__ cons, int E, int H, int W, int h, int w, int auto_draw, int reduce,
                 uint8_t* __restrict__ fg, uint8_t* __restrict__ sh_out, uint8_t* __restrict__ ori_out,
                 uint8_t* __restrict__ comp_lat, uint8_t* __restrict__ lvar_lat) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= E * h * w) return;
  const int e = t / (h * w), ij = t - e * h * w, i = ij / w, j = ij - i * w;
  const int ry = H / h, rx = W / w;                       // integer ratios (checked on the host)
  const size_t base = (size_t)e * H * W;
  const uint8_t* s_ = shifted + base;
  const uint8_t* o_ = ori + base;
  const uint8_t* d_ = draw ? draw + base : nullptr;
  const uint8_t* c_ = cons ? cons + base : nullptr;
  // ---- full-resolution outputs of my r x r block (elementwise)
  for (int dy = 0; dy < ry; ++dy) {
    const int y = i * ry + dy;
    for (int dx = 0; dx < rx; ++dx) {
      const int idx = y * W + j * rx + dx;
      const uint8_t sh = bin(s_, idx);
      uint8_t f = sh;
      if (!auto_draw) {
        const uint8_t flex = (uint8_t)(bin(d_, idx) * (uint8_t)(1 - sh));
        f = (uint8_t)((uint8_t)(flex + sh) > 0);
      }
      fg[base + idx] = f;
      sh_out[base + idx] = sh;
      ori_out[base + idx] = bin(o_, idx);
    }
  }
  // ---- latent-resolution outputs: nearest down-sampling reads full-resolution pixel (i*ry, j*rx)
  const int y = i * ry, x = j * rx, idx = y * W + x;
  const uint8_t sh = bin(s_, idx), orib = bin(o_, idx);
  uint8_t comp, lvar;
  if (!auto_draw) {
    const uint8_t flex = (uint8_t)(bin(d_, idx) * (uint8_t)(1 - sh));
    comp = flex;
    if (!reduce) {
      lvar = flex;
    } else {
      const uint8_t dil = window_or(o_, H, W, y, x, 30);
      const uint8_t v = (uint8_t)((uint8_t)((uint8_t)(1 - bin(c_, idx)) * (uint8_t)(1 - sh)) * dil + flex);
      lvar = (uint8_t)(v > 0);
    }
  } else {
    const uint8_t c2 = (uint8_t)(bin(c_, idx) - orib);                       // wraps to 255 (quirk Q1)
    const uint8_t gate = (uint8_t)((uint8_t)(1 - c2) * (uint8_t)(1 - sh));   // (1 - 255) = 2
    if (!reduce) {
      comp = (uint8_t)(gate * window_or(s_, H, W, y, x, 15));
    } else {
      uint8_t u = (uint8_t)(window_or(o_, H, W, y, x, 30) + window_or(s_, H, W, y, x, 15));
      u = (uint8_t)(u > 0);
      comp = (uint8_t)(u * gate);
    }
    lvar = comp;
  }
  comp_lat[(size_t)e * h * w + ij] = comp;
  lvar_lat[(size_t)e * h * w + ij] = lvar;
}

}  // namespace

extern "C" int ff_mask_prep(const uint8_t* shifted, const uint8_t* ori, const uint8_t* draw, const uint8_t* cons,
                            int32_t E, int32_t H, int32_t W, int32_t h, int32_t w, int32_t use_auto_draw,
                            int32_t reduce_inp_artifacts, uint8_t* fg, uint8_t* shifted_out, uint8_t* ori_out,
                            uint8_t* comp_lat, uint8_t* lvar_lat, void* stream) {
  FF_REQUIRE(shifted && ori && fg && shifted_out && ori_out && comp_lat && lvar_lat, "ff_mask_prep: null pointer");
  FF_REQUIRE(E > 0 && H > 0 && W > 0 && h > 0 && w > 0, "ff_mask_prep: bad shape");
  FF_REQUIRE(H % h == 0 && W % w == 0, "ff_mask_prep: the latent grid must divide the image (H=%d h=%d W=%d w=%d)", H, h, W, w);
  FF_REQUIRE(use_auto_draw || draw, "ff_mask_prep: draw mask needed without use_auto_draw");
  FF_REQUIRE(!(use_auto_draw || reduce_inp_artifacts) || cons, "ff_mask_prep: cons_area needed (reference asserts it)");
  FF_REQUIRE(W % 4 == 0 && (reinterpret_cast<uintptr_t>(shifted) & 3u) == 0 && (reinterpret_cast<uintptr_t>(ori) & 3u) == 0,
             "ff_mask_prep: masks must be 4-byte aligned with W %% 4 == 0");
  const int total = E * h * w;
  mask_prep_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      shifted, ori, draw, cons, E, H, W, h, w, use_auto_draw, reduce_inp_artifacts, fg, shifted_out, ori_out, comp_lat, lvar_lat);
  return ff::check_launch("ff_mask_prep");
}

// Stand-alone square dilation (dilate_mask, model.py:927-934 / vis_utils.py:340-347): out = cv2.dilate(mask != 0, ones(k,k)).
namespace {
__global__ void __launch_bounds__(256)
dilate_kernel(const uint8_t* __restrict__ m, uint8_t* __restrict__ out, int N, int H, int W, int k) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N * H * W) return;
  const int n = t / (H * W), p = t - n * H * W, y = p / W, x = p - y * W;
  out[t] = window_or(m + (size_t)n * H * W, H, W, y, x, k);
}
}  // namespace

extern "C" int ff_dilate_mask(const uint8_t* mask, uint8_t* out, int32_t N, int32_t H, int32_t W, int32_t k, void* stream) {
  FF_REQUIRE(mask && out, "ff_dilate_mask: null pointer");
  FF_REQUIRE(N > 0 && H > 0 && W > 0 && k > 0 && k <= 255, "ff_dilate_mask: bad shape / kernel size");
  FF_REQUIRE(W % 4 == 0 && (reinterpret_cast<uintptr_t>(mask) & 3u) == 0, "ff_dilate_mask: mask must be 4-byte aligned, W %% 4 == 0");
  const long long total = (long long)N * H * W;
  dilate_kernel<<<(int)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, out, N, H, W, k);
  return ff::check_launch("ff_dilate_mask");
}
