// Region-mask preparation for the fused attention: nearest down-sample of full-resolution uint8 masks to a
// layer's token grid and packing into bit-vectors (+ population count).
//
// Reference: Attention_Modulator.process_mask_before_attention (src/utils/attention.py:841-855) followed by
// .flatten() in prepare_various_attention_mask (:862-889).  Integer/index work: bit-exact.
//   - if mask.max() > 1 the reference computes (mask / max).to(uint8): only pixels equal to the max stay 1;
//   - F.interpolate(mode='nearest') source index = min(floor(i * (in/out)) , in-1) with the ratio in fp32 (ATen).
// The [B*heads,S,S] additive masks the reference builds from these vectors are never materialised.
#include "ff_common.cuh"

namespace {

__device__ __forceinline__ int nearest_src(int dst, int in_size, int out_size) {
  if (in_size == out_size) return dst;
  if (out_size == 2 * in_size) return dst >> 1;
  const float scale = (float)in_size / (float)out_size;
  const int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}

__global__ void __launch_bounds__(1024)
mask_downsample_pack_kernel(const uint8_t* __restrict__ masks, int H, int W, int h, int w,
                            uint32_t* __restrict__ bits, int words, int* __restrict__ popcount) {
  __shared__ int s_max;
  __shared__ int s_cnt;
  const uint8_t* m = masks + (size_t)blockIdx.x * H * W;
  if (threadIdx.x == 0) { s_max = 0; s_cnt = 0; }
  __syncthreads();
  // pass 1: max over the FULL-resolution mask (the reference normalises before it down-samples)
  int mx = 0;
  const int total = H * W;
  if ((reinterpret_cast<uintptr_t>(m) & 15u) == 0) {
    const uint4* m4 = reinterpret_cast<const uint4*>(m);
    for (int i = threadIdx.x; i < total / 16; i += blockDim.x) {
      uint4 v = __ldg(m4 + i);
      uint32_t a = __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w));
      a = max(max(a & 0xff, (a >> 8) & 0xff), max((a >> 16) & 0xff, a >> 24));
      mx = max(mx, (int)a);
    }
    for (int i = (total / 16) * 16 + threadIdx.x; i < total; i += blockDim.x) mx = max(mx, (int)m[i]);
  } else {
    for (int i = threadIdx.x; i < total; i += blockDim.x) mx = max(mx, (int)m[i]);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) atomicMax(&s_max, mx);
  __syncthreads();
  const int vmax = s_max;
  // pass 2: one warp per 32 tokens -> one word
  const int S = h * w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  int cnt = 0;
  for (int wd = warp; wd < words; wd += nwarp) {
    const int tok = wd * 32 + lane;
    bool on = false;
    if (tok < S) {
      const int oy = tok / w, ox = tok - oy * w;
      const uint8_t v = m[(size_t)nearest_src(oy, H, h) * W + nearest_src(ox, W, w)];
      on = vmax > 1 ? (v == vmax) : (v != 0);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) {
      bits[(size_t)blockIdx.x * words + wd] = word;
      cnt += __popc(word);
    }
  }
  if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) popcount[blockIdx.x] = s_cnt;
}

}  // namespace

extern "C" int ff_mask_downsample_pack(const uint8_t* masks, int32_t n, int32_t H, int32_t W, int32_t h, int32_t w,
                                       uint32_t* bits, int32_t words, int32_t* popcount, void* stream) {
  FF_REQUIRE(masks && bits && popcount, "ff_mask_downsample_pack: null pointer");
  FF_REQUIRE(n > 0 && H > 0 && W > 0 && h > 0 && w > 0, "ff_mask_downsample_pack: bad shape");
  FF_REQUIRE(words >= (h * w + 31) / 32, "ff_mask_downsample_pack: words=%d < ceil(%d/32)", words, h * w);
  mask_downsample_pack_kernel<<<n, 1024, 0, static_cast<cudaStream_t>(stream)>>>(masks, H, W, h, w, bits, words,
                                                                                 popcount);
  return ff::check_launch("ff_mask_downsample_pack");
}
