// Local cross-attention modulation epilogue (modulate_local_cross_attn, src/utils/attention.py:1381-1383):
// after the 77-key cross attention of the 4 streams [u_e, u_r, c_e, c_r] of an edit,
//   c_e' = region*c_e + (1-region)*u_e   (region in {0,1}: an exact select),   c_r' = u_r.
// In place on hs [n_edits, 4, S, C]; region = bit-vector of S bits per edit.  128-bit accesses.
#include "ff_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
cross_region_blend_kernel(uint4* __restrict__ hs, const uint32_t* __restrict__ bitmasks, int mask_words,
                          const int* __restrict__ region_mask, int n_edits, int S, int row_vec) {
  // row_vec = C*elem_size/16 vectors per token row
  const long long per_stream = (long long)S * row_vec;
  const long long total = (long long)n_edits * per_stream;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / per_stream);
    const long long r = i - (long long)e * per_stream;
    const int tok = (int)(r / row_vec);
    const uint32_t word = __ldg(bitmasks + (size_t)__ldg(region_mask + e) * mask_words + (tok >> 5));
    const bool in_region = (word >> (tok & 31)) & 1u;
    uint4* base = hs + (long long)e * 4 * per_stream + r;
    const uint4 ur = base[per_stream];
    if (!in_region) base[2 * per_stream] = base[0];
    base[3 * per_stream] = ur;
  }
}

}  // namespace

extern "C" int ff_cross_region_blend(void* hs, const uint32_t* bitmasks, int32_t mask_words,
                                     const int32_t* region_mask, int32_t n_edits, int32_t S, int32_t C, int32_t dtype,
                                     void* stream) {
  FF_REQUIRE(hs && bitmasks && region_mask, "ff_cross_region_blend: null pointer");
  FF_REQUIRE(n_edits > 0 && S > 0 && C > 0, "ff_cross_region_blend: bad shape");
  FF_REQUIRE(dtype == FF_DT_F32 || dtype == FF_DT_BF16, "ff_cross_region_blend: dtype must be f32 or bf16");
  const int es = dtype == FF_DT_F32 ? 4 : 2;
  FF_REQUIRE((C * es) % 16 == 0 && ff::aligned16(hs), "ff_cross_region_blend: rows must be 16-byte multiples");
  FF_REQUIRE(mask_words >= (S + 31) / 32, "ff_cross_region_blend: mask_words too small");
  const int row_vec = C * es / 16;
  const long long total = (long long)n_edits * S * row_vec;
  long long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  cross_region_blend_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(hs), bitmasks, mask_words, region_mask, n_edits, S, row_vec);
  return ff::check_launch("ff_cross_region_blend");
}
