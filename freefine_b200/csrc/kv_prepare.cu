// K/V staging for ff_attn_masked_kv: ONE pass that (i) gathers the token rows of K and V in "mask bits first" order
// (FF_PASS_KEY_PREFIX: softmax is invariant under a permutation of the keys, src/utils/attention.py:774-806 sees the
// same set) and (ii) converts V from bf16 to fp16, the format of the single-operand P.V contraction, into a per-head
// padded layout whose first padding channel is 1.0 (so that P.V also yields the softmax denominator).  bf16 -> fp16 is
// exact for 2^-14 <= |v| <= 65504; larger magnitudes saturate to +-65504, smaller ones round to fp16 subnormals
// (absolute error < 2^-25).  HBM-bound: 2 tensors read + 2 written, 128-bit accesses, rows are whole 16-byte multiples.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ff_common.cuh"

namespace {

__device__ __forceinline__ uint32_t bf16x2_to_f16x2_sat(uint32_t b) {
  float lo = __uint_as_float(b << 16), hi = __uint_as_float(b & 0xffff0000u);
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);   // (NaN stays NaN: fminf/fmaxf return the non-NaN operand -> clamp of NaN
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);   //  gives -65504; inputs are finite activations)
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One thread per 16-byte chunk.  K: rows * (heads*d/8) chunks, plain gather.  V: rows * heads * (vhs/8) chunks of the
// padded layout [rows, heads, vhs]: chunks below d come from V, the chunk at d holds {1,0,..,0}, the rest zeros.
template <bool F16>
__global__ void __launch_bounds__(256)
kv_gather_cast_kernel(const uint4* __restrict__ k, const uint4* __restrict__ v, const long long* __restrict__ idx,
                      uint4* __restrict__ k_out, uint4* __restrict__ v_out, long long rows, int heads, int dv, int vv) {
  // dv = head_dim/8 chunks per head in the source, vv = v_head_stride/8 chunks per head in the staging
  const int krow = heads * dv, vrow = heads * vv;
  const long long total_v = rows * vrow;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_v;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vrow;
    const int c = (int)(i - r * vrow);
    const int h = c / vv, w = c - h * vv;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (w < dv) {
      const long long src = idx ? __ldg(idx + r) : r;
      const uint4 x = __ldg(v + src * krow + h * dv + w);
      if (F16) {
        o.x = bf16x2_to_f16x2_sat(x.x);
        o.y = bf16x2_to_f16x2_sat(x.y);
        o.z = bf16x2_to_f16x2_sat(x.z);
        o.w = bf16x2_to_f16x2_sat(x.w);
      } else {
        o = x;
      }
      if (k_out) k_out[r * krow + h * dv + w] = __ldg(k + src * krow + h * dv + w);
    } else if (w == dv) {
      o.x = F16 ? 0x00003c00u : 0x00003f80u;   // 1.0 (fp16 / bf16) at channel head_dim
    }
    v_out[i] = o;
  }
}

}  // namespace

extern "C" int ff_kv_gather_cast(const void* k, const void* v, const int64_t* row_index, void* k_out, void* v_out,
                                 int32_t v_out_dtype, int64_t rows, int32_t heads, int32_t head_dim, void* stream) {
  void* v_out_f16 = v_out;
  FF_REQUIRE(v && v_out, "ff_kv_gather_cast: null pointer");
  FF_REQUIRE(v_out_dtype == FF_DT_F16 || v_out_dtype == FF_DT_BF16, "ff_kv_gather_cast: v_out_dtype must be f16 or bf16");
  FF_REQUIRE((k == nullptr) == (k_out == nullptr), "ff_kv_gather_cast: k and k_out go together");
  FF_REQUIRE(rows > 0 && heads > 0 && head_dim >= 8 && head_dim % 8 == 0 && head_dim <= 160,
             "ff_kv_gather_cast: rows=%lld heads=%d head_dim=%d (head_dim must be a multiple of 8, <= 160)",
             (long long)rows, heads, head_dim);
  FF_REQUIRE(ff::aligned16(v) && ff::aligned16(v_out_f16) && ff::aligned16(k) && ff::aligned16(k_out),
             "ff_kv_gather_cast: pointers must be 16-byte aligned");
  const int vhs = ff_attn_v_head_stride(head_dim);
  const long long total = rows * heads * (vhs / 8);
  long long grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (v_out_dtype == FF_DT_F16)
    kv_gather_cast_kernel<true><<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(k), static_cast<const uint4*>(v), reinterpret_cast<const long long*>(row_index),
        static_cast<uint4*>(k_out), static_cast<uint4*>(v_out), rows, heads, head_dim / 8, vhs / 8);
  else
    kv_gather_cast_kernel<false><<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(k), static_cast<const uint4*>(v), reinterpret_cast<const long long*>(row_index),
        static_cast<uint4*>(k_out), static_cast<uint4*>(v_out), rows, heads, head_dim / 8, vhs / 8);
  return ff::check_launch("ff_kv_gather_cast");
}
