// (c) fused classifier-free guidance + DDIM / local-DDPM step, and the DDIM inversion step.
//
// Reference arithmetic: src/demo/model.py:605-611 (local CFG), :134-198 (ctrl_step), :109-132 (inv_step).
// HBM-bound elementwise work: one pass over eps_u, eps_c, x, noise (+2 byte masks) -> x_prev (+pred_x0), 128-bit
// accesses, every load issued before the first use.  Arithmetic is fp32 with the reference's operation ORDER and
// explicit round-to-nearest intrinsics (no FMA contraction) so results are bit-identical to the CPU oracle.
#include "ff_common.cuh"

namespace {

struct StepCoef {
  float gs, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma;
};

// one element of ctrl_step; m is the uint8 variance mask of this stream (ref stream: 1), sd its sigma
__device__ __forceinline__ void step_elem(float eu, float ec, float x, float nz, bool do_cfg, bool has_cfg_mask,
                                          uint8_t cfgm,
                                          uint8_t m, float sd, float cd, bool has_noise, const StepCoef& k,
                                          float& x_prev, float& x0) {
  // eps = eps_u + gs*(eps_c-eps_u)*cfg_mask                                   (model.py:608 / :610-611)
  float eps = eu;
  if (do_cfg) {
    float g = __fmul_rn(k.gs, __fsub_rn(ec, eu));
    if (has_cfg_mask) g = __fmul_rn(g, (float)cfgm);
    eps = __fadd_rn(eu, g);
  }
  // pred_x0 = (x - (1-a_t)**.5 * eps) / a_t**.5                               (model.py:162-163)
  x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sqrt_1m_at, eps)), k.sqrt_at);
  const uint8_t om = (uint8_t)(1 - m);  // uint8 wrap-around: 1-2 = 255 (quirk Q1, model.py:179-180)
  const float pdm = __fmul_rn(__fmul_rn(cd, eps), (float)m);                     // :177-178
  const float dir = __fadd_rn(__fmul_rn(__fmul_rn(k.c_ddim, eps), (float)om), pdm);  // :179-180
  float xp = __fadd_rn(__fmul_rn(k.sqrt_ap, x0), dir);                           // :183
  if (has_noise) xp = __fadd_rn(xp, __fmul_rn(__fmul_rn(sd, nz), (float)m));     // :186-196
  x_prev = xp;
}

template <int V>  // V = 4 (float4 path, hw % 4 == 0) or 1
__global__ void __launch_bounds__(256)
ddim_cfg_step_kernel(const float* __restrict__ eps4, const float* __restrict__ x, const float* __restrict__ noise,
                     const uint8_t* __restrict__ cfg_mask, const uint8_t* __restrict__ var_mask, StepCoef k,
                     float* __restrict__ x_prev, float* __restrict__ pred_x0, int n_edits, int C, int hw, int spe,
                     int so, int cs) {
  // spe = streams per edit in eps4, so = latent streams per edit (2: [edit, ref], 1: edit only), cs = distance from an
  // unconditional stream to its conditional partner (0: eps is already guidance-combined):
  //   edit / bg-gen  spe 4, so 2, cs 2   [u_e,u_r,c_e,c_r]          (model.py:594-617)
  //   ctrl_step only spe 2, so 2, cs 0
  //   compose        spe N+2, so 1, cs N+1   [e, r_1..r_N, c_e]     (model.py:407-431)
  const int hwv = hw / V;
  const long long total = (long long)n_edits * so * C * hwv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % hwv);
    long long r = i / hwv;
    const int c = (int)(r % C);
    r /= C;
    const int s = so == 2 ? (int)(r & 1) : 0;
    const int e = so == 2 ? (int)(r >> 1) : (int)r;
    const long long off_u = (((long long)(e * spe + s) * C + c) * hw) + (long long)p * V;
    const long long off_c = (((long long)(e * spe + s + cs) * C + c) * hw) + (long long)p * V;
    const long long off_x = (((long long)(e * so + s) * C + c) * hw) + (long long)p * V;
    const long long off_m = (long long)e * hw + (long long)p * V;
    float eu[V], ec[V], xv[V], nz[V], xp[V], x0[V];
    uint8_t cm[V], vm[V];
    // loads first (memory-level parallelism), arithmetic after
    if (V == 4) {
      *reinterpret_cast<float4*>(eu) = __ldg(reinterpret_cast<const float4*>(eps4 + off_u));
      *reinterpret_cast<float4*>(ec) = __ldg(reinterpret_cast<const float4*>(eps4 + off_c));
      *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(x + off_x));
      if (noise) *reinterpret_cast<float4*>(nz) = __ldg(reinterpret_cast<const float4*>(noise + off_x));
      if (cfg_mask) *reinterpret_cast<uint32_t*>(cm) = __ldg(reinterpret_cast<const uint32_t*>(cfg_mask + off_m));
      if (s == 0) *reinterpret_cast<uint32_t*>(vm) = __ldg(reinterpret_cast<const uint32_t*>(var_mask + off_m));
    } else {
      eu[0] = __ldg(eps4 + off_u);
      ec[0] = __ldg(eps4 + off_c);
      xv[0] = __ldg(x + off_x);
      if (noise) nz[0] = __ldg(noise + off_x);
      if (cfg_mask) cm[0] = __ldg(cfg_mask + off_m);
      if (s == 0) vm[0] = __ldg(var_mask + off_m);
    }
    // ref stream: mask = ones, sigma = 0, (1-a_prev-0)**.5 == c_ddim         (model.py:169-174)
    const float sd = s == 0 ? k.sigma : 0.f;
    const float cd = s == 0 ? k.c_ddpm : k.c_ddim;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const uint8_t m = s == 0 ? vm[j] : (uint8_t)1;
      step_elem(eu[j], ec[j], xv[j], noise ? nz[j] : 0.f, cs != 0, cfg_mask != nullptr, cfg_mask ? cm[j] : (uint8_t)1,
                m, sd, cd, noise != nullptr, k, xp[j], x0[j]);
    }
    if (V == 4) {
      *reinterpret_cast<float4*>(x_prev + off_x) = *reinterpret_cast<float4*>(xp);
      if (pred_x0) *reinterpret_cast<float4*>(pred_x0 + off_x) = *reinterpret_cast<float4*>(x0);
    } else {
      x_prev[off_x] = xp[0];
      if (pred_x0) pred_x0[off_x] = x0[0];
    }
  }
}

template <int V>
__global__ void __launch_bounds__(256)
ddim_inv_step_kernel(const float* __restrict__ eps, const float* __restrict__ x, float sqrt_1m_at, float sqrt_at,
                     float sqrt_an, float c_next, float* __restrict__ x_next, float* __restrict__ pred_x0,
                     long long nv) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float e[V], xv[V], xn[V], x0[V];
    if (V == 4) {
      *reinterpret_cast<float4*>(e) = __ldg(reinterpret_cast<const float4*>(eps) + i);
      *reinterpret_cast<float4*>(xv) = __ldg(reinterpret_cast<const float4*>(x) + i);
    } else {
      e[0] = __ldg(eps + i);
      xv[0] = __ldg(x + i);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      x0[j] = __fdiv_rn(__fsub_rn(xv[j], __fmul_rn(sqrt_1m_at, e[j])), sqrt_at);   // model.py:128
      xn[j] = __fadd_rn(__fmul_rn(sqrt_an, x0[j]), __fmul_rn(c_next, e[j]));       // :129-130
    }
    if (V == 4) {
      reinterpret_cast<float4*>(x_next)[i] = *reinterpret_cast<float4*>(xn);
      if (pred_x0) reinterpret_cast<float4*>(pred_x0)[i] = *reinterpret_cast<float4*>(x0);
    } else {
      x_next[i] = xn[0];
      if (pred_x0) pred_x0[i] = x0[0];
    }
  }
}

inline int grid_for(long long work_items, int block) {
  // persistent-style grid: a multiple of the 148 SMs, 8 resident CTAs of 256 threads each at most
  long long need = (work_items + block - 1) / block;
  const long long cap = 148LL * 8;
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

}  // namespace

static int launch_step(const char* what, const float* eps, int spe, int so, int cs, const float* x, const float* noise,
                       const uint8_t* cfg_mask, const uint8_t* var_mask, const StepCoef& k, float* x_prev,
                       float* pred_x0, int32_t n_edits, int32_t C, int32_t h, int32_t w, void* stream) {
  if (!(eps && x && var_mask && x_prev)) return ff::fail(FF_E_INVALID, "%s: null pointer", what);
  if (!(n_edits > 0 && C > 0 && h > 0 && w > 0))
    return ff::fail(FF_E_INVALID, "%s: bad shape %d,%d,%d,%d", what, n_edits, C, h, w);
  const int hw = h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (hw % 4 == 0) && ff::aligned16(eps) && ff::aligned16(x) && ff::aligned16(x_prev) &&
                   (!noise || ff::aligned16(noise)) && (!pred_x0 || ff::aligned16(pred_x0)) &&
                   (reinterpret_cast<uintptr_t>(var_mask) % 4 == 0) &&
                   (!cfg_mask || reinterpret_cast<uintptr_t>(cfg_mask) % 4 == 0);
  if (vec) {
    const long long items = (long long)n_edits * so * C * (hw / 4);
    ddim_cfg_step_kernel<4><<<grid_for(items, 256), 256, 0, st>>>(eps, x, noise, cfg_mask, var_mask, k, x_prev,
                                                                   pred_x0, n_edits, C, hw, spe, so, cs);
  } else {
    const long long items = (long long)n_edits * so * C * hw;
    ddim_cfg_step_kernel<1><<<grid_for(items, 256), 256, 0, st>>>(eps, x, noise, cfg_mask, var_mask, k, x_prev,
                                                                   pred_x0, n_edits, C, hw, spe, so, cs);
  }
  return ff::check_launch(what);
}

extern "C" int ff_ddim_cfg_step(const float* eps4, const float* x, const float* noise, const uint8_t* cfg_mask,
                                const uint8_t* var_mask, float guidance_scale, float sqrt_1m_at, float sqrt_at,
                                float sqrt_ap, float c_ddim, float c_ddpm, float sigma, float* x_prev,
                                float* pred_x0, int32_t n_edits, int32_t C, int32_t h, int32_t w, void* stream) {
  StepCoef k{guidance_scale, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma};
  return launch_step("ff_ddim_cfg_step", eps4, 4, 2, 2, x, noise, cfg_mask, var_mask, k, x_prev, pred_x0, n_edits, C, h,
                     w, stream);
}

extern "C" int ff_ddim_step(const float* eps2, const float* x, const float* noise, const uint8_t* var_mask,
                            float sqrt_1m_at, float sqrt_at, float sqrt_ap, float c_ddim, float c_ddpm, float sigma,
                            float* x_prev, float* pred_x0, int32_t n_edits, int32_t C, int32_t h, int32_t w,
                            void* stream) {
  StepCoef k{0.f, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma};
  return launch_step("ff_ddim_step", eps2, 2, 2, 0, x, noise, nullptr, var_mask, k, x_prev, pred_x0, n_edits, C, h, w,
                     stream);
}

extern "C" int ff_ddim_cfg_step_compose(const float* eps, int32_t streams_per_edit, const float* x, const float* noise,
                                        const uint8_t* cfg_mask, const uint8_t* var_mask, float guidance_scale,
                                        float sqrt_1m_at, float sqrt_at, float sqrt_ap, float c_ddim, float c_ddpm,
                                        float sigma, float* x_prev, float* pred_x0, int32_t n_edits, int32_t C, int32_t h,
                                        int32_t w, void* stream) {
  FF_REQUIRE(streams_per_edit >= 2, "ff_ddim_cfg_step_compose: streams_per_edit=%d < 2", streams_per_edit);
  StepCoef k{guidance_scale, sqrt_1m_at, sqrt_at, sqrt_ap, c_ddim, c_ddpm, sigma};
  return launch_step("ff_ddim_cfg_step_compose", eps, streams_per_edit, 1, streams_per_edit - 1, x, noise, cfg_mask,
                     var_mask, k, x_prev, pred_x0, n_edits, C, h, w, stream);
}

extern "C" int ff_ddim_inv_step(const float* eps, const float* x, float sqrt_1m_at, float sqrt_at, float sqrt_an,
                                float c_next, float* x_next, float* pred_x0, int64_t n, void* stream) {
  FF_REQUIRE(eps && x && x_next, "ff_ddim_inv_step: null pointer");
  FF_REQUIRE(n > 0, "ff_ddim_inv_step: n must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (n % 4 == 0) && ff::aligned16(eps) && ff::aligned16(x) && ff::aligned16(x_next) &&
                   (!pred_x0 || ff::aligned16(pred_x0));
  if (vec) {
    ddim_inv_step_kernel<4><<<grid_for(n / 4, 256), 256, 0, st>>>(eps, x, sqrt_1m_at, sqrt_at, sqrt_an, c_next,
                                                                   x_next, pred_x0, n / 4);
  } else {
    ddim_inv_step_kernel<1><<<grid_for(n, 256), 256, 0, st>>>(eps, x, sqrt_1m_at, sqrt_at, sqrt_an, c_next, x_next,
                                                               pred_x0, n);
  }
  return ff::check_launch("ff_ddim_inv_step");
}
