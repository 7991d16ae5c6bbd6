// Sharded AdamW over the flat adapter parameter vector (decoupled weight decay,
// bias-corrected, fp32 master + rounded working copy).  The reference delegates this to
// DeepSpeed ZeRO-1/2 with torch AdamW semantics (VisualText/zero_stage2_config.json:23-34,
// AudioVisualText/deepspeed/stage2-offload.json:9 "torch_adam": true).
#include "moka_common.cuh"
#include "moka_kernels.h"

namespace moka {

template <typename P>
__global__ void __launch_bounds__(256)
k_adamw(int64_t n, float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
        const float* __restrict__ g, P* __restrict__ out, float lr, float b1, float b2, float eps,
        float wd, float bc1, float bc2_sqrt, float gscale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    float wi = w[i];
    wi *= (1.f - lr * wd);                       // torch.optim.AdamW: decoupled decay first
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    wi -= (lr / bc1) * (mi / denom);
    w[i] = wi;
    if (out) out[i] = static_cast<P>(wi);
  }
}

int launch_adamw(int64_t n, float* master, float* m, float* v, const float* grad, void* param_out, int dtype,
                 float lr, float b1, float b2, float eps, float wd, int step, float grad_scale, cudaStream_t st) {
  if (n <= 0) return MOKA_OK;
  const float bc1 = 1.f - powf(b1, (float)step);
  const float bc2s = sqrtf(1.f - powf(b2, (float)step));
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  if (dtype == MOKA_BF16)
    k_adamw<__nv_bfloat16><<<grid, 256, 0, st>>>(n, master, m, v, grad, static_cast<__nv_bfloat16*>(param_out), lr, b1, b2, eps, wd, bc1, bc2s, grad_scale);
  else
    k_adamw<float><<<grid, 256, 0, st>>>(n, master, m, v, grad, static_cast<float*>(param_out), lr, b1, b2, eps, wd, bc1, bc2s, grad_scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}


// Rotary position embedding applied IN PLACE on a [B, S, H, D] bf16 tensor (half-split
// convention of HF Llama: pairs (i, i + D/2)).  sign = +1 forward, -1 backward (the transpose of
// a rotation is the rotation by the opposite angle).  Helper for the frozen Llama caller in
// moka_b200/llama.py -- the reference does this with eager ops (modeling_llama.py apply_rotary_pos_emb).
__global__ void __launch_bounds__(256)
k_rope(const __nv_bfloat16* src, __nv_bfloat16* dst, int S, int H, int D, float log2_theta, float sign, long long npairs8) {
  const int half = D >> 1;
  const int gpr = half >> 3;                     // groups of 8 pairs per (token, head)
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < npairs8; g += (long long)gridDim.x * blockDim.x) {
    const int gi = (int)(g % gpr);
    const long long th = g / gpr;                // (b*S + s) * H + h
    const int spos = (int)((th / H) % S);
    const __nv_bfloat16* sbase = src + th * D + gi * 8;
    __nv_bfloat16* base = dst + th * D + gi * 8;
    uint4 lo = *reinterpret_cast<const uint4*>(sbase);
    uint4 hi = *reinterpret_cast<const uint4*>(sbase + half);
    uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
    uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int i = gi * 8 + e;
      const float inv = exp2f(-log2_theta * (float)(2 * i) / (float)D);
      float sn, cs;
      sincosf((float)spos * inv, &sn, &cs);
      sn *= sign;
      const float a = (e & 1) ? bf16hi(lw[e >> 1]) : bf16lo(lw[e >> 1]);
      const float b = (e & 1) ? bf16hi(hw[e >> 1]) : bf16lo(hw[e >> 1]);
      const __nv_bfloat16 ra = __float2bfloat16_rn(a * cs - b * sn);
      const __nv_bfloat16 rb = __float2bfloat16_rn(b * cs + a * sn);
      reinterpret_cast<__nv_bfloat16*>(lw)[e] = ra;
      reinterpret_cast<__nv_bfloat16*>(hw)[e] = rb;
    }
    *reinterpret_cast<uint4*>(base) = lo;
    *reinterpret_cast<uint4*>(base + half) = hi;
  }
}

int launch_rope(const void* src, void* x, long long BS, int S, int H, int D, float theta, float sign, cudaStream_t st) {
  const long long npairs8 = BS * H * (D / 16);
  const int grid = (int)((npairs8 + 255) / 256 < 148 * 16 ? (npairs8 + 255) / 256 : 148 * 16);
  k_rope<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(x), S, H, D, log2f(theta), sign, npairs8);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}

}  // namespace moka
