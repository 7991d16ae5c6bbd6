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

}  // namespace moka
