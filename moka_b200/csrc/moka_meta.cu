// Masks -> token metadata, one CTA per sample, no host synchronisation.
//
// Consumes the reference's mask tensors as produced
// (AudioVisualText/models/unified_arch.py:159-204,336: four [B,S,1] int32 tensors;
//  VisualText/train/train.py:206-231: three [B,S] bool tensors) and replaces the
// per-layer `torch.where(question_mask[i]==1)` / `nonzero` of lora.py:489-491 and
// layer.py:603-637 by one pass per training step.
#include <climits>
#include "moka_common.cuh"
#include "moka_kernels.h"

namespace moka {

__device__ __forceinline__ int read_mask(const void* p, int eb, size_t i) {
  if (!p) return 0;
  if (eb == 1) return static_cast<const uint8_t*>(p)[i] != 0;
  if (eb == 4) return static_cast<const int32_t*>(p)[i] != 0;
  return static_cast<const long long*>(p)[i] != 0;
}

struct MaskPtrs { const void* p[4]; };

__global__ void __launch_bounds__(256)
k_build_meta(int variant, int S, int Lk, MaskPtrs mk, int eb, uint8_t* tokmeta, int32_t* kidx,
             int32_t* klen, int32_t* status) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ int s_first, s_last, s_err, s_base;
  __shared__ int s_warp[8];
  if (tid == 0) { s_first = INT_MAX; s_last = -1; s_err = 0; s_base = 0; }
  __syncthreads();
  int err = 0, first = INT_MAX, last = -1;
  for (int i = tid; i < S; i += 256) {
    const size_t g = (size_t)b * S + i;
    const int tx = read_mask(mk.p[0], eb, g), vi = read_mask(mk.p[1], eb, g), au = read_mask(mk.p[2], eb, g);
    const int q = read_mask(mk.p[3], eb, g);
    if (tx + vi + au > 1) err |= MOKA_META_OVERLAP;
    const int code = tx ? MOKA_MOD_TEXT : (vi ? MOKA_MOD_VIS : (au ? MOKA_MOD_AUD : MOKA_MOD_NONE));
    const bool kv = (variant == MOKA_VARIANT_AVT) ? (q && tx) : (q && code != MOKA_MOD_NONE);
    tokmeta[g] = (uint8_t)(code | (kv ? MOKA_KEY_VALID : 0));
    if (q) { first = min(first, i); last = max(last, i); }
  }
  if (err) atomicOr(&s_err, err);
  if (last >= 0) { atomicMin(&s_first, first); atomicMax(&s_last, last); }
  __syncthreads();
  int n = 0;
  if (variant == MOKA_VARIANT_AVT) {
    // contiguous span first..last (lora.py:489-491); empty -> the reference raises IndexError
    if (s_last < 0) {
      if (tid == 0) atomicOr(&s_err, MOKA_META_EMPTY_QUESTION);
    } else {
      n = s_last - s_first + 1;
      for (int k = tid; k < min(n, Lk); k += 256) kidx[(size_t)b * Lk + k] = s_first + k;
    }
  } else {
    // exact, ordered gather of the question positions (layer.py:635)
    const int lane = tid & 31, warp = tid >> 5;
    for (int base = 0; base < S; base += 256) {
      const int i = base + tid;
      const int q = (i < S) ? read_mask(mk.p[3], eb, (size_t)b * S + i) : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, q);
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      int off = s_base;
      for (int w = 0; w < warp; ++w) off += s_warp[w];
      const int k = off + __popc(bal & ((1u << lane) - 1));
      if (q && k < Lk) kidx[(size_t)b * Lk + k] = i;
      __syncthreads();
      if (tid == 0) { int tot = 0; for (int w = 0; w < 8; ++w) tot += s_warp[w]; s_base += tot; }
      __syncthreads();
    }
    n = s_base;
  }
  for (int k = max(min(n, Lk), 0) + tid; k < Lk; k += 256) kidx[(size_t)b * Lk + k] = -1;
  if (tid == 0) {
    int e = s_err;
    if (n > Lk) { e |= MOKA_META_KEY_OVERFLOW; }
    klen[b] = min(n, Lk);
    if (e) atomicOr(&status[0], e);
    atomicMax(&status[1], n);
  }
}

int launch_build_meta(int variant, int B, int S, int Lk, const void* const* masks, int eb,
                      uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status, cudaStream_t st) {
  MaskPtrs mk{{masks[0], masks[1], masks[2], masks[3]}};
  k_build_meta<<<B, 256, 0, st>>>(variant, S, Lk, mk, eb, tokmeta, kidx, klen, status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}

}  // namespace moka
