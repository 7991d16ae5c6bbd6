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

// ---------------------------------------------------------------------------- N1: metadata without mask tensors
// VT: straight from the token ids and labels the collator emits (VisualText/train/train.py:206-231, 259-318):
//   image    = ids == <image>                                  (:210)
//   text     = ids != <image>, inside the sample's length      (:211; the collator right-pads every mask with False)
//   question = !image & label == -100 & pos > last image pos   (:220-231)
// keys = exact ordered gather of the question positions (layer.py:635).  No mask tensors, no host sync.
__global__ void __launch_bounds__(256)
k_build_meta_ids(int S, int Lk, const long long* __restrict__ ids, const long long* __restrict__ labels,
                 long long image_id, long long ignore_index, const int32_t* __restrict__ lengths,
                 uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ int s_last_img, s_base;
  __shared__ int s_warp[8];
  if (tid == 0) { s_last_img = -1; s_base = 0; }
  __syncthreads();
  const int len = lengths ? min(max(lengths[b], 0), S) : S;
  int last = -1;
  for (int i = tid; i < len; i += 256)
    if (ids[(size_t)b * S + i] == image_id) last = i;          // (ascending per thread: the last hit is the largest)
  if (last >= 0) atomicMax(&s_last_img, last);
  __syncthreads();
  const int last_img = s_last_img;
  const int lane = tid & 31, warp = tid >> 5;
  for (int base = 0; base < S; base += 256) {
    const int i = base + tid;
    int q = 0;
    if (i < S) {
      const size_t g = (size_t)b * S + i;
      const bool in = i < len;
      const bool img = in && ids[g] == image_id;
      // train.py:216-229: without any image token the position mask is all False -> no question token
      q = (in && !img && labels[g] == ignore_index && last_img >= 0 && i > last_img) ? 1 : 0;
      const int code = !in ? MOKA_MOD_NONE : (img ? MOKA_MOD_VIS : MOKA_MOD_TEXT);
      tokmeta[g] = (uint8_t)(code | (q ? MOKA_KEY_VALID : 0));
    }
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
  const int n = s_base;
  for (int k = min(n, Lk) + tid; k < Lk; k += 256) kidx[(size_t)b * Lk + k] = -1;
  if (tid == 0) {
    klen[b] = min(n, Lk);
    if (n > Lk) atomicOr(&status[0], MOKA_META_KEY_OVERFLOW);
    atomicMax(&status[1], n);
  }
}

// AVT: from the per-sample SEGMENT TABLE that prepare_multimodal_inputs walks
// (AudioVisualText/models/unified_arch.py:139-240): segments in order, each (kind, length) with kind
// MOKA_SEG_TEXT (text = 1), MOKA_SEG_VIDEO (video / image features -> video mask, :197-225), MOKA_SEG_AUDIO (:186-195),
// MOKA_SEG_QUESTION (the text segment in front of <question_end>: text = 1 and question = 1, :153-166); special
// tokens are not part of any segment (:230).  Samples are LEFT-padded to S (:304-312): the first S - sum(length)
// positions are in no mask.  keys = contiguous span first..last question position (lora.py:489-491).
__global__ void __launch_bounds__(256)
k_build_meta_segments(int S, int Lk, int max_seg, const int32_t* __restrict__ segs, uint8_t* tokmeta, int32_t* kidx,
                      int32_t* klen, int32_t* status) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ int s_start[65], s_kind[64];
  __shared__ int s_nseg, s_first, s_last, s_err;
  if (tid == 0) {
    int tot = 0, n = 0, err = 0;
    for (int i = 0; i < max_seg && i < 64; ++i) {
      const int kind = segs[((size_t)b * max_seg + i) * 2], len = segs[((size_t)b * max_seg + i) * 2 + 1];
      if (kind == MOKA_SEG_END) break;
      if (kind < MOKA_SEG_TEXT || kind > MOKA_SEG_QUESTION || len < 0) { err |= MOKA_META_BAD_SEGMENT; break; }
      s_kind[n] = kind; s_start[n] = tot; tot += len; ++n;
    }
    if (tot > S) { err |= MOKA_META_BAD_SEGMENT; n = 0; tot = 0; }
    s_start[n] = tot;
    const int pad = S - tot;                                   // left padding
    for (int i = 0; i <= n; ++i) s_start[i] += pad;
    s_nseg = n; s_first = INT_MAX; s_last = -1; s_err = err;
  }
  __syncthreads();
  const int n = s_nseg;
  int first = INT_MAX, last = -1;
  for (int i = tid; i < S; i += 256) {
    int code = MOKA_MOD_NONE, q = 0;
    if (n > 0 && i >= s_start[0]) {
      int sg = 0;
      while (sg + 1 < n && i >= s_start[sg + 1]) ++sg;
      const int kind = s_kind[sg];
      code = (kind == MOKA_SEG_VIDEO) ? MOKA_MOD_VIS : (kind == MOKA_SEG_AUDIO ? MOKA_MOD_AUD : MOKA_MOD_TEXT);
      q = kind == MOKA_SEG_QUESTION;
    }
    tokmeta[(size_t)b * S + i] = (uint8_t)(code | (q ? MOKA_KEY_VALID : 0));
    if (q) { first = min(first, i); last = max(last, i); }
  }
  if (last >= 0) { atomicMin(&s_first, first); atomicMax(&s_last, last); }
  __syncthreads();
  int nk = 0;
  if (s_last < 0) {
    if (tid == 0) atomicOr(&s_err, MOKA_META_EMPTY_QUESTION);   // the reference raises IndexError (lora.py:490)
  } else {
    nk = s_last - s_first + 1;
    for (int k = tid; k < min(nk, Lk); k += 256) kidx[(size_t)b * Lk + k] = s_first + k;
  }
  for (int k = max(min(nk, Lk), 0) + tid; k < Lk; k += 256) kidx[(size_t)b * Lk + k] = -1;
  __syncthreads();
  if (tid == 0) {
    int e = s_err;
    if (nk > Lk) e |= MOKA_META_KEY_OVERFLOW;
    klen[b] = min(nk, Lk);
    if (e) atomicOr(&status[0], e);
    atomicMax(&status[1], nk);
  }
}

int launch_build_meta_ids(int B, int S, int Lk, const long long* ids, const long long* labels, long long image_id,
                          long long ignore_index, const int32_t* lengths, uint8_t* tokmeta, int32_t* kidx,
                          int32_t* klen, int32_t* status, cudaStream_t st) {
  k_build_meta_ids<<<B, 256, 0, st>>>(S, Lk, ids, labels, image_id, ignore_index, lengths, tokmeta, kidx, klen, status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
}

int launch_build_meta_segments(int B, int S, int Lk, int max_seg, const int32_t* segs, uint8_t* tokmeta, int32_t* kidx,
                               int32_t* klen, int32_t* status, cudaStream_t st) {
  k_build_meta_segments<<<B, 256, 0, st>>>(S, Lk, max_seg, segs, tokmeta, kidx, klen, status);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_cuda_error(e); return MOKA_ECUDA; }
  note_launch();
  return MOKA_OK;
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
