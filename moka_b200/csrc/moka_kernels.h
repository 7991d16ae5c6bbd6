// Internal launch interfaces between the ABI layer and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "moka_common.cuh"

namespace moka {

struct SimtArgs {
  const void* x; const void* y0; void* y; const void* dy; void* dx;
  const void* A[3]; const void* Bw;
  float* dA[3]; float* dB;
  const uint8_t* tokmeta; const int32_t* kidx; const int32_t* klen;
  float* h_stash;
  float* hpart;     // forward workspace [2, T, r]: partial h of a split down-projection (null: none)
  int* counters;    // backward workspace [B]: per-sample arrival counters of k_attn_pack (zeroed by k_bwd_prep)
  float* dh;        // workspace [T, r]
  float* dh_att;    // workspace [T, r]: attention-backward contributions, added to dh by the wgrad stage
  float* dK;        // workspace [B, Lk, r]
  void* pack;       // workspace for the packed bf16 dh operands of k_wgrad_tc (tc_pack_bytes)
  int klen_hint;    // upper bound on klen (kidx row stride)
};

// number of partial dG buffers the persistent k_dgrad_tc may write for this shape (>= 2)
int tc_dgrad_parts_max(int T, int d_out, int num_sms);

// bytes of the packed-operand workspace the tensor-core wgrad kernel needs
size_t tc_pack_bytes(int T, int r, int M);

enum { kStageDown = 1, kStageUp = 2, kStageDgrad = 4, kStageWgrad = 8, kStageAll = 15 };

// `stages` selects which of the four streaming kernels run (the rest are done by the tcgen05 path)
int simt_forward(const DevDesc& d, int dtype, const SimtArgs& a, int num_sms, cudaStream_t st, int stages = kStageAll);
int simt_backward(const DevDesc& d, int dtype, const SimtArgs& a, int num_sms, cudaStream_t st, int stages = kStageAll);
int launch_keep_mask(const DevDesc& d, uint8_t* keep, cudaStream_t st);

int launch_build_meta(int variant, int B, int S, int Lk, const void* const* masks, int elem_bytes,
                      uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status, cudaStream_t st);

int launch_build_meta_ids(int B, int S, int Lk, const long long* ids, const long long* labels, long long image_id,
                          long long ignore_index, const int32_t* lengths, uint8_t* tokmeta, int32_t* kidx,
                          int32_t* klen, int32_t* status, cudaStream_t st);
int launch_build_meta_segments(int B, int S, int Lk, int max_seg, const int32_t* segs, uint8_t* tokmeta, int32_t* kidx,
                               int32_t* klen, int32_t* status, cudaStream_t st);

int launch_adamw(int64_t n, float* master, float* m, float* v, const float* grad, void* param_out, int dtype,
                 float lr, float b1, float b2, float eps, float wd, int step, float grad_scale, cudaStream_t st);

int launch_rope(const void* src, void* dst, long long BS, int S, int H, int D, float theta, float sign, cudaStream_t st);

int debug_read(unsigned long long* host, int n);
int debug_mark(int slot, cudaStream_t st);
int tc_dbg_enabled();
int debug_set_trap_buffer(unsigned long long* mapped);

// tensor-core (tcgen05 + TMA) path; returns MOKA_EUNSUPPORTED when the shape does not qualify
// Returns the mask of stages the tcgen05 path implements for this shape (0 = none).
int tc_stage_mask(const DevDesc& d, int dtype);
int tc_forward(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages);
int tc_backward(const DevDesc& d, const SimtArgs& a, int num_sms, cudaStream_t st, int stages, bool fused);
bool tc_backward_fused_ok(const DevDesc& d, int klen_hint, bool has_dx);

}  // namespace moka
