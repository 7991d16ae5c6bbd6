// extern "C" entry points of libmoka.so (see include/moka.h for the contract).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "moka_common.cuh"
#include "moka_kernels.h"

namespace moka {

static std::atomic<int64_t> g_launches{0};
static thread_local char g_errbuf[256] = "";

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_cuda_error(cudaError_t e) {
  snprintf(g_errbuf, sizeof(g_errbuf), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
}

static int num_sms() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

static int validate(const MokaDesc* m) {
  if (!m) return MOKA_EINVAL;
  if (m->B <= 0 || m->S <= 0 || m->d_in <= 0 || m->d_out <= 0) return MOKA_EINVAL;
  if ((m->d_in & 7) || (m->d_out & 7)) return MOKA_EINVAL;
  if (m->r < 1 || m->r > 256) return MOKA_EINVAL;
  if (m->M < 1 || m->M > 3) return MOKA_EINVAL;
  if (m->dtype != MOKA_BF16 && m->dtype != MOKA_F32) return MOKA_EINVAL;
  if (m->Lk < 0) return MOKA_EINVAL;
  if (!(m->dropout_p >= 0.f && m->dropout_p < 1.f)) return MOKA_EINVAL;
  if ((long long)m->B * m->S > 0x7fffffffLL / 256) return MOKA_EINVAL;
  return MOKA_OK;
}

static DevDesc make_dev(const MokaDesc* m, uint64_t seed, uint64_t offset) {
  DevDesc d;
  memset(&d, 0, sizeof(d));
  d.B = m->B; d.S = m->S; d.T = m->B * m->S;
  d.d_in = m->d_in; d.d_out = m->d_out; d.r = m->r; d.M = m->M; d.Lk = m->Lk;
  d.pre = m->pre;
  d.post[0] = 0.f; d.post[1] = m->post[0]; d.post[2] = m->post[1]; d.post[3] = m->post[2];
  d.attn_w = m->attn_w;
  d.sm_scale = 1.0f / sqrtf((float)m->r);
  if (m->dropout_p > 0.f) {
    long t = lrintf(m->dropout_p * 4096.0f);
    if (t < 1) t = 1;
    if (t > 4095) t = 4095;
    d.drop_thresh = (uint32_t)t;
    // the generator drops with probability t / 4096: scale by the probability actually realised, so that
    // E[u] = x exactly (moka_effective_dropout_p reports it)
    d.drop_scale = 1.0f / (1.0f - (float)t / 4096.0f);
    for (int i = 0; i < 12; ++i) d.drop_sel[i] = ((t >> i) & 1) ? 0xffffffffu : 0u;
  } else {
    d.drop_scale = 1.0f; d.drop_thresh = 0;
  }
  d.seed = seed; d.offset = offset;
  {
    // splitmix64 over (seed, offset) -> the two 32-bit keys of the dropout generator
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + offset + 0xD1B54A32D192ED03ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    d.key_lo = (uint32_t)z;
    d.key_odd = (uint32_t)(z >> 32) | 1u;
  }
  return d;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t ws_dh_bytes(const MokaDesc* m) { return align_up((size_t)m->B * m->S * m->r * 4, 256); }
static size_t ws_dk_bytes(const MokaDesc* m) { return align_up((size_t)m->B * (m->Lk > 0 ? m->Lk : 1) * m->r * 4, 256); }

static size_t ws_cnt_bytes(const MokaDesc* m) { return align_up((size_t)m->B * 4, 256); }
// number of [T, r] fp32 buffers at the head of the workspace: the partial dG buffers of the persistent k_dgrad_tc
// (one per CTA that shares a token tile; at least the two the other paths use)
static size_t ws_parts(const MokaDesc* m) {
  return (size_t)tc_dgrad_parts_max(m->B * m->S, m->d_out, num_sms());
}

// stages the tcgen05 path takes for this call (flags bits 8..11 veto individual stages)
static int tc_stages(const MokaDesc* m, const DevDesc& d) {
  if (m->flags & MOKA_FLAG_FORCE_SIMT) return 0;
  return tc_stage_mask(d, m->dtype) & ~((m->flags >> 8) & kStageAll);
}

}  // namespace moka

using namespace moka;

extern "C" {

size_t moka_workspace_bytes(const MokaDesc* desc) {
  if (validate(desc) != MOKA_OK) return 0;
  // the packed-operand region is shared by B^T [rk, d_out] (dgrad) and the dh operands (wgrad)
  size_t pack = tc_pack_bytes(desc->B * desc->S, desc->r, desc->M);
  const size_t bt = (size_t)((desc->r + 15) / 16 * 16) * desc->d_out * 2;
  if (bt > pack) pack = bt;
  // [dG partials (>= 2: dh | dh_att) | dK | counters | pack]; the forward uses the first 2 T r floats for the partial h
  return ws_parts(desc) * ws_dh_bytes(desc) + ws_dk_bytes(desc) + ws_cnt_bytes(desc) + align_up(pack, 256) + 256;
}

int moka_uses_tensor_cores(const MokaDesc* desc) {
  if (validate(desc) != MOKA_OK) return 0;
  DevDesc d = make_dev(desc, 0, 0);
  return tc_stages(desc, d) ? 1 : 0;
}

int moka_supports_delta(const MokaDesc* desc, int has_dx) {
  if (validate(desc) != MOKA_OK) return 0;
  DevDesc d = make_dev(desc, 0, 0);
  const int tcs = tc_stages(desc, d);
  if ((tcs & kStageAll) != kStageAll) return 0;
  return tc_backward_fused_ok(d, desc->Lk, has_dx != 0) ? 1 : 0;
}

int moka_forward(const MokaDesc* desc, const void* x, const void* y0, void* y,
                 const void* const* A, const void* Bw,
                 const uint8_t* tokmeta, const int32_t* kidx, const int32_t* klen,
                 float* h_stash, uint64_t seed, uint64_t offset,
                 void* workspace, size_t workspace_bytes, moka_stream_t stream) {
  if (int e = validate(desc)) return e;
  const bool delta = (desc->flags & MOKA_FLAG_DELTA_OUT) != 0;
  if (delta) y0 = y;                                           // write-only output: y0 is not read
  if (!x || !y0 || !y || !A || !Bw || !tokmeta || !klen || !h_stash) return MOKA_EINVAL;
  if (desc->Lk > 0 && !kidx) return MOKA_EINVAL;
  for (int m = 0; m < desc->M; ++m) if (!A[m]) return MOKA_EINVAL;
  if (!aligned16(x) || !aligned16(y0) || !aligned16(y) || !aligned16(Bw) || !aligned16(h_stash)) return MOKA_EALIGN;
  for (int m = 0; m < desc->M; ++m) if (!aligned16(A[m])) return MOKA_EALIGN;
  if (!workspace || !aligned16(workspace)) return workspace ? MOKA_EALIGN : MOKA_EWORKSPACE;
  if (workspace_bytes < moka_workspace_bytes(desc)) return MOKA_EWORKSPACE;
  DevDesc d = make_dev(desc, seed, offset);
  SimtArgs a;
  memset(&a, 0, sizeof(a));
  a.x = x; a.y0 = y0; a.y = y; a.Bw = Bw;
  for (int m = 0; m < desc->M; ++m) a.A[m] = A[m];
  a.tokmeta = tokmeta; a.kidx = kidx; a.klen = klen; a.h_stash = h_stash; a.klen_hint = desc->Lk;
  a.hpart = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tcs = tc_stages(desc, d) & (kStageDown | kStageUp);
  if ((desc->flags & MOKA_FLAG_FORCE_TC) && tcs != (kStageDown | kStageUp)) return MOKA_EUNSUPPORTED;
  const bool force = (desc->flags & MOKA_FLAG_FORCE_TC) != 0;
  const int skip = (desc->flags >> MOKA_FLAG_DBG_SKIP_SHIFT) & 0xff;     // timing aid only
  d.dbg_skip = skip;
  d.dbg = tc_dbg_enabled();
  d.delta = delta ? 1 : 0;
  if (delta && tcs != (kStageDown | kStageUp)) return MOKA_EUNSUPPORTED;   // write-only outputs: tcgen05 kernels only
  if (tcs == (kStageDown | kStageUp)) {
    // both stages on the tensor-core path: one call so that they can agree on the split-reduction layout
    int e = tc_forward(d, a, num_sms(), st, tcs);
    if (e != MOKA_EUNSUPPORTED || force || delta) return e;
    // a shape the tensor-core kernels cannot tile (shared-memory budget): redo stage by stage below
    // (the down-projection is idempotent)
  }
  a.hpart = nullptr;
  // stage order matters: down (h) before up (y)
  for (int stage : {kStageDown, kStageUp}) {
    int e = MOKA_EUNSUPPORTED;
    if (tcs & stage) e = tc_forward(d, a, num_sms(), st, stage);
    // a shape the tensor-core kernel cannot tile (shared-memory budget) takes the warp-FMA kernel
    if (e == MOKA_EUNSUPPORTED && !force) e = simt_forward(d, desc->dtype, a, num_sms(), st, stage);
    if (e) return e;
  }
  return MOKA_OK;
}

int moka_backward(const MokaDesc* desc, const void* dy, const void* x, const float* h_stash,
                  const void* const* A, const void* Bw,
                  const uint8_t* tokmeta, const int32_t* kidx, const int32_t* klen,
                  void* dx, float* const* dA, float* dB,
                  uint64_t seed, uint64_t offset,
                  void* workspace, size_t workspace_bytes, moka_stream_t stream) {
  if (int e = validate(desc)) return e;
  if (!dy || !x || !h_stash || !A || !Bw || !tokmeta || !klen || !dA || !dB || !workspace) return MOKA_EINVAL;
  if (desc->Lk > 0 && !kidx) return MOKA_EINVAL;
  for (int m = 0; m < desc->M; ++m) if (!A[m] || !dA[m]) return MOKA_EINVAL;
  if (!aligned16(dy) || !aligned16(x) || !aligned16(Bw) || !aligned16(h_stash) || !aligned16(dB) ||
      !aligned16(workspace) || (dx && !aligned16(dx))) return MOKA_EALIGN;
  for (int m = 0; m < desc->M; ++m) if (!aligned16(A[m]) || !aligned16(dA[m])) return MOKA_EALIGN;
  if (workspace_bytes < moka_workspace_bytes(desc)) return MOKA_EWORKSPACE;
  DevDesc d = make_dev(desc, seed, offset);
  SimtArgs a;
  memset(&a, 0, sizeof(a));
  a.dy = dy; a.x = x; a.dx = dx; a.Bw = Bw; a.dB = dB;
  for (int m = 0; m < desc->M; ++m) { a.A[m] = A[m]; a.dA[m] = dA[m]; }
  a.tokmeta = tokmeta; a.kidx = kidx; a.klen = klen; a.h_stash = const_cast<float*>(h_stash);
  a.klen_hint = desc->Lk;
  char* ws = static_cast<char*>(workspace);
  a.dh = reinterpret_cast<float*>(ws);
  a.dh_att = reinterpret_cast<float*>(ws + ws_dh_bytes(desc));
  const size_t parts_bytes = ws_parts(desc) * ws_dh_bytes(desc);
  a.dK = reinterpret_cast<float*>(ws + parts_bytes);
  a.counters = reinterpret_cast<int*>(ws + parts_bytes + ws_dk_bytes(desc));
  a.pack = ws + parts_bytes + ws_dk_bytes(desc) + ws_cnt_bytes(desc);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tcs = tc_stages(desc, d) & (kStageDgrad | kStageWgrad);
  if ((desc->flags & MOKA_FLAG_FORCE_TC) && tcs != (kStageDgrad | kStageWgrad)) return MOKA_EUNSUPPORTED;
  const bool force = (desc->flags & MOKA_FLAG_FORCE_TC) != 0;
  d.dbg_skip = (desc->flags >> MOKA_FLAG_DBG_SKIP_SHIFT) & 0xff;          // timing aid only
  d.dbg = tc_dbg_enabled();
  d.delta = ((desc->flags & MOKA_FLAG_DELTA_OUT) && dx) ? 1 : 0;
  if (d.delta && !(tcs == (kStageDgrad | kStageWgrad) && tc_backward_fused_ok(d, desc->Lk, true)))
    return MOKA_EUNSUPPORTED;                                  // write-only dx: fused tcgen05 chain only
  if (tcs == (kStageDgrad | kStageWgrad) && tc_backward_fused_ok(d, desc->Lk, dx != nullptr)) {
    // whole backward on the tensor-core path: four kernels chained by programmatic dependent launch,
    // the first of which zero-fills what needs it (no memset node in the chain)
    // (tc_backward_fused_ok has already checked that every stage can tile this shape)
    return tc_backward(d, a, num_sms(), st, tcs, true);
  }
  (void)force;
  // one memset covers dh (the tensor-core dgrad reduces into it), dh_att and dK
  cudaError_t ce = cudaMemsetAsync(a.dh, 0, parts_bytes + ws_dk_bytes(desc) + ws_cnt_bytes(desc), st);
  if (ce != cudaSuccess) { set_cuda_error(ce); return MOKA_ECUDA; }
  for (int stage : {kStageDgrad, kStageWgrad}) {
    int e = MOKA_EUNSUPPORTED;
    if (tcs & stage) e = tc_backward(d, a, num_sms(), st, stage, false);
    if (e == MOKA_EUNSUPPORTED && !force) e = simt_backward(d, desc->dtype, a, num_sms(), st, stage);
    if (e) return e;
  }
  return MOKA_OK;
}

int moka_build_token_meta(int32_t variant, int32_t B, int32_t S, int32_t Lk,
                          const void* const* masks, int32_t mask_elem_bytes,
                          uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status,
                          moka_stream_t stream) {
  if (variant != MOKA_VARIANT_AVT && variant != MOKA_VARIANT_VT) return MOKA_EINVAL;
  if (B <= 0 || S <= 0 || Lk <= 0 || !masks || !tokmeta || !kidx || !klen || !status) return MOKA_EINVAL;
  if (mask_elem_bytes != 1 && mask_elem_bytes != 4 && mask_elem_bytes != 8) return MOKA_EINVAL;
  if (!masks[0] || !masks[1] || !masks[3]) return MOKA_EINVAL;
  return launch_build_meta(variant, B, S, Lk, masks, mask_elem_bytes, tokmeta, kidx, klen, status,
                           static_cast<cudaStream_t>(stream));
}

int moka_build_token_meta_from_ids(int32_t B, int32_t S, int32_t Lk, const int64_t* input_ids, const int64_t* labels,
                                   int64_t image_token_id, int64_t ignore_index, const int32_t* lengths,
                                   uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status,
                                   moka_stream_t stream) {
  if (B <= 0 || S <= 0 || Lk <= 0 || !input_ids || !labels || !tokmeta || !kidx || !klen || !status) return MOKA_EINVAL;
  return launch_build_meta_ids(B, S, Lk, reinterpret_cast<const long long*>(input_ids),
                               reinterpret_cast<const long long*>(labels), (long long)image_token_id,
                               (long long)ignore_index, lengths, tokmeta, kidx, klen, status,
                               static_cast<cudaStream_t>(stream));
}

int moka_build_token_meta_from_segments(int32_t B, int32_t S, int32_t Lk, int32_t max_segments, const int32_t* segments,
                                        uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status,
                                        moka_stream_t stream) {
  if (B <= 0 || S <= 0 || Lk <= 0 || max_segments <= 0 || max_segments > 64 || !segments || !tokmeta || !kidx || !klen ||
      !status)
    return MOKA_EINVAL;
  return launch_build_meta_segments(B, S, Lk, max_segments, segments, tokmeta, kidx, klen, status,
                                    static_cast<cudaStream_t>(stream));
}

int moka_dropout_keep_mask(const MokaDesc* desc, uint64_t seed, uint64_t offset, uint8_t* keep,
                           moka_stream_t stream) {
  if (int e = validate(desc)) return e;
  if (!keep) return MOKA_EINVAL;
  DevDesc d = make_dev(desc, seed, offset);
  return launch_keep_mask(d, keep, static_cast<cudaStream_t>(stream));
}

float moka_effective_dropout_p(float p) {
  if (!(p > 0.f)) return 0.f;
  long t = lrintf(p * 4096.0f);
  if (t < 1) t = 1;
  if (t > 4095) t = 4095;
  return (float)t / 4096.0f;
}

int moka_adamw_step(int64_t n, float* master, float* exp_avg, float* exp_avg_sq, const float* grad,
                    void* param_out, int32_t dtype, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int32_t step, float grad_scale, moka_stream_t stream) {
  if (n < 0 || !master || !exp_avg || !exp_avg_sq || !grad || step < 1) return MOKA_EINVAL;
  if (dtype != MOKA_BF16 && dtype != MOKA_F32) return MOKA_EINVAL;
  return launch_adamw(n, master, exp_avg, exp_avg_sq, grad, param_out, dtype, lr, beta1, beta2, eps,
                      weight_decay, step, grad_scale, static_cast<cudaStream_t>(stream));
}

int moka_rope(const void* src, void* dst, int64_t tokens, int32_t S, int32_t H, int32_t D, float theta,
              int32_t backward, moka_stream_t stream) {
  if (!src || !dst || tokens <= 0 || S <= 0 || H <= 0 || D <= 0 || (D % 16) || !aligned16(src) || !aligned16(dst))
    return MOKA_EINVAL;
  return launch_rope(src, dst, tokens, S, H, D, theta, backward ? -1.f : 1.f, static_cast<cudaStream_t>(stream));
}

int moka_debug_read(unsigned long long* host, int32_t n) {
  if (!host || n <= 0) return MOKA_EINVAL;
  return debug_read(host, n);
}

int moka_debug_mark(int32_t slot, moka_stream_t stream) { return debug_mark(slot, static_cast<cudaStream_t>(stream)); }

int moka_debug_trap_buffer(unsigned long long* host_mapped) { return debug_set_trap_buffer(host_mapped); }

int64_t moka_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* moka_strerror(int code) {
  switch (code) {
    case MOKA_OK: return "ok";
    case MOKA_EINVAL: return "invalid descriptor or null pointer";
    case MOKA_EWORKSPACE: return "workspace too small (see moka_workspace_bytes)";
    case MOKA_ECUDA: return g_errbuf[0] ? g_errbuf : "CUDA launch failed";
    case MOKA_EUNSUPPORTED: return "requested kernel path not available for this shape";
    case MOKA_EALIGN: return "pointer not 16-byte aligned";
    default: return "unknown error";
  }
}

int moka_version(void) { return 100; }

}  // extern "C"
