/*
 * moka.h -- C ABI of libmoka.so, the Blackwell (sm_100a) MokA adapter-layer library.
 *
 * The reference has no FFI on this path: the layer is a Python nn.Module swap
 * (AudioVisualText/peft_hyper/tuners/lora.py:125-188, VisualText/modified_peft/tuners/lora/
 * model.py:173-228).  These entry points are what a maintainer binds with ctypes from the
 * replaced ``Linear.forward``; each one cites the reference code it
 * replaces.
 *
 * Conventions: every function returns 0 (MOKA_OK) or a negative MOKA_E* code, never throws,
 * never allocates device memory, never synchronises the device; all work is ordered on the
 * caller's ``stream``; every pointer is caller-owned DEVICE memory unless stated otherwise;
 * the library keeps no per-call state and is thread-safe.
 */
#ifndef MOKA_H_
#define MOKA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* opaque CUDA stream handle (cudaStream_t) */
typedef void* moka_stream_t;

enum {
  MOKA_OK = 0,
  MOKA_EINVAL = -1,        /* bad descriptor / null pointer / unsupported shape */
  MOKA_EWORKSPACE = -2,    /* workspace too small (see moka_workspace_bytes) */
  MOKA_ECUDA = -3,         /* a CUDA launch failed; moka_strerror gives the CUDA text */
  MOKA_EUNSUPPORTED = -4,  /* requested path (e.g. forced tensor-core path) not available for this shape */
  MOKA_EALIGN = -5         /* pointer or leading dimension not 16-byte aligned */
};

enum { MOKA_BF16 = 0, MOKA_F32 = 1 };

/* key-list semantics used by moka_build_token_meta */
enum {
  MOKA_VARIANT_AVT = 0,    /* contiguous first..last question span, zero rows inside (lora.py:489-491) */
  MOKA_VARIANT_VT = 1      /* exact gather of question positions (layer.py:635-644) */
};

/* token metadata byte: low 3 bits modality code, bit 3 = usable key row */
enum {
  MOKA_MOD_NONE = 0, MOKA_MOD_TEXT = 1, MOKA_MOD_VIS = 2, MOKA_MOD_AUD = 3,
  MOKA_MOD_MASK = 0x7, MOKA_KEY_VALID = 0x8
};

/* MokaDesc.flags */
enum {
  MOKA_FLAG_FORCE_SIMT = 1,   /* use the warp-FMA kernels even when the tcgen05 path applies */
  MOKA_FLAG_FORCE_TC = 2,     /* fail with MOKA_EUNSUPPORTED instead of falling back to warp-FMA */
  /* Frozen-GEMM epilogue fusion: the outputs are WRITE-ONLY.  moka_forward stores the adapter
   * term alone into y (y0 is ignored and may be NULL; tokens in no mask get zeros), moka_backward stores the
   * adapter's input gradient alone into dx (previous contents ignored).  The caller then runs the frozen GEMM
   * with beta = 1 onto that buffer (y.addmm_(x, W0^T), dx.addmm_(dy, W0)), so the residual read happens in the
   * GEMM's epilogue instead of in these kernels.  Only on the tcgen05 path: MOKA_EUNSUPPORTED otherwise
   * (ask moka_supports_delta first). */
  MOKA_FLAG_DELTA_OUT = 4,
  /* per-stage vetoes of the tcgen05 path (debugging / A-B measurements) */
  MOKA_FLAG_NO_TC_DOWN = 1 << 8, MOKA_FLAG_NO_TC_UP = 1 << 9,
  MOKA_FLAG_NO_TC_DGRAD = 1 << 10, MOKA_FLAG_NO_TC_WGRAD = 1 << 11,
  /* timing aid (results are NOT valid): bits 16.. skip individual kernels of the tensor-core path:
     +0 down, +1 up, +2 backward prep, +3 dgrad, +4 attention backward / pack, +5 wgrad,
     +6 forward attention */
  MOKA_FLAG_DBG_SKIP_SHIFT = 16
};

/* status words written by moka_build_token_meta (int32 status[4]) */
enum {
  MOKA_META_EMPTY_QUESTION = 1,   /* AVT: a sample has no question token (reference: IndexError, lora.py:490) */
  MOKA_META_OVERLAP = 2,          /* a token is in more than one modality mask */
  MOKA_META_KEY_OVERFLOW = 4,     /* a sample has more key rows than Lk */
  MOKA_META_BAD_SEGMENT = 8       /* moka_build_token_meta_from_segments: unknown kind / negative length / sum > S */
};

/* segment kinds of moka_build_token_meta_from_segments */
enum {
  MOKA_SEG_END = 0,               /* terminates a sample's segment list */
  MOKA_SEG_TEXT = 1,              /* text tokens (unified_arch.py:168-180) */
  MOKA_SEG_VIDEO = 2,             /* <video> / <image> features -> video mask (:197-225) */
  MOKA_SEG_AUDIO = 3,             /* <audio> features (:186-195) */
  MOKA_SEG_QUESTION = 4           /* the text segment in front of <question_end>: text AND question (:153-166) */
};

typedef struct MokaDesc {
  int32_t B;            /* samples */
  int32_t S;            /* tokens per sample; x is [B*S, d_in] row-major, contiguous */
  int32_t d_in;         /* multiple of 8 */
  int32_t d_out;        /* multiple of 8 */
  int32_t r;            /* adapter rank, 1..256, the same for every modality (lora.py:525 needs it) */
  int32_t M;            /* number of A matrices: 2 (VT) or 3 (AVT); 1 = plain LoRA */
  int32_t dtype;        /* MOKA_BF16 / MOKA_F32: storage type of x, y0, y, dy, dx, A_m, B */
  int32_t Lk;           /* row stride (capacity) of kidx */
  float pre;            /* scale on A_m x           AVT alpha/r (lora.py:477)   VT 1 */
  float post[3];        /* scale after B, per code  AVT 1                       VT alpha/r (layer.py:662,668) */
  float attn_w;         /* AVT blc_weight (lora.py:499), VT attn_weight (layer.py:653) */
  float dropout_p;      /* lora_dropout; 0 disables (lora.py:267-270, layer.py:102-105) */
  int32_t flags;        /* MOKA_FLAG_* */
} MokaDesc;

/* Bytes of scratch moka_forward / moka_backward of this shape need (16-byte aligned base; the
 * same buffer may serve both calls and every layer of the same shape on one stream). */
size_t moka_workspace_bytes(const MokaDesc* desc);

/*
 * Forward of the adapter: y = y0 + post_m * B (h + w * attn(h, K_b)),  h = pre * A_m drop(x).
 * Replaces lora.py:460-532 (AVT train / :385-457 prefill) and layer.py:589-671 (VT); with
 * every token MOKA_MOD_TEXT and klen == 0 it is the plain-LoRA decode branch
 * (lora.py:373-381, layer.py:672-678).  ``y0`` is the frozen base output W0 x (+bias) and may
 * alias ``y`` (in-place residual add, like layer.py:663).  ``A`` is a HOST array of M device
 * pointers to [r, d_in] matrices, ``Bw`` is [d_out, r].  ``h_stash`` is fp32 [2*B*S*r + B*S]: it
 * receives the pre-attention projections h [B*S, r], the post-attention rows g [B*S, r] and the
 * per-token log-sum-exp of the attention scores [B*S] that moka_backward re-uses.
 */
int moka_forward(const MokaDesc* desc, const void* x, const void* y0, void* y,
                 const void* const* A, const void* Bw,
                 const uint8_t* tokmeta, const int32_t* kidx, const int32_t* klen,
                 float* h_stash, uint64_t seed, uint64_t offset,
                 void* workspace, size_t workspace_bytes, moka_stream_t stream);

/*
 * Backward of the adapter (what autograd derives from the same reference lines):
 *   dx  += pre * drop^T(A_m^T dh)       (dx already holds the frozen-base term W0^T dy)
 *   dA[m] += pre * sum_t dh_t (x) u_t    fp32 [r, d_in], accumulated (HOST array of M device pointers)
 *   dB  += sum_t post_m dy_t (x) g_t     fp32 [d_out, r], accumulated
 * ``dx`` may be NULL when the input needs no gradient.
 */
int moka_backward(const MokaDesc* desc, const void* dy, const void* x, const float* h_stash,
                  const void* const* A, const void* Bw,
                  const uint8_t* tokmeta, const int32_t* kidx, const int32_t* klen,
                  void* dx, float* const* dA, float* dB,
                  uint64_t seed, uint64_t offset,
                  void* workspace, size_t workspace_bytes, moka_stream_t stream);

/*
 * Masks -> token metadata, once per step, on the device, no host sync.  Replaces the
 * per-call ``torch.where`` / ``nonzero`` / ``.any()`` of lora.py:489 and layer.py:603-637 and
 * consumes the mask tensors exactly as their producers emit them
 * (AudioVisualText/models/unified_arch.py:159-204,336; VisualText/train/train.py:206-231).
 * ``masks`` is a HOST array {text, vis, aud, question} of device pointers to [B*S]
 * elements of ``mask_elem_bytes`` (1 bool/uint8, 4 int32, 8 int64) each; aud may be NULL.
 * Outputs: tokmeta uint8 [B*S], kidx int32 [B, Lk], klen int32 [B], status int32 [4] =
 * {MOKA_META_* bits, max klen, 0, 0} (status must be zeroed by the caller).
 */
int moka_build_token_meta(int32_t variant, int32_t B, int32_t S, int32_t Lk,
                          const void* const* masks, int32_t mask_elem_bytes,
                          uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status,
                          moka_stream_t stream);

/*
 * SURVEY.md 8(f) N1 -- the metadata straight from what the data pipeline holds, with no mask tensors at all.
 *
 * VT (VisualText/train/train.py:206-231 + the collator :259-318): ``input_ids`` / ``labels`` int64 [B, S] as the
 * collator pads them (image positions still carry ``image_token_id``; the dataset zeroes them only after it has
 * derived its masks, :213), ``lengths`` int32 [B] = unpadded sample lengths (NULL: every sample is S long):
 *   image = ids == image_token_id;  text = the other tokens inside the length;
 *   question = !image && label == ignore_index && pos > last image position   (none when a sample has no image).
 */
int moka_build_token_meta_from_ids(int32_t B, int32_t S, int32_t Lk, const int64_t* input_ids, const int64_t* labels,
                                   int64_t image_token_id, int64_t ignore_index, const int32_t* lengths,
                                   uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status,
                                   moka_stream_t stream);

/*
 * AVT (AudioVisualText/models/unified_arch.py:139-240, 304-312): ``segments`` int32 [B, max_segments, 2] =
 * (MOKA_SEG_* kind, length) in sequence order, MOKA_SEG_END-terminated; samples are LEFT-padded to S like the
 * reference pads them.  Same outputs and status word as moka_build_token_meta.
 */
int moka_build_token_meta_from_segments(int32_t B, int32_t S, int32_t Lk, int32_t max_segments, const int32_t* segments,
                                        uint8_t* tokmeta, int32_t* kidx, int32_t* klen, int32_t* status,
                                        moka_stream_t stream);

/* Test hook: the dropout keep mask (1 = kept) moka_forward/backward use for (seed, offset),
 * uint8 [B*S, d_in].  Lets the oracle be run on the very same mask. */
int moka_dropout_keep_mask(const MokaDesc* desc, uint64_t seed, uint64_t offset,
                           uint8_t* keep, moka_stream_t stream);

/* The drop probability the kernels realise for lora_dropout = p: round(4096 p) / 4096 (clamped to
 * [1, 4095] / 4096), e.g. 0.05 -> 205/4096 = 0.050049.  Kept elements are scaled by 1 / (1 - that value), so the
 * expectation of the dropped input is exact (nn.Dropout semantics, lora.py:267-270, at a 12-bit quantised rate). */
float moka_effective_dropout_p(float p);

/*
 * Sharded AdamW on the flat adapter parameter vector (ZeRO-1/2 optimizer shard; the
 * reference delegates this to DeepSpeed: VisualText/zero_stage2_config.json:23-34).
 * master/m/v are fp32 [n]; grad is fp32 [n] (already averaged over ranks);
 * param_out (bf16 or fp32 per ``dtype``) receives the rounded working copy.
 */
int moka_adamw_step(int64_t n, float* master, float* exp_avg, float* exp_avg_sq,
                    const float* grad, void* param_out, int32_t dtype,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    int32_t step, float grad_scale, moka_stream_t stream);

/*
 * Caller-side helper (not part of the adapter path): rotary position embedding of a bf16
 * [tokens = B*S, H, D] tensor (dst may alias src), HF-Llama half-split convention; backward != 0
 * applies the inverse rotation (the gradient).  Used by the frozen Llama wrapper of the benchmark.
 */
int moka_rope(const void* src, void* dst, int64_t tokens, int32_t S, int32_t H, int32_t D, float theta,
              int32_t backward, moka_stream_t stream);

/* Tuning aid: copies the in-kernel cycle stamps recorded when MOKA_DEBUG_TIMING=1 (host buffer). */
int moka_debug_read(unsigned long long* host, int32_t n);

/* Tuning aid: enqueues a one-thread kernel on `stream` that writes the GPU's globaltimer into slot `slot` (0..127) of
 * the last row of the debug stamp buffer - a stream-ordered time marker next to the kernels' own stamps. */
int moka_debug_mark(int32_t slot, moka_stream_t stream);

/* Tuning / post-mortem aid: registers two words of pinned, device-mapped host memory; a bounded in-kernel wait
 * that times out writes (source line << 32 | block << 12 | thread) there before it traps. */
int moka_debug_trap_buffer(unsigned long long* host_mapped);

/* Number of kernels this library has launched in this process (monotonic). */
int64_t moka_launch_count(void);

/* 1 when the tcgen05/TMA path would be taken for this descriptor, else 0. */
int moka_uses_tensor_cores(const MokaDesc* desc);

/* 1 when moka_forward AND moka_backward (with a dx buffer iff has_dx) accept MOKA_FLAG_DELTA_OUT for this
 * descriptor, else 0. */
int moka_supports_delta(const MokaDesc* desc, int has_dx);

const char* moka_strerror(int code);
int moka_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MOKA_H_ */
