// Rank-r cross-modal attention of the MokA layer, forward and backward, on smem-resident
// token rows.  One warp per query token; keys of the sample staged in shared memory
// (lane <-> key for the scores, lane <-> rank index for the weighted sums); softmax
// reductions by warp shuffles.  A generic slow path (keys read from the global stash)
// covers key lists that do not fit in shared memory.
//
// Semantics: AudioVisualText/peft_hyper/tuners/lora.py:480-521 and
// VisualText/modified_peft/tuners/lora/layer.py:627-653 (see oracle/moka_oracle.py,
// fused_forward / fused_backward, for the restatement this follows).
#pragma once
#include "moka_common.cuh"

namespace moka {

constexpr int kKeyBlocksMax = 8;   // fast path handles up to 256 keys per sample

struct AttnCtx {
  const uint8_t* tokmeta;   // global [T]
  const int32_t* kidx;      // global [B, Lk]
  const int32_t* klen;      // global [B]
  const float* h_stash;     // global [T, r] pre-attention projections
  float* dK_glob;           // global [B, Lk, r] (backward only)
  float* Ks;                // smem  [kcap, ks]
  float* dKs;               // smem  [kcap, ks] (backward only)
  int ks;                   // row stride of Ks (r + 1: conflict-free lane<->key reads)
  int kcap;                 // keys that fit in Ks
  float* lse_out;           // global [T] log2-sum-exp of the scaled scores (forward; may be null)
  const float* h_stash2 = nullptr;   // optional second partial of h (added to h_stash rows when loading keys)
};

// ------------------------------------------------------------------ fast path (smem keys)
template <bool BWD>
__device__ __forceinline__ void attn_token_fast(const DevDesc& d, const AttnCtx& c, int L,
                                                const float* hq, float* gq, float* dgq, int lane,
                                                float* lse = nullptr) {
  const int r = d.r;
  const int nb = (L + 31) >> 5;
  float p[kKeyBlocksMax];
  float mx = -INFINITY;
#pragma unroll
  for (int kb = 0; kb < kKeyBlocksMax; ++kb) {
    p[kb] = -INFINITY;
    if (kb < nb) {
      const int k = kb * 32 + lane;
      if (k < L) {
        const float* Kr = c.Ks + k * c.ks;
        float s = 0.f;
        for (int j = 0; j < r; ++j) s = fmaf(hq[j], Kr[j], s);
        p[kb] = s * d.sm_scale * kLog2e;
      }
      mx = fmaxf(mx, p[kb]);
    }
  }
  mx = warp_max(mx);
  float l = 0.f;
#pragma unroll
  for (int kb = 0; kb < kKeyBlocksMax; ++kb) {
    if (kb < nb) {
      const int k = kb * 32 + lane;
      p[kb] = (k < L) ? exp2f(p[kb] - mx) : 0.f;
      l += p[kb];
    } else {
      p[kb] = 0.f;
    }
  }
  l = warp_sum(l);
  const float inv = 1.f / l;
  if (!BWD && lse && lane == 0) *lse = mx + log2f(l);
#pragma unroll
  for (int kb = 0; kb < kKeyBlocksMax; ++kb) p[kb] *= inv;

  if (!BWD) {
    __syncwarp();                      // gq may alias hq: every lane has finished reading the row
    for (int jb = 0; jb < r; jb += 32) {
      const int jj = jb + lane;
      const bool act = jj < r;
      float a = 0.f;
#pragma unroll
      for (int kb = 0; kb < kKeyBlocksMax; ++kb) {
        if (kb < nb) {
          const int kend = min(32, L - kb * 32);
          for (int kk = 0; kk < kend; ++kk) {
            const float pk = __shfl_sync(0xffffffffu, p[kb], kk);
            if (act) a = fmaf(pk, c.Ks[(kb * 32 + kk) * c.ks + jj], a);
          }
        }
      }
      if (act) gq[jj] = hq[jj] + d.attn_w * a;
    }
    return;
  }

  // ---- backward: da = w * dg ; dp_k = da . K_k ; ds = p (dp - <p,dp>)
  float ds[kKeyBlocksMax];
  float Dsum = 0.f;
#pragma unroll
  for (int kb = 0; kb < kKeyBlocksMax; ++kb) {
    ds[kb] = 0.f;
    if (kb < nb) {
      const int k = kb * 32 + lane;
      if (k < L) {
        const float* Kr = c.Ks + k * c.ks;
        float s = 0.f;
        for (int j = 0; j < r; ++j) s = fmaf(dgq[j], Kr[j], s);
        ds[kb] = s * d.attn_w;          // dp_k
        Dsum = fmaf(p[kb], ds[kb], Dsum);
      }
    }
  }
  Dsum = warp_sum(Dsum);
#pragma unroll
  for (int kb = 0; kb < kKeyBlocksMax; ++kb) ds[kb] = p[kb] * (ds[kb] - Dsum);
  __syncwarp();
  for (int jb = 0; jb < r; jb += 32) {
    const int jj = jb + lane;
    const bool act = jj < r;
    const float dgv = act ? dgq[jj] : 0.f;
    const float hv = act ? hq[jj] : 0.f;
    const float dav = d.attn_w * dgv;
    float acc = 0.f;
#pragma unroll
    for (int kb = 0; kb < kKeyBlocksMax; ++kb) {
      if (kb < nb) {
        const int kend = min(32, L - kb * 32);
        for (int kk = 0; kk < kend; ++kk) {
          const float pk = __shfl_sync(0xffffffffu, p[kb], kk);
          const float dsk = __shfl_sync(0xffffffffu, ds[kb], kk);
          if (act) {
            const int off = (kb * 32 + kk) * c.ks + jj;
            acc = fmaf(dsk, c.Ks[off], acc);
            atomicAdd(&c.dKs[off], fmaf(pk, dav, d.sm_scale * dsk * hv));
          }
        }
      }
    }
    if (act) dgq[jj] = dgv + d.sm_scale * acc;      // dh of the query token, in place
  }
}

// ------------------------------------------------------------------ thread-per-token forward
// One thread owns one query token: h and the running weighted sum live in registers, the key
// rows are broadcast reads from shared memory (row stride = r), softmax is the online form with
// the rescale skipped unless the running max moves.  Used when r <= 32 and the keys fit in smem.
template <int R>
__device__ __forceinline__ void attn_thread_fwd(const DevDesc& d, const float* __restrict__ Ks, int L,
                                                float* row, float* lse) {
  const int r = d.r;
  float h[R], acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) { h[j] = (j < r) ? row[j] : 0.f; acc[j] = 0.f; }
  const float sl = d.sm_scale * kLog2e;
  float m = -INFINITY, l = 0.f;
  const uint32_t ks_addr = static_cast<uint32_t>(__cvta_generic_to_shared(Ks));
  constexpr int KB = 4;                         // keys per step: one rescale and four exps per step
  for (int k0 = 0; k0 < L; k0 += KB) {
    float kv[KB][R];
    float s[KB];
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const int k = min(k0 + u, L - 1);         // clamp: the duplicate of the last key is masked below
      if (r == R) {
#pragma unroll
        for (int j = 0; j < R; j += 4) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ks_addr + (uint32_t)(k * r + j) * 4u));
          kv[u][j] = v.x; kv[u][j + 1] = v.y; kv[u][j + 2] = v.z; kv[u][j + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < R; ++j) kv[u][j] = (j < r) ? Ks[k * r + j] : 0.f;
      }
    }
    float smax = -INFINITY;
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j = 0; j < R; j += 4) {
        s0 = fmaf(h[j], kv[u][j], s0);
        s1 = fmaf(h[j + 1], kv[u][j + 1], s1);
        s2 = fmaf(h[j + 2], kv[u][j + 2], s2);
        s3 = fmaf(h[j + 3], kv[u][j + 3], s3);
      }
      s[u] = (k0 + u < L) ? ((s0 + s1) + (s2 + s3)) * sl : -INFINITY;
      smax = fmaxf(smax, s[u]);
    }
    if (smax > m) {
      const float corr = exp2f(m - smax);
      l *= corr;
#pragma unroll
      for (int j = 0; j < R; ++j) acc[j] *= corr;
      m = smax;
    }
#pragma unroll
    for (int u = 0; u < KB; ++u) {
      const float e = exp2f(s[u] - m);          // exp2(-inf) = 0 for the masked tail
      l += e;
#pragma unroll
      for (int j = 0; j < R; ++j) acc[j] = fmaf(e, kv[u][j], acc[j]);
    }
  }
  const float sc = d.attn_w / l;
  if (lse) *lse = m + log2f(l);
#pragma unroll
  for (int j = 0; j < R; ++j)
    if (j < r) row[j] = fmaf(sc, acc[j], h[j]);
}

// ------------------------------------------------------------------ slow path (global keys)
// lanes <-> rank index (up to 8 per lane, r <= 256); one key at a time.
template <bool BWD>
__device__ __forceinline__ void attn_token_slow(const DevDesc& d, const AttnCtx& c, int b, int L,
                                                const float* hq, float* gq, float* dgq, int lane,
                                                float* lse = nullptr) {
  const int r = d.r;
  float hv[8], dav[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int jj = lane + 32 * i;
    hv[i] = jj < r ? hq[jj] : 0.f;
    dav[i] = (BWD && jj < r) ? d.attn_w * dgq[jj] : 0.f;
    acc[i] = 0.f;
  }
  const float sl = d.sm_scale * kLog2e;
  auto load_key = [&](int k, float (&kv)[8]) -> bool {
    const int pos = c.kidx[(size_t)b * d.Lk + k];
    const size_t t = (size_t)b * d.S + pos;
    const bool valid = (c.tokmeta[t] & MOKA_KEY_VALID) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int jj = lane + 32 * i;
      kv[i] = (valid && jj < r) ? (c.h_stash[t * r + jj] + (c.h_stash2 ? c.h_stash2[t * r + jj] : 0.f)) : 0.f;
    }
    return valid;
  };
  auto dot8 = [&](const float (&a)[8], const float (&kv)[8]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(a[i], kv[i], s);
    return warp_sum(s);
  };
  float kv[8];
  float mx = -INFINITY;
  for (int k = 0; k < L; ++k) { load_key(k, kv); mx = fmaxf(mx, dot8(hv, kv) * sl); }
  float l = 0.f, E = 0.f;
  for (int k = 0; k < L; ++k) {
    load_key(k, kv);
    const float e = exp2f(dot8(hv, kv) * sl - mx);
    l += e;
    if (!BWD) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(e, kv[i], acc[i]);
    } else {
      E = fmaf(e, dot8(dav, kv), E);
    }
  }
  const float inv = 1.f / l;
  if (!BWD) {
    if (lse && lane == 0) *lse = mx + log2f(l);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int jj = lane + 32 * i;
      if (jj < r) gq[jj] = hv[i] + d.attn_w * acc[i] * inv;
    }
    return;
  }
  const float Dsum = E * inv;
  for (int k = 0; k < L; ++k) {
    const bool valid = load_key(k, kv);
    const float pk = exp2f(dot8(hv, kv) * sl - mx) * inv;
    const float dsk = pk * (dot8(dav, kv) - Dsum);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int jj = lane + 32 * i;
      acc[i] = fmaf(dsk, kv[i], acc[i]);
      if (valid && jj < r)
        red_add_f32(&c.dK_glob[((size_t)b * d.Lk + k) * r + jj], fmaf(pk, dav[i], d.sm_scale * dsk * hv[i]));
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int jj = lane + 32 * i;
    if (jj < r) dgq[jj] = dgq[jj] + d.sm_scale * acc[i];
  }
}

// ------------------------------------------------------------------ per-tile driver
// A group of threads (the whole CTA, or the worker warps of a warp-specialised kernel) that
// synchronises on its own named barrier.
struct SyncGroup {
  int tid, nthr, bar_id;
  int* or_slot;                 // one int of shared memory for sync_or
  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, %1;" :: "r"(bar_id), "r"(nthr) : "memory");
  }
  __device__ __forceinline__ int sync_or(int pred) const {
    if (tid == 0) *or_slot = 0;
    sync();
    if (pred) atomicOr(or_slot, 1);
    sync();
    const int v = *or_slot;
    sync();
    return v;
  }
};

// hs : [nt, r] pre-attention rows (read-only here)
// gs : forward  -> receives g rows of query tokens (may alias hs: a row is only touched by its own warp)
// dgs: backward -> holds dg rows, overwritten with dh rows of query tokens
// meta_s: tile-local tokmeta bytes.  Must be called by every thread of the group.
template <bool BWD, int MAXR = 16>
__device__ __forceinline__ void attention_tile(const DevDesc& d, const AttnCtx& c, const SyncGroup& g, int t0, int nt,
                                               const uint8_t* meta_s, const float* hs, float* gs, float* dgs) {
  const int tid = g.tid, nthr = g.nthr;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int r = d.r;
  const int b_first = t0 / d.S, b_last = (t0 + nt - 1) / d.S;
  for (int b = b_first; b <= b_last; ++b) {
    const int lo = max(t0, b * d.S) - t0;
    const int hi = min(t0 + nt, (b + 1) * d.S) - t0;
    int q = 0;
    for (int i = lo + tid; i < hi; i += nthr) q |= ((meta_s[i] & MOKA_MOD_MASK) >= MOKA_MOD_VIS);
    if (!g.sync_or(q)) continue;
    const int L = c.klen[b];
    if (L <= 0) continue;
    const bool fast = (L <= c.kcap) && (L <= 32 * kKeyBlocksMax);
    // forward with a small rank: one thread per query token, keys packed with row stride r
    const bool per_thread = !BWD && fast && r <= MAXR;
    const int ks = per_thread ? r : c.ks;
    if (fast) {
      // one thread per key row: the index / validity lookups (two dependent global loads) happen
      // once per key, then the row is copied with independent loads
      for (int k = tid; k < L; k += nthr) {
        const int pos = c.kidx[(size_t)b * d.Lk + k];
        const size_t t = (size_t)b * d.S + pos;
        const bool valid = (c.tokmeta[t] & MOKA_KEY_VALID) != 0;
        const float* src = c.h_stash + t * r;
        const float* src2 = c.h_stash2 ? c.h_stash2 + t * r : nullptr;
        float* dst = c.Ks + k * ks;
        if ((r & 3) == 0 && (ks & 3) == 0) {
          for (int j = 0; j < r; j += 4) {
            float4 v = valid ? *reinterpret_cast<const float4*>(src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid && src2) {
              const float4 w = *reinterpret_cast<const float4*>(src2 + j);
              v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
            *reinterpret_cast<float4*>(dst + j) = v;
          }
        } else {
          for (int j = 0; j < r; ++j) dst[j] = valid ? (src[j] + (src2 ? src2[j] : 0.f)) : 0.f;
        }
        if (BWD) for (int j = 0; j < r; ++j) c.dKs[k * ks + j] = 0.f;
      }
      g.sync();
    }
    if (per_thread) {
      for (int tt = lo + tid; tt < hi; tt += nthr) {
        if ((meta_s[tt] & MOKA_MOD_MASK) < MOKA_MOD_VIS) continue;
        float* rowp = gs + (size_t)tt * r;
        float* lsep = c.lse_out ? c.lse_out + t0 + tt : nullptr;
        if (r <= 8) attn_thread_fwd<8>(d, c.Ks, L, rowp, lsep);
        else if (r <= 16) attn_thread_fwd<16>(d, c.Ks, L, rowp, lsep);
        else if constexpr (MAXR >= 32) attn_thread_fwd<32>(d, c.Ks, L, rowp, lsep);
      }
      g.sync();
      continue;
    }
    for (int tt = lo + warp; tt < hi; tt += nwarps) {
      if ((meta_s[tt] & MOKA_MOD_MASK) < MOKA_MOD_VIS) continue;
      const float* hq = hs + (size_t)tt * r;
      float* gq = BWD ? nullptr : gs + (size_t)tt * r;
      float* dgq = BWD ? dgs + (size_t)tt * r : nullptr;
      float* lsep = (!BWD && c.lse_out) ? c.lse_out + t0 + tt : nullptr;
      if (fast) attn_token_fast<BWD>(d, c, L, hq, gq, dgq, lane, lsep);
      else attn_token_slow<BWD>(d, c, b, L, hq, gq, dgq, lane, lsep);
    }
    if (fast) {
      g.sync();
      if (BWD) {
        for (int idx = tid; idx < L * r; idx += nthr) {
          const int k = idx / r, j = idx - k * r;
          const int pos = c.kidx[(size_t)b * d.Lk + k];
          const size_t t = (size_t)b * d.S + pos;
          if (c.tokmeta[t] & MOKA_KEY_VALID)
            red_add_f32(&c.dK_glob[((size_t)b * d.Lk + k) * r + j], c.dKs[k * c.ks + j]);
        }
        g.sync();
      }
    }
  }
}

}  // namespace moka
