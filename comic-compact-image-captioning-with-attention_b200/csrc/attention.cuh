// attention.cuh -- fused multi-head additive-LN attention for one decoder step.
//
// Replaces the TF ops of MultiHeadAddLN.__call__ (common/ops_rnn.py:531-565) and
// the context part of MultiHeadAttentionWrapperV3.call (common/ops_rnn.py:
// 692-716, 741-744) -- ~12 memory-bound Eigen kernels over [N, 196, 512]
// intermediates in the reference -- with ONE kernel, one CTA per image:
//
//   phase 1  scores: a warp owns one feature-map position at a time, keeps the
//            (row-centred) key row in registers and scores it against the k beam
//            queries of the image: LN over R channels (two-pass variance on the
//            centred row), tanh, * v, per-head sums, / T.  The key row is read
//            from HBM once per step and shared by the k beams (the reference
//            tiles the keys k times).  A lane owns R/32 CONTIGUOUS channels, so a
//            head lives in D/(R/32) adjacent lanes (2 shuffles for 8 heads) and
//            the LN / head reductions of up to 4 beams are interleaved for ILP.
//   phase 2  softmax / signorm over the M positions of every (beam, head) in
//            shared memory (+ attention-map dropout), alignment-history write.
//   phase 3  context: ctx[beam, c] = sum_m alpha[beam, head(c), m] * values[m, c],
//            values re-read from L2 (the image's 400-650 KB tile was just
//            streamed by phase 1 when values == keys; otherwise one HBM pass).
//
// tanh(y) is evaluated as 1 - 2 / (exp2(2*log2(e)*y) + 1) with ex2.approx /
// rcp.approx (~1e-6 abs), the 2*log2(e) factor folded into gamma / beta, and
//   sum_j v_j tanh_j = sum_j v_j - 2 sum_j v_j r_j
// so the inner loop is FADD FFMA | FMUL FFMA EX2 FADD RCP FFMA per element.
// FAST (precision mode 2) uses the single-MUFU tanh.approx.f32 instead.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "gemm_f32.cuh"

namespace comic {

__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct AttnArgs {
  const float* keys;      // [B, M, R]
  const float* values;    // [B, M, VAL]
  const float* lq;        // [N, ld_lq], query at column q_off
  int ld_lq, q_off;
  const float* gamma;     // [R]   (add_LN)
  const float* beta;      // [R]
  const float* vvec;      // [R]
  const float* temperature;
  float* ctx_out;         // [N, ld_ctx]
  int ld_ctx;
  float* hist_t;          // [N, H*M] or nullptr
  const float* att_mask;  // [N, H*M] 0/1 or nullptr
  float att_keep;
  int k, M, VAL, prob_fn;
  const int* fin_count;   // decode loops: skip when every row finished in step t-1
  int t, n_rows;
};

constexpr int kAttnThreads = 512;
constexpr int kAttnBeamChunk = 4;

// Shared-memory carve-up shared by host and device.
struct AttnSmem {
  int tpc, nsplit, qfloats;
  size_t bytes;
};
__host__ __device__ inline AttnSmem attn_smem_layout(int k, int R, int H, int M, int VAL) {
  AttnSmem L;
  L.tpc = (VAL / 4 + 31) / 32 * 32;                  // phase 3: threads per position group
  L.nsplit = kAttnThreads / L.tpc;
  if (L.nsplit < 1) L.nsplit = 1;
  int red = (L.nsplit - 1) * kAttnBeamChunk * VAL;   // phase-3 partial sums alias the query block
  L.qfloats = k * R > red ? k * R : red;
  L.bytes = ((size_t)L.qfloats + (size_t)k * H * M) * sizeof(float);
  return L;
}

// scores of one position against KB beams (add_LN).  kc = centred key slice of this lane.
template <int CPL, int KB, bool FAST>
__device__ __forceinline__ void ln_tanh_scores(const float (&kc)[CPL], const float* __restrict__ qs, int R,
                                               int lane, const float (&gm)[CPL], const float (&bt)[CPL],
                                               const float (&vv)[CPL], float sv, float inv_R,
                                               float (&out)[KB]) {
  constexpr int G4 = CPL / 4;
  float ss[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 q = *reinterpret_cast<const float4*>(qs + (size_t)j * R + (g * 32 + lane) * 4);
      float d0 = kc[g * 4 + 0] + q.x, d1 = kc[g * 4 + 1] + q.y, d2 = kc[g * 4 + 2] + q.z, d3 = kc[g * 4 + 3] + q.w;
      acc = fmaf(d0, d0, acc); acc = fmaf(d1, d1, acc); acc = fmaf(d2, d2, acc); acc = fmaf(d3, d3, acc);
    }
    ss[j] = acc;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int j = 0; j < KB; ++j) ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
  }
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    const float rstd = rsqrtf(ss[j] * inv_R + 1e-12f);
    float acc = FAST ? 0.f : sv;
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 q = *reinterpret_cast<const float4*>(qs + (size_t)j * R + (g * 32 + lane) * 4);
      const float qa[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = g * 4 + e;
        float y = fmaf((kc[c] + qa[e]) * rstd, gm[c], bt[c]);
        if (FAST) acc = fmaf(tanh_approx(y), vv[c], acc);
        else acc = fmaf(rcp_approx(ex2_approx(y) + 1.0f), vv[c], acc);   // vv = -2 v, y pre-scaled by 2 log2 e
      }
    }
    out[j] = acc;
  }
}

// dynamic shared memory: q_c [k][R] (lane-permuted) | alpha [k][H][M]; phase-3 partials alias q_c
template <int R, int H, int MODE, bool FAST, int KB>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fused_kernel(const AttnArgs a) {
  if (a.fin_count != nullptr && a.t > 0 && a.fin_count[a.t - 1] >= a.n_rows) return;
  constexpr int CPL = R / 32;      // contiguous channels per lane
  constexpr int G4 = CPL / 4;      // float4 per lane
  constexpr int D = R / H;         // head width
  constexpr int LPH = (D >= CPL) ? D / CPL : 1;   // lanes per head
  static_assert(D % CPL == 0 || CPL % D == 0, "head width vs lane slice");
  static_assert(D >= CPL, "more than 32 heads per warp row is not built");
  constexpr float kTwoLog2e = 2.885390081777927f;
  extern __shared__ __align__(16) float sm[];
  const int k = a.k, M = a.M, VAL = a.VAL;
  const AttnSmem L = attn_smem_layout(k, R, H, M, VAL);
  float* sm_q = sm;                                 // [k][G4][32] float4: lane-permuted centred queries
  float* sm_s = sm + L.qfloats;                     // [k][H][M] scores -> alpha
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kAttnThreads / 32;
  const int c0 = lane * CPL;                        // first channel of this lane

  // ---- queries: load, centre (add_LN), store lane-permuted ----
  for (int beam = warp; beam < k; beam += NW) {
    const float* q = a.lq + (size_t)(b * k + beam) * a.ld_lq + a.q_off + c0;
    float4 v[G4];
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      v[g] = ldg4(q + g * 4);
      s += (v[g].x + v[g].y) + (v[g].z + v[g].w);
    }
    float mean = (MODE == 0) ? wsum(s) * (1.0f / R) : 0.f;
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 c = make_float4(v[g].x - mean, v[g].y - mean, v[g].z - mean, v[g].w - mean);
      *reinterpret_cast<float4*>(sm_q + (size_t)beam * R + (g * 32 + lane) * 4) = c;
    }
  }
  // per-lane constants (add_LN): gamma', beta' pre-scaled by 2 log2(e); vv = -2 v; sv = sum v
  float gm[CPL], bt[CPL], vv[CPL];
  float sv = 0.f;
  if (MODE == 0) {
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 g4 = ldg4(a.gamma + c0 + g * 4), b4 = ldg4(a.beta + c0 + g * 4), v4 = ldg4(a.vvec + c0 + g * 4);
      const float ga[4] = {g4.x, g4.y, g4.z, g4.w}, ba[4] = {b4.x, b4.y, b4.z, b4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        gm[g * 4 + e] = FAST ? ga[e] : ga[e] * kTwoLog2e;
        bt[g * 4 + e] = FAST ? ba[e] : ba[e] * kTwoLog2e;
        vv[g * 4 + e] = FAST ? va[e] : -2.0f * va[e];
        sv += va[e];
      }
    }
  }
  const float out_scale = (MODE == 0) ? (1.0f / a.temperature[0]) : (1.0f / sqrtf((float)D));
  __syncthreads();

  // ---- phase 1: scores ----
  const float* kbase = a.keys + (size_t)b * M * R + c0;
  for (int m = warp; m < M; m += NW) {
    float kc[CPL];
    {
      const float* kr = kbase + (size_t)m * R;
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        float4 t = ldg4(kr + g * 4);
        kc[g * 4 + 0] = t.x; kc[g * 4 + 1] = t.y; kc[g * 4 + 2] = t.z; kc[g * 4 + 3] = t.w;
        s += (t.x + t.y) + (t.z + t.w);
      }
      if (MODE == 0) {
        float mean = wsum(s) * (1.0f / R);
#pragma unroll
        for (int c = 0; c < CPL; ++c) kc[c] -= mean;
      }
    }
    // KB beams at a time; a short last chunk re-scores the final beams (results identical, stores idempotent)
    for (int beam0 = 0; beam0 < k; beam0 += KB) {
      const int bs = min(beam0, k - KB);               // chunk start, clamped so that bs + KB <= k
      float part[KB];
      const float* qs = sm_q + (size_t)bs * R;
      if (MODE == 0) {
        ln_tanh_scores<CPL, KB, FAST>(kc, qs, R, lane, gm, bt, vv, sv, 1.0f / R, part);
      } else {
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          float acc = 0.f;
#pragma unroll
          for (int g = 0; g < G4; ++g) {
            float4 q = *reinterpret_cast<const float4*>(qs + (size_t)j * R + (g * 32 + lane) * 4);
            acc = fmaf(kc[g * 4 + 0], q.x, acc); acc = fmaf(kc[g * 4 + 1], q.y, acc);
            acc = fmaf(kc[g * 4 + 2], q.z, acc); acc = fmaf(kc[g * 4 + 3], q.w, acc);
          }
          part[j] = acc;
        }
      }
      // head sums: a head occupies LPH adjacent lanes
#pragma unroll
      for (int o = LPH / 2; o; o >>= 1) {
#pragma unroll
        for (int j = 0; j < KB; ++j) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
      }
      if ((lane % LPH) == 0) {
        const int hh = lane / LPH;
#pragma unroll
        for (int j = 0; j < KB; ++j) sm_s[((size_t)(bs + j) * H + hh) * M + m] = part[j] * out_scale;
      }
    }
  }
  __syncthreads();

  // ---- phase 2: probability fn over M, dropout, history ----
  const int npair = k * H;
  for (int pr = warp; pr < npair; pr += NW) {
    float* s = sm_s + (size_t)pr * M;
    float sum = 0.f;
    if (a.prob_fn == 0) {
      float mx = -INFINITY;
      for (int m = lane; m < M; m += 32) mx = fmaxf(mx, s[m]);
      mx = wmax(mx);
      for (int m = lane; m < M; m += 32) {
        float e = expf(s[m] - mx);
        s[m] = e;
        sum += e;
      }
    } else {
      for (int m = lane; m < M; m += 32) {
        float e = 1.0f / (1.0f + expf(-s[m]));
        s[m] = e;
        sum += e;
      }
    }
    sum = wsum(sum);
    const size_t grow = ((size_t)b * npair + pr) * M;
    const float* mk = a.att_mask ? a.att_mask + grow : nullptr;
    for (int m = lane; m < M; m += 32) {
      float al = s[m] / sum;
      if (mk) al = (al / a.att_keep) * mk[m];
      s[m] = al;
      if (a.hist_t) a.hist_t[grow + m] = al;
    }
  }
  __syncthreads();

  // ---- phase 3: context ----
  // thread -> 4 value channels; the CTA's threads are split into `nsplit` position groups;
  // partial sums are combined through shared memory (aliases the query block).
  const int tpc = L.tpc, nsplit = L.nsplit;
  const int grp = tid / tpc, ct = tid - grp * tpc;
  const int c = ct * 4;
  const bool active = grp < nsplit && c < VAL;
  const int dv = VAL / H;
  const int hd = active ? c / dv : 0;
  const float* vb = a.values + (size_t)b * M * VAL + c;
  const int m_lo = (int)(((long long)M * grp) / nsplit), m_hi = (int)(((long long)M * (grp + 1)) / nsplit);
  float* red = sm_q;                                  // [nsplit-1][4][VAL]
  for (int beam0 = 0; beam0 < k; beam0 += kAttnBeamChunk) {
    const int nb = min(kAttnBeamChunk, k - beam0);
    float4 acc[kAttnBeamChunk];
#pragma unroll
    for (int j = 0; j < kAttnBeamChunk; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const float* a0 = sm_s + ((size_t)beam0 * H + hd) * M;
#pragma unroll 4
      for (int m = m_lo; m < m_hi; ++m) {
        float4 v = ldg4(vb + (size_t)m * VAL);
#pragma unroll
        for (int j = 0; j < kAttnBeamChunk; ++j) {
          if (j < nb) {
            float al = a0[(size_t)j * H * M + m];
            acc[j].x = fmaf(al, v.x, acc[j].x); acc[j].y = fmaf(al, v.y, acc[j].y);
            acc[j].z = fmaf(al, v.z, acc[j].z); acc[j].w = fmaf(al, v.w, acc[j].w);
          }
        }
      }
    }
    if (nsplit > 1) {
      __syncthreads();                                // query block no longer needed / previous chunk consumed
      if (active && grp > 0) {
#pragma unroll
        for (int j = 0; j < kAttnBeamChunk; ++j)
          *reinterpret_cast<float4*>(red + ((size_t)(grp - 1) * kAttnBeamChunk + j) * VAL + c) = acc[j];
      }
      __syncthreads();
      if (active && grp == 0) {
        for (int g2 = 1; g2 < nsplit; ++g2) {
#pragma unroll
          for (int j = 0; j < kAttnBeamChunk; ++j) {
            float4 p = *reinterpret_cast<const float4*>(red + ((size_t)(g2 - 1) * kAttnBeamChunk + j) * VAL + c);
            acc[j].x += p.x; acc[j].y += p.y; acc[j].z += p.z; acc[j].w += p.w;
          }
        }
      }
    }
    if (active && grp == 0) {
#pragma unroll
      for (int j = 0; j < kAttnBeamChunk; ++j)
        if (j < nb)
          *reinterpret_cast<float4*>(a.ctx_out + (size_t)(b * k + beam0 + j) * a.ld_ctx + c) = acc[j];
    }
  }
}

inline size_t attn_fused_smem(int k, int R, int H, int M, int VAL) { return attn_smem_layout(k, R, H, M, VAL).bytes; }

}  // namespace comic
