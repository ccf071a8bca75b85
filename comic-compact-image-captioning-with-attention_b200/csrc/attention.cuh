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
//            tiles the keys k times).
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

constexpr int kAttnThreads = 256;

// dynamic shared memory: q_c [k][R] | alpha [k][H][M] | (phase 3 partials alias q_c.. when split)
template <int R, int H, int MODE, bool FAST>
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fused_kernel(const AttnArgs a) {
  if (a.fin_count != nullptr && a.t > 0 && a.fin_count[a.t - 1] >= a.n_rows) return;
  constexpr int G = R / 128;   // float4 groups per lane
  constexpr int D = R / H;     // head width
  constexpr float kTwoLog2e = 2.885390081777927f;
  extern __shared__ __align__(16) float sm[];
  const int k = a.k, M = a.M;
  // phase-3 geometry (also fixes the shared-memory carve-up)
  const int VAL = a.VAL;
  const int tpc = (VAL / 4 + 31) / 32 * 32;          // threads per position group
  const int nsplit = (kAttnThreads / tpc) > 0 ? (kAttnThreads / tpc) : 1;
  const int qfloats = max(k * R, (nsplit - 1) * 4 * VAL);
  float* sm_q = sm;                                 // [k][R] centred queries (phase 3: partial sums)
  float* sm_s = sm + qfloats;                       // [k][H][M] scores -> alpha
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kAttnThreads / 32;

  // ---- queries: load, centre (add_LN) ----
  for (int beam = warp; beam < k; beam += NW) {
    const float* q = a.lq + (size_t)(b * k + beam) * a.ld_lq + a.q_off;
    float4 v[G];
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      v[g] = ldg4(q + g * 128 + lane * 4);
      s += (v[g].x + v[g].y) + (v[g].z + v[g].w);
    }
    float mean = (MODE == 0) ? wsum(s) * (1.0f / R) : 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float4 c = make_float4(v[g].x - mean, v[g].y - mean, v[g].z - mean, v[g].w - mean);
      *reinterpret_cast<float4*>(sm_q + (size_t)beam * R + g * 128 + lane * 4) = c;
    }
  }
  // per-lane constants
  float4 g4[G], b4[G], v4[G];
  float sv[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    sv[g] = 0.f;
    if (MODE == 0) {
      int c = g * 128 + lane * 4;
      g4[g] = ldg4(a.gamma + c);
      b4[g] = ldg4(a.beta + c);
      v4[g] = ldg4(a.vvec + c);
      if (!FAST) {
        g4[g].x *= kTwoLog2e; g4[g].y *= kTwoLog2e; g4[g].z *= kTwoLog2e; g4[g].w *= kTwoLog2e;
        b4[g].x *= kTwoLog2e; b4[g].y *= kTwoLog2e; b4[g].z *= kTwoLog2e; b4[g].w *= kTwoLog2e;
        sv[g] = (v4[g].x + v4[g].y) + (v4[g].z + v4[g].w);
        v4[g].x *= -2.0f; v4[g].y *= -2.0f; v4[g].z *= -2.0f; v4[g].w *= -2.0f;
      }
    }
  }
  const float out_scale = (MODE == 0) ? (1.0f / a.temperature[0]) : (1.0f / sqrtf((float)D));
  __syncthreads();

  // ---- phase 1: scores ----
  const float* kbase = a.keys + (size_t)b * M * R;
  for (int m = warp; m < M; m += NW) {
    float4 key[G];
    const float* kr = kbase + (size_t)m * R;
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      key[g] = ldg4(kr + g * 128 + lane * 4);
      s += (key[g].x + key[g].y) + (key[g].z + key[g].w);
    }
    if (MODE == 0) {
      float mean = wsum(s) * (1.0f / R);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        key[g].x -= mean; key[g].y -= mean; key[g].z -= mean; key[g].w -= mean;
      }
    }
    for (int beam = 0; beam < k; ++beam) {
      const float* q = sm_q + (size_t)beam * R;
      float part[G];
      if (MODE == 0) {
        float4 d[G];
        float ss = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float4 qq = *reinterpret_cast<const float4*>(q + g * 128 + lane * 4);
          d[g].x = key[g].x + qq.x; d[g].y = key[g].y + qq.y;
          d[g].z = key[g].z + qq.z; d[g].w = key[g].w + qq.w;
          ss = fmaf(d[g].x, d[g].x, ss); ss = fmaf(d[g].y, d[g].y, ss);
          ss = fmaf(d[g].z, d[g].z, ss); ss = fmaf(d[g].w, d[g].w, ss);
        }
        float var = wsum(ss) * (1.0f / R);
        float rstd = rsqrtf(var + 1e-12f);
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float yx = fmaf(d[g].x * rstd, g4[g].x, b4[g].x);
          float yy = fmaf(d[g].y * rstd, g4[g].y, b4[g].y);
          float yz = fmaf(d[g].z * rstd, g4[g].z, b4[g].z);
          float yw = fmaf(d[g].w * rstd, g4[g].w, b4[g].w);
          if (FAST) {
            part[g] = (tanh_approx(yx) * v4[g].x + tanh_approx(yy) * v4[g].y) +
                      (tanh_approx(yz) * v4[g].z + tanh_approx(yw) * v4[g].w);
          } else {
            // tanh = 1 - 2r, r = 1/(2^y' + 1);  sum v*tanh = sum v + sum (-2v)*r   (v4 holds -2v, sv the sum)
            float rx = rcp_approx(ex2_approx(yx) + 1.0f);
            float ry = rcp_approx(ex2_approx(yy) + 1.0f);
            float rz = rcp_approx(ex2_approx(yz) + 1.0f);
            float rw = rcp_approx(ex2_approx(yw) + 1.0f);
            part[g] = sv[g] + ((rx * v4[g].x + ry * v4[g].y) + (rz * v4[g].z + rw * v4[g].w));
          }
        }
      } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float4 qq = *reinterpret_cast<const float4*>(q + g * 128 + lane * 4);
          part[g] = (key[g].x * qq.x + key[g].y * qq.y) + (key[g].z * qq.z + key[g].w * qq.w);
        }
      }
      float* srow = sm_s + (size_t)beam * H * M + m;
      if (D >= 128) {
        float hs[H];
#pragma unroll
        for (int hh = 0; hh < H; ++hh) hs[hh] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) hs[(g * 128) / D] += wsum(part[g]);
        if (lane == 0) {
#pragma unroll
          for (int hh = 0; hh < H; ++hh) srow[(size_t)hh * M] = hs[hh] * out_scale;
        }
      } else {
        constexpr int LPH = D / 4;   // lanes per head inside a 128-channel group
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float v = part[g];
#pragma unroll
          for (int o = LPH / 2; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if ((lane % LPH) == 0) srow[(size_t)((g * 128 + lane * 4) / D) * M] = v * out_scale;
        }
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
  // thread -> 4 value channels; the CTA's 256 threads are split into `nsplit` position
  // groups when VAL/4 <= 128; partial sums are combined through shared memory (sm_q).
  const int grp = tid / tpc, ct = tid - grp * tpc;
  const int c = ct * 4;
  const bool active = grp < nsplit && c < VAL;
  const int dv = VAL / H;
  const int hd = active ? c / dv : 0;
  const float* vb = a.values + (size_t)b * M * VAL + c;
  const int m_lo = (int)(((long long)M * grp) / nsplit), m_hi = (int)(((long long)M * (grp + 1)) / nsplit);
  float* red = sm_q;                                  // [nsplit-1][4][VAL] scratch (k*R >= needed, checked on host)
  for (int beam0 = 0; beam0 < k; beam0 += 4) {
    const int nb = min(4, k - beam0);
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const float* a0 = sm_s + ((size_t)beam0 * H + hd) * M;
#pragma unroll 4
      for (int m = m_lo; m < m_hi; ++m) {
        float4 v = ldg4(vb + (size_t)m * VAL);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < nb) {
            float al = a0[(size_t)j * H * M + m];
            acc[j].x = fmaf(al, v.x, acc[j].x); acc[j].y = fmaf(al, v.y, acc[j].y);
            acc[j].z = fmaf(al, v.z, acc[j].z); acc[j].w = fmaf(al, v.w, acc[j].w);
          }
        }
      }
    }
    if (nsplit > 1) {
      __syncthreads();                                // sm_q no longer needed / previous chunk consumed
      if (active && grp > 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(red + ((size_t)(grp - 1) * 4 + j) * VAL + c) = acc[j];
      }
      __syncthreads();
      if (active && grp == 0) {
        for (int g2 = 1; g2 < nsplit; ++g2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4 p = *reinterpret_cast<const float4*>(red + ((size_t)(g2 - 1) * 4 + j) * VAL + c);
            acc[j].x += p.x; acc[j].y += p.y; acc[j].z += p.z; acc[j].w += p.w;
          }
        }
      }
    }
    if (active && grp == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nb)
          *reinterpret_cast<float4*>(a.ctx_out + (size_t)(b * k + beam0 + j) * a.ld_ctx + c) = acc[j];
    }
  }
}

// Shared memory the fused kernel needs; the phase-3 scratch aliases the query block.
inline size_t attn_fused_smem(int k, int R, int H, int M, int VAL) {
  int tpc = (VAL / 4 + 31) / 32 * 32;
  int nsplit = kAttnThreads / tpc;
  if (nsplit < 1) nsplit = 1;
  size_t qfloats = (size_t)k * R;
  size_t red = (size_t)(nsplit - 1) * 4 * VAL;
  if (red > qfloats) qfloats = red;
  return (qfloats + (size_t)k * H * M) * sizeof(float);
}

}  // namespace comic
