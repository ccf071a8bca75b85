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
//   sum_j v_j tanh_j = sum_j v_j - 2 sum_j v_j r_j.
// Measured (ncu, profiles/): the kernel is bound by the MUFU (XU) pipe, whose
// EX2/RCP rate is far below the FMA pipe's, so four reciprocals share one
// MUFU.RCP (1.25 MUFU per element).  FAST (precision mode 2) uses the
// single-MUFU tanh.approx.f32 instead.
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
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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
  float* hist_t;          // [N, H*M] or nullptr (post-dropout alignments = the reference's history)
  float* hist_pre;        // [N, H*M] or nullptr (pre-dropout alignments, training tape)
  const float* att_mask;  // [N, H*M] 0/1 or nullptr
  float att_keep;
  int k, M, VAL, prob_fn;
  const int* fin_count;   // decode loops: skip when every row finished in step t-1
  int t, n_rows;
};

#ifndef COMIC_ATTN_THREADS
#define COMIC_ATTN_THREADS 384
#endif
constexpr int kAttnThreads = COMIC_ATTN_THREADS;   // 384: two co-resident CTAs per SM (finer image granularity; measured 146 vs 160 us
                                                   // per step at 512 images against 768 threads / one CTA per SM)
constexpr int kAttnMinBlocks = (kAttnThreads <= 384) ? 2 : 1;
constexpr int kAttnRing = 2;        // key-row slots per warp (prefetch distance 1)
constexpr int kAttnBeamChunk = 4;

// Shared-memory carve-up shared by host and device.
struct AttnSmem {
  int tpc, nsplit, qfloats, sfloats;
  size_t bytes;
};
__host__ __device__ inline AttnSmem attn_smem_layout(int k, int R, int H, int M, int VAL) {
  AttnSmem L;
  L.tpc = (VAL / 4 + 31) / 32 * 32;                  // phase 3: threads per position group
  L.nsplit = kAttnThreads / L.tpc;
  if (L.nsplit < 1) L.nsplit = 1;
  int red = (L.nsplit - 1) * kAttnBeamChunk * VAL;   // phase-3 partial sums alias the query block
  int qf = 2 * k * R + ((k + 31) & ~31);           // centred queries | * gamma' | sum of squares
  L.qfloats = qf > red ? qf : red;
  L.sfloats = (k * H * M + 3) & ~3;
  // + LN constants [3][R] + per-warp key-row ring (cp.async staging, lane-private layout)
  L.bytes = ((size_t)L.qfloats + (size_t)L.sfloats + 3 * (size_t)R + (size_t)(kAttnThreads / 32) * kAttnRing * R) *
            sizeof(float);
  return L;
}

// scores of one position against KB beams (add_LN).
//   kc = centred key slice of this lane, kg = kc * gamma' (gamma' = gamma * 2 log2 e; FAST: gamma);
//   qs / qgs = lane-permuted centred queries and centred queries * gamma';
//   sqq[j] = sum_c qc_j[c]^2, skk = sum_c kc[c]^2 (both warp-uniform);
//   cs = lane-permuted constants [gamma' (used by the caller for kg) | beta' | vv].
// Both operands are centred, so sum_c (kc+qc) = 0 and the LN variance is
//   (skk + sqq + 2 <kc, qc>) / R  -- one FFMA per element instead of add + FFMA,
// and the normalised argument is  y' = rstd * (kg + qg) + beta'  -- add + FFMA.
// tanh(y) = 1 - 2/(2^y' + 1); the four reciprocals of a float4 share ONE MUFU.RCP with the
// v-weights folded into the numerators:
//   sum_e vv_e / x_e = (p23 (vv0 x1 + vv1 x0) + p01 (vv2 x3 + vv3 x2)) / (p01 p23),  p01 = x0 x1, p23 = x2 x3
// (y' clamped at 30, products <= 2^120).  Per element: 7.5 FP32 + 1.25 MUFU instructions, which
// balances the FMA-pipe issue rate against the MUFU pipe (8 clk per warp instruction).
template <int CPL, int KB, bool FAST>
__device__ __forceinline__ void ln_tanh_scores(const float (&kc)[CPL], const float (&kg)[CPL],
                                               const float* __restrict__ qs, const float* __restrict__ qgs,
                                               const float* __restrict__ cs, const float* __restrict__ sqq, float skk,
                                               int R, int lane, float sv, float inv_R, float (&out)[KB]) {
  // The arithmetic runs on the packed fp32 pipe (add/mul/fma.f32x2: two lanes of work per issue slot; a plain
  // FFMA issues every other cycle per scheduler on sm_100), which moves the bound from the FMA pipe to MUFU.
  constexpr int G4 = CPL / 4;
  float2 dot2[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) dot2[j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int g = 0; g < G4; ++g) {
    const float2 k01 = make_float2(kc[g * 4 + 0], kc[g * 4 + 1]), k23 = make_float2(kc[g * 4 + 2], kc[g * 4 + 3]);
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const float4 q = *reinterpret_cast<const float4*>(qs + (size_t)j * R + (g * 32 + lane) * 4);
      dot2[j] = __ffma2_rn(k01, make_float2(q.x, q.y), dot2[j]);
      dot2[j] = __ffma2_rn(k23, make_float2(q.z, q.w), dot2[j]);
    }
  }
  float dot[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) dot[j] = dot2[j].x + dot2[j].y;
#pragma unroll
  for (int o = 16; o; o >>= 1) {
#pragma unroll
    for (int j = 0; j < KB; ++j) dot[j] += __shfl_xor_sync(0xffffffffu, dot[j], o);
  }
  float2 rstd2[KB];
#pragma unroll
  for (int j = 0; j < KB; ++j) {
    const float ss = fmaxf(fmaf(2.0f, dot[j], skk + sqq[j]), 0.f);
    float rs;      // argument >= 1e-12: never denormal, so the plain MUFU.RSQ without range fix-ups
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(ss * inv_R + 1e-12f));
    rstd2[j] = make_float2(rs, rs);
    out[j] = FAST ? 0.f : sv;
  }
  const float2 one2 = make_float2(1.0f, 1.0f);
#pragma unroll
  for (int g = 0; g < G4; ++g) {
    const float4 bt = *reinterpret_cast<const float4*>(cs + R + (g * 32 + lane) * 4);
    const float4 vv = *reinterpret_cast<const float4*>(cs + 2 * R + (g * 32 + lane) * 4);
    const float2 kg01 = make_float2(kg[g * 4 + 0], kg[g * 4 + 1]), kg23 = make_float2(kg[g * 4 + 2], kg[g * 4 + 3]);
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const float4 q = *reinterpret_cast<const float4*>(qgs + (size_t)j * R + (g * 32 + lane) * 4);
      const float2 y01 = __ffma2_rn(__fadd2_rn(kg01, make_float2(q.x, q.y)), rstd2[j], make_float2(bt.x, bt.y));
      const float2 y23 = __ffma2_rn(__fadd2_rn(kg23, make_float2(q.z, q.w)), rstd2[j], make_float2(bt.z, bt.w));
      if (FAST) {
        out[j] = fmaf(tanh_approx(y01.x), vv.x, out[j]); out[j] = fmaf(tanh_approx(y01.y), vv.y, out[j]);
        out[j] = fmaf(tanh_approx(y23.x), vv.z, out[j]); out[j] = fmaf(tanh_approx(y23.y), vv.w, out[j]);
      } else {
        // pairs (x0, x2) and (x1, x3): p = (x0 x1, x2 x3), n = (x0 vv1 + x1 vv0, x2 vv3 + x3 vv2)
        const float2 xa = __fadd2_rn(make_float2(ex2_approx(fminf(y01.x, 30.0f)), ex2_approx(fminf(y23.x, 30.0f))), one2);
        const float2 xb = __fadd2_rn(make_float2(ex2_approx(fminf(y01.y, 30.0f)), ex2_approx(fminf(y23.y, 30.0f))), one2);
        const float2 p = __fmul2_rn(xa, xb);
        // (!FAST: vv is stored as (v1, v3, v0, v2) so both operand pairs are adjacent registers)
        const float2 n = __ffma2_rn(xa, make_float2(vv.x, vv.y), __fmul2_rn(xb, make_float2(vv.z, vv.w)));
        const float rp = rcp_approx(p.x * p.y);
        out[j] = fmaf(rp, fmaf(p.x, n.y, p.y * n.x), out[j]);
      }
    }
  }
}

// dynamic shared memory: q_c [k][R], q_c*gamma' [k][R] (lane-permuted), sqq [k] | alpha [k][H][M] | LN constants [3][R] |
// key ring [warps][2][R];
// phase-3 partials alias q_c
template <int R, int H, int MODE, bool FAST, int KB>
__global__ void __launch_bounds__(kAttnThreads, kAttnMinBlocks)
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
  float* sm_c = sm + L.qfloats + L.sfloats;          // [3][R] lane-permuted gamma', beta', vv
  float* ring = sm_c + 3 * R + (size_t)warp * kAttnRing * R;   // this warp's key-row slots
  const int c0 = lane * CPL;                        // first channel of this lane

  // ---- queries: load, centre (add_LN), store lane-permuted (+ gamma'-scaled copy and sum of squares) ----
  float* sm_qg = sm_q + (size_t)k * R;               // [k][G4][32] float4: centred queries * gamma'
  float* sm_sqq = sm_q + 2 * (size_t)k * R;          // [k]
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
    float sq = 0.f;
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 c = make_float4(v[g].x - mean, v[g].y - mean, v[g].z - mean, v[g].w - mean);
      *reinterpret_cast<float4*>(sm_q + (size_t)beam * R + (g * 32 + lane) * 4) = c;
      if (MODE == 0) {
        sq = fmaf(c.x, c.x, sq); sq = fmaf(c.y, c.y, sq); sq = fmaf(c.z, c.z, sq); sq = fmaf(c.w, c.w, sq);
        const float sc = FAST ? 1.0f : kTwoLog2e;
        const float4 g4 = ldg4(a.gamma + c0 + g * 4);
        *reinterpret_cast<float4*>(sm_qg + (size_t)beam * R + (g * 32 + lane) * 4) =
            make_float4(c.x * (g4.x * sc), c.y * (g4.y * sc), c.z * (g4.z * sc), c.w * (g4.w * sc));
      }
    }
    if (MODE == 0) {
      sq = wsum(sq);
      if (lane == 0) sm_sqq[beam] = sq;
    }
  }
  // LN constants (add_LN) in the lane-permuted layout: gamma', beta' pre-scaled by 2 log2(e);
  // vv = -2 v; sv = sum of this lane's v
  float sv = 0.f;
  if (MODE == 0) {
    if (warp == NW - 1) {
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        float4 g4 = ldg4(a.gamma + c0 + g * 4), b4 = ldg4(a.beta + c0 + g * 4), v4 = ldg4(a.vvec + c0 + g * 4);
        const float sc = FAST ? 1.0f : kTwoLog2e, vs = FAST ? 1.0f : -2.0f;
        *reinterpret_cast<float4*>(sm_c + (g * 32 + lane) * 4) = make_float4(g4.x * sc, g4.y * sc, g4.z * sc, g4.w * sc);
        *reinterpret_cast<float4*>(sm_c + R + (g * 32 + lane) * 4) = make_float4(b4.x * sc, b4.y * sc, b4.z * sc, b4.w * sc);
        // exact-tanh path: pair order (v1, v3 | v0, v2) as consumed by the shared-reciprocal numerators
        *reinterpret_cast<float4*>(sm_c + 2 * R + (g * 32 + lane) * 4) =
            FAST ? make_float4(v4.x * vs, v4.y * vs, v4.z * vs, v4.w * vs) : make_float4(v4.y * vs, v4.w * vs, v4.x * vs, v4.z * vs);
      }
    }
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 v4 = ldg4(a.vvec + c0 + g * 4);
      sv += (v4.x + v4.y) + (v4.z + v4.w);
    }
  }
  const float out_scale = (MODE == 0) ? (1.0f / a.temperature[0]) : (1.0f / sqrtf((float)D));
  __syncthreads();

  // ---- phase 1: scores ----
  // Key rows are staged through a per-warp 2-slot cp.async ring (the next row is in flight
  // while the current one is scored); every lane copies exactly the 16-byte chunks it reads
  // back, in a bank-conflict-free [g][lane] layout, so no cross-lane synchronisation is needed.
  const float* kbase = a.keys + (size_t)b * M * R + c0;
  auto stage_row = [&](int m, int slot) {
    if (m < M) {
      const float* kr = kbase + (size_t)m * R;
      float* dst = ring + (size_t)slot * R + lane * 4;
#pragma unroll
      for (int g = 0; g < G4; ++g) cp_async16(dst + g * 128, kr + g * 4);
    }
    cp_async_commit();
  };
  stage_row(warp, 0);
  int slot = 0;
  for (int m = warp; m < M; m += NW) {
    float kc[CPL], kg[CPL];
    float skk = 0.f;
    cp_async_wait<0>();
    {
      const float* src = ring + (size_t)slot * R + lane * 4;
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        float4 t = *reinterpret_cast<const float4*>(src + g * 128);
        kc[g * 4 + 0] = t.x; kc[g * 4 + 1] = t.y; kc[g * 4 + 2] = t.z; kc[g * 4 + 3] = t.w;
        s += (t.x + t.y) + (t.z + t.w);
      }
      if (MODE == 0) {
        float mean = wsum(s) * (1.0f / R);
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          kc[c] -= mean;
          sq = fmaf(kc[c], kc[c], sq);
        }
        skk = wsum(sq);
#pragma unroll
        for (int g = 0; g < G4; ++g) {
          const float4 gm = *reinterpret_cast<const float4*>(sm_c + (g * 32 + lane) * 4);
          kg[g * 4 + 0] = kc[g * 4 + 0] * gm.x; kg[g * 4 + 1] = kc[g * 4 + 1] * gm.y;
          kg[g * 4 + 2] = kc[g * 4 + 2] * gm.z; kg[g * 4 + 3] = kc[g * 4 + 3] * gm.w;
        }
      }
    }
    slot ^= 1;
    stage_row(m + NW, slot);   // next row lands in the slot read one iteration ago while this row is scored
    // KB beams at a time; a short last chunk re-scores the final beams (results identical, stores idempotent)
    for (int beam0 = 0; beam0 < k; beam0 += KB) {
      const int bs = min(beam0, k - KB);               // chunk start, clamped so that bs + KB <= k
      float part[KB];
      const float* qs = sm_q + (size_t)bs * R;
      if (MODE == 0) {
        ln_tanh_scores<CPL, KB, FAST>(kc, kg, qs, sm_qg + (size_t)bs * R, sm_c, sm_sqq + bs, skk, R, lane, sv,
                                      1.0f / R, part);
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
      if (a.hist_pre) a.hist_pre[grow + m] = al;
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
  for (int beam0 = 0; beam0 < k; beam0 += KB) {
    const int bs = min(beam0, k - KB);                // clamped chunk start (a short last chunk recomputes)
    float4 acc[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      const float* a0 = sm_s + ((size_t)bs * H + hd) * M;
      const float* vp = vb + (size_t)m_lo * VAL;
#pragma unroll 7
      for (int m = m_lo; m < m_hi; ++m, vp += VAL) {
        float4 v = ldg4(vp);
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          float al = a0[(size_t)j * H * M + m];
          acc[j].x = fmaf(al, v.x, acc[j].x); acc[j].y = fmaf(al, v.y, acc[j].y);
          acc[j].z = fmaf(al, v.z, acc[j].z); acc[j].w = fmaf(al, v.w, acc[j].w);
        }
      }
    }
    if (nsplit > 1) {
      __syncthreads();                                // query block no longer needed / previous chunk consumed
      if (active && grp > 0) {
#pragma unroll
        for (int j = 0; j < KB; ++j)
          *reinterpret_cast<float4*>(red + ((size_t)(grp - 1) * kAttnBeamChunk + j) * VAL + c) = acc[j];
      }
      __syncthreads();
      if (active && grp == 0) {
        for (int g2 = 1; g2 < nsplit; ++g2) {
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            float4 p = *reinterpret_cast<const float4*>(red + ((size_t)(g2 - 1) * kAttnBeamChunk + j) * VAL + c);
            acc[j].x += p.x; acc[j].y += p.y; acc[j].z += p.z; acc[j].w += p.w;
          }
        }
      }
    }
    if (active && grp == 0) {
#pragma unroll
      for (int j = 0; j < KB; ++j)
        *reinterpret_cast<float4*>(a.ctx_out + (size_t)(b * k + bs + j) * a.ld_ctx + c) = acc[j];
    }
  }
}

inline size_t attn_fused_smem(int k, int R, int H, int M, int VAL) { return attn_smem_layout(k, R, H, M, VAL).bytes; }

}  // namespace comic
