// search_steps.cuh -- per-step selection of the decode loops as device functions,
// shared by the one-launch-per-step kernels (decoder.cu) and the persistent
// whole-loop kernel (persistent.cu) so that both produce the same bits:
//   beam_step_block   TF r1.9 _beam_search_step (BeamSearchDecoder.step tail) under
//                     rnn_decoder_beam_search, common/ops_rnn.py:49-112
//   greedy_step_block GreedyEmbeddingHelper.sample + BasicDecoder bookkeeping under
//                     rnn_decoder_search, common/ops_rnn.py:115-180
//   lstm_cell         BasicLSTMCell pointwise part (gate order i, j, f, o; forget_bias 1)
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

namespace comic {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Loop gate: step t runs only while not every row had finished after step t-1.
__device__ __forceinline__ bool step_stopped(const int* fin_count, int t, int n_rows) {
  return fin_count != nullptr && t > 0 && fin_count[t - 1] >= n_rows;
}

// c' = c * sigmoid(f + 1) + sigmoid(i) * tanh(j);  h' = tanh(c') * sigmoid(o)
__device__ __forceinline__ void lstm_cell(float gi, float gj, float gf, float go, float cp, float* cn, float* hn) {
  float c = cp * sigmoidf_(gf + 1.0f) + sigmoidf_(gi) * tanhf(gj);
  *cn = c;
  *hn = tanhf(c) * sigmoidf_(go);
}

// Barrier among the first `n` threads of a CTA (n a multiple of 32): lets a 256-thread
// selection run inside a larger CTA while its other warps skip the call.
__device__ __forceinline__ void group_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  // descending value, ties -> lower flat index (nn.top_k)
  return (v > bv) || (v == bv && i < bi);
}

__device__ __forceinline__ float length_penalty_dev(long long len, float w) {
  return powf(5.0f + (float)len, w) / powf(6.0f, w);
}

struct BeamStepSmem {
  static constexpr int KMAX = 16;
  float s_max[KMAX], s_lse[KMAX], s_cum[KMAX];
  unsigned char s_fin[KMAX];
  long long s_len[KMAX];
  float s_rv[8];
  int s_ri[8];
  float s_selv[KMAX];
  int s_seli[KMAX];
  // large vocabularies: candidates that reach the block's score threshold (see beam_step_block)
  static constexpr int CAND_CAP = 128;
  int s_cnt;
  float s_theta;
  float s_cv[CAND_CAP];
  int s_ci[CAND_CAP];
};

// One image's beam step, executed by EXACTLY the first 256 threads of the CTA (tid < 256);
// CG = true reads the logits with ld.global.cg (written by other CTAs of the same launch).
// `stage` (capacity stage_cap floats, shared memory): when the image's k logits rows fit they are copied
// there once and every later pass reads shared memory; the arithmetic and its order do not change.
template <bool CG>
__device__ __forceinline__ void beam_step_block(BeamStepSmem& S, float* stage, int stage_cap, int tid, int b,
                                                const float* __restrict__ logits, int ld,
                                                int k, int V, int eos, float lpw, float* __restrict__ log_probs,
                                                uint8_t* __restrict__ finished, long long* __restrict__ lengths,
                                                float* __restrict__ scores_out, int* __restrict__ word_out,
                                                int* __restrict__ parent_out, int* __restrict__ tok_next,
                                                int* __restrict__ src_next, int* fin_count, int t) {
  const int lane = tid & 31, warp = tid >> 5;
  const bool staged = stage != nullptr && k * V <= stage_cap;
  const float* base = logits + (size_t)b * k * ld;     // row j of this image: base + j * ldr
  int ldr = ld;
  if (staged) {
    for (int j = 0; j < k; ++j)
      for (int i = tid; i < V; i += 256) stage[j * V + i] = CG ? __ldcg(base + (size_t)j * ld + i) : base[(size_t)j * ld + i];
    base = stage;
    ldr = V;
  }
  auto ld_logit = [&](const float* p) -> float { return (CG && !staged) ? __ldcg(p) : *p; };
  if (tid < k) {
    S.s_cum[tid] = log_probs[b * k + tid];
    S.s_fin[tid] = finished[b * k + tid];
    S.s_len[tid] = lengths[b * k + tid];
  }
  if (staged) group_sync(1, 256);
  // log-softmax statistics per beam row: max, log(sum(exp(x - max)))
  float tbest = -INFINITY;          // large vocabularies: the best score among one real candidate per row of this thread
  for (int j = 0; j < k; ++j) {
    const float* row = base + (size_t)j * ldr;
    float mx = -INFINITY;
    int amx = -1;                   // (non-staged) column of this thread's largest logit in row j
    if (staged) {
      for (int i = tid; i < V; i += 256) mx = fmaxf(mx, ld_logit(row + i));
    } else {
      // word vocabularies read the row from L2: 8 independent loads in flight per thread instead of one dependent load
      // per iteration (39 round trips per pass at V = 10,000); the order of every accumulation is unchanged
      for (int i0 = tid; i0 < V; i0 += 256 * 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (i0 + 256 * u < V) ? ld_logit(row + i0 + 256 * u) : -INFINITY;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (x[u] > mx) { mx = x[u]; amx = i0 + 256 * u; }
      }
    }
    const float tmx = mx;           // this thread's own row maximum (before the block reduction)
    mx = warp_max(mx);
    if (lane == 0) S.s_rv[warp] = mx;
    group_sync(1, 256);
    float m2 = S.s_rv[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m2 = fmaxf(m2, S.s_rv[w]);
    group_sync(1, 256);
    float sm = 0.f;
    if (staged) {
      for (int i = tid; i < V; i += 256) sm += expf(ld_logit(row + i) - m2);
    } else {
      // word vocabularies: 10^4 exponentials per row -- ex2.approx-based exp (relative error ~1e-6 per
      // term, far below the fp32 rounding of the log-prob it feeds)
      for (int i0 = tid; i0 < V; i0 += 256 * 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (i0 + 256 * u < V) ? ld_logit(row + i0 + 256 * u) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + 256 * u < V) sm += __expf(x[u] - m2);
      }
    }
    sm = warp_sum(sm);
    if (lane == 0) S.s_rv[warp] = sm;
    group_sync(1, 256);
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += S.s_rv[w];
      S.s_max[j] = m2;
      S.s_lse[j] = logf(tot);
    }
    group_sync(1, 256);
    if (!staged && amx >= 0) {
      // the score of candidate (j, amx), with exactly the arithmetic of the selection scan below
      const bool finj = S.s_fin[j] != 0;
      float lp;
      if (finj) lp = (amx == eos) ? 0.0f : -FLT_MAX;
      else lp = (tmx - S.s_max[j]) - S.s_lse[j];
      const float tot = S.s_cum[j] + lp;
      float sc = tot;
      if (lpw != 0.0f) sc = tot / length_penalty_dev(S.s_len[j] + ((!finj && amx != eos) ? 1 : 0), lpw);
      if (sc > tbest) tbest = sc;
    }
  }
  const int ncand = k * V;
  auto total_of = [&](int idx) -> float {
    int j = idx / V, w = idx - j * V;
    float lp;
    if (S.s_fin[j]) lp = (w == eos) ? 0.0f : -FLT_MAX;
    else lp = (ld_logit(base + (size_t)j * ldr + w) - S.s_max[j]) - S.s_lse[j];
    return S.s_cum[j] + lp;
  };
  auto score_of = [&](int idx, float tot) -> float {
    if (lpw == 0.0f) return tot;
    int j = idx / V, w = idx - j * V;
    long long len = S.s_len[j] + ((!S.s_fin[j] && w != eos) ? 1 : 0);
    return tot / length_penalty_dev(len, lpw);
  };
  bool selected = false;
  if (!staged && k <= 8) {
    // Large vocabularies.  The k-th largest of the 256 threads' `tbest` values is a lower bound (theta) of the k-th best
    // score of the image -- they are the scores of k distinct real candidates -- so one scan that keeps only candidates
    // with score >= theta (a handful out of k * V) followed by an exact top-k over those few gives the same winners as
    // the k-pass selection below, in the same total order (score descending, flat index ascending).  The per-thread
    // sorted top-8 list this replaces ran its insertion code with 1-3 active lanes on almost every iteration
    // (ncu, V = 10,000: 64 M warp instructions per launch at 14 of 32 threads active, issue-bound at 98 us;
    // profiles/r12w_ncu_beam_word_summary.txt).
    {
      float v = tbest;
      float theta = -INFINITY;
      for (int sel = 0; sel < k; ++sel) {
        float bv = v;
        int bi = tid;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { S.s_rv[warp] = bv; S.s_ri[warp] = bi; }
        group_sync(1, 256);
        bv = S.s_rv[0]; bi = S.s_ri[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
          if (better(S.s_rv[w], S.s_ri[w], bv, bi)) { bv = S.s_rv[w]; bi = S.s_ri[w]; }
        group_sync(1, 256);
        if (bi == tid) v = -INFINITY;            // pop the winner
        theta = bv;
      }
      if (tid == 0) { S.s_theta = theta; S.s_cnt = 0; }
      group_sync(1, 256);
    }
    const float theta = S.s_theta;
    for (int j = 0; j < k; ++j) {
      const float* row = base + (size_t)j * ldr;
      const float mxj = S.s_max[j], lsej = S.s_lse[j], cumj = S.s_cum[j];
      const bool finj = S.s_fin[j] != 0;
      const long long lenj = S.s_len[j];
      const float pen_live = (lpw == 0.0f) ? 1.0f : length_penalty_dev(lenj + (finj ? 0 : 1), lpw);
      const float pen_eos = (lpw == 0.0f) ? 1.0f : length_penalty_dev(lenj, lpw);
      const int off = j * V;
      for (int w0 = tid; w0 < V; w0 += 256 * 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (!finj && w0 + 256 * u < V) ? ld_logit(row + w0 + 256 * u) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int w = w0 + 256 * u;
          if (w >= V) break;
          float lp;
          if (finj) lp = (w == eos) ? 0.0f : -FLT_MAX;
          else lp = (x[u] - mxj) - lsej;
          const float tot = cumj + lp;
          const float sc = (lpw == 0.0f) ? tot : tot / ((w == eos) ? pen_eos : pen_live);
          if (sc >= theta) {
            const int pos = atomicAdd(&S.s_cnt, 1);
            if (pos < BeamStepSmem::CAND_CAP) { S.s_cv[pos] = sc; S.s_ci[pos] = off + w; }
          }
        }
      }
    }
    group_sync(1, 256);
    const int cnt = S.s_cnt;
    if (cnt <= BeamStepSmem::CAND_CAP) {         // (more: massive ties -- the k-pass selection below handles them)
      selected = true;
      if (warp == 0) {
        float pv = INFINITY;
        int pi = -1;
        for (int sel = 0; sel < k; ++sel) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int e = lane; e < cnt; e += 32) {
            const float sc = S.s_cv[e];
            const int idx = S.s_ci[e];
            const bool eligible = (sc < pv) || (sc == pv && idx > pi);
            if (eligible && better(sc, idx, bv, bi)) { bv = sc; bi = idx; }
          }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
          }
          pv = bv; pi = bi;
          if (lane == 0) { S.s_selv[sel] = bv; S.s_seli[sel] = bi; }
        }
      }
    }
  }
  if (!selected) {
  float pv = INFINITY;
  int pi = -1;
  for (int sel = 0; sel < k; ++sel) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (staged) {
      for (int idx = tid; idx < ncand; idx += 256) {
        float sc = score_of(idx, total_of(idx));
        bool eligible = (sc < pv) || (sc == pv && idx > pi);
        if (eligible && better(sc, idx, bv, bi)) { bv = sc; bi = idx; }
      }
    } else {
      // large vocabularies (word models): same arithmetic per candidate, but walked row by row so the
      // flat index never has to be divided by V and the per-row terms are hoisted out of the scan
      for (int j = 0; j < k; ++j) {
        const float* row = base + (size_t)j * ldr;
        const float mxj = S.s_max[j], lsej = S.s_lse[j], cumj = S.s_cum[j];
        const bool finj = S.s_fin[j] != 0;
        const long long lenj = S.s_len[j];
        const float pen_live = (lpw == 0.0f) ? 1.0f : length_penalty_dev(lenj + (finj ? 0 : 1), lpw);
        const float pen_eos = (lpw == 0.0f) ? 1.0f : length_penalty_dev(lenj, lpw);
        const int off = j * V;
        for (int w = tid; w < V; w += 256) {
          float lp;
          if (finj) lp = (w == eos) ? 0.0f : -FLT_MAX;
          else lp = (ld_logit(row + w) - mxj) - lsej;
          const float tot = cumj + lp;
          const float sc = (lpw == 0.0f) ? tot : tot / ((w == eos) ? pen_eos : pen_live);
          const int idx = off + w;
          const bool eligible = (sc < pv) || (sc == pv && idx > pi);
          if (eligible && better(sc, idx, bv, bi)) { bv = sc; bi = idx; }
        }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { S.s_rv[warp] = bv; S.s_ri[warp] = bi; }
    group_sync(1, 256);
    bv = S.s_rv[0]; bi = S.s_ri[0];
#pragma unroll
    for (int w = 1; w < 8; ++w)
      if (better(S.s_rv[w], S.s_ri[w], bv, bi)) { bv = S.s_rv[w]; bi = S.s_ri[w]; }
    group_sync(1, 256);
    pv = bv; pi = bi;
    if (tid == 0) { S.s_selv[sel] = bv; S.s_seli[sel] = bi; }
  }
  }   // k-pass selection
  group_sync(1, 256);
  if (tid < k) {
    int idx = S.s_seli[tid];
    if (idx == 0x7fffffff) idx = 0;   // only if every candidate is NaN
    int par = idx / V, w = idx - par * V;
    float tot = total_of(idx);
    bool pfin = S.s_fin[par] != 0;
    bool nfin = pfin || (w == eos);
    long long nlen = S.s_len[par] + (pfin ? 0 : 1);
    log_probs[b * k + tid] = tot;
    finished[b * k + tid] = nfin ? 1 : 0;
    lengths[b * k + tid] = nlen;
    scores_out[b * k + tid] = S.s_selv[tid];
    word_out[b * k + tid] = w;
    parent_out[b * k + tid] = par;
    if (tok_next) tok_next[b * k + tid] = w;
    if (src_next) src_next[b * k + tid] = b * k + par;
    if (fin_count && nfin) atomicAdd(&fin_count[t], 1);
  }
}

struct GreedyStepSmem {
  float s_rv[4];
  int s_ri[4];
};

// One row's greedy step, executed by EXACTLY the first 128 threads of the CTA.
// argmax = first maximum (int32); outputs are NOT masked after EOS (impute_finished=False).
template <bool CG>
__device__ __forceinline__ void greedy_step_block(GreedyStepSmem& S, int tid, int n, const float* __restrict__ logits,
                                                  int ld, int V, int eos, int* __restrict__ ids_t,
                                                  float* __restrict__ logits_t, int* __restrict__ tok_next,
                                                  uint8_t* __restrict__ finished, int* fin_count, int t) {
  const int lane = tid & 31, warp = tid >> 5;
  const float* row = logits + (size_t)n * ld;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < V; i += 128) {
    float v = CG ? __ldcg(row + i) : row[i];
    if (logits_t) logits_t[(size_t)n * V + i] = v;
    if (better(v, i, bv, bi)) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  if (lane == 0) { S.s_rv[warp] = bv; S.s_ri[warp] = bi; }
  group_sync(2, 128);
  if (tid == 0) {
    for (int w = 1; w < 4; ++w)
      if (better(S.s_rv[w], S.s_ri[w], bv, bi)) { bv = S.s_rv[w]; bi = S.s_ri[w]; }
    if (bi == 0x7fffffff) bi = 0;
    ids_t[n] = bi;
    tok_next[n] = bi;
    bool f = finished[n] || (bi == eos);
    finished[n] = f ? 1 : 0;
    if (f) atomicAdd(&fin_count[t], 1);
  }
}

}  // namespace comic
