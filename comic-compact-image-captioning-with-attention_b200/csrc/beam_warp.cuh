// beam_warp.cuh -- one image's beam-search step by ONE warp (small vocabularies: k * V candidates fit a warp's shared-memory
// slice).  TF r1.9 _beam_search_step under rnn_decoder_beam_search (common/ops_rnn.py:49-112): log-softmax of the k beam rows,
// _mask_probs for finished beams, length penalty, top-k over the flattened [k * V] scores with lowest-flat-index tie-break,
// parent / word / finished / lengths bookkeeping.  Shared by beam_step_warp_kernel (decoder.cu) and the finaliser warp of the
// streaming attention kernel (attention2.cuh), which runs the step of an image beside its attention instead of in a separate
// launch: same code, same bits.
#pragma once
#include "search_steps.cuh"

namespace comic {

struct BeamWarpArgs {
  const float* logits;      // [B * k][ld], this step's logits in the first V columns
  int ld, k, V, eos;
  float lpw;
  float* log_probs;         // [B * k] cumulative log-probabilities (in / out)
  uint8_t* finished;        // [B * k]
  long long* lengths;       // [B * k]
  float* scores_out;        // [B * k] this step's top-k scores
  int* word_out;            // [B * k]
  int* parent_out;          // [B * k]
  int* tok_next;            // [B * k] or nullptr
  int* src_next;            // [B * k] or nullptr
  int* fin_count;           // [T + 1] or nullptr: fin_count[t] counts the rows finished after this step
  int t;
};

// sc: k * V floats of shared memory owned by the calling warp.
__device__ __forceinline__ void beam_step_one_warp(const BeamWarpArgs& g, int b, float* sc, int lane) {
  const float* __restrict__ logits = g.logits;
  const int ld = g.ld, k = g.k, V = g.V, eos = g.eos, t = g.t;
  const float lpw = g.lpw;
  float* log_probs = g.log_probs;
  uint8_t* finished = g.finished;
  long long* lengths = g.lengths;
  float* scores_out = g.scores_out;
  int *word_out = g.word_out, *parent_out = g.parent_out, *tok_next = g.tok_next, *src_next = g.src_next, *fin_count = g.fin_count;
  const int ncand = k * V;
  const float* base = logits + (size_t)b * k * ld;
  // lane j < k keeps the state and the log-softmax statistics of beam row j
  float cum = 0.f, mxr = 0.f, lser = 0.f;
  int fin = 0;
  long long len = 0;
  if (lane < k) {
    cum = log_probs[b * k + lane];
    fin = finished[b * k + lane];
    len = lengths[b * k + lane];
  }
  for (int j = 0; j < k; ++j) {
    const float* row = base + (size_t)j * ld;
    float* srow = sc + j * V;
    float mx = -INFINITY;
    for (int i = lane; i < V; i += 32) {
      const float x = row[i];
      srow[i] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float sm = 0.f;
    for (int i = lane; i < V; i += 32) sm += expf(srow[i] - mx);
    sm = warp_sum(sm);
    const float lse = logf(sm);
    if (lane == j) { mxr = mx; lser = lse; }
    const float cumj = __shfl_sync(0xffffffffu, cum, j);
    const bool finj = __shfl_sync(0xffffffffu, fin, j) != 0;
    const long long lenj = __shfl_sync(0xffffffffu, len, j);
    const float pen_live = (lpw == 0.0f) ? 1.0f : length_penalty_dev(lenj + (finj ? 0 : 1), lpw);
    const float pen_eos = (lpw == 0.0f) ? 1.0f : length_penalty_dev(lenj, lpw);
    for (int i = lane; i < V; i += 32) {
      float lp;
      if (finj) lp = (i == eos) ? 0.0f : -FLT_MAX;
      else lp = (srow[i] - mx) - lse;
      const float tot = cumj + lp;
      srow[i] = (lpw == 0.0f) ? tot : tot / ((i == eos) ? pen_eos : pen_live);
    }
  }
  __syncwarp();
  float pv = INFINITY, myv = 0.f;
  int pi = -1, myi = 0x7fffffff;
  for (int sel = 0; sel < k; ++sel) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int idx = lane; idx < ncand; idx += 32) {
      const float s = sc[idx];
      const bool eligible = (s < pv) || (s == pv && idx > pi);
      if (eligible && better(s, idx, bv, bi)) { bv = s; bi = idx; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    pv = bv; pi = bi;
    if (lane == sel) { myv = bv; myi = bi; }
  }
  // lane `sel` < k finishes selection `sel`
  int idx = myi;
  if (idx == 0x7fffffff) idx = 0;                      // only if every candidate is NaN
  const int par = (lane < k) ? idx / V : 0, w = idx - par * V;
  const float cum_p = __shfl_sync(0xffffffffu, cum, par), mx_p = __shfl_sync(0xffffffffu, mxr, par);
  const float lse_p = __shfl_sync(0xffffffffu, lser, par);
  const int fin_p = __shfl_sync(0xffffffffu, fin, par);
  const long long len_p = __shfl_sync(0xffffffffu, len, par);
  if (lane < k) {
    float lp;
    if (fin_p) lp = (w == eos) ? 0.0f : -FLT_MAX;
    else lp = (base[(size_t)par * ld + w] - mx_p) - lse_p;
    const bool nfin = fin_p || (w == eos);
    log_probs[b * k + lane] = cum_p + lp;
    finished[b * k + lane] = nfin ? 1 : 0;
    lengths[b * k + lane] = len_p + (fin_p ? 0 : 1);
    scores_out[b * k + lane] = myv;
    word_out[b * k + lane] = w;
    parent_out[b * k + lane] = par;
    if (tok_next) tok_next[b * k + lane] = w;
    if (src_next) src_next[b * k + lane] = b * k + par;
    if (fin_count && nfin) atomicAdd(&fin_count[t], 1);
  }
}

}  // namespace comic
