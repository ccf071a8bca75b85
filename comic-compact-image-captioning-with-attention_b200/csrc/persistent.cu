// persistent.cu -- the whole beam / greedy decode loop as ONE cooperative kernel
// for small row counts (N = batch x beam <= 32), sm_100a.
//
// Replaces, for one decode call, the while_loop that TF builds for
//   rnn_decoder_beam_search / rnn_decoder_search   common/ops_rnn.py:49-180
// around MultiHeadAttentionWrapperV3.call           common/ops_rnn.py:660-755
// (~150 TF ops per step; 8 kernel launches per step on the per-step path of
// decoder.cu).  At N = 24 rows a step moves ~16 MB and is pure latency, so the
// loop runs inside one launch with the operands that do not change between
// steps resident in shared memory, and the CTAs meet at four grid barriers per
// step:
//
//   phase A  gates + LSTM cell: CTA a owns UPC hidden units = 4*UPC columns of
//            the [W+A+R, 4R] LSTM kernel, RESIDENT in smem for the whole loop
//            (80 KB at COMIC-256); x = [emb(tok) ; ctx[src] ; h[src]] streams in
//            through a cp.async double buffer; 16 warps split K, fixed-order
//            reduction; the cell update is applied in place -> c', h'.
//   phase B  [logits | query] = h' . [W_o | W_q] + b: 8-column blocks per CTA.
//   phase C1 attention scores: CTA (image, position slice) keeps its CENTRED key
//            rows resident in smem (keys never re-read from HBM after step 0) and
//            scores them against the k beam queries (ln_tanh_scores, attention.cuh);
//            meanwhile B other CTAs run the beam step (log-softmax, top-k,
//            backpointers: beam_step_block, search_steps.cuh) on the logits.
//   phase C2 softmax over the 196 positions + context: CTA (image, channel slice)
//            with its value columns resident in smem; alignment history write.
//
// Cross-CTA data inside the launch is read with ld.global.cg / cp.async.cg (L2)
// and published with __threadfence() before the barrier arrive.  The barrier is a
// monotonic counter; a spin that exceeds ~2 s sets an abort flag that every CTA
// honours (T_out becomes -1 and the host raises), so a bug cannot hang the GPU.
#include "attention.cuh"
#include "comic_internal.cuh"
#include "search_steps.cuh"

namespace comic {

constexpr int kPT = 512;            // threads per CTA
constexpr int kPW = kPT / 32;       // warps
constexpr int kPMaxRPL = 4;         // rows per lane group: N <= 8 * 4
constexpr int kPKC = 256;           // phase-A K chunk (16 warps x 16)
constexpr int kPMaxHeads = 4;       // heads a CTA may need in phase C2

struct PersistArgs {
  // weights
  const float* lstm_kernel;   // [KX][4R]
  const float* lstm_bias;     // [4R]
  const float* outq;          // [R][LQ] = [W_o | pad | W_q]
  const float* outq_bias;     // [LQ]
  const float* emb;           // [V][W]
  const float *gamma, *beta, *vvec, *temperature;
  // per-call inputs
  const float* keys;          // [B][M][R]
  const float* values;        // [B][M][VAL]
  const float* c0;            // [B][R]
  const float* h0;
  // dims
  int B, k, N, W, A, KX, V, LQ, q_off, M, VAL, prob_fn, eos, max_it, greedy;
  float lpw;
  // state (global, caller workspace)
  float* c[2];
  float* h[2];
  float* ctx[2];
  float* lq;                  // [N][LQ]
  float* scores;              // [N][H][M]
  float* hist;                // [T][N][H*M] or nullptr
  int* tok;
  int* src;
  float* cum;
  uint8_t* fin;
  long long* len;
  int* fin_count;
  int* step_ids;              // beam: [T][N] words; greedy: ids_out [T][N]
  int* parents;               // beam
  float* sc;                  // beam: [T][N] scores
  float* logits_out;          // greedy: [T][N][V] or nullptr
  unsigned* bar;              // [0] barrier counter, [1] abort flag
  // partition / smem plan (floats)
  int nA, UPC, CB, nB, S, rps, cps, nh_max, keys_res, vals_res, n_sel;
  int off_wA, off_c, off_keys, off_skk, off_vals, off_tr;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (cooperative) grid meet here.  Returns false when the launch was aborted.
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned& target, unsigned nblocks, int* s_abort) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    int ab = 0;
    while (ld_acquire_u32(bar) < target) {
      if (ld_acquire_u32(bar + 1) != 0u) { ab = 1; break; }
      if (clock64() - t0 > 4000000000ll) {   // ~2 s at 1.9 GHz: publish the abort and leave
        atomicExch(bar + 1, 1u);
        ab = 1;
        break;
      }
    }
    __threadfence();
    *s_abort = ab;
  }
  __syncthreads();
  return *s_abort == 0;
}

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(sz)
               : "memory");
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

template <int H, int KB>
__global__ void __launch_bounds__(kPT, 1)
decode_loop_kernel(const PersistArgs a) {
  constexpr int R = 512;
  constexpr int CPL = R / 32, G4 = CPL / 4;
  constexpr int D = R / H;
  constexpr int LPH = (D >= CPL) ? D / CPL : 1;
  constexpr float kTwoLog2e = 2.885390081777927f;
  extern __shared__ __align__(16) float sm[];
  __shared__ BeamStepSmem s_beam;
  __shared__ GreedyStepSmem s_greedy;
  __shared__ const float* s_xp[3][32];   // x segment row pointers (nullptr = zero row)
  __shared__ int s_crow[32];             // c_prev row of each output row
  __shared__ int s_abort;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = blockIdx.x;
  const unsigned G = gridDim.x;
  const int N = a.N, k = a.k, M = a.M, KX = a.KX, LQ = a.LQ, VAL = a.VAL;
  const int ng = lane >> 2, cg = lane & 3;
  const int rpl = (N + 7) >> 3;
  unsigned bar_target = 0;

  float* wA = sm + a.off_wA;          // [KX][4*UPC]
  float* sm_c = sm + a.off_c;         // [3][R] lane-permuted gamma', beta', vv
  float* sm_keys = sm + a.off_keys;   // [rps][R] centred key rows, lane-permuted
  float* sm_skk = sm + a.off_skk;     // [rps]
  float* sm_vals = sm + a.off_vals;   // [M][cps]
  float* tr = sm + a.off_tr;          // transient region (per-phase layouts)

  const int NC = 4 * a.UPC;           // gate columns of this CTA
  const bool has_A = cta < a.nA;
  const bool has_B = cta < a.nB;
  const bool has_C = cta < a.B * a.S;
  const int img = has_C ? cta / a.S : 0, slc = has_C ? cta % a.S : 0;
  const int m0 = slc * a.rps, m1 = min(M, m0 + a.rps);
  const int ch0 = slc * a.cps, ch1 = min(VAL, ch0 + a.cps);
  const bool sel_cta = cta >= a.B * a.S && cta < a.B * a.S + a.n_sel;
  const int c0l = lane * CPL;

  // ------------------------------------------------------------------ setup
  if (has_A) {
    // column j = g*UPC + u  <-  global column g*R + (cta*UPC + u)
    const int tot = KX * NC;
    for (int i = tid; i < tot; i += kPT) {
      int kk = i / NC, j = i - kk * NC;
      int g = j / a.UPC, u = j - g * a.UPC;
      wA[i] = __ldg(a.lstm_kernel + (size_t)kk * 4 * R + g * R + cta * a.UPC + u);
    }
  }
  float sv = 0.f;
  {
    if (warp == 0) {
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        float4 g4 = ldg4(a.gamma + c0l + g * 4), b4 = ldg4(a.beta + c0l + g * 4), v4 = ldg4(a.vvec + c0l + g * 4);
        *reinterpret_cast<float4*>(sm_c + (g * 32 + lane) * 4) =
            make_float4(g4.x * kTwoLog2e, g4.y * kTwoLog2e, g4.z * kTwoLog2e, g4.w * kTwoLog2e);
        *reinterpret_cast<float4*>(sm_c + R + (g * 32 + lane) * 4) =
            make_float4(b4.x * kTwoLog2e, b4.y * kTwoLog2e, b4.z * kTwoLog2e, b4.w * kTwoLog2e);
        *reinterpret_cast<float4*>(sm_c + 2 * R + (g * 32 + lane) * 4) =
            make_float4(v4.x * -2.0f, v4.y * -2.0f, v4.z * -2.0f, v4.w * -2.0f);
      }
    }
#pragma unroll
    for (int g = 0; g < G4; ++g) {
      float4 v4 = ldg4(a.vvec + c0l + g * 4);
      sv += (v4.x + v4.y) + (v4.z + v4.w);
    }
  }
  if (has_C && a.keys_res) {
    for (int r = warp; r < m1 - m0; r += kPW) {
      const float* kr = a.keys + ((size_t)img * M + m0 + r) * R + c0l;
      float4 v[G4];
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        v[g] = ldg4(kr + g * 4);
        s += (v[g].x + v[g].y) + (v[g].z + v[g].w);
      }
      const float mean = wsum(s) * (1.0f / R);
      float sq = 0.f;
#pragma unroll
      for (int g = 0; g < G4; ++g) {
        float4 c = make_float4(v[g].x - mean, v[g].y - mean, v[g].z - mean, v[g].w - mean);
        sq = fmaf(c.x, c.x, sq); sq = fmaf(c.y, c.y, sq); sq = fmaf(c.z, c.z, sq); sq = fmaf(c.w, c.w, sq);
        *reinterpret_cast<float4*>(sm_keys + (size_t)r * R + (g * 32 + lane) * 4) = c;
      }
      sq = wsum(sq);
      if (lane == 0) sm_skk[r] = sq;
    }
  }
  if (has_C && a.vals_res && ch0 < VAL) {
    const int nc4 = (ch1 - ch0) >> 2;
    for (int i = tid; i < M * nc4; i += kPT) {
      int m = i / nc4, c4 = i - m * nc4;
      *reinterpret_cast<float4*>(sm_vals + (size_t)m * a.cps + c4 * 4) =
          ldg4(a.values + ((size_t)img * M + m) * VAL + ch0 + c4 * 4);
    }
  }
  const float out_scale = 1.0f / __ldg(a.temperature);
  __syncthreads();

  // ------------------------------------------------------------------ step loop
  for (int t = 0; t < a.max_it; ++t) {
    const int cur = t & 1;
    const float* c_prev = (t == 0) ? a.c0 : a.c[cur];
    const float* h_prev = (t == 0) ? a.h0 : a.h[cur];
    const float* ctx_prev = a.ctx[cur];
    float* c_new = a.c[cur ^ 1];
    float* h_new = a.h[cur ^ 1];
    float* ctx_new = a.ctx[cur ^ 1];

    // ============================ phase A: gates + LSTM cell
    if (has_A) {
      if (tid < N) {
        const int n = tid;
        const int tk = __ldcg(a.tok + n);
        int sr;
        if (a.greedy) sr = n;
        else sr = (t == 0) ? n / k : __ldcg(a.src + n);
        const int lim = (t == 0 && !a.greedy) ? a.B : N;
        const bool ok = sr >= 0 && sr < lim;
        s_xp[0][n] = (tk >= 0 && tk < a.V) ? a.emb + (size_t)tk * a.W : nullptr;
        s_xp[1][n] = ok ? ctx_prev + (size_t)sr * a.A : nullptr;
        s_xp[2][n] = ok ? h_prev + (size_t)sr * R : nullptr;
        s_crow[n] = ok ? sr : -1;
      }
      __syncthreads();
      const int nchunk = (KX + kPKC - 1) / kPKC;
      const int XLD = kPKC + 4;                       // padded row stride (bank shift of one float4 per row)
      auto issue = [&](int c) {
        float* buf = tr + (size_t)(c & 1) * N * XLD;
        const int kc0 = c * kPKC;
        const int pieces = N * (kPKC / 4);
        for (int p = tid; p < pieces; p += kPT) {
          const int n = p / (kPKC / 4), q = p - n * (kPKC / 4);
          const int kg = kc0 + q * 4;
          const float* srcp = nullptr;
          if (kg < KX) {
            if (kg < a.W) { const float* b0 = s_xp[0][n]; srcp = b0 ? b0 + kg : nullptr; }
            else if (kg < a.W + a.A) { const float* b1 = s_xp[1][n]; srcp = b1 ? b1 + (kg - a.W) : nullptr; }
            else { const float* b2 = s_xp[2][n]; srcp = b2 ? b2 + (kg - a.W - a.A) : nullptr; }
          }
          cp_async16_zfill(buf + (size_t)n * XLD + q * 4, srcp ? (const void*)srcp : (const void*)a.emb, srcp != nullptr);
        }
        cp_async_commit();
      };
      float acc[kPMaxRPL][4];
#pragma unroll
      for (int r = 0; r < kPMaxRPL; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
      issue(0);
      for (int c = 0; c < nchunk; ++c) {
        if (c + 1 < nchunk) {
          issue(c + 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
        const float* buf = tr + (size_t)(c & 1) * N * XLD;
        const int kc0 = c * kPKC;
        // a lane owns rows {ng, ng+8, ...} and, for every group of 4 consecutive columns jg, the 4 columns of it
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kl = warp * 16 + j * 4;
          if (kc0 + kl < KX) {
            float4 x4[kPMaxRPL];
#pragma unroll
            for (int r = 0; r < kPMaxRPL; ++r) {
              const int n = ng + 8 * r;
              x4[r] = (r < rpl && n < N) ? *reinterpret_cast<const float4*>(buf + (size_t)n * XLD + kl)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int jg = cg; jg * 4 < NC; jg += 4) {
              // UPC == 4: NC = 16 and every lane has exactly one column group (jg = cg)
              const float* wp = wA + (size_t)(kc0 + kl) * NC + jg * 4;
              const float4 w0 = *reinterpret_cast<const float4*>(wp);
              const float4 w1 = *reinterpret_cast<const float4*>(wp + NC);
              const float4 w2 = *reinterpret_cast<const float4*>(wp + 2 * NC);
              const float4 w3 = *reinterpret_cast<const float4*>(wp + 3 * NC);
#pragma unroll
              for (int r = 0; r < kPMaxRPL; ++r) {
                if (r < rpl) {
                  acc[r][0] = fmaf(x4[r].x, w0.x, acc[r][0]); acc[r][1] = fmaf(x4[r].x, w0.y, acc[r][1]);
                  acc[r][2] = fmaf(x4[r].x, w0.z, acc[r][2]); acc[r][3] = fmaf(x4[r].x, w0.w, acc[r][3]);
                  acc[r][0] = fmaf(x4[r].y, w1.x, acc[r][0]); acc[r][1] = fmaf(x4[r].y, w1.y, acc[r][1]);
                  acc[r][2] = fmaf(x4[r].y, w1.z, acc[r][2]); acc[r][3] = fmaf(x4[r].y, w1.w, acc[r][3]);
                  acc[r][0] = fmaf(x4[r].z, w2.x, acc[r][0]); acc[r][1] = fmaf(x4[r].z, w2.y, acc[r][1]);
                  acc[r][2] = fmaf(x4[r].z, w2.z, acc[r][2]); acc[r][3] = fmaf(x4[r].z, w2.w, acc[r][3]);
                  acc[r][0] = fmaf(x4[r].w, w3.x, acc[r][0]); acc[r][1] = fmaf(x4[r].w, w3.y, acc[r][1]);
                  acc[r][2] = fmaf(x4[r].w, w3.z, acc[r][2]); acc[r][3] = fmaf(x4[r].w, w3.w, acc[r][3]);
                }
              }
            }
          }
        }
        __syncthreads();
      }
      // fixed-order reduction over the 16 K slices: red[warp][n][NC] (aliases the x buffers)
      float* red = tr;
#pragma unroll
      for (int r = 0; r < kPMaxRPL; ++r) {
        const int n = ng + 8 * r;
        if (r < rpl && n < N && cg * 4 < NC)
          *reinterpret_cast<float4*>(red + ((size_t)warp * N + n) * NC + cg * 4) =
              make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      }
      __syncthreads();
      if (tid < N * a.UPC) {
        const int n = tid / a.UPC, u = tid - n * a.UPC;
        const int unit = cta * a.UPC + u;
        float g[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float s = 0.f;
          for (int w = 0; w < kPW; ++w) s += red[((size_t)w * N + n) * NC + q * a.UPC + u];
          g[q] = s + __ldg(a.lstm_bias + q * R + unit);
        }
        const int cr = s_crow[n];
        const float cp = cr >= 0 ? __ldcg(c_prev + (size_t)cr * R + unit) : 0.f;
        float cn, hn;
        lstm_cell(g[0], g[1], g[2], g[3], cp, &cn, &hn);
        c_new[(size_t)n * R + unit] = cn;
        h_new[(size_t)n * R + unit] = hn;
      }
    }
    if (!grid_barrier(a.bar, bar_target, G, &s_abort)) return;

    // ============================ phase B: [logits | q] = h' . [W_o | W_q] + bias
    if (has_B) {
      const int HLD = R + 4;
      float* hs = tr;                                  // [N][R+4]
      float* wB = tr + (size_t)N * HLD;                // [R][8]
      float* red = wB + (size_t)R * 8;                 // [kPW][N][8]
      for (int p = tid; p < N * (R / 4); p += kPT) {
        const int n = p / (R / 4), q = p - n * (R / 4);
        cp_async16(hs + (size_t)n * HLD + q * 4, h_new + (size_t)n * R + q * 4);
      }
      cp_async_commit();
      const int nblk = a.CB / 8;
      for (int blk = 0; blk < nblk; ++blk) {
        const int col0 = cta * a.CB + blk * 8;
        if (col0 >= LQ) break;
        for (int p = tid; p < R * 2; p += kPT) {
          const int kk = p >> 1, hlf = p & 1;
          const int col = col0 + hlf * 4;
          cp_async16_zfill(wB + (size_t)kk * 8 + hlf * 4, col < LQ ? a.outq + (size_t)kk * LQ + col : a.outq, col < LQ);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        float acc2[kPMaxRPL][2];
#pragma unroll
        for (int r = 0; r < kPMaxRPL; ++r) acc2[r][0] = acc2[r][1] = 0.f;
#pragma unroll
        for (int j = 0; j < R / kPW / 4; ++j) {
          const int kk = warp * (R / kPW) + j * 4;
          float2 w2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) w2[q] = *reinterpret_cast<const float2*>(wB + (size_t)(kk + q) * 8 + cg * 2);
#pragma unroll
          for (int r = 0; r < kPMaxRPL; ++r) {
            const int n = ng + 8 * r;
            if (r < rpl && n < N) {
              const float4 x = *reinterpret_cast<const float4*>(hs + (size_t)n * HLD + kk);
              acc2[r][0] = fmaf(x.x, w2[0].x, acc2[r][0]); acc2[r][1] = fmaf(x.x, w2[0].y, acc2[r][1]);
              acc2[r][0] = fmaf(x.y, w2[1].x, acc2[r][0]); acc2[r][1] = fmaf(x.y, w2[1].y, acc2[r][1]);
              acc2[r][0] = fmaf(x.z, w2[2].x, acc2[r][0]); acc2[r][1] = fmaf(x.z, w2[2].y, acc2[r][1]);
              acc2[r][0] = fmaf(x.w, w2[3].x, acc2[r][0]); acc2[r][1] = fmaf(x.w, w2[3].y, acc2[r][1]);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kPMaxRPL; ++r) {
          const int n = ng + 8 * r;
          if (r < rpl && n < N)
            *reinterpret_cast<float2*>(red + ((size_t)warp * N + n) * 8 + cg * 2) = make_float2(acc2[r][0], acc2[r][1]);
        }
        __syncthreads();
        if (tid < N * 8) {
          const int n = tid >> 3, cc = tid & 7;
          const int col = col0 + cc;
          if (col < LQ) {
            float s = 0.f;
            for (int w = 0; w < kPW; ++w) s += red[((size_t)w * N + n) * 8 + cc];
            a.lq[(size_t)n * LQ + col] = s + __ldg(a.outq_bias + col);
          }
        }
        __syncthreads();
      }
      cp_async_wait<0>();
    }
    if (!grid_barrier(a.bar, bar_target, G, &s_abort)) return;

    // ============================ phase C1: attention scores  ||  beam / greedy selection
    if (has_C) {
      float* sm_q = tr;                         // [k][R] centred queries (lane-permuted)
      float* sm_qg = tr + (size_t)k * R;        // [k][R] * gamma'
      float* sm_sqq = tr + 2 * (size_t)k * R;   // [k]
      for (int beam = warp; beam < k; beam += kPW) {
        const float* q = a.lq + (size_t)(img * k + beam) * LQ + a.q_off + c0l;
        float4 v[G4];
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < G4; ++g) {
          v[g] = ldcg4(q + g * 4);
          s += (v[g].x + v[g].y) + (v[g].z + v[g].w);
        }
        const float mean = wsum(s) * (1.0f / R);
        float sq = 0.f;
#pragma unroll
        for (int g = 0; g < G4; ++g) {
          float4 c = make_float4(v[g].x - mean, v[g].y - mean, v[g].z - mean, v[g].w - mean);
          *reinterpret_cast<float4*>(sm_q + (size_t)beam * R + (g * 32 + lane) * 4) = c;
          sq = fmaf(c.x, c.x, sq); sq = fmaf(c.y, c.y, sq); sq = fmaf(c.z, c.z, sq); sq = fmaf(c.w, c.w, sq);
          const float4 gm = *reinterpret_cast<const float4*>(sm_c + (g * 32 + lane) * 4);
          *reinterpret_cast<float4*>(sm_qg + (size_t)beam * R + (g * 32 + lane) * 4) =
              make_float4(c.x * gm.x, c.y * gm.y, c.z * gm.z, c.w * gm.w);
        }
        sq = wsum(sq);
        if (lane == 0) sm_sqq[beam] = sq;
      }
      __syncthreads();
      for (int r = warp; r < m1 - m0; r += kPW) {
        const int m = m0 + r;
        float kc[CPL], kg[CPL];
        float skk;
        if (a.keys_res) {
#pragma unroll
          for (int g = 0; g < G4; ++g) {
            const float4 tq = *reinterpret_cast<const float4*>(sm_keys + (size_t)r * R + (g * 32 + lane) * 4);
            kc[g * 4 + 0] = tq.x; kc[g * 4 + 1] = tq.y; kc[g * 4 + 2] = tq.z; kc[g * 4 + 3] = tq.w;
          }
          skk = sm_skk[r];
        } else {
          const float* kr = a.keys + ((size_t)img * M + m) * R + c0l;
          float s = 0.f;
#pragma unroll
          for (int g = 0; g < G4; ++g) {
            const float4 tq = ldg4(kr + g * 4);
            kc[g * 4 + 0] = tq.x; kc[g * 4 + 1] = tq.y; kc[g * 4 + 2] = tq.z; kc[g * 4 + 3] = tq.w;
            s += (tq.x + tq.y) + (tq.z + tq.w);
          }
          const float mean = wsum(s) * (1.0f / R);
          float sq = 0.f;
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            kc[c] -= mean;
            sq = fmaf(kc[c], kc[c], sq);
          }
          skk = wsum(sq);
        }
#pragma unroll
        for (int g = 0; g < G4; ++g) {
          const float4 gm = *reinterpret_cast<const float4*>(sm_c + (g * 32 + lane) * 4);
          kg[g * 4 + 0] = kc[g * 4 + 0] * gm.x; kg[g * 4 + 1] = kc[g * 4 + 1] * gm.y;
          kg[g * 4 + 2] = kc[g * 4 + 2] * gm.z; kg[g * 4 + 3] = kc[g * 4 + 3] * gm.w;
        }
        for (int beam0 = 0; beam0 < k; beam0 += KB) {
          const int bs = min(beam0, k - KB);
          float part[KB];
          ln_tanh_scores<CPL, KB, false>(kc, kg, sm_q + (size_t)bs * R, sm_qg + (size_t)bs * R, sm_c, sm_sqq + bs, skk,
                                         R, lane, sv, 1.0f / R, part);
#pragma unroll
          for (int o = LPH / 2; o; o >>= 1) {
#pragma unroll
            for (int j = 0; j < KB; ++j) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
          }
          if ((lane % LPH) == 0) {
            const int hh = lane / LPH;
#pragma unroll
            for (int j = 0; j < KB; ++j)
              a.scores[((size_t)(img * k + bs + j) * H + hh) * M + m] = part[j] * out_scale;
          }
        }
      }
    } else if (sel_cta) {
      const int si = cta - a.B * a.S;
      if (!a.greedy) {
        if (tid < 256)
          beam_step_block<true>(s_beam, tid, si, a.lq, LQ, k, a.V, a.eos, a.lpw, a.cum, a.fin, a.len,
                                a.sc + (size_t)t * N, a.step_ids + (size_t)t * N, a.parents + (size_t)t * N, a.tok,
                                a.src, a.fin_count, t);
      } else {
        if (tid < 128)
          greedy_step_block<true>(s_greedy, tid, si, a.lq, LQ, a.V, a.eos, a.step_ids + (size_t)t * N,
                                  a.logits_out ? a.logits_out + (size_t)t * N * a.V : nullptr, a.tok, a.fin,
                                  a.fin_count, t);
      }
    }
    if (!grid_barrier(a.bar, bar_target, G, &s_abort)) return;

    // ============================ phase C2: softmax over positions, history, context
    if (has_C) {
      const int dv = VAL / H;
      const bool active = ch0 < VAL;
      // heads this CTA needs: those its channel slice touches + those whose history it writes (h % S == slc)
      int heads[kPMaxHeads];
      int nh = 0;
      if (active)
        for (int hh = ch0 / dv; hh <= (ch1 - 1) / dv && nh < kPMaxHeads; ++hh) heads[nh++] = hh;
      const int n_touched = nh;
      if (a.hist)
        for (int hh = slc; hh < H && nh < kPMaxHeads; hh += a.S) {
          bool dup = false;
          for (int i = 0; i < n_touched; ++i) dup |= (heads[i] == hh);
          if (!dup) heads[nh++] = hh;
        }
      float* sm_al = tr;                                   // [k][nh_max][M]
      float* redc = tr + (size_t)k * a.nh_max * M;          // [groups][tasks] float4
      redc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(redc) + 15) & ~uintptr_t(15));
      for (int pr = warp; pr < k * nh; pr += kPW) {
        const int beam = pr / nh, hs_ = pr - beam * nh;
        const int hh = heads[hs_];
        const size_t grow = ((size_t)(img * k + beam) * H + hh) * M;
        float* s = sm_al + ((size_t)beam * a.nh_max + hs_) * M;
        float sum = 0.f;
        if (a.prob_fn == 0) {
          float mx = -INFINITY;
          for (int m = lane; m < M; m += 32) {
            const float x = __ldcg(a.scores + grow + m);
            s[m] = x;
            mx = fmaxf(mx, x);
          }
          mx = wmax(mx);
          for (int m = lane; m < M; m += 32) {
            float e = expf(s[m] - mx);
            s[m] = e;
            sum += e;
          }
        } else {
          for (int m = lane; m < M; m += 32) {
            float e = 1.0f / (1.0f + expf(-__ldcg(a.scores + grow + m)));
            s[m] = e;
            sum += e;
          }
        }
        sum = wsum(sum);
        const bool writes_hist = a.hist != nullptr && (hh % a.S) == slc;
        float* hrow = writes_hist ? a.hist + (size_t)t * N * H * M + grow : nullptr;
        for (int m = lane; m < M; m += 32) {
          const float al = s[m] / sum;
          s[m] = al;
          if (hrow) hrow[m] = al;
        }
      }
      __syncthreads();
      if (active) {
        const int nc4 = (ch1 - ch0) >> 2;
        const int ntask = k * nc4;                          // (beam, 4 channels)
        const int ngrp = kPT / ntask;                       // position groups
        const int task = tid % ntask, grp = tid / ntask;
        const int beam = task / nc4, c4 = task - beam * nc4;
        float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grp < ngrp) {
          const int hh = (ch0 + c4 * 4) / dv;
          int hs_ = 0;
          for (int i = 0; i < n_touched; ++i)
            if (heads[i] == hh) hs_ = i;
          const float* al = sm_al + ((size_t)beam * a.nh_max + hs_) * M;
          const int ma = (int)(((long long)M * grp) / ngrp), mb = (int)(((long long)M * (grp + 1)) / ngrp);
          for (int m = ma; m < mb; ++m) {
            const float4 v = a.vals_res ? *reinterpret_cast<const float4*>(sm_vals + (size_t)m * a.cps + c4 * 4)
                                        : ldg4(a.values + ((size_t)img * M + m) * VAL + ch0 + c4 * 4);
            const float w = al[m];
            acc4.x = fmaf(w, v.x, acc4.x); acc4.y = fmaf(w, v.y, acc4.y);
            acc4.z = fmaf(w, v.z, acc4.z); acc4.w = fmaf(w, v.w, acc4.w);
          }
          *reinterpret_cast<float4*>(redc + ((size_t)grp * ntask + task) * 4) = acc4;
        }
        __syncthreads();
        if (tid < ntask) {
          float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int g2 = 0; g2 < ngrp; ++g2) {
            const float4 p = *reinterpret_cast<const float4*>(redc + ((size_t)g2 * ntask + tid) * 4);
            s4.x += p.x; s4.y += p.y; s4.z += p.z; s4.w += p.w;
          }
          const int bm = tid / nc4, cc4 = tid - bm * nc4;
          *reinterpret_cast<float4*>(ctx_new + (size_t)(img * k + bm) * a.A + ch0 + cc4 * 4) = s4;
        }
      }
    }
    if (!grid_barrier(a.bar, bar_target, G, &s_abort)) return;

    // every row finished in this step -> the loop ends (dynamic_decode's all(finished))
    if (__ldcg(a.fin_count + t) >= N) break;
  }
}

// ---------------------------------------------------------------------------
// Host side: applicability, partition and shared-memory plan, cooperative launch.
// ---------------------------------------------------------------------------
struct PersistPlan {
  bool ok = false;
  int G, nA, UPC, CB, nB, S, rps, cps, nh_max, keys_res, vals_res;
  int off_wA, off_c, off_keys, off_skk, off_vals, off_tr;
  size_t smem_bytes;
};

static PersistPlan persist_plan(comic_handle_t h, int B, int k, bool greedy) {
  PersistPlan p;
  const int N = B * k, R = h->R, M = h->M, VAL = h->VAL, H = h->H;
  if (h->persist_max_rows <= 0 || N > h->persist_max_rows || N > 8 * kPMaxRPL) return p;
  if (R != 512 || h->cfg.alignment != 0 || h->cfg.context_layer || k > 16) return p;
  if (!(H == 1 || H == 2 || H == 4 || H == 8 || H == 16) || VAL % (4 * H) != 0 || h->KX % 4 != 0 || h->W % 4 != 0 ||
      h->A % 4 != 0 || h->LQ % 4 != 0 || h->Vp % 4 != 0 || h->A != VAL)
    return p;
  p.G = h->num_sms;
  const int n_sel = B;   // beam: one CTA per image; greedy: one per row (k == 1)
  if (2 * B > p.G || p.G - n_sel < B) return p;
  p.UPC = 1;
  while (R / p.UPC > p.G) p.UPC *= 2;
  if (p.UPC > 4) return p;                              // phase-A lanes own one 4-column group: 4*UPC <= 16
  p.nA = R / p.UPC;
  p.CB = round_up((h->LQ + p.G - 1) / p.G, 8);
  p.nB = (h->LQ + p.CB - 1) / p.CB;
  p.S = (p.G - n_sel) / B;
  if (p.S > M) p.S = M;
  p.rps = (M + p.S - 1) / p.S;
  p.S = (M + p.rps - 1) / p.rps;
  p.cps = round_up((VAL + p.S - 1) / p.S, 4);
  const int dv = VAL / H;
  p.nh_max = 1;
  for (int s = 0; s < p.S; ++s) {
    int ch0 = s * p.cps, ch1 = ch0 + p.cps < VAL ? ch0 + p.cps : VAL;
    int nh = 0;
    if (ch0 < VAL) nh = (ch1 - 1) / dv - ch0 / dv + 1;
    for (int hh = s; hh < H; hh += p.S)
      if (!(ch0 < VAL && hh >= ch0 / dv && hh <= (ch1 - 1) / dv)) ++nh;
    if (nh > p.nh_max) p.nh_max = nh;
  }
  if (p.nh_max > kPMaxHeads) return p;
  if (k * (p.cps / 4) > kPT) return p;
  // transient region: max over the phase layouts (floats)
  size_t trA = (size_t)2 * N * (kPKC + 4), trAr = (size_t)kPW * N * 4 * p.UPC;
  size_t trB = (size_t)N * (R + 4) + (size_t)R * 8 + (size_t)kPW * N * 8;
  size_t trC1 = (size_t)2 * k * R + 32;
  size_t trC2 = (size_t)k * p.nh_max * M + 4 + (size_t)kPT * 4;
  size_t tr = trA;
  if (trAr > tr) tr = trAr;
  if (trB > tr) tr = trB;
  if (trC1 > tr) tr = trC1;
  if (trC2 > tr) tr = trC2;
  const size_t budget = (size_t)(220 * 1024) / sizeof(float);
  for (int attempt = 0; attempt < 3; ++attempt) {
    p.keys_res = attempt < 2;
    p.vals_res = attempt < 1;
    size_t off = 0;
    p.off_wA = (int)off; off += (size_t)h->KX * 4 * p.UPC;
    p.off_c = (int)off; off += 3 * (size_t)R;
    p.off_keys = (int)off; off += p.keys_res ? (size_t)p.rps * R : 0;
    p.off_skk = (int)off; off += p.keys_res ? (size_t)round_up(p.rps, 4) : 0;
    p.off_vals = (int)off; off += p.vals_res ? (size_t)M * p.cps : 0;
    p.off_tr = (int)off; off += tr;
    if (off <= budget) {
      p.smem_bytes = off * sizeof(float);
      p.ok = true;
      return p;
    }
  }
  return p;
}

template <int H, int KB>
static cudaError_t launch_persist(const PersistArgs& pa, int G, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(decode_loop_kernel<H, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         222 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  void* args[] = {const_cast<PersistArgs*>(&pa)};
  return cudaLaunchCooperativeKernel((const void*)decode_loop_kernel<H, KB>, dim3(G), dim3(kPT), args, smem, st);
}

template <int H>
static cudaError_t launch_persist_h(const PersistArgs& pa, int G, size_t smem, cudaStream_t st) {
  if (pa.k == 1) return launch_persist<H, 1>(pa, G, smem, st);
  if (pa.k == 2 || pa.k == 4) return launch_persist<H, 2>(pa, G, smem, st);
  return launch_persist<H, 3>(pa, G, smem, st);
}

bool persist_applicable(comic_handle_t h, int B, int k, bool greedy) { return persist_plan(h, B, k, greedy).ok; }

// Returns 1 when the loop was enqueued as one cooperative launch, 0 when the configuration is
// not covered (the caller runs the per-step path), < 0 on error.
int decode_persistent(comic_handle_t h, const PersistCall& pc, cudaStream_t st) {
  PersistPlan p = persist_plan(h, pc.B, pc.k, pc.greedy != 0);
  if (!p.ok) return 0;
  PersistArgs a{};
  a.lstm_kernel = h->w.lstm_kernel; a.lstm_bias = h->w.lstm_bias;
  a.outq = h->pk.outq; a.outq_bias = h->pk.outq_bias; a.emb = h->w.embedding_map;
  a.gamma = h->w.ln_gamma; a.beta = h->w.ln_beta; a.vvec = h->w.attention_v; a.temperature = h->w.temperature;
  a.keys = pc.keys; a.values = pc.values; a.c0 = pc.c0; a.h0 = pc.h0;
  a.B = pc.B; a.k = pc.k; a.N = pc.B * pc.k; a.W = h->W; a.A = h->A; a.KX = h->KX; a.V = h->V; a.LQ = h->LQ;
  a.q_off = h->Vp; a.M = h->M; a.VAL = h->VAL; a.prob_fn = h->cfg.prob_fn; a.eos = h->cfg.eos_id;
  a.max_it = pc.max_it; a.greedy = pc.greedy; a.lpw = pc.lpw;
  for (int i = 0; i < 2; ++i) { a.c[i] = pc.c[i]; a.h[i] = pc.h[i]; a.ctx[i] = pc.ctx[i]; }
  a.lq = pc.lq; a.scores = pc.scores; a.hist = pc.hist; a.tok = pc.tok; a.src = pc.src; a.cum = pc.cum;
  a.fin = pc.fin; a.len = pc.len; a.fin_count = pc.fin_count; a.step_ids = pc.step_ids; a.parents = pc.parents;
  a.sc = pc.sc; a.logits_out = pc.logits_out; a.bar = pc.bar;
  a.nA = p.nA; a.UPC = p.UPC; a.CB = p.CB; a.nB = p.nB; a.S = p.S; a.rps = p.rps; a.cps = p.cps;
  a.nh_max = p.nh_max; a.keys_res = p.keys_res; a.vals_res = p.vals_res; a.n_sel = pc.B;
  a.off_wA = p.off_wA; a.off_c = p.off_c; a.off_keys = p.off_keys; a.off_skk = p.off_skk; a.off_vals = p.off_vals;
  a.off_tr = p.off_tr;
  COMIC_CHECK_CUDA(cudaMemsetAsync(pc.bar, 0, 2 * sizeof(unsigned), st));
  cudaError_t e = cudaErrorInvalidValue;
  {
    Prof pf(h, T_PERSIST, st);
    switch (h->H) {
      case 1: e = launch_persist_h<1>(a, p.G, p.smem_bytes, st); break;
      case 2: e = launch_persist_h<2>(a, p.G, p.smem_bytes, st); break;
      case 4: e = launch_persist_h<4>(a, p.G, p.smem_bytes, st); break;
      case 8: e = launch_persist_h<8>(a, p.G, p.smem_bytes, st); break;
      case 16: e = launch_persist_h<16>(a, p.G, p.smem_bytes, st); break;
      default: break;
    }
  }
  COMIC_CHECK_CUDA(e);
  return 1;
}

}  // namespace comic
