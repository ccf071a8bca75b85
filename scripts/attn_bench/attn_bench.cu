// Standalone check + timing of the attention kernels (development harness, not shipped):
//   attn_bench [B] [k] [iters]
// Fills random keys / queries / LN constants, runs attention.cuh's fused kernel (reference) and
// attention2.cuh's streaming kernel on the same inputs, prints max differences and CUDA-event times.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>
#define COMIC_A2_WATCHDOG 1
#ifndef HARNESS_GAMMA_SCALE
#define HARNESS_GAMMA_SCALE 1.0f
#endif
#include "../../comic-compact-image-captioning-with-attention_b200/csrc/attention2.cuh"

using namespace comic;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float urand() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (float)((rng_state >> 40) & 0xffffff) / 16777216.0f; }
static float nrand() { float u1 = urand() + 1e-7f, u2 = urand(); return sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2); }

int main(int argc, char** argv) {
  int B = argc > 1 ? atoi(argv[1]) : 512;
  int k = argc > 2 ? atoi(argv[2]) : 3;
  int iters = argc > 3 ? atoi(argv[3]) : 20;
  const int scale_mode = argc > 4 ? atoi(argv[4]) : 1;   // 1: history unnormalised + 1 / sum side buffer (decode loops); 0: rescaled in place
  const int R = 512, H = 8, M = 196, N = B * k, LQ = 512 + 256, QOFF = 256;
  printf("attn_bench B=%d k=%d\n", B, k);
  size_t nk = (size_t)B * M * R;
  std::vector<float> hk(nk), hq((size_t)N * LQ), hg(R), hb(R), hv(R);
  for (size_t i = 0; i < nk; ++i) hk[i] = 0.7f * nrand() + 0.3f;
  for (auto& x : hq) x = 0.8f * nrand() + 0.1f;
  for (int c = 0; c < R; ++c) { hg[c] = (1.0f + 0.2f * nrand()) * HARNESS_GAMMA_SCALE; hb[c] = 0.1f * nrand(); hv[c] = 0.2f * (urand() - 0.5f); }
  float hT = 5.0f;
  float *dk, *dq, *dg, *db, *dv, *dT, *dks, *dbound, *ctx0, *ctx1, *hist0, *hist1;
  CK(cudaMalloc(&dk, nk * 4)); CK(cudaMalloc(&dq, hq.size() * 4));
  CK(cudaMalloc(&dg, R * 4)); CK(cudaMalloc(&db, R * 4)); CK(cudaMalloc(&dv, R * 4)); CK(cudaMalloc(&dT, 4));
  CK(cudaMalloc(&dks, (size_t)B * M * 2 * 4)); CK(cudaMalloc(&dbound, 16 * 4));
  CK(cudaMalloc(&ctx0, (size_t)N * R * 4)); CK(cudaMalloc(&ctx1, (size_t)N * R * 4));
  CK(cudaMalloc(&hist0, (size_t)N * H * M * 4)); CK(cudaMalloc(&hist1, (size_t)N * H * M * 4));
  CK(cudaMemcpy(dk, hk.data(), nk * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dq, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg, hg.data(), R * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), R * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), R * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dT, &hT, 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(ctx1, 0xff, (size_t)N * R * 4)); CK(cudaMemset(hist1, 0xff, (size_t)N * H * M * 4));
  int dev = 0, sms = 148;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

  // ---- reference: attention.cuh ----
  AttnArgs aa{};
  aa.keys = dk; aa.values = dk; aa.lq = dq; aa.ld_lq = LQ; aa.q_off = QOFF; aa.gamma = dg; aa.beta = db; aa.vvec = dv;
  aa.temperature = dT; aa.ctx_out = ctx0; aa.ld_ctx = R; aa.hist_t = hist0; aa.k = k; aa.M = M; aa.VAL = R; aa.prob_fn = 0;
  aa.n_rows = N;
  size_t smem0 = attn_fused_smem(k, R, H, M, R);
  auto run_old = [&]() {
    if (k == 3) { CK(cudaFuncSetAttribute(attn_fused_kernel<512, 8, 0, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_fused_kernel<512, 8, 0, false, 3><<<B, kAttnThreads, smem0>>>(aa); }
    else if (k == 2) { CK(cudaFuncSetAttribute(attn_fused_kernel<512, 8, 0, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_fused_kernel<512, 8, 0, false, 2><<<B, kAttnThreads, smem0>>>(aa); }
    else { CK(cudaFuncSetAttribute(attn_fused_kernel<512, 8, 0, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attn_fused_kernel<512, 8, 0, false, 1><<<B, kAttnThreads, smem0>>>(aa); }
  };
  run_old();
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 3; ++i) run_old();
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) run_old();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms0; CK(cudaEventElapsedTime(&ms0, e0, e1));
  printf("old fused: %.2f us / launch\n", ms0 * 1000.f / iters);

  // ---- new: attention2.cuh ----
  CK(a2::launch_key_stats(dk, (long long)B * M, dks, dv, dT, dg, db, dbound, 0));
  CK(cudaDeviceSynchronize());
  float hbound[12];
  CK(cudaMemcpy(hbound, dbound, 48, cudaMemcpyDeviceToHost));
  printf("bound: %.3f %.3f ... %.3f  exponent shift %.0f feasible %.0f\n", hbound[0], hbound[1], hbound[7], hbound[8], hbound[9]);
  a2::Args a{};
  a.keys = dk; a.kstats = dks; a.bound = dbound; a.lq = dq; a.ld_lq = LQ; a.q_off = QOFF; a.gamma = dg; a.beta = db; a.vvec = dv;
  a.temperature = dT; a.ctx_out = ctx1; a.ld_ctx = R; a.hist_t = hist1; a.B = B; a.M = M; a.n_rows = N;
  float* dscratch; int* dcount;
  CK(cudaMalloc(&dscratch, a2::scratch_floats(sms) * 4)); CK(cudaMalloc(&dcount, (size_t)B * 4));
  CK(cudaMemset(dcount, 0, (size_t)B * 4));
  a.scratch = dscratch; a.counters = dcount;
  float* dscale = nullptr;
  if (scale_mode) { CK(cudaMalloc(&dscale, (size_t)N * H * 4)); a.hist_scale = dscale; }
#if COMIC_A2_TRACE
  const int kWarps = COMIC_A2_NSW + a2::kCtxWarps2 + 1 + a2::kStatWarps;
  size_t ntr = (size_t)sms * kWarps * a2::kTraceSlices * 8;
  long long* dtr;
  CK(cudaMalloc(&dtr, ntr * 8));
  CK(cudaMemset(dtr, 0, ntr * 8));
  a.trace = dtr;
#endif
  CK(a2::launch(a, k, sms, dev, 0));
  CK(cudaDeviceSynchronize());
  for (int i = 0; i < 3; ++i) CK(a2::launch(a, k, sms, dev, 0));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK(a2::launch(a, k, sms, dev, 0));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms1; CK(cudaEventElapsedTime(&ms1, e0, e1));
  printf("new attn2: %.2f us / launch  (NSW=%d STAGES=%d)\n", ms1 * 1000.f / iters, COMIC_A2_NSW, COMIC_A2_STAGES);
  CK(cudaEventRecord(e0));
  CK(a2::launch_key_stats(dk, (long long)B * M, dks, dv, dT, dg, db, dbound, 0));
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms2; CK(cudaEventElapsedTime(&ms2, e0, e1));
  printf("key_stats: %.2f us\n", ms2 * 1000.f);

#if COMIC_A2_TRACE
  {
    // one more launch into a clean trace buffer, then per-phase averages (cycles)
    CK(cudaMemset(dtr, 0, ntr * 8));
    CK(a2::launch(a, k, sms, dev, 0));
    CK(cudaDeviceSynchronize());
    std::vector<long long> tr(ntr);
    CK(cudaMemcpy(tr.data(), dtr, ntr * 8, cudaMemcpyDeviceToHost));
    const int grid = (long long)B * 49 < sms ? B * 49 : sms;
    double sc[5] = {0, 0, 0, 0, 0}; long long nsc = 0;
    double cx[3] = {0, 0, 0}; long long ncx = 0, nfin = 0; double fin = 0;
    long long span_max = 0; double span_sum = 0; double fz = 0; long long nfz = 0;
    for (int b = 0; b < grid; ++b) {
      long long tmin = (1ll << 62), tmax = 0;
      for (int w = 0; w < kWarps; ++w) {
        const long long* p = tr.data() + ((size_t)b * kWarps + w) * a2::kTraceSlices * 8;
        for (int i = 0; i < a2::kTraceSlices; ++i) {
          const long long* e = p + i * 8;
          if (w < COMIC_A2_NSW) {
            if (e[5] == 0) break;
            for (int q = 0; q < 5; ++q) sc[q] += (double)(e[q + 1] - e[q]);
            ++nsc;
            tmin = std::min(tmin, e[0]); tmax = std::max(tmax, e[5]);
          } else if (w == COMIC_A2_NSW + 1 || w == COMIC_A2_NSW + 2) {   // context warps (2 statistics warps layout)
            if (e[2] == 0 && e[3] == 0) break;
            if (e[2] != 0) { cx[0] += (double)(e[1] - e[0]); cx[1] += (double)(e[2] - e[1]); ++ncx; }
            if (e[3] != 0 && i > 0 && p[(i - 1) * 8 + 2] != 0) { fin += (double)(e[3] - p[(i - 1) * 8 + 2]); ++nfin; }
            tmax = std::max(tmax, std::max(e[2], e[3]));
          } else if (w == COMIC_A2_NSW + 3) {   // finaliser
            if (e[2] == 0) break;
            tmax = std::max(tmax, e[2]);   // finaliser: end of a segment's finalisation
            fz += (double)(e[2] - e[1]); ++nfz;
          }
        }
      }
      span_max = std::max(span_max, tmax - tmin); span_sum += (double)(tmax - tmin);
    }
    {
      // whole-warp residency (last trace record of every warp): entry -> end of prologue -> exit
      long long g0 = (1ll << 62), g1 = 0; double pro = 0, res = 0; long long res_max = 0; int res_max_w = -1, res_max_b = -1; long long n = 0;
      std::vector<double> by_w(kWarps, 0.0);
      for (int b = 0; b < grid; ++b)
        for (int w = 0; w < kWarps; ++w) {
          const long long* e = tr.data() + (((size_t)b * kWarps + w) * a2::kTraceSlices + (a2::kTraceSlices - 1)) * 8;
          if (e[2] == 0) continue;
          g0 = std::min(g0, e[3]); g1 = std::max(g1, e[4]);
          pro += (double)(e[1] - e[0]); res += (double)(e[2] - e[0]); by_w[w] += (double)(e[2] - e[0]); ++n;
          if (e[2] - e[0] > res_max) { res_max = e[2] - e[0]; res_max_w = w; res_max_b = b; }
        }
      printf("residency: kernel %lld ns on the global timer; per warp prologue %.0f, entry->exit mean %.0f max %lld cycles (CTA %d warp %d)\n",
             g1 - g0, pro / n, res / n, res_max, res_max_b, res_max_w);
      if (res_max_b >= 0) {
        const long long* p = tr.data() + ((size_t)res_max_b * kWarps + (kWarps - 1)) * a2::kTraceSlices * 8;
        const long long t0 = p[(a2::kTraceSlices - 1) * 8];
        printf("  finaliser of CTA %d, per segment (cycles after entry): wait-imgdone-from waited parked counted combined hist end\n", res_max_b);
        for (int i = 0; i < 6; ++i) {
          const long long* e = p + i * 8;
          if (e[0] == 0) break;
          auto rel = [&](long long v) { return v ? v - t0 : 0; };
          printf("    seg %d: %lld %lld %lld %lld %lld %lld %lld\n", i, rel(e[0]), rel(e[1]), rel(e[6]), rel(e[3]), rel(e[4]), rel(e[5]), rel(e[2]));
        }
      }
      {
        // per CTA on the global timer: entry of its first warp and exit of its last warp, ns after the first entry of the grid
        std::vector<long long> ent, ext, dur;
        for (int b = 0; b < grid; ++b) {
          long long e0 = (1ll << 62), e1 = 0;
          for (int w = 0; w < kWarps; ++w) {
            const long long* e = tr.data() + (((size_t)b * kWarps + w) * a2::kTraceSlices + (a2::kTraceSlices - 1)) * 8;
            if (e[2] == 0) continue;
            e0 = std::min(e0, e[3]); e1 = std::max(e1, e[4]);
          }
          if (e1) { ent.push_back(e0 - g0); ext.push_back(e1 - g0); dur.push_back(e1 - e0); }
        }
        auto pr = [&](const char* nm, std::vector<long long> v) {
          std::sort(v.begin(), v.end());
          const size_t n2 = v.size();
          printf("  %s (ns): min %lld p10 %lld p25 %lld p50 %lld p75 %lld p90 %lld max %lld\n", nm, v[0], v[n2 / 10], v[n2 / 4], v[n2 / 2],
                 v[3 * n2 / 4], v[9 * n2 / 10], v[n2 - 1]);
        };
        pr("CTA entry", ent); pr("CTA exit ", ext); pr("CTA resident", dur);
        printf("  CTA resident ns by block index (every 4th):");
        for (size_t b = 0; b < dur.size(); b += 4) printf(" %lld", dur[b]);
        printf("\n");
      }
      printf("  entry->exit mean by warp:");
      for (int w = 0; w < kWarps; ++w) printf(" %.0f", by_w[w] / grid);
      printf("\n");
    }
    {
      // distributions: per slice index of the warp, and bucketed
      const int NB = 8; const long long edge[NB] = {250, 500, 1000, 2000, 4000, 8000, 16000, 1ll << 60};
      long long hist[5][NB] = {};
      std::vector<double> byidx(a2::kTraceSlices * 5, 0.0); std::vector<long long> nidx(a2::kTraceSlices, 0);
      for (int b = 0; b < grid; ++b)
        for (int w = 0; w < COMIC_A2_NSW; ++w) {
          const long long* p = tr.data() + ((size_t)b * kWarps + w) * a2::kTraceSlices * 8;
          for (int i = 0; i < a2::kTraceSlices; ++i) {
            const long long* e = p + i * 8;
            if (e[5] == 0) break;
            for (int q = 0; q < 5; ++q) {
              const long long d = e[q + 1] - e[q];
              int bk = 0; while (d >= edge[bk]) ++bk;
              ++hist[q][bk];
              byidx[i * 5 + q] += (double)d;
            }
            ++nidx[i];
          }
        }
      printf("  buckets (<250 <500 <1k <2k <4k <8k <16k more):\n");
      const char* nm[5] = {"grab+setup", "wait key", "pass1", "pass2", "exp+store"};
      for (int q = 0; q < 5; ++q) { printf("    %-10s", nm[q]); for (int k2 = 0; k2 < NB; ++k2) printf(" %7lld", hist[q][k2]); printf("\n"); }
      printf("  by slice index of the warp (n | grab wait pass1 pass2 exp):\n");
      for (int i = 0; i < a2::kTraceSlices && nidx[i]; ++i)
        printf("    %2d %6lld | %6.0f %6.0f %6.0f %6.0f %6.0f\n", i, nidx[i], byidx[i * 5] / nidx[i], byidx[i * 5 + 1] / nidx[i],
               byidx[i * 5 + 2] / nidx[i], byidx[i * 5 + 3] / nidx[i], byidx[i * 5 + 4] / nidx[i]);
      // timeline of CTA 7: per warp, slice g and stamp times relative to the CTA's first stamp
      const int bb = 7 < grid ? 7 : 0;
      long long t0 = 1ll << 62;
      for (int w = 0; w < kWarps; ++w) { const long long* p = tr.data() + ((size_t)bb * kWarps + w) * a2::kTraceSlices * 8; if (p[0] && p[0] < t0) t0 = p[0]; }
      printf("  timeline CTA %d (warp: [g t0 grab wait p1 p2 exp] ...):\n", bb);
      for (int w = 0; w < COMIC_A2_NSW; ++w) {
        const long long* p = tr.data() + ((size_t)bb * kWarps + w) * a2::kTraceSlices * 8;
        printf("    w%02d:", w);
        for (int i = 0; i < a2::kTraceSlices; ++i) { const long long* e = p + i * 8; if (e[5] == 0) break;
          printf(" [%lld @%lld %lld %lld %lld %lld %lld]", e[6], e[0] - t0, e[1] - e[0], e[2] - e[1], e[3] - e[2], e[4] - e[3], e[5] - e[4]); }
        printf("\n");
      }
      for (int w = COMIC_A2_NSW; w < kWarps; ++w) {
        const long long* p = tr.data() + ((size_t)bb * kWarps + w) * a2::kTraceSlices * 8;
        printf("    w%02d raw:", w);
        for (int i = 0; i < a2::kTraceSlices; ++i) { const long long* e = p + i * 8; if (e[0] == 0 && e[1] == 0 && e[2] == 0 && e[3] == 0) break;
          printf(" [%lld %lld %lld %lld | %lld %lld %lld]", e[0] ? e[0] - t0 : 0, e[1] ? e[1] - t0 : 0, e[2] ? e[2] - t0 : 0, e[3] ? e[3] - t0 : 0,
                 e[6] ? e[6] - t0 : 0, e[4] ? e[4] - t0 : 0, e[5] ? e[5] - t0 : 0); }
        printf("\n");
      }
    }
    printf("trace (cycles, averages): CTA span mean %.0f max %lld\n", span_sum / grid, span_max);
    printf("  score warp per slice (%lld slices): grab+setup %.0f | wait key slice %.0f | pass1+stats %.0f | pass2 %.0f | exp+store+arrive %.0f\n",
           nsc, sc[0] / nsc, sc[1] / nsc, sc[2] / nsc, sc[3] / nsc, sc[4] / nsc);
    printf("  finaliser per segment (%lld): %.0f\n", nfz, nfz ? fz / nfz : 0.0);
    printf("  ctx warp per slice (%lld): wait scored %.0f | accumulate+release %.0f | image finalise %.0f (x%lld)\n", ncx, cx[0] / ncx,
           cx[1] / ncx, nfin ? fin / nfin : 0.0, nfin);
  }
#endif
  // ---- compare ----
  std::vector<float> c0((size_t)N * R), c1((size_t)N * R), h0((size_t)N * H * M), h1((size_t)N * H * M);
  CK(cudaMemcpy(c0.data(), ctx0, c0.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(c1.data(), ctx1, c1.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h0.data(), hist0, h0.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h1.data(), hist1, h1.size() * 4, cudaMemcpyDeviceToHost));
  if (scale_mode) {
    std::vector<float> sc((size_t)N * H);
    CK(cudaMemcpy(sc.data(), dscale, sc.size() * 4, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < (size_t)N * H; ++r) for (int m = 0; m < M; ++m) h1[r * M + m] *= sc[r];
  }
  double dc = 0, mc = 0, dh = 0, mh = 0, drel = 0;
  size_t bad = 0;
  for (size_t i = 0; i < c0.size(); ++i) { double d = fabs((double)c0[i] - c1[i]); if (!(d == d)) ++bad; dc = std::max(dc, d); mc = std::max(mc, (double)fabs(c0[i])); }
  for (size_t i = 0; i < h0.size(); ++i) { double d = fabs((double)h0[i] - h1[i]); if (!(d == d)) ++bad; dh = std::max(dh, d); mh = std::max(mh, (double)fabs(h0[i]));
    if (h0[i] > 1e-6) drel = std::max(drel, d / h0[i]); }
  printf("ctx:  max |diff| %.3e (max |ref| %.3e)\nhist: max |diff| %.3e (max |ref| %.3e), max rel diff %.3e, NaNs %zu\n", dc, mc, dh, mh, drel, bad);
  printf("%s\n", (bad == 0 && dc < 1e-4 * mc + 1e-6 && drel < 1e-4) ? "PARITY OK" : "PARITY FAIL");
  return 0;
}
