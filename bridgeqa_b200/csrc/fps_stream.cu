// fps_stream.cu -- furthest point sampling of a cell-sorted scene on ONE SM (sm_100a).
// Same contract and bit-exact result as fps.cu / fps_sorted.cu
// (replaces /root/reference/lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-173).
//
// fps_sorted.cu keeps a whole scene on chip -- 40k points need the registers / shared memory of
// three to six SMs, held for the ~2047 iterations of the chain although (after pruning) only a
// few per cent of the points are touched per iteration; with several batches in flight that
// SM-time is what bounds the step.  Here only what CHANGES lives on chip:
//   * the running min-distances of all points, in shared memory (4 bytes per point: 40k points
//     = 160 KB of one SM),
//   * per run of 32*PPL consecutive sorted points: its bounding box, its current best
//     (value, tie key, coordinates) -- in the registers of the run's owner thread.
// The coordinates {x, y, z, original index} stay where the ball query's cell grid put them
// (ball_query_grid.cu) and are read from L2 only for the runs the new sample can reach: the same
// warp-uniform box rule as fps_sorted.cu, but at a granularity of 64 points instead of 448-576,
// so a 40k-point scene touches ~1.6k points (25 KB) per iteration.  An L2 hit costs about what
// the DSMEM hop of the cluster kernels cost, and there is no cluster any more: one CTA, one
// __syncthreads per iteration.
//
// Per iteration (x1, y1, z1 = the newest sample, known to every thread):
//   1. every thread tests the box of the run it owns; a warp processes the runs of its lanes that
//      cannot be skipped, one after the other (runs are dealt round-robin over the warps, so the
//      few neighbouring runs a sample reaches land in different warps): 32 lanes x PPL points,
//      d = fma(dz,dz,fma(dx,dx,dy*dy)), temp = min(d, temp) back to shared memory, warp argmax by
//      (value, tie key) with redux.sync, result into the owner lane's registers;
//   2. warp argmax over the lanes' run records -> one 32-byte post per warp;
//   3. __syncthreads; every warp folds the posts (double-buffered, so no second barrier).
// Tie-break exactly as fps_sorted.cu: 27-bit ~tie_key(k) of the reference's pairwise tree.
//
// Measured and rejected (B200, 16 x 40k -> 2048; this kernel: 1.22 us per iteration with 20 warps):
//  * 32 warps, one run at a time: 1.39 us -- the per-iteration fold / box test of every warp is issue-bound;
//  * 16 warps, two owner slots per lane: 1.73 us -- more serial work on the busiest warp;
//  * a single-run fast path with a lazily computed tie key (ballot + shuffle instead of the second
//    redux): 1.58 us;
//  * a shared-memory WORK LIST (owners only test boxes; ballot -> atomic slot -> every warp takes one
//    entry, so the ~21 active runs are processed by 21 warps; second CTA barrier, records through shared
//    memory): bit-exact, 1.30 us with 20 warps, 1.54 us with 32 -- the busiest warp's 2-3 runs are not
//    what bounds an iteration: with ~4500 warp-instructions per iteration the SM's four schedulers are
//    ~50 % busy (ncu) and the rest is the dependent chain fold -> box test -> L2 load -> reduce -> post;
//  * TWO scenes per CTA (min-distances in global memory too -- the grid's spent cell_of words --, 16 warps
//    and a named barrier per scene, so two chains share the schedulers): bit-exact; with 128-point runs
//    2.21 us per iteration for both scenes = 36 SM-ms per batch instead of 40 at 1.8x the latency of a
//    call; with 64-point runs and two owner slots per lane 2.66 us = 43 SM-ms.  Not worth the latency.
//  * SEVERAL SAMPLES PER ROUND: take the 8 best run records (two posted per warp, exact up to the first
//    second-record), let warp j compute -- read-only -- the best of run(c_j) after c_j itself, and certify the
//    prefix c_1 .. c_s in which no accepted c_j can change run(c_i) (the box rule) and every such U_j ranks below
//    c_i; then apply the s samples in one pass.  tools/fps_batch_sim.py replays the rule on the CPU: 4.7 samples
//    per round, sequence identical to the oracle's.  Implemented, bit-exact on every size on the first run
//    (2047 samples in 506-512 rounds), and SLOWER: 2.99-3.06 ms against 2.50 (phase trace per round, 12.4-13.3k
//    cycles: box tests 0.7k, the one-pass update 4.5k for ~4 runs per warp plus ~3k at the barrier for the warp
//    that got 7, posts 0.6k, ranking 40 posts ~1.7k, speculation ~1.5k, certification 0.15k on one warp).  The
//    reason is the one above: the update of the ~20 runs a sample reaches is ~2.5k warp-instructions either
//    way and the SM issues at ~55 % of its rate on this dependent code, so only the ~2k instructions of
//    fold / box test / post per sample can be amortised -- at best ~20 %, which the serial rank + speculate +
//    certify section (8 of 20 warps busy) eats.  Removed again; the replay tool stays.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace bqa {

const float4 *ball_query_grid_sorted(const void *grid, int b, int n);   // ball_query_grid.cu

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) Post {      // 32-byte slot
  uint32_t vb;                   // 0 = no selectable point, else float bits + 1
  uint32_t nk;                   // (27-bit ~tie key) << 5; larger wins
  float x, y, z;
  uint32_t pad[3];
};

// 27-bit tie key for bs = 512: (bitrev9(k & 511) << 18) | (k >> 9); smaller wins.  k < 2^27.
__device__ __forceinline__ uint32_t nkey_of(uint32_t k) {
  // (~key & 0x7ffffff) << 5 with key = bitrev9(k & 511) << 18 | k >> 9: the reversed low 9 bits are the top 9
  // bits of brev(k), and (k >> 9) << 5 = (k >> 4) with its low 5 bits cleared
  return ((__brev(k) & 0xff800000u) | ((k >> 4) & 0x007fffe0u)) ^ 0xffffffe0u;
}
__device__ __forceinline__ uint32_t index_of(uint32_t nk) {
  const uint32_t key = ~(nk >> 5) & 0x7ffffffu;
  return ((key & 0x3ffffu) << 9) | (__brev(key >> 18) >> 23);
}

#ifdef BQA_FPS_STATS
__device__ unsigned long long g_stream_active_runs;
#endif

// -DBQA_SA_TRACE (tools/build_stats.sh): per-warp cycle sums of the phases of CTA 0
#ifdef BQA_SA_TRACE
__device__ long long g_stream_trace[32][8];
__device__ long long g_stream_trace2[8];
#define ST_T(i) const long long tt##i = clock64();
#define ST_ACC(k, a, b) tr[k] += tt##b - tt##a;
#else
#define ST_T(i)
#define ST_ACC(k, a, b)
#endif

// Two runs of 32 * PPL points against the newest sample, phase by phase so that their dependent
// chains (distance, min, two warp reductions each) overlap: new min-distances back to shared memory,
// each run's best (value bits + 1, tie key) warp-uniform in (wv[u], wk[u]), its coordinates into
// res[r[u]].  has_b is warp-uniform; without a second run its slot is computed on zeros and dropped.
template <int PPL>
__device__ __forceinline__ void stream_pair(bool has_b, const int (&r)[2], const float4 (&v)[2][PPL], float x1,
                                            float y1, float z1, float *td, float4 *res, int lane, uint32_t nk0,
                                            uint32_t (&wv)[2], uint32_t (&wk)[2]
#ifdef BQA_SA_TRACE
                                            , long long *ptr
#endif
                                            ) {
#ifdef BQA_SA_TRACE
  const long long pa = clock64();
#endif
  float t[2][PPL];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int p = 0; p < PPL; ++p) t[u][p] = (u == 0 || has_b) ? td[r[u] * (32 * PPL) + p * 32 + lane] : -INFINITY;
  uint32_t vb[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    float best = -INFINITY;
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
      const float dx = v[u][p].x - x1, dy = v[u][p].y - y1, dz = v[u][p].z - z1;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      const float d2 = fminf(d, t[u][p]);                  // -inf (skip set, padding) stays -inf
      t[u][p] = d2;
      best = fmaxf(best, d2);
    }
    vb[u] = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
  }
#pragma unroll
  for (int u = 0; u < 2; ++u)
    if (u == 0 || has_b)
#pragma unroll
      for (int p = 0; p < PPL; ++p) td[r[u] * (32 * PPL) + p * 32 + lane] = t[u][p];
#ifdef BQA_SA_TRACE
  const long long pb = clock64();
#endif
  wv[0] = __reduce_max_sync(kFull, vb[0]);
  wv[1] = __reduce_max_sync(kFull, vb[1]);
#ifdef BQA_SA_TRACE
  const long long pc_ = clock64();
#endif
  uint32_t cand[2];
  int pc[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const float wf = __uint_as_float(wv[u] - 1u);          // wv == 0: NaN pattern, matches nothing
    cand[u] = 0u; pc[u] = 0;
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
      const uint32_t k = (wv[u] && t[u][p] == wf) ? nkey_of((uint32_t)__float_as_int(v[u][p].w)) : 0u;
      if (k > cand[u]) { cand[u] = k; pc[u] = p; }
    }
  }
  wk[0] = __reduce_max_sync(kFull, cand[0]);
  wk[1] = __reduce_max_sync(kFull, cand[1]);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    if (wv[u]) {                                           // else res[r] keeps point 0 (a run without
      if (cand[u] == wk[u]) {                              // selectable points never gets one);
        float4 c = v[u][0];                                // exactly one lane: keys are distinct, non-zero
#pragma unroll
        for (int p = 1; p < PPL; ++p)
          if (pc[u] == p) c = v[u][p];
        res[r[u]] = c;
      }
    } else {
      wk[u] = nk0;
    }
  }
#ifdef BQA_SA_TRACE
  const long long pd = clock64();
  ptr[0] += pb - pa; ptr[1] += pc_ - pb; ptr[2] += pd - pc_; ptr[3] += 1;
#endif
}

// Launch bound: 64-point runs need at most 27 warps (835 runs, ~53k points).  Compiled for 1024 threads the kernel
// is capped at 64 registers and measures 2.7 % slower (2.484 vs 2.419 ms at 16 x 40000) than with this bound
// (68 registers); a bound of 640 -- what a 40k-point scene launches -- gave 69 registers and 2.478 ms.
template <int PPL>
__global__ void __launch_bounds__(PPL == 2 ? 864 : 1024, 1)
fps_stream_kernel(int n, int m, int nr, const float4 *__restrict__ sorted_all,
                  const float *__restrict__ xyz_all, int *__restrict__ idx_all,
                  float *__restrict__ new_xyz_all) {
  constexpr int kRun = 32 * PPL;
  extern __shared__ float4 dyn[];        // res[1024] then td[nr * kRun]
  float4 *res = dyn;                     // coordinates of every run's current best point
  float *td = reinterpret_cast<float *>(dyn + 1024);       // running min-distances, run-major
  __shared__ Post part[2][32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nw = blockDim.x >> 5;
  const int scene = blockIdx.x;
  const float4 *sorted = sorted_all + (size_t)scene * n;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  int *idxs = idx_all + (size_t)scene * m;
  float *new_xyz = new_xyz_all ? new_xyz_all + (size_t)scene * m * 3 : nullptr;

  float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];            // idxs[0] = 0, sampling_gpu.cu:85-86
  const uint32_t nk0 = nkey_of(0u);

  // Lane l of warp w owns run l * nw + w: runs are dealt round-robin over the warps, so the few
  // neighbouring runs a sample reaches land in different warps.
  const int my_run = lane * nw + wid;
  const bool own = my_run < nr;
  float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f;
  // "nothing selectable here": loses to every real candidate; if the whole scene is like that the
  // answer is index 0 with point 0's coordinates (sampling_gpu.cu: besti stays 0)
  uint32_t rvb = 0u, rnk = own ? nk0 : 0u;
  float rmax = -INFINITY;                                  // the run's current max min-distance

  // ---- boxes and initial min-distances ------------------------------------------------------
  for (int o = 0; o < 32; ++o) {
    const int r = o * nw + wid;
    if (r >= nr) break;                                    // warp-uniform
    float bl[3] = {INFINITY, INFINITY, INFINITY}, bh[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
      const int q = r * kRun + p * 32 + lane;
      float t = -INFINITY;                                 // -inf: never updated, never selected
      if (q < n) {
        const float4 v = sorted[q];
        const float mag = __fmaf_rn(v.z, v.z, __fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y)));
        if (!((double)mag <= 1e-3)) t = 1e10f;             // sampling_gpu.cu:100-101, sampling.cpp:74-76
        bl[0] = fminf(bl[0], v.x); bh[0] = fmaxf(bh[0], v.x);   // fminf/fmaxf drop NaNs
        bl[1] = fminf(bl[1], v.y); bh[1] = fmaxf(bh[1], v.y);
        bl[2] = fminf(bl[2], v.z); bh[2] = fmaxf(bh[2], v.z);
      }
      td[q] = t;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
        bl[a] = fminf(bl[a], __shfl_xor_sync(kFull, bl[a], sh));
        bh[a] = fmaxf(bh[a], __shfl_xor_sync(kFull, bh[a], sh));
      }
    }
    if (lane == o) { lox = bl[0]; loy = bl[1]; loz = bl[2]; hix = bh[0]; hiy = bh[1]; hiz = bh[2]; }
    if (lane == 0) res[r] = make_float4(x1, y1, z1, 0.f);
  }
  if (tid == 0 && m > 0) {
    idxs[0] = 0;
    if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
  }
  if (lane == 0) {                                         // both post buffers valid from the start
    Post q;
    q.vb = 0u; q.nk = nk0; q.x = x1; q.y = y1; q.z = z1; q.pad[0] = q.pad[1] = q.pad[2] = 0u;
    part[0][wid] = q; part[1][wid] = q;
  }
  __syncwarp();                                            // td[] / res[] of a run are only touched by its own warp
  int dirty = 0;                                           // post buffers that still hold an older record

#ifdef BQA_SA_TRACE
  long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long ptr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
  for (int j = 1; j < m; ++j) {
    ST_T(0)
    // ---- 1. which of my warp's runs can the new sample change? --------------------------------
    // (lb * (1 - 1e-5) >= the run's max  =>  min(d, temp) == temp for all its points, see
    //  fps_sorted.cu; the first iteration evaluates every run, whatever its box)
    const float ex = fmaxf(fmaxf(lox - x1, x1 - hix), 0.f);
    const float ey = fmaxf(fmaxf(loy - y1, y1 - hiy), 0.f);
    const float ez = fmaxf(fmaxf(loz - z1, z1 - hiz), 0.f);
    const float lb = ex * ex + ey * ey + ez * ez;
    const bool act = own && (j == 1 || !(lb * 0.99999f >= rmax));
    unsigned todo = __ballot_sync(kFull, act);
#ifdef BQA_FPS_STATS
    if (lane == 0 && todo) atomicAdd(&g_stream_active_runs, (unsigned long long)__popc(todo));
#endif
    ST_T(1)
#ifdef BQA_SA_TRACE
    tr[5] += __popc(todo);
    if (todo) tr[6] += 1;
#endif
    if (todo) dirty = 2;
    // two runs at a time (their dependent chains -- L2 load, distance, two warp reductions -- overlap)
    while (todo) {
      const int oa = __ffs(todo) - 1;
      todo &= todo - 1;
      const int ob = todo ? __ffs(todo) - 1 : -1;
      todo &= todo - 1;                                    // 0 stays 0
      const int r2[2] = {oa * nw + wid, (ob >= 0 ? ob : oa) * nw + wid};
      float4 v2[2][PPL];
#pragma unroll
      for (int p = 0; p < PPL; ++p) {
        const int qa = r2[0] * kRun + p * 32 + lane, qb = r2[1] * kRun + p * 32 + lane;
        v2[0][p] = qa < n ? __ldg(&sorted[qa]) : make_float4(0.f, 0.f, 0.f, 0.f);
        v2[1][p] = (ob >= 0 && qb < n) ? __ldg(&sorted[qb]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      uint32_t wv[2], wk[2];
#ifdef BQA_SA_TRACE
      const long long q0 = clock64();
      stream_pair<PPL>(ob >= 0, r2, v2, x1, y1, z1, td, res, lane, nk0, wv, wk, ptr);
      ptr[4] += q0 - tt1;
#else
      stream_pair<PPL>(ob >= 0, r2, v2, x1, y1, z1, td, res, lane, nk0, wv, wk);
#endif
      if (lane == oa) { rvb = wv[0]; rnk = wk[0]; rmax = wv[0] ? __uint_as_float(wv[0] - 1u) : -INFINITY; }
      if (lane == ob) { rvb = wv[1]; rnk = wk[1]; rmax = wv[1] ? __uint_as_float(wv[1] - 1u) : -INFINITY; }
    }
    ST_T(2)
    // ---- 2. the warp's best run -> its post (only while a buffer still holds an older record) ---
    if (dirty) {
      --dirty;
      const uint32_t mv = __reduce_max_sync(kFull, rvb);
      const uint32_t mk = __reduce_max_sync(kFull, rvb == mv ? rnk : 0u);
      __syncwarp();                                        // res[] written by other lanes of this warp
      if (rvb == mv && rnk == mk) {                        // one lane, or several with identical records
        const float4 c = res[my_run];
        Post q;
        q.vb = mv; q.nk = mk; q.x = c.x; q.y = c.y; q.z = c.z;
        q.pad[0] = q.pad[1] = q.pad[2] = 0u;
        part[j & 1][wid] = q;
      }
    }
    ST_T(3)
    __syncthreads();
    ST_T(4)
    // ---- 3. fold the posts ---------------------------------------------------------------------
    {
      const Post *pb = part[j & 1];
      uint2 c = lane < nw ? *reinterpret_cast<const uint2 *>(&pb[lane]) : make_uint2(0u, 0u);
      const uint32_t mv = __reduce_max_sync(kFull, c.x);
      const uint32_t mk = __reduce_max_sync(kFull, c.x == mv ? c.y : 0u);
      const int w = __ffs(__ballot_sync(kFull, c.x == mv && c.y == mk && lane < nw)) - 1;
      const Post *win = &pb[w];
      const float2 xy = *reinterpret_cast<const float2 *>(&win->x);
      x1 = xy.x; y1 = xy.y; z1 = win->z;
      if (tid == 0) {
        idxs[j] = (int)index_of(mk);                       // sampling_gpu.cu:170-171
        if (new_xyz) { new_xyz[j * 3 + 0] = x1; new_xyz[j * 3 + 1] = y1; new_xyz[j * 3 + 2] = z1; }
      }
    }
    ST_T(5)
    ST_ACC(0, 0, 1) ST_ACC(1, 1, 2) ST_ACC(2, 2, 3) ST_ACC(3, 3, 4) ST_ACC(4, 4, 5)
  }
#ifdef BQA_SA_TRACE
  if (blockIdx.x == 0 && lane == 0)
    for (int i = 0; i < 8; ++i) g_stream_trace[wid][i] = tr[i];
  if (blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 8; ++i) g_stream_trace2[i] = ptr[i];
#endif
}

constexpr size_t kStreamResBytes = 1024 * sizeof(float4);
constexpr size_t kStreamSmemMax = 227 * 1024 - 2 * 32 * sizeof(Post) - kStreamResBytes - 256;

// warps per CTA: every run needs an owner lane (nr <= 32 * warps); more warps spread the runs a
// sample reaches over more schedulers, fewer keep the per-iteration fold and box tests short
int stream_warps(int nr) {
  static const int forced = [] { const char *e = getenv("BQA_FPS_STREAM_WARPS"); return e ? atoi(e) : 0; }();
  int w = ceil_div(nr, 32);
  if (w < 8) w = 8;
  if (forced && forced >= w && forced <= 27) w = forced;   // (27 warps: the launch bound of the 64-point-run kernel)
  return w;
}

template <int PPL>
int launch_stream(int b, int n, int m, const float4 *sorted, const float *xyz, int *idxs, float *new_xyz,
                  cudaStream_t stream) {
  auto *kernel = fps_stream_kernel<PPL>;
  const int nr = ceil_div(n, 32 * PPL);
  const int warps = stream_warps(nr);
  const size_t smem = kStreamResBytes + (size_t)nr * 32 * PPL * sizeof(float);
  BQA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef BQA_FPS_STATS
  unsigned long long zero = 0;
  cudaMemcpyToSymbol(g_stream_active_runs, &zero, sizeof(zero));
#endif
  kernel<<<b, warps * 32, smem, stream>>>(n, m, nr, sorted, xyz, idxs, new_xyz);
#ifdef BQA_FPS_STATS
  unsigned long long act = 0;
  cudaMemcpyFromSymbol(&act, g_stream_active_runs, sizeof(act));
  fprintf(stderr, "[bqa fps stream] b=%d n=%d m=%d run=%d: %.1f active runs (%.0f points) per scene-iteration\n",
          b, n, m, 32 * PPL, (double)act / ((double)b * (m - 1)), 32.0 * PPL * act / ((double)b * (m - 1)));
#endif
#ifdef BQA_SA_TRACE
  {
    long long t[32][8];
    cudaMemcpyFromSymbol(t, g_stream_trace, sizeof(t));
    const double it = m > 1 ? m - 1 : 1;
    long long mx1 = 0, mx5 = 0;
    for (int w = 0; w < warps; ++w) { mx1 = t[w][1] > mx1 ? t[w][1] : mx1; mx5 = t[w][5] > mx5 ? t[w][5] : mx5; }
    fprintf(stderr, "[bqa fps stream trace] n=%d run=%d warps=%d | warp 0 cycles/iter: box %.0f process %.0f post %.0f barrier %.0f fold %.0f"
            " | runs/iter warp0 %.2f, busiest warp %.2f (its process cycles/iter %.0f); warp0 busy in %.0f %% of iterations\n",
            n, 32 * PPL, warps, t[0][0] / it, t[0][1] / it, t[0][2] / it, t[0][3] / it, t[0][4] / it, t[0][5] / it, mx5 / it,
            mx1 / it, 100.0 * t[0][6] / it);
    long long t2[8];
    cudaMemcpyFromSymbol(t2, g_stream_trace2, sizeof(t2));
    const double np = t2[3] ? (double)t2[3] : 1.0;
    fprintf(stderr, "[bqa fps stream trace] warp 0 per pair (%.0f pairs): setup+issue %.0f | load+distance %.0f | 2 redux %.0f | keys+2 redux+res %.0f cycles\n",
            np, t2[4] / np, t2[0] / np, t2[1] / np, t2[2] / np);
  }
#endif
  count_launch();
  return check_launch("fps_stream_kernel");
}

int stream_ppl(int n) {
  // shortest runs whose count fits one owner thread each and whose min-distances fit shared memory
  static const int forced = [] { const char *e = getenv("BQA_FPS_STREAM_PPL"); return e ? atoi(e) : 0; }();
  for (int ppl = 1; ppl <= 2; ppl *= 2) {
    if (forced && ppl != forced) continue;
    const int nr = ceil_div(n, 32 * ppl);
    if (nr <= 1024 && (size_t)nr * 32 * ppl * sizeof(float) <= kStreamSmemMax) return ppl;
  }
  return 0;
}

}  // namespace

// scenes the one-SM kernel takes: the tie key assumes the reference's block size 512 (n >= 512), and
// the running min-distances of the scene must fit one SM's shared memory (n <= ~53k)
bool fps_stream_supported(int n, int m) { return n >= 512 && m >= 1 && stream_ppl(n) != 0; }

int fps_stream_dispatch(int b, int n, int m, const float *xyz, const void *grid, int *idxs,
                        float *new_xyz, cudaStream_t stream) {
  const float4 *sorted = ball_query_grid_sorted(grid, b, n);
  switch (stream_ppl(n)) {
    case 1: return launch_stream<1>(b, n, m, sorted, xyz, idxs, new_xyz, stream);
    case 2: return launch_stream<2>(b, n, m, sorted, xyz, idxs, new_xyz, stream);
  }
  return set_error(BQA_ERR_UNSUPPORTED, "fps (stream): n=%d not supported", n);
}

}  // namespace bqa
