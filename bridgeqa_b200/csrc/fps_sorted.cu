// fps_sorted.cu -- furthest point sampling over a CELL-SORTED copy of the scene, with
// warp-level pruning (sm_100a).  Same contract and bit-exact result as fps.cu
// (replaces /root/reference/lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-173).
//
// fps.cu updates all n running min-distances in every one of the npoint-1 iterations, like
// the reference.  After the first few dozen samples almost none of them can change: a new
// sample s only lowers temp[k] for points closer to s than their current min-distance, and
// those lie in a small ball around s.  Here the points come in the ball query's cell-sorted
// order (ball_query_grid.cu: {x, y, z, original index} sorted by cell, built anyway for the
// grouping), so the 32*P points a warp keeps in registers are spatial neighbours with a
// tight bounding box, and per iteration a warp first asks one warp-uniform question:
//     lb = squared distance from the new sample to my box;   lb * (1 - 1e-5) >= my current max ?
// If yes, every fp32 distance the update would compute is >= every temp it would be min-ed
// with (d_fp32 >= D_true * (1 - 2^-21) >= lb_true * (1 - 2^-21) and lb is itself evaluated to
// within 2^-22), so min(d, temp) == temp for all 32*P points: nothing changes, the warp's posted
// candidate is still its argmax, and the warp goes straight to the barrier.  Typically 3-8 % of
// the warps do real work per iteration.  Consecutive runs of the sorted order are dealt to the
// cluster's CTAs round-robin so the few active warps land on different SMs.
//
// Everything else is fps.cu's chain: warp candidate -> CTA fold by warp 0 -> st.async to every
// CTA of the cluster -> fold of the cs candidates.  Because points are no longer held in index
// order, the tie-break is carried explicitly: every point keeps nkey = ~tie_key(k) (the
// reference's pairwise tree keeps the lower slot: minimal (bitrev9(k mod 512), k) wins), a
// thread's/warp's candidate is the lexicographic max of (value bits, nkey), and the four low
// bits of the key word carry the posting warp / CTA so the coordinates of the winner are found
// without another vote.
//
// Measured and rejected: replacing the reduction tree by a table of every warp's candidate
// replicated in every CTA (active warps push {value, key, xyz} straight into all copies with
// st.async; every CTA derives the number of pushes to expect from the same skip test; every warp
// folds the whole table; a second mbarrier hand-shake keeps pushes of iteration j+1 out of tables
// still being folded for iteration j).  Bit-exact and deadlock-free on all tests, but 1.15 us per
// iteration against 0.61 us here (2.35 vs 1.26 ms): 144 small DSMEM stores plus the hand-shake per
// iteration cost more than the CTA barrier + warp-0 fold they remove.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace bqa {

const float4 *ball_query_grid_sorted(const void *grid, int b, int n);   // ball_query_grid.cu
bool fps_stream_supported(int n, int m);                                 // fps_stream.cu
int fps_stream_dispatch(int b, int n, int m, const float *xyz, const void *grid, int *idxs,
                        float *new_xyz, cudaStream_t stream);

namespace {

constexpr int kT = 512;
constexpr int kNW = kT / 32;
constexpr int kMaxCluster = 16;
constexpr int kMaxP = 20;
constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) Cand {      // 32-byte slot, two 16-byte halves (st.async v4 + b32)
  uint32_t vb;                   // 0 = no selectable point, else float bits + 1
  uint32_t nk;                   // (27-bit ~tie key) << 5 | slot or poster id; larger wins
  float x, y;
  float z;
  uint32_t pad[3];
};

// 27-bit tie key for bs = 512: (bitrev9(k & 511) << 18) | (k >> 9); smaller wins.  k < 2^27.
__device__ __forceinline__ uint32_t nkey_of(uint32_t k) {
  // (~key & 0x7ffffff) << 5: the reversed low 9 bits are the top 9 bits of brev(k), and (k >> 9) << 5 is
  // (k >> 4) with its low 5 bits cleared -- three instructions less per point than the literal form
  return ((__brev(k) & 0xff800000u) | ((k >> 4) & 0x007fffe0u)) ^ 0xffffffe0u;
}
__device__ __forceinline__ uint32_t index_of(uint32_t nk) {
  const uint32_t key = ~(nk >> 5) & 0x7ffffffu;
  return ((key & 0x3ffffu) << 9) | (__brev(key >> 18) >> 23);
}

#ifdef BQA_FPS_STATS
__device__ unsigned long long g_active_warp_iters;
#endif

// kLean = false: coordinates, running min-distances and tie keys in registers (125 registers x 512
// threads: the CTA owns its SM), shared memory only holds a copy for the winner's lookup -- the
// lowest latency per iteration.
// kLean = true (the throughput variant, T = 768): only the running min-distances stay in
// registers; coordinates and tie keys are read from the shared-memory copy {x, y, z, key} by the
// few warps whose box the new sample can reach.  A 40k-point scene then fits THREE SMs (13824
// points x 16 bytes = 221 KB each) instead of six, at ~35 % more latency per iteration: with
// several batches in flight (graphs.InFlight) the chain is not the bottleneck, SM-time is.
template <int P, int T, bool kLean>
__global__ void __launch_bounds__(T, 1)
fps_sorted_kernel(int n, int m, int cs, const float4 *__restrict__ sorted_all,
                  const float *__restrict__ xyz_all, int *__restrict__ idx_all,
                  float *__restrict__ new_xyz_all) {
  constexpr int kT = T, kNW = T / 32;
  constexpr int PR = kLean ? 1 : P;      // register copies of coordinates / keys (unused when lean)
  __shared__ Cand recv[2][kMaxCluster];
  __shared__ Cand part[kNW];
  __shared__ __align__(8) uint64_t bars[2];
  extern __shared__ float4 pts[];        // [P][kT] coordinates (+ tie key in .w when lean): the winner's lookup by slot

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / cs;
  const float4 *sorted = sorted_all + (size_t)scene * n;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  int *idxs = idx_all + (size_t)scene * m;
  float *new_xyz = new_xyz_all ? new_xyz_all + (size_t)scene * m * 3 : nullptr;

  // run of 32*P consecutive sorted points per warp; runs dealt round-robin to the CTAs
  const int run = wid * cs + (int)rank;
  const int base = run * (32 * P);
  float px[PR], py[PR], pz[PR], td[P];
  uint32_t nkey[PR];
  float lox = INFINITY, loy = INFINITY, loz = INFINITY, hix = -INFINITY, hiy = -INFINITY, hiz = -INFINITY;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int s = base + p * 32 + lane;
    float x = 0.f, y = 0.f, z = 0.f, t = -INFINITY;      // -inf: never updated, never selected
    uint32_t nk = 0u;
    if (s < n) {
      const float4 v = sorted[s];
      x = v.x; y = v.y; z = v.z;
      nk = nkey_of((uint32_t)__float_as_int(v.w)) | (uint32_t)p;      // low 5 bits: my slot
      const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      if (!((double)mag <= 1e-3)) t = 1e10f;             // sampling_gpu.cu:100-101, sampling.cpp:74-76
      lox = fminf(lox, x); hix = fmaxf(hix, x);          // fminf/fmaxf drop NaNs
      loy = fminf(loy, y); hiy = fmaxf(hiy, y);
      loz = fminf(loz, z); hiz = fmaxf(hiz, z);
    }
    td[p] = t;
    if constexpr (!kLean) { px[p] = x; py[p] = y; pz[p] = z; nkey[p] = nk; }
    pts[p * kT + tid] = make_float4(x, y, z, __uint_as_float(nk));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lox = fminf(lox, __shfl_xor_sync(kFull, lox, o)); hix = fmaxf(hix, __shfl_xor_sync(kFull, hix, o));
    loy = fminf(loy, __shfl_xor_sync(kFull, loy, o)); hiy = fmaxf(hiy, __shfl_xor_sync(kFull, hiy, o));
    loz = fminf(loz, __shfl_xor_sync(kFull, loz, o)); hiz = fmaxf(hiz, __shfl_xor_sync(kFull, hiz, o));
  }

  float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];            // idxs[0] = 0, sampling_gpu.cu:85-86
  if (lane == 0) {
    // "nothing selectable here": loses to every real candidate; if the whole scene is like that
    // the answer is index 0 with point 0's coordinates (sampling_gpu.cu: besti stays 0)
    Cand c;
    c.vb = 0u; c.nk = nkey_of(0u) | (uint32_t)wid; c.x = x1; c.y = y1; c.z = z1;
    c.pad[0] = c.pad[1] = c.pad[2] = 0u;
    part[wid] = c;
  }
  const uint32_t bar0 = smem_u32(&bars[0]);
  if (cs > 1) {
    if (tid == 0) {
      mbar_init(bar0, 1);
      mbar_init(bar0 + 8, 1);
      fence_mbar_init_cluster();
    }
    cluster_sync_all();
  } else {
    __syncthreads();
  }
  if (rank == 0 && tid == 0 && m > 0) {
    idxs[0] = 0;
    if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
  }

  float wmax = INFINITY;           // the warp's current max min-distance (warp-uniform); inf = not evaluated yet
  for (int j = 1; j < m; ++j) {
    // ---- 0. can the new sample change anything in this warp's box? ------------------------
    const float ex = fmaxf(fmaxf(lox - x1, x1 - hix), 0.f);
    const float ey = fmaxf(fmaxf(loy - y1, y1 - hiy), 0.f);
    const float ez = fmaxf(fmaxf(loz - z1, z1 - hiz), 0.f);
    const float lb = ex * ex + ey * ey + ez * ez;
    if (!(lb * 0.99999f >= wmax)) {
#ifdef BQA_FPS_STATS
      if (lane == 0) atomicAdd(&g_active_warp_iters, 1ull);
#endif
      // ---- 1. update, warp max of the values ----------------------------------------------
      // (a tree-shaped max and a key pass predicated on `best == wf` were both measured slower:
      //  1.32 vs 1.26 ms -- the serial FMNMX chain hides under the distance arithmetic)
      float best = -INFINITY;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        float qx, qy, qz;
        if constexpr (kLean) {
          // at most 6 shared-memory loads in flight: more would spill, and with 221 KB of shared
          // memory the L1 that local memory lives in is a few KB (a spill costs an L2 round trip)
          if (p % 6 == 0 && p) asm volatile("" ::: "memory");
          const float4 v = pts[p * kT + tid]; qx = v.x; qy = v.y; qz = v.z;
        }
        else { qx = px[p]; qy = py[p]; qz = pz[p]; }
        const float dx = qx - x1, dy = qy - y1, dz = qz - z1;
        const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        const float d2 = fminf(d, td[p]);
        td[p] = d2;
        best = fmaxf(best, d2);
      }
      const uint32_t vb = best >= 0.f ? __float_as_uint(best) + 1u : 0u;
      const uint32_t wvb = __reduce_max_sync(kFull, vb);
      if (wvb) {
        // ---- 2. tie-break among the points that hold the max, then the winner's coordinates
        const float wf = __uint_as_float(wvb - 1u);
        uint32_t cand = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if constexpr (kLean) { if (td[p] == wf) cand = max(cand, __float_as_uint(pts[p * kT + tid].w)); }
          else cand = max(cand, td[p] == wf ? nkey[p] : 0u);
        }
        const uint32_t wnk = __reduce_max_sync(kFull, cand);
        if (cand == wnk) {                                 // exactly one lane (keys are distinct, non-zero)
          const float4 w = pts[(wnk & 31u) * kT + tid];    // own slot: no barrier needed
          Cand c;
          c.vb = wvb; c.nk = (wnk & ~31u) | (uint32_t)wid; c.x = w.x; c.y = w.y; c.z = w.z;
          c.pad[0] = c.pad[1] = c.pad[2] = 0u;
          part[wid] = c;
        }
        wmax = wf;
      } else {
        wmax = -INFINITY;                                  // nothing selectable: the initial post stands
      }
    }
    __syncthreads();
    uint32_t win_nk;
    if (cs == 1) {
      uint2 c = lane < kNW ? *reinterpret_cast<const uint2 *>(&part[lane]) : make_uint2(0u, 0u);
      const uint32_t mv = __reduce_max_sync(kFull, c.x);
      win_nk = __reduce_max_sync(kFull, c.x == mv ? c.y : 0u);
      const Cand *w = &part[win_nk & 31u];
      x1 = w->x; y1 = w->y; z1 = w->z;
      __syncthreads();             // part[] is rewritten in the next iteration
    } else {
      const int jj = j - 1;
      const uint32_t bar = bar0 + 8u * (jj & 1);
      if (wid == 0) {
        uint2 c = lane < kNW ? *reinterpret_cast<const uint2 *>(&part[lane]) : make_uint2(0u, 0u);
        const uint32_t mv = __reduce_max_sync(kFull, c.x);
        const uint32_t mk = __reduce_max_sync(kFull, c.x == mv ? c.y : 0u);
        const Cand *w = &part[mk & 31u];
        const float wx = w->x, wy = w->y, wz = w->z;
        const uint32_t slot = smem_u32(&recv[jj & 1][rank]);
        if (lane == 0) mbar_arrive_expect_tx(bar, 20u * cs);
        if (lane < cs) {
          st_async_v4(mapa_shared(slot, lane), mv, (mk & ~31u) | rank, __float_as_uint(wx), __float_as_uint(wy),
                      mapa_shared(bar, lane));
        } else if (lane < 2 * cs) {
          st_async_b32(mapa_shared(slot + 16u, lane - cs), __float_as_uint(wz), mapa_shared(bar, lane - cs));
        }
      }
      mbar_wait(bar, (jj >> 1) & 1);
      const Cand *rb = recv[jj & 1];
      uint2 gc = make_uint2(0u, 0u);
      if (lane < cs) gc = *reinterpret_cast<const uint2 *>(&rb[lane]);
      const uint32_t mv = __reduce_max_sync(kFull, gc.x);
      win_nk = __reduce_max_sync(kFull, gc.x == mv ? gc.y : 0u);
      const Cand *w = &rb[win_nk & 15u];
      const float2 xy = *reinterpret_cast<const float2 *>(&w->x);
      x1 = xy.x; y1 = xy.y; z1 = w->z;
    }
    if (rank == 0 && tid == 0) {
      idxs[j] = (int)index_of(win_nk);                     // sampling_gpu.cu:170-171
      if (new_xyz) { new_xyz[j * 3 + 0] = x1; new_xyz[j * 3 + 1] = y1; new_xyz[j * 3 + 2] = z1; }
    }
  }
  if (cs > 1) cluster_sync_all();  // nobody exits while a peer may still write into it
}

template <int P, int T = kT, bool kLean = false>
int launch_sorted(int b, int n, int m, int cs, const float4 *sorted, const float *xyz, int *idxs,
                  float *new_xyz, cudaStream_t stream) {
  constexpr int kT = T, kNW = T / 32;
  auto *kernel = fps_sorted_kernel<P, T, kLean>;
  if (cs > 8)
    BQA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t{};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3((unsigned)(b * cs));
  cfg.blockDim = dim3(kT);
  cfg.dynamicSmemBytes = sizeof(float4) * P * kT;
  BQA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)cfg.dynamicSmemBytes));
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#ifdef BQA_FPS_STATS
  unsigned long long zero = 0;
  cudaMemcpyToSymbol(g_active_warp_iters, &zero, sizeof(zero));
#endif
  BQA_CUDA(cudaLaunchKernelEx(&cfg, kernel, n, m, cs, sorted, xyz, idxs, new_xyz));
#ifdef BQA_FPS_STATS
  unsigned long long act = 0;
  cudaMemcpyFromSymbol(&act, g_active_warp_iters, sizeof(act));
  fprintf(stderr, "[bqa fps sorted] b=%d n=%d m=%d cs=%d P=%d: %.2f%% of warp-iterations active (%.1f warps per scene-iteration of %d)\n",
          b, n, m, cs, P, 100.0 * act / ((double)b * cs * kNW * (m - 1)), (double)act / ((double)b * (m - 1)), cs * kNW);
#endif
  count_launch();
  (void)kNW;
  return check_launch("fps_sorted_kernel");
}

}  // namespace

// scenes the sorted kernel takes: the tie key assumes the reference's block size 512 (n >= 512)
// and the scene must fit the registers of a cluster of <= 16 CTAs
bool fps_sorted_supported(int n, int m) {
  return n >= 512 && m >= 1 && (long long)n <= 16ll * kT * kMaxP;
}

// throughput variant: scenes of up to 8 x 13824 points (a portable cluster of lean CTAs)
constexpr int kLeanT = 768, kLeanMaxP = 18;
bool fps_lean_supported(int n, int m) {
  return n >= 512 && m >= 1 && (long long)n <= 8ll * kLeanT * kLeanMaxP;
}

int fps_sorted_dispatch(int b, int n, int m, const float *xyz, const void *grid, int *idxs,
                        float *new_xyz, int lean, cudaStream_t stream) {
  if (!fps_sorted_supported(n, m))
    return set_error(BQA_ERR_UNSUPPORTED, "fps (sorted): n=%d not supported", n);
  static const int env_lean = [] { const char *e = getenv("BQA_FPS_LEAN"); return e ? atoi(e) : -1; }();
  if (env_lean >= 0) lean = env_lean;                    // developer override
  // throughput variant of choice: the one-SM kernel (fps_stream.cu).  BQA_FPS_STREAM=0 keeps the
  // cluster kernels below, =2 also routes the latency variant through it (measurement only).
  static const int env_stream = [] { const char *e = getenv("BQA_FPS_STREAM"); return e ? atoi(e) : 1; }();
  if (((lean && env_stream) || env_stream == 2) && fps_stream_supported(n, m))
    return fps_stream_dispatch(b, n, m, xyz, grid, idxs, new_xyz, stream);
  if (lean && fps_lean_supported(n, m)) {
    // fewest CTAs whose shared memory holds the scene; the tail of the last run is padding
    const int cs = ceil_div(n, kLeanT * kLeanMaxP);
    const int per = ceil_div(n, cs * kLeanT);
    const float4 *sorted = ball_query_grid_sorted(grid, b, n);
#define BQA_LEAN_CASE(PP) if (per <= PP) return launch_sorted<PP, kLeanT, true>(b, n, m, cs, sorted, xyz, idxs, new_xyz, stream);
    BQA_LEAN_CASE(3) BQA_LEAN_CASE(6) BQA_LEAN_CASE(10) BQA_LEAN_CASE(14) BQA_LEAN_CASE(18)
#undef BQA_LEAN_CASE
  }
  // cluster size: same trade-off as fps.cu's plan (co-resident clusters on B200: 22 of 6 CTAs,
  // 33 of 4, 15 of 8), but pruned iterations cost little per point, so fewer, fuller CTAs
  static const int forced = [] { const char *e = getenv("BQA_FPS_SORTED_CS"); return e ? atoi(e) : 0; }();
  static const int kSizes[] = {1, 2, 4, 6, 8, 16};
  static const int kCoResident[] = {1 << 20, 74, 33, 22, 15, 7};
  int cs = 0, per = 0;
  double best_t = 0;
  for (int i = 0; i < 6; ++i) {
    const int c = kSizes[i];
    const int p = ceil_div(n, c * kT);
    if (p > kMaxP) continue;
    const double t = ceil_div(b, kCoResident[i]) * ((c == 1 ? 600.0 : 1000.0) + 10.0 * p);
    if (!cs || t < best_t) { cs = c; per = p; best_t = t; }
  }
  if (forced > 0 && forced <= 16 && ceil_div(n, forced * kT) <= kMaxP) { cs = forced; per = ceil_div(n, cs * kT); }
  const float4 *sorted = ball_query_grid_sorted(grid, b, n);
  int rc = BQA_ERR_UNSUPPORTED;
#define BQA_SORTED_CASE(PP) if (per <= PP) { rc = launch_sorted<PP>(b, n, m, cs, sorted, xyz, idxs, new_xyz, stream); } else
  BQA_SORTED_CASE(1) BQA_SORTED_CASE(2) BQA_SORTED_CASE(4) BQA_SORTED_CASE(6) BQA_SORTED_CASE(8)
  BQA_SORTED_CASE(10) BQA_SORTED_CASE(12) BQA_SORTED_CASE(14) BQA_SORTED_CASE(16) BQA_SORTED_CASE(18) BQA_SORTED_CASE(20)
  { rc = set_error(BQA_ERR_UNSUPPORTED, "fps (sorted): %d points per thread", per); }
#undef BQA_SORTED_CASE
  return rc;
}

}  // namespace bqa
