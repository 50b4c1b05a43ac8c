// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces /root/reference/lib/pointnet2/_ext_src/src/sampling_gpu.cu:69-229
// (furthest_point_sampling_kernel<512>, grid = B, every point re-read from global
// memory on each of the npoint-1 serial iterations, 9 block barriers per iteration).
//
// Design (one scene = one thread-block cluster):
//   * the scene's points AND their running min-distances live in REGISTERS, P points
//     per thread, spread over CS CTAs x 512 threads (40k points: CS=8, P=10);
//   * per iteration every thread updates its P min-distances and keeps its best;
//     a warp finds its winner with two redux.sync, the CTA with one __syncthreads
//     and two more redux.sync in warp 0;
//   * CTA winners (value, tie-key, x, y, z) are pushed into every CTA of the cluster
//     with st.async (DSMEM store that completes a transaction on the receiver's
//     mbarrier), so there is no cluster-wide barrier in the loop; each warp reduces
//     the CS candidates it received and carries on;
//   * the winner's coordinates ride along, so the gather of new_xyz is free.
//
// Bit-exact contract with the reference:
//   d = fma(dz,dz,fma(dx,dx,dy*dy)), temp = min(d,temp), points with
//   (double)|p|^2 <= 1e-3 never update and are never selected, and among equal maxima
//   the winner minimises (bitrev(k mod bs), k) where bs = opt_n_threads(n) is the block
//   size the reference would have used -- that is what its pairwise tree with
//   "ties keep the lower slot" (sampling_gpu.cu:59-65,115-168) computes.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace bqa {

namespace {

constexpr int kMaxCluster = 16;
constexpr unsigned kFull = 0xffffffffu;

// What one CTA tells the cluster about its winner (32-byte slot, two 16-byte halves).
struct __align__(16) Candidate {
  uint32_t valbits;               // 0 = "no selectable point", else float bits + 1
  uint32_t nkey;                  // ~tie_key: larger is better
  float x, y;
  float z;
  uint32_t pad[3];
};
static_assert(sizeof(Candidate) == 32, "candidate slot is 32 bytes");

// tie key: smaller wins.  bits = log2(bs).
__device__ __forceinline__ uint32_t tie_key(uint32_t k, int bits) {
  const uint32_t low = k & ((1u << bits) - 1u);
  const uint32_t rev = bits ? (__brev(low) >> (32 - bits)) : 0u;
  return (rev << 22) | (k >> bits);  // k >> bits < 2^22 for every n this kernel takes
}
__device__ __forceinline__ uint32_t key_to_index(uint32_t key, int bits) {
  const uint32_t rev = key >> 22;
  const uint32_t low = bits ? (__brev(rev) >> (32 - bits)) : 0u;
  return ((key & 0x3fffffu) << bits) | low;
}

// (valbits, nkey) lexicographic max over the warp; every lane gets the result.
__device__ __forceinline__ void warp_argmax(uint32_t &vb, uint32_t &nk) {
  const uint32_t m = __reduce_max_sync(kFull, vb);
  const uint32_t l = __reduce_max_sync(kFull, vb == m ? nk : 0u);
  vb = m;
  nk = l;
}

#ifdef BQA_FPS_TRACE
}  // namespace
unsigned long long *g_fps_trace = nullptr;
namespace {
#define FPS_TRACE_ARG , unsigned long long *trace
#define FPS_TRACE_PASS , g_fps_trace
#define FPS_TRACE_BEGIN const bool tr = trace && blockIdx.x == 0 && threadIdx.x == 0; long long tc = clock64(), tl0 = tc;
#define FPS_TRACE(i) if (tr) { long long now = clock64(); trace[i] += (unsigned long long)(now - tc); tc = now; }
#define FPS_TRACE_TOTAL(i) if (tr) { long long now = clock64(); trace[i] += (unsigned long long)(now - tl0); tl0 = now; tc = now; }
#else
#define FPS_TRACE_ARG
#define FPS_TRACE_PASS
#define FPS_TRACE_BEGIN
#define FPS_TRACE(i)
#define FPS_TRACE_TOTAL(i)
#endif

// P points per thread, T threads per CTA, cs CTAs per scene (runtime; cs_magic =
// ceil(2^32 / cs) so that q / cs == __umulhi(q, cs_magic) for the small q used here).
//
// One iteration, in dependent-latency order (cycle counts measured on B200 with
// tools/fps_trace.cu and tools/ubench.cu):
//   1. every thread updates its P min-distances (6 FMA-pipe ops per point: the floor of
//      this kernel is P * 6 * warps-per-scheduler cycles) and keeps (value, slot) of its max;
//   2. warp argmax = two redux.sync (~28 cycles each); lane 0 posts (value, key);
//   3. one __syncthreads; the 16 posts are folded with two more redux.sync -- by warp 0 only
//      when the scene spans a cluster, by every warp (no second barrier) when it does not;
//      the winner's coordinates come from the CTA's shared-memory copy of its points;
//   4. cluster: warp 0 pushes (value, key, x, y, z) into every CTA of the cluster with
//      st.async (complete_tx on the receiver's mbarrier, ~340 cycles one way, no cluster
//      barrier); every warp folds the cs candidates that landed in its own shared memory.
// Variants that were measured and lost: packed fp32x2 math for step 1 (FFMA2 issues at half
// rate: 43 vs 22 cycles per point), one warp folding all 512 thread candidates from shared
// memory (574 cycles), handing results over through shared memory instead of redux/shfl, and
// two scenes interleaved per 8-CTA cluster with a 17th communication warp (bit-exact, 28 %
// fewer SM-cycles per scene, but 2024 cycles per pair-iteration vs 1517 for one scene on a
// 6-CTA cluster, so the batch of 16 finishes later: 2.06 ms vs 1.58 ms), and every warp pushing
// its own candidate to every CTA (no CTA barrier, no warp-0 fold, each warp folds cs * 16
// candidates; 192 small DSMEM stores per CTA and iteration instead of 12: 2.28 ms vs 1.58 ms).
template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_cluster_kernel(int n, int m, int j_begin, int j_end, int cs, uint32_t cs_magic, int bits,
                   const float *__restrict__ xyz_all, int *__restrict__ idx_all,
                   float *__restrict__ new_xyz_all, float *__restrict__ state_all,
                   const int *__restrict__ run_flags FPS_TRACE_ARG) {
  // run_flags (optional): scenes whose flag is 0 already hold their answer (the identity prefix,
  // see fps_prefix_check_kernel) -- the whole cluster leaves before touching any barrier.
  if (run_flags && run_flags[blockIdx.x / cs] == 0) return;
  // Samples j_begin .. j_end-1 are produced by this launch (1 <= j_begin <= j_end <= m).  A
  // launch that does not start at 1 resumes from the running min-distances a previous launch
  // left in state_all (b,n), one that does not end at m leaves them there: the sampling can be
  // issued in slices so that consumers of the first centres run under the later slices.
  static_assert(T == 512, "slot <-> index arithmetic below assumes 512 threads (= max bs)");
  constexpr int NW = T / 32;
  constexpr int kLogT = 9;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ Candidate recv[2][kMaxCluster] | uint2 part[2][NW] | u64 bar[2] | sx,sy,sz ]
  Candidate *recv = reinterpret_cast<Candidate *>(smem_raw);
  uint2 *part = reinterpret_cast<uint2 *>(recv + 2 * kMaxCluster);
  uint64_t *bars = reinterpret_cast<uint64_t *>(part + 2 * NW);
  float *sx = reinterpret_cast<float *>(bars + 2);
  float *sy = sx + P * T;
  float *sz = sy + P * T;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wid = tid >> 5;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / cs;
  const int t_total = cs * T;
  const int g = rank * T + tid;  // this thread's slot in the scene-wide thread grid

  const float *xyz = xyz_all + (size_t)scene * n * 3;
  int *idxs = idx_all + (size_t)scene * m;
  float *new_xyz = new_xyz_all ? new_xyz_all + (size_t)scene * m * 3 : nullptr;

  // point k = p * t_total + g  ->  all of a thread's points share k mod bs (t_total is a
  // multiple of bs) and are visited in ascending k, like one reference thread's.
  float px[P], py[P], pz[P], td[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int k = p * t_total + g;
    float x = 0.f, y = 0.f, z = 0.f, t = -INFINITY;  // -inf: never updated, never selected
    if (k < n) {
      x = xyz[(size_t)k * 3 + 0];
      y = xyz[(size_t)k * 3 + 1];
      z = xyz[(size_t)k * 3 + 2];
      const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
      if (!((double)mag <= 1e-3)) t = 1e10f;  // sampling_gpu.cu:100-101, sampling.cpp:74-76
      if (j_begin > 1) t = state_all[(size_t)scene * n + k];   // resume (-inf marks skipped points)
    }
    px[p] = x; py[p] = y; pz[p] = z; td[p] = t;
    sx[p * T + tid] = x;
    sy[p * T + tid] = y;
    sz[p * T + tid] = z;
  }
  // tie key of slot p is key0 + p * kstep (t_total is a multiple of bs = 1 << bits)
  const uint32_t key0 = tie_key((uint32_t)g, bits);
  const uint32_t kstep = (uint32_t)t_total >> bits;

  const uint32_t bar0 = smem_u32(&bars[0]);
  if (cs > 1) {
    if (tid == 0) {
      mbar_init(bar0, 1);
      mbar_init(bar0 + 8, 1);
      fence_mbar_init_cluster();
    }
    cluster_sync_all();  // peers' barriers are initialised and their smem is live
  } else {
    __syncthreads();
  }

  const int last = j_begin > 1 ? idxs[j_begin - 1] : 0;       // written by the previous slice
  float x1 = xyz[(size_t)last * 3], y1 = xyz[(size_t)last * 3 + 1], z1 = xyz[(size_t)last * 3 + 2];
  if (rank == 0 && tid == 0 && m > 0 && j_begin == 1) {
    idxs[0] = 0;  // sampling_gpu.cu:85-86
    if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
  }

  FPS_TRACE_BEGIN
  for (int j = j_begin; j < j_end; ++j) {
    const int buf = j & 1;
    // ---- 1. update the P running min-distances, keep the thread's max --------------------
    float best = -1.f;
    int bslot = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const float dx = px[p] - x1, dy = py[p] - y1, dz = pz[p] - z1;
      const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
      const float d2 = fminf(d, td[p]);
      td[p] = d2;
      if (d2 > best) { best = d2; bslot = p; }   // strict: lowest k wins inside a thread
    }
    const bool valid = best >= 0.f;
    uint32_t vb = valid ? __float_as_uint(best) + 1u : 0u;
    uint32_t nk = valid ? ~(key0 + (uint32_t)bslot * kstep) : 0xffffffffu;   // invalid decodes to k = 0
    // ---- 2. warp argmax -> post --------------------------------------------------------
    warp_argmax(vb, nk);
    if (lane == 0) part[buf * NW + wid] = make_uint2(vb, nk);
    FPS_TRACE(0)
    __syncthreads();
    FPS_TRACE(1)
    uint32_t old;
    if (cs == 1) {
      // ---- 3. every warp folds the NW posts; slot index == point index when cs == 1 ------
      uint2 c = lane < NW ? part[buf * NW + lane] : make_uint2(0u, 0u);
      warp_argmax(c.x, c.y);
      old = key_to_index(~c.y, bits);
      x1 = sx[old]; y1 = sy[old]; z1 = sz[old];
      FPS_TRACE(2)
    } else {
      const int jj = j - j_begin;
      const uint32_t bar = bar0 + 8u * (jj & 1);
      if (wid == 0) {
        // ---- 3. warp 0 folds the posts and 4a. pushes the CTA's candidate to every peer ---
        uint2 c = lane < NW ? part[buf * NW + lane] : make_uint2(0u, 0u);
        warp_argmax(c.x, c.y);
        const uint32_t key = ~c.y;                       // (bitrev(tid) << 22) | (p * cs + rank)
        const int loc = (int)__umulhi(key & 0x3fffffu, cs_magic) * T + (int)(__brev(key >> 22) >> 23);
        const float wx = sx[loc], wy = sy[loc], wz = sz[loc];   // own point, or slot 0 if none
        const uint32_t slot = smem_u32(&recv[(jj & 1) * kMaxCluster + rank]);
        if (lane == 0) mbar_arrive_expect_tx(bar, 20u * cs);
        if (lane < cs) {
          st_async_v4(mapa_shared(slot, lane), c.x, c.y, __float_as_uint(wx), __float_as_uint(wy),
                      mapa_shared(bar, lane));
        } else if (lane < 2 * cs) {
          st_async_b32(mapa_shared(slot + 16u, lane - cs), __float_as_uint(wz),
                       mapa_shared(bar, lane - cs));
        }
      }
      FPS_TRACE(2)
      // ---- 4b. fold the cs candidates that landed in this CTA (every warp, no barrier) ----
      mbar_wait(bar, (jj >> 1) & 1);
      FPS_TRACE(3)
      const Candidate *rb = &recv[(jj & 1) * kMaxCluster];
      uint2 gc = make_uint2(0u, 0u);
      if (lane < cs) gc = *reinterpret_cast<const uint2 *>(&rb[lane]);
      warp_argmax(gc.x, gc.y);
      // owner CTA of the winner: key & 0x3fffff = k >> 9 = p * cs + rank (bs is 512 whenever a
      // scene spans a cluster); all-invalid decodes to k = 0, owned by rank 0
      const uint32_t q = ~gc.y & 0x3fffffu;
      const uint32_t owner = q - __umulhi(q, cs_magic) * (uint32_t)cs;
      const float2 xy = *reinterpret_cast<const float2 *>(&rb[owner].x);
      x1 = xy.x; y1 = xy.y; z1 = rb[owner].z;
      old = gc.y;                                        // decoded off the critical path
      FPS_TRACE(4)
    }
    FPS_TRACE_TOTAL(5)
    if (rank == 0 && tid == 0) {
      idxs[j] = (int)(cs == 1 ? old : key_to_index(~old, bits));   // sampling_gpu.cu:170-171
      if (new_xyz) { new_xyz[j * 3 + 0] = x1; new_xyz[j * 3 + 1] = y1; new_xyz[j * 3 + 2] = z1; }
    }
  }
  if (j_end < m) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const int k = p * t_total + g;
      if (k < n) state_all[(size_t)scene * n + k] = td[p];
    }
  }
  if (cs > 1) cluster_sync_all();  // nobody exits while a peer may still write into it
}

// Any-n fallback (n beyond what a 16-CTA cluster's registers hold, > 131072 points):
// streams xyz and the min-distances from global/L2 like the reference, with the same
// redux-based reduction as above.  temp lives in a caller-provided scratch (b,n).
__global__ void __launch_bounds__(1024, 1)
fps_global_kernel(int n, int m, int bits, const float *__restrict__ xyz_all,
                  float *__restrict__ temp_all, int *__restrict__ idx_all,
                  float *__restrict__ new_xyz_all, const int *__restrict__ run_flags) {
  __shared__ uint2 part[32];
  __shared__ uint32_t winner;
  if (run_flags && run_flags[blockIdx.x] == 0) return;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int scene = blockIdx.x;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  float *temp = temp_all + (size_t)scene * n;
  int *idxs = idx_all + (size_t)scene * m;
  float *new_xyz = new_xyz_all ? new_xyz_all + (size_t)scene * m * 3 : nullptr;
  for (int k = tid; k < n; k += 1024) {
    const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
    const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
    temp[k] = ((double)mag <= 1e-3) ? -INFINITY : 1e10f;
  }
  uint32_t old = 0;
  if (tid == 0 && m > 0) {
    idxs[0] = 0;
    if (new_xyz) { new_xyz[0] = xyz[0]; new_xyz[1] = xyz[1]; new_xyz[2] = xyz[2]; }
  }
  __syncthreads();
  for (int j = 1; j < m; ++j) {
    const float x1 = xyz[(size_t)old * 3], y1 = xyz[(size_t)old * 3 + 1], z1 = xyz[(size_t)old * 3 + 2];
    float best = -1.f;
    uint32_t bk = 0;
    for (int k = tid; k < n; k += 1024) {  // 1024 is a multiple of bs: same k mod bs per thread
      const float d = sqdist3(xyz[(size_t)k * 3], xyz[(size_t)k * 3 + 1], xyz[(size_t)k * 3 + 2], x1, y1, z1);
      const float d2 = fminf(d, temp[k]);
      temp[k] = d2;
      if (d2 > best) { best = d2; bk = k; }
    }
    uint32_t vb = 0u, nk = 0xffffffffu;
    if (best >= 0.f) { vb = __float_as_uint(best) + 1u; nk = ~tie_key(bk, bits); }
    warp_argmax(vb, nk);
    if (lane == 0) part[wid] = make_uint2(vb, nk);
    __syncthreads();
    if (wid == 0) {
      uint2 c = part[lane];
      warp_argmax(c.x, c.y);
      if (lane == 0) winner = key_to_index(~c.y, bits);
    }
    __syncthreads();
    old = winner;
    if (tid == 0) {
      idxs[j] = (int)old;
      if (new_xyz) {
        new_xyz[j * 3 + 0] = xyz[(size_t)old * 3];
        new_xyz[j * 3 + 1] = xyz[(size_t)old * 3 + 1];
        new_xyz[j * 3 + 2] = xyz[(size_t)old * 3 + 2];
      }
    }
  }
}

constexpr int kT = 512;     // threads per CTA of the register-resident kernel
constexpr int kMaxP = 20;   // points per thread: 4 registers each, 128 registers per thread at 512 threads

template <int P>
size_t fps_smem_bytes() {
  return sizeof(Candidate) * 2 * kMaxCluster + sizeof(uint2) * 2 * (kT / 32) + 16 +
         sizeof(float) * 3 * P * kT;
}

// `exclusive`: ask for so much shared memory that no other CTA fits on the SM, so kernels that
// run concurrently with a sampling slice do not steal issue slots from its latency chain
constexpr size_t kExclusiveSmem = 208 * 1024;

template <int P>
cudaError_t fps_config(cudaLaunchConfig_t *cfg, cudaLaunchAttribute *attr, int b, int cs,
                       cudaStream_t stream, bool exclusive = false) {
  size_t smem = fps_smem_bytes<P>();
  if (exclusive && smem < kExclusiveSmem) smem = kExclusiveSmem;
  cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<P, kT>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (cs > 8) {
    e = cudaFuncSetAttribute(fps_cluster_kernel<P, kT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3((unsigned)(b * cs));
  cfg->blockDim = dim3(kT);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  return cudaSuccess;
}

template <int P>
int launch_fps(int b, int n, int m, int j_begin, int j_end, int cs, int bits, const float *xyz,
               int *idxs, float *new_xyz, float *state, bool exclusive, const int *run_flags,
               cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  BQA_CUDA(fps_config<P>(&cfg, attr, b, cs, stream, exclusive));
  const uint32_t cs_magic = (uint32_t)((0x100000000ull + (unsigned)cs - 1) / (unsigned)cs);
  BQA_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<P, kT>, n, m, j_begin, j_end, cs, cs_magic, bits,
                              xyz, idxs, new_xyz, state, run_flags FPS_TRACE_PASS));
  count_launch();
  return check_launch("fps_cluster_kernel");
}

template <int P>
int max_clusters(int cs) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  if (fps_config<P>(&cfg, attr, 1, cs, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, fps_cluster_kernel<P, kT>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return ncl;
}

#define BQA_FPS_DISPATCH(per_thread, EXPR)            \
  do {                                                \
    if ((per_thread) <= 1) { constexpr int PP_ = 1; EXPR; }        \
    else if ((per_thread) <= 2) { constexpr int PP_ = 2; EXPR; }   \
    else if ((per_thread) <= 4) { constexpr int PP_ = 4; EXPR; }   \
    else if ((per_thread) <= 6) { constexpr int PP_ = 6; EXPR; }   \
    else if ((per_thread) <= 8) { constexpr int PP_ = 8; EXPR; }   \
    else if ((per_thread) <= 10) { constexpr int PP_ = 10; EXPR; } \
    else if ((per_thread) <= 12) { constexpr int PP_ = 12; EXPR; } \
    else if ((per_thread) <= 14) { constexpr int PP_ = 14; EXPR; } \
    else if ((per_thread) <= 16) { constexpr int PP_ = 16; EXPR; } \
    else { constexpr int PP_ = 20; EXPR; }                         \
  } while (0)

}  // namespace

// How many CTAs share one scene.  A scene must fit in the registers of its cluster
// (cs * 512 threads * <= 20 points); among the sizes that fit, pick the one with the lowest
// modelled time  waves * (fixed + per-point work / cs)  where waves = ceil(b / clusters that
// are co-resident on this GPU) -- measured on B200: 8-CTA clusters co-reside 15 at a time,
// 6-CTA 22, 4-CTA 33, so 16 scenes of 40k points run best on 6-CTA clusters.
struct FpsPlan { int cs; int per_thread; };

static FpsPlan fps_plan(int b, int n) {
  static int cached_clusters[17] = {0};
  static const int kSizes[] = {1, 2, 4, 6, 8, 16};
  FpsPlan best = {0, kMaxP + 1};
  double best_t = 0;
  for (int cs : kSizes) {
    const int per_thread = ceil_div(n, cs * kT);
    if (per_thread > kMaxP) continue;
    int ncl;
    if (cs == 1) {
      ncl = 1 << 20;
    } else {
      if (!cached_clusters[cs]) {
        int v = 0;
        BQA_FPS_DISPATCH(kMaxP, v = max_clusters<PP_>(cs));   // worst-case shared memory
        cached_clusters[cs] = v > 0 ? v : -1;
      }
      ncl = cached_clusters[cs];
      if (ncl <= 0) continue;
    }
    // cycles per iteration (measured): ~1000 of argmax chain (~600 without the cluster hop)
    // + ~36 per point held by a thread
    const double t = ceil_div(b, ncl) * ((cs == 1 ? 600.0 : 1000.0) + 36.0 * per_thread);
    if (!best.cs || t < best_t) { best = {cs, per_thread}; best_t = t; }
  }
  return best;
}

long long fps_scratch_bytes(int b, int n) {
  // conservative: scenes beyond what a 16-CTA cluster holds need the global-memory variant
  return (long long)n > 16ll * kT * kMaxP ? (long long)sizeof(float) * b * n : 0;
}

int fps_dispatch(int b, int n, int m, int j_begin, int j_end, const float *xyz, int *idxs,
                 float *new_xyz, float *scratch, bool exclusive, const int *run_flags,
                 cudaStream_t stream) {
  const bool sliced = j_begin > 1 || j_end < m;
  const int bs = ref_opt_n_threads(n);
  int bits = 0;
  while ((1 << bits) < bs) ++bits;
  if ((long long)n >= (1ll << 22) * bs) return set_error(BQA_ERR_UNSUPPORTED, "fps: n=%d too large", n);

  FpsPlan plan = fps_plan(b, n);
  static const char *dbg = getenv("BQA_FPS_DEBUG");
  if (dbg) {
    // developer override of the cluster size (e.g. BQA_FPS_DEBUG=8) and a one-line plan report
    const int ocs = atoi(dbg);
    if (ocs > 0 && ocs <= 16 && ceil_div(n, ocs * kT) <= kMaxP) plan = {ocs, ceil_div(n, ocs * kT)};
    fprintf(stderr, "[bqa fps] b=%d n=%d m=%d -> cs=%d per_thread=%d\n", b, n, m, plan.cs, plan.per_thread);
  }
  const int cs = plan.cs, per_thread = plan.per_thread;
  if (per_thread > kMaxP) {
    if (sliced)
      return set_error(BQA_ERR_UNSUPPORTED, "fps: n=%d is too large for sliced sampling", n);
    if (!scratch)
      return set_error(BQA_ERR_INVALID_ARG,
                       "fps: n=%d needs %lld bytes of scratch (see bqa_fps_scratch_bytes)", n,
                       (long long)sizeof(float) * b * n);
    fps_global_kernel<<<b, 1024, 0, stream>>>(n, m, bits, xyz, scratch, idxs, new_xyz, run_flags);
    count_launch();
    return check_launch("fps_global_kernel");
  }
  int rc = BQA_ERR_UNSUPPORTED;
  if (sliced && !scratch)
    return set_error(BQA_ERR_INVALID_ARG, "fps: sliced sampling needs a (b,n) float state buffer");
  BQA_FPS_DISPATCH(per_thread, rc = launch_fps<PP_>(b, n, m, j_begin, j_end, cs, bits, xyz, idxs, new_xyz,
                                                    scratch, exclusive, run_flags, stream));
  return rc;
}

// ---- sampling a cloud that is itself in sampling order -------------------------------------
//
// SA2-4 sample the centres the previous level sampled (models/backbone_module.py:52-86), i.e. a
// cloud already in furthest-point order.  Re-sampling such a cloud returns the prefix
// 0,1,...,m-1 -- the reference's own comment says so (backbone_module.py:111) -- unless two
// candidates tie, and then the reference's tie-break decides.  The serial chain (m-1 dependent
// argmax steps, 0.7 ms for the three levels) can therefore be replaced by a PARALLEL proof:
//   V[j]    = min(1e10, min_{i<j} d(x_j, x_i))        what the chain would hold for point j at step j
//   R(k, j) = min(1e10, min_{i<j} d(x_k, x_i))        ... and for any later point k
// If V[j] > 0 and R(k, j) < V[j] STRICTLY for every j in [1, m) and every k in (j, n), and no
// point k >= 1 is in the reference's skip set (|x|^2 <= 1e-3), then at every step point j is the
// unique maximum (earlier points sit at 0), so every reduction order and tie-break returns j:
// the sampling is the identity prefix, and so is the sampling of any prefix of the cloud to any
// m' <= m.  Same fp32 distance (sqdist3) and the same 1e10 start value as the chain.  Scenes that
// fail any test (ties, duplicates, skipped points, NaNs) keep run_flag = 1 and go through the
// real chain, so the result is the reference's in every case.
constexpr int kVerT = 256;

__global__ void __launch_bounds__(kVerT)
fps_prefix_v_kernel(int n, int m, const float *__restrict__ xyz_all, float *__restrict__ v_all,
                    int *__restrict__ run_flags) {
  extern __shared__ float spts[];      // the first m points, xyz interleaved
  const int scene = blockIdx.y;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  for (int e = threadIdx.x; e < m * 3; e += kVerT) spts[e] = xyz[e];
  __syncthreads();
  const int j = blockIdx.x * kVerT + threadIdx.x;
  if (j >= m) return;
  const float x = spts[j * 3], y = spts[j * 3 + 1], z = spts[j * 3 + 2];
  float v = 1e10f;
  for (int i = 0; i < j; ++i) v = fminf(v, sqdist3(x, y, z, spts[i * 3], spts[i * 3 + 1], spts[i * 3 + 2]));
  v_all[(size_t)scene * m + j] = v;
  if (j >= 1 && !(v > 0.f)) run_flags[scene] = 1;
}

__global__ void __launch_bounds__(kVerT)
fps_prefix_check_kernel(int n, int m, const float *__restrict__ xyz_all,
                        const float *__restrict__ v_all, int *__restrict__ run_flags) {
  extern __shared__ float spts[];      // [3m] points, then [m] V
  float *sv = spts + 3 * m;
  const int scene = blockIdx.y;
  const float *xyz = xyz_all + (size_t)scene * n * 3;
  for (int e = threadIdx.x; e < m * 3; e += kVerT) spts[e] = xyz[e];
  for (int e = threadIdx.x; e < m; e += kVerT) sv[e] = v_all[(size_t)scene * m + e];
  __syncthreads();
  const int k = blockIdx.x * kVerT + threadIdx.x;
  if (k < 1 || k >= n) return;
  const float x = xyz[(size_t)k * 3], y = xyz[(size_t)k * 3 + 1], z = xyz[(size_t)k * 3 + 2];
  const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
  bool ok = !((double)mag <= 1e-3);                      // sampling_gpu.cu:100-101 skip set
  const int jmax = min(k, m);
  float r = 1e10f;
  for (int j = 1; j < jmax; ++j) {
    r = fminf(r, sqdist3(x, y, z, spts[(j - 1) * 3], spts[(j - 1) * 3 + 1], spts[(j - 1) * 3 + 2]));
    ok = ok && (r < sv[j]);
  }
  if (!ok) run_flags[scene] = 1;
}

__global__ void __launch_bounds__(256)
fps_identity_fill_kernel(int n, int m, const float *__restrict__ xyz_all, const int *__restrict__ run_flags,
                         int *__restrict__ idx_all, float *__restrict__ new_xyz_all) {
  const int scene = blockIdx.y;
  if (run_flags[scene] != 0) return;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e < m) idx_all[(size_t)scene * m + e] = e;
  if (new_xyz_all && e < m * 3) new_xyz_all[(size_t)scene * m * 3 + e] = xyz_all[(size_t)scene * n * 3 + e];
}

// prefix lengths whose coordinates + V fit the check kernel's shared memory
bool fps_prefix_check_supported(int n, int m) { return m >= 1 && m <= n && m <= 8192; }

int fps_prefix_check_dispatch(int b, int n, int m, const float *xyz, float *v_scratch, int *run_flags,
                              cudaStream_t stream) {
  if (b > 65535) return set_error(BQA_ERR_UNSUPPORTED, "fps prefix check: batch too large");
  BQA_CUDA(cudaMemsetAsync(run_flags, 0, sizeof(int) * (size_t)b, stream));
  static bool attr_done = false;
  if (!attr_done) {
    BQA_CUDA(cudaFuncSetAttribute(fps_prefix_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 12));
    BQA_CUDA(cudaFuncSetAttribute(fps_prefix_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 16));
    attr_done = true;
  }
  fps_prefix_v_kernel<<<dim3((unsigned)ceil_div(m, kVerT), (unsigned)b), kVerT, (size_t)m * 12, stream>>>(
      n, m, xyz, v_scratch, run_flags);
  count_launch();
  if (int rc = check_launch("fps_prefix_v_kernel")) return rc;
  fps_prefix_check_kernel<<<dim3((unsigned)ceil_div(n, kVerT), (unsigned)b), kVerT, (size_t)m * 16, stream>>>(
      n, m, xyz, v_scratch, run_flags);
  count_launch();
  return check_launch("fps_prefix_check_kernel");
}

int fps_identity_fill_dispatch(int b, int n, int m, const float *xyz, const int *run_flags, int *idxs,
                               float *new_xyz, cudaStream_t stream) {
  fps_identity_fill_kernel<<<dim3((unsigned)ceil_div(m * 3, 256), (unsigned)b), 256, 0, stream>>>(
      n, m, xyz, run_flags, idxs, new_xyz);
  count_launch();
  return check_launch("fps_identity_fill_kernel");
}

}  // namespace bqa
