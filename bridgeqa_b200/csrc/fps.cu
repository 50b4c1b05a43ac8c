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
#include "common.cuh"

namespace bqa {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxCluster = 16;
constexpr unsigned kFull = 0xffffffffu;

struct __align__(16) Candidate {  // what one CTA tells the cluster about its winner
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

template <int P>
__global__ void __launch_bounds__(kThreads, 1)
fps_cluster_kernel(int n, int m, int cs, int bits, const float *__restrict__ xyz_all,
                   int *__restrict__ idx_all, float *__restrict__ new_xyz_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ Candidate recv[2][kMaxCluster] | uint2 part[kWarps] | uint64 bar[2] | float4 bcast | sx,sy,sz ]
  Candidate *recv = reinterpret_cast<Candidate *>(smem_raw);
  uint2 *part = reinterpret_cast<uint2 *>(recv + 2 * kMaxCluster);
  uint64_t *bars = reinterpret_cast<uint64_t *>(part + kWarps);
  float4 *bcast = reinterpret_cast<float4 *>(bars + 2);
  float *sx = reinterpret_cast<float *>(bcast + 1);
  float *sy = sx + P * kThreads;
  float *sz = sy + P * kThreads;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int wid = tid >> 5;
  const uint32_t rank = cs > 1 ? cluster_ctarank() : 0u;
  const int scene = blockIdx.x / cs;
  const int t_total = cs * kThreads;
  const int g = rank * kThreads + tid;  // this thread's slot in the scene-wide thread grid

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
    }
    px[p] = x; py[p] = y; pz[p] = z; td[p] = t;
    sx[p * kThreads + tid] = x;
    sy[p * kThreads + tid] = y;
    sz[p * kThreads + tid] = z;
  }

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

  float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
  if (rank == 0 && tid == 0 && m > 0) {
    idxs[0] = 0;  // sampling_gpu.cu:85-86
    if (new_xyz) { new_xyz[0] = x1; new_xyz[1] = y1; new_xyz[2] = z1; }
  }

  for (int j = 1; j < m; ++j) {
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
    uint32_t vb = 0u, nk = 0xffffffffu;          // "nothing selectable" decodes to k = 0
    if (best >= 0.f) {
      vb = __float_as_uint(best) + 1u;
      nk = ~tie_key((uint32_t)(bslot * t_total + g), bits);
    }
    warp_argmax(vb, nk);
    if (lane == 0) part[wid] = make_uint2(vb, nk);
    __syncthreads();

    if (wid == 0) {
      uint2 c = lane < kWarps ? part[lane] : make_uint2(0u, 0u);
      warp_argmax(c.x, c.y);
      const uint32_t k = key_to_index(~c.y, bits);
      // this CTA's winner is one of its own points (or k = 0 when it has none)
      const int loc = (int)(k / (uint32_t)t_total) * kThreads + (int)(k % (uint32_t)kThreads);
      const float wx = sx[loc], wy = sy[loc], wz = sz[loc];
      if (cs == 1) {
        if (lane == 0) *bcast = make_float4(wx, wy, wz, __uint_as_float(k));
      } else {
        const int jj = j - 1;
        const uint32_t bar = bar0 + 8u * (jj & 1);
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
    }

    uint32_t old;
    if (cs == 1) {
      __syncthreads();
      const float4 w = *bcast;
      x1 = w.x; y1 = w.y; z1 = w.z; old = __float_as_uint(w.w);
    } else {
      const int jj = j - 1;
      mbar_wait(bar0 + 8u * (jj & 1), (jj >> 1) & 1);
      uint32_t cv = 0u, ck = 0u;
      float cx = 0.f, cy = 0.f, cz = 0.f;
      if (lane < cs) {
        const Candidate *c = &recv[(jj & 1) * kMaxCluster + lane];
        const uint4 q = *reinterpret_cast<const uint4 *>(c);
        cv = q.x; ck = q.y; cx = __uint_as_float(q.z); cy = __uint_as_float(q.w);
        cz = c->z;
      }
      uint32_t mv = cv, mk = ck;
      warp_argmax(mv, mk);
      // all-invalid: every CTA reports (0, ~0) and the lowest rank (owner of k=0) wins
      const int src = __ffs(__ballot_sync(kFull, lane < cs && cv == mv && ck == mk)) - 1;
      x1 = __shfl_sync(kFull, cx, src);
      y1 = __shfl_sync(kFull, cy, src);
      z1 = __shfl_sync(kFull, cz, src);
      old = key_to_index(~mk, bits);
    }
    if (rank == 0 && tid == 0) {
      idxs[j] = (int)old;  // sampling_gpu.cu:170-171
      if (new_xyz) { new_xyz[j * 3 + 0] = x1; new_xyz[j * 3 + 1] = y1; new_xyz[j * 3 + 2] = z1; }
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
                  float *__restrict__ new_xyz_all) {
  __shared__ uint2 part[32];
  __shared__ uint32_t winner;
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

template <int P>
size_t fps_smem_bytes() {
  return sizeof(Candidate) * 2 * kMaxCluster + sizeof(uint2) * kWarps + 16 + 16 +
         sizeof(float) * 3 * P * kThreads;
}

template <int P>
int launch_fps(int b, int n, int m, int cs, int bits, const float *xyz, int *idxs, float *new_xyz,
               cudaStream_t stream) {
  const size_t smem = fps_smem_bytes<P>();
  BQA_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
  if (cs > 8)
    BQA_CUDA(cudaFuncSetAttribute(fps_cluster_kernel<P>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * cs));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BQA_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<P>, n, m, cs, bits, xyz, idxs, new_xyz));
  count_launch();
  return check_launch("fps_cluster_kernel");
}

}  // namespace

// Largest per-thread point count instantiated.
static const int kMaxP = 16;

static void fps_plan(int n, int *cs_out, int *per_thread_out) {
  // smallest cluster whose threads hold the scene with <= 10 points each (more CTAs per
  // scene shorten the per-iteration update, but the hop across the cluster costs
  // latency, so small scenes stay in one CTA); 16-CTA clusters are non-portable and only
  // used when 8 CTAs x 16 points cannot hold the scene.
  int cs = 1;
  if (n > kThreads * kMaxP) {
    cs = 2;
    while (cs < 8 && (long long)cs * kThreads * 10 < n) cs *= 2;
    if ((long long)cs * kThreads * kMaxP < n) cs = 16;
  }
  *cs_out = cs;
  *per_thread_out = ceil_div(n, cs * kThreads);
}

long long fps_scratch_bytes(int b, int n) {
  int cs, per_thread;
  fps_plan(n, &cs, &per_thread);
  return per_thread > kMaxP ? (long long)sizeof(float) * b * n : 0;
}

int fps_dispatch(int b, int n, int m, const float *xyz, int *idxs, float *new_xyz, float *scratch,
                 cudaStream_t stream) {
  const int bs = ref_opt_n_threads(n);
  int bits = 0;
  while ((1 << bits) < bs) ++bits;
  if ((long long)n >= (1ll << 22) * bs) return set_error(BQA_ERR_UNSUPPORTED, "fps: n=%d too large", n);

  int cs, per_thread;
  fps_plan(n, &cs, &per_thread);
  if (per_thread > kMaxP) {
    if (!scratch)
      return set_error(BQA_ERR_INVALID_ARG,
                       "fps: n=%d needs bqa_fps_scratch_bytes() = %lld bytes of scratch", n,
                       fps_scratch_bytes(b, n));
    fps_global_kernel<<<b, 1024, 0, stream>>>(n, m, bits, xyz, scratch, idxs, new_xyz);
    count_launch();
    return check_launch("fps_global_kernel");
  }
#define BQA_FPS_CASE(PP) \
  if (per_thread <= PP) return launch_fps<PP>(b, n, m, cs, bits, xyz, idxs, new_xyz, stream);
  BQA_FPS_CASE(1)
  BQA_FPS_CASE(2)
  BQA_FPS_CASE(3)
  BQA_FPS_CASE(4)
  BQA_FPS_CASE(5)
  BQA_FPS_CASE(6)
  BQA_FPS_CASE(8)
  BQA_FPS_CASE(10)
  BQA_FPS_CASE(12)
  BQA_FPS_CASE(16)
#undef BQA_FPS_CASE
  return set_error(BQA_ERR_UNSUPPORTED, "fps: internal dispatch error");
}

}  // namespace bqa
