// fp_fused.cu -- fused feature-propagation layer for inference on sm_100a:
//   three_nn -> inverse-distance weights -> three_interpolate -> concat with the skip
//   features -> 2 x [1x1 conv + folded BN + ReLU] on tcgen05 tensor cores.
//
// Replaces, for eval-mode forward, PointnetFPModule.forward
// (/root/reference/lib/pointnet2/pointnet2_modules.py:376-421): three_nn_kernel,
// sqrt / reciprocal / sum / div elementwise kernels, three_interpolate_kernel, torch.cat,
// and 2 x [cuDNN conv, BN, ReLU] -- the interpolated (B, C2, n) tensor, the concatenated
// (B, C1+C2, n) tensor and the hidden activation never touch HBM.
//
// One CTA of 128 threads owns a tile of 128 unknown points (thread = point = TMEM lane):
//   * three_nn over the scene's known set staged in shared memory (same strict-`<` cascade in
//     ascending index as the stand-alone op, so the neighbours are bit-identical);
//   * layer 1, K = C_known + C_skip = 512 in 8 chunks of 64: the thread builds its row of the
//     A operand (interpolated channels fma(p3,w3,fma(p1,w1,p2*w2)), then its own skip
//     channels) as 16-bit K-major vectors while the bulk-copy engine streams the matching
//     32 KB slice of W1 into the other stage (cp.async.bulk -> mbarrier) and the tensor core
//     works on the previous chunk (tcgen05.commit -> mbarrier per stage);
//   * epilogue 1 keeps relu(D1 + b1) in shared memory as the A operand of layer 2, whose
//     4 weight slices flow through the same two stages;
//   * epilogue 2 writes relu(D2 + b2) both channel-major (the reference layout) and
//     point-major (for the next fused layer), each coalesced.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bqa {
namespace {

constexpr int kRows = 128;
constexpr int kChunk = 64;         // K per weight slice
constexpr int kWidth = 256;        // C1 == C2 == 256 (both FP layers of the backbone)
constexpr int kKnownTile = 512;    // known points staged per pass of three_nn (as float4: 8 KB)

// two fp32 -> one packed 16-bit pair in one instruction (fp16 saturates at +-65504; see sa_fused.cu)
__device__ __forceinline__ uint32_t pack16(float lo, float hi, int fp16) {
  uint32_t d;
  if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack16_relu(float lo, float hi, int fp16) {
  uint32_t d;
  if (fp16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// 8 consecutive floats of a feature row: one 256-bit load (LDG.E.256, sm_100) when the row is 32-byte aligned
__device__ __forceinline__ void ldg_row32(const float *p, float (&v)[8], int wide) {
  if (wide) {
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
        : "l"(p));
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p));
    const float4 b = __ldg(reinterpret_cast<const float4 *>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

struct FpParams {
  int b, n, m, c_known, c_skip;
  int known_stride, skip_stride;          // floats between consecutive points
  int skip16_stride;                      // 16-bit elements between consecutive points of skip16
  const float *unknown, *known;           // (b,n,3), (b,m,3)
  const float *known_feat, *skip_feat;    // point-major f32
  const uint16_t *skip16;                 // optional 16-bit point-major twin of the skip features (else NULL)
  const uint4 *w;                         // W1 image [(ck+cs)/8][256] then W2 image [256/8][256], 16-byte vectors
  const float *b1, *b2;
  float *out_cm, *out_pm;
  int fp16;
  int wide;                               // known_feat rows are 32-byte aligned: one 256-bit load per lane
  int num_tiles, tiles_per_scene;
#ifdef BQA_SA_TRACE
  int dbg;
#endif
};

#ifdef BQA_SA_TRACE
__device__ long long g_fp_trace[64];
#define FP_TRACE(i) if (blockIdx.x == 0 && threadIdx.x == 0) g_fp_trace[i] = clock64();
#else
#define FP_TRACE(i)
#endif

constexpr int kThreads = 512;      // 16 warps: four threads per row (warps w, w + 4, w + 8, w + 12 share TMEM lane quarter w % 4)
constexpr int kParts = kThreads / kRows;        // threads per row: each scans 1 / kParts of the known points and
                                                // converts 1 / kParts of the epilogue columns
constexpr int kRowsPerWarp = kRows / (kThreads / 32);   // rows a warp gathers (8)

struct NnRow { int i1, i2, i3; float w1, w2, w3; };

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// lexicographic (distance, index) "less": the order the reference's strict-< cascade over ascending
// indices produces (interpolate_gpu.cu:9-58)
__device__ __forceinline__ bool nn_less(float da, int ia, float db, int ib) {
  return da < db || (da == db && ia < ib);
}

__global__ void __launch_bounds__(kThreads)
fp_mlp_kernel(const FpParams P) {
  constexpr int kAVecs = kChunk / 8 * kRows;          // 1024 x 16 B = 16 KB per stage
  constexpr int kWVecs = kChunk / 8 * kWidth;         // 2048 x 16 B = 32 KB per stage
  constexpr int kXVecs = kWidth / 8 * kRows;          // 4096 x 16 B = 64 KB
  extern __shared__ __align__(1024) unsigned char smem_unaligned[];
  // 1024-byte alignment: the A stages hold 128-byte swizzle atoms
  // (pointer arithmetic on the array, not an integer round trip: the accesses stay LDS / STS)
  unsigned char *smem_raw = smem_unaligned + ((1024u - (smem_u32(smem_unaligned) & 1023u)) & 1023u);
  uint4 *a_st = reinterpret_cast<uint4 *>(smem_raw);                  // [2][kAVecs]
  uint4 *w_st = a_st + 2 * kAVecs;                                    // [2][kWVecs]
  uint4 *x1 = w_st + 2 * kWVecs;                                      // [kXVecs]
  float *s_known = reinterpret_cast<float *>(x1 + kXVecs);            // [kKnownTile] float4 {x, y, z, -}
  float *s_b1 = s_known + kKnownTile * 4;
  float *s_b2 = s_b1 + kWidth;
  NnRow *s_nn = reinterpret_cast<NnRow *>(s_b2 + kWidth);             // [kRows]: neighbours + weights of every row
  float *s_merge = reinterpret_cast<float *>(s_nn + kRows);           // [kParts - 1][kRows][6]: the other parts' candidates
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_merge + (kParts - 1) * kRows * 6);  // wfull[2], mdone[2]
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3;            // TMEM lane quarter of this warp
  const int half = warp >> 2;              // 0 .. kParts-1: which part of the known set / of the columns this thread takes
  const int row = quarter * 32 + lane;     // row of the tile this thread shares with kParts - 1 others
  for (int i = tid; i < kWidth; i += kThreads) { s_b1[i] = P.b1[i]; s_b2[i] = P.b2[i]; }
  const uint32_t wfull0 = smem_u32(&s_bar[0]), mdone0 = smem_u32(&s_bar[2]);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
    fence_mbar_init_cluster();
  }
  if (warp == 0) umma::tmem_alloc(smem_u32(s_tmem), 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *s_tmem;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + kWidth;
  const uint32_t a_addr = smem_u32(a_st), w_addr = smem_u32(w_st), x_addr = smem_u32(x1);
  const uint32_t idesc = umma::instr_desc_16b_f32(128, kWidth, !P.fp16);
  const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

  const int nk1 = (P.c_known + P.c_skip) / kChunk;   // weight slices of layer 1
  const int nk2 = kWidth / kChunk;                   // of layer 2
  const int nchunks = nk1 + nk2;
  uint32_t gchunk = 0;                               // running slice counter (stage / parity)

  for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
    const int scene = tile / P.tiles_per_scene;
    const int row0 = (tile % P.tiles_per_scene) * kRows;
    const bool live = row0 + row < P.n;
    const int j = live ? row0 + row : P.n - 1;       // clamp: dead rows compute, never store

    // first weight slice can fly while three_nn runs (both stages are idle between tiles)
    if (tid == 0) {
      const uint32_t s = gchunk & 1;
      mbar_arrive_expect_tx(wfull0 + 8 * s, kWVecs * 16);
      bulk_load(w_addr + s * kWVecs * 16, P.w, kWVecs * 16, wfull0 + 8 * s);
    }

    // ---- the skip half of the layer-1 operand does not depend on the neighbours: when the previous fused
    // layer left a 16-bit twin, ALL its chunks are copied now (cp.async, landing under three_nn and the
    // interpolation) into the X1 buffer, which is idle until epilogue 1 -- which only runs after the last
    // layer-1 MMA has read them
    const bool skip_early = P.skip16 != nullptr && P.c_skip * kRows * 2 <= kXVecs * 16;
    if (skip_early) {
      // a quarter warp copies 128 contiguous bytes (64 channels) of one row into the 128-byte-swizzled layout:
      // K block kb = 64 channels = 16 KB [128 rows][128 B], chunk ^ (row & 7)
      const int rqe = lane >> 3, q8e = lane & 7;
#pragma unroll 1
      for (int it = 0; it < kRowsPerWarp / 4; ++it) {
        const int r = warp * kRowsPerWarp + it * 4 + rqe;
        const int jr = min(row0 + r, P.n - 1);
        const uint16_t *src = P.skip16 + ((size_t)scene * P.n + jr) * P.skip16_stride + q8e * 8;
        const uint32_t d = x_addr + (uint32_t)(r * 128 + ((q8e ^ (r & 7)) << 4));
        for (int kb = 0; kb < P.c_skip / kChunk; ++kb) cp_async16(d + kb * (kRows * 128), src + kb * kChunk);
      }
    }
    FP_TRACE(0)
    // ---- three_nn (interpolate_gpu.cu:9-58 semantics) + weights (pointnet2_modules.py:399-402).
    // The two threads of a row scan one half of every staged block of known points each; the reference's
    // result is the three smallest (distance, index) pairs, so the halves merge exactly.
    float ux, uy, uz;
    {
      const float *u = P.unknown + ((size_t)scene * P.n + j) * 3;
      ux = u[0]; uy = u[1]; uz = u[2];
    }
    float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;
    int i1 = 0, i2 = 0, i3 = 0;
    const float *known = P.known + (size_t)scene * P.m * 3;
    float4 *s_known4 = reinterpret_cast<float4 *>(s_known);
    for (int base = 0; base < P.m; base += kKnownTile) {
      const int tn = min(kKnownTile, P.m - base);
      __syncthreads();
      for (int t = tid; t < tn; t += kThreads) {
        const float *kp = known + (size_t)(base + t) * 3;
        s_known4[t] = make_float4(kp[0], kp[1], kp[2], 0.f);
      }
      __syncthreads();
      const int hn = (tn + kParts - 1) / kParts;
      const int t0 = min(tn, half * hn), t1 = min(tn, t0 + hn);
      // Ascending k inside this thread's range, so "strictly smaller than the current j-th best" is the
      // reference's cascade; the insert is branch-free, and skipped for the whole warp when no lane's
      // third-best improves (the common case after the first few dozen points).
      // Four points per step: their distances are independent (the loop is bound by instruction latency,
      // two warps per scheduler), ONE vote decides whether any lane improves on any of the four, and the
      // inserts -- no-ops for a point that does not improve -- run in ascending order.
      auto insert = [&](float d, int k) {
        const bool l1 = d < best1, l2 = d < best2, l3 = d < best3;
        best3 = l2 ? best2 : (l3 ? d : best3); i3 = l2 ? i2 : (l3 ? k : i3);
        best2 = l1 ? best1 : (l2 ? d : best2); i2 = l1 ? i1 : (l2 ? k : i2);
        best1 = l1 ? d : best1;                i1 = l1 ? k : i1;
      };
      int t = t0;
      for (; t + 4 <= t1; t += 4) {
        float d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 kq = s_known4[t + e];
          d[e] = sqdist3(ux, uy, uz, kq.x, kq.y, kq.z);
        }
        const float dm = fminf(fminf(d[0], d[1]), fminf(d[2], d[3]));
        // (a NaN distance never improves anything: d < best is false, and fminf drops it here)
        if (__any_sync(0xffffffffu, dm < best3)) {
#pragma unroll
          for (int e = 0; e < 4; ++e) insert(d[e], base + t + e);
        }
      }
      for (; t < t1; ++t) {
        const float4 kq = s_known4[t];
        const float d = sqdist3(ux, uy, uz, kq.x, kq.y, kq.z);
        if (__any_sync(0xffffffffu, d < best3)) insert(d, base + t);
      }
    }
    if (half >= 1) {
      float *mg = s_merge + ((half - 1) * kRows + row) * 6;
      mg[0] = best1; mg[1] = best2; mg[2] = best3;
      mg[3] = __int_as_float(i1); mg[4] = __int_as_float(i2); mg[5] = __int_as_float(i3);
    }
    __syncthreads();
    if (half == 0) {
      // merge the parts' sorted triples one after the other: each is the three smallest (distance, index)
      // pairs of its range, so the three smallest of the union are the reference's result
      float od[3] = {best1, best2, best3};
      int oi[3] = {i1, i2, i3};
#pragma unroll
      for (int part = 1; part < kParts; ++part) {
        const float *mg = s_merge + ((part - 1) * kRows + row) * 6;
        const float ad[3] = {od[0], od[1], od[2]}, bd[3] = {mg[0], mg[1], mg[2]};
        const int ai[3] = {oi[0], oi[1], oi[2]};
        const int bi[3] = {__float_as_int(mg[3]), __float_as_int(mg[4]), __float_as_int(mg[5])};
        int pa = 0, pb = 0;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          // entries never filled keep (inf, 0) on both sides: either choice writes (inf, 0) like the reference
          const float da = pa == 0 ? ad[0] : pa == 1 ? ad[1] : ad[2];
          const float db = pb == 0 ? bd[0] : pb == 1 ? bd[1] : bd[2];
          const int ia = pa == 0 ? ai[0] : pa == 1 ? ai[1] : ai[2];
          const int ib = pb == 0 ? bi[0] : pb == 1 ? bi[1] : bi[2];
          const bool take_b = nn_less(db, ib, da, ia);
          od[t] = take_b ? db : da; oi[t] = take_b ? ib : ia;
          if (take_b) ++pb; else ++pa;
        }
      }
      const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(od[0]), 1e-8f));
      const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(od[1]), 1e-8f));
      const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(od[2]), 1e-8f));
      const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
      NnRow nr;
      nr.i1 = oi[0]; nr.i2 = oi[1]; nr.i3 = oi[2];
      nr.w1 = __fdiv_rn(r1, norm); nr.w2 = __fdiv_rn(r2, norm); nr.w3 = __fdiv_rn(r3, norm);
      s_nn[row] = nr;
    }
    __syncthreads();

    FP_TRACE(1)
    // ---- gather geometry: a warp serves 8 rows of the tile
    const int g4 = lane >> 3;                  // skip channels (canonical layout): lane = (row lane & 7, chunk group g4)
    const int rq = lane >> 3, q8 = lane & 7;   // interpolated channels: lane = (row of a quad, 16-byte chunk)
    for (int c = 0; c < nchunks; ++c) {
      const uint32_t g = gchunk + c, s = g & 1;
      // stage s^1 is free once the MMAs of slice g-1 are done: prefetch the next weights
      if (c + 1 < nchunks) {
        if (c >= 1) mbar_wait(mdone0 + 8 * (s ^ 1), ((g - 1) >> 1) & 1);
        if (tid == 0) {
          mbar_arrive_expect_tx(wfull0 + 8 * (s ^ 1), kWVecs * 16);
          bulk_load(w_addr + (s ^ 1) * kWVecs * 16, P.w + (size_t)(c + 1) * kWVecs, kWVecs * 16,
                    wfull0 + 8 * (s ^ 1));
        }
      }
      if (c < nk1) {
        // ---- layer-1 A operand, channels [c*64, c*64+64) of all 128 rows ---------------------
        uint4 *dst = a_st + s * kAVecs;
        const int k0 = c * kChunk;
        if (k0 < P.c_known) {
          // interpolated channels: a quarter warp takes one row and reads 256 contiguous bytes (two full
          // lines) of each of its three neighbour rows, 32 bytes per lane; the warp's 8 rows are two such
          // passes, all 6 loads of a lane in flight before the first use.  The 16-byte results go to the
          // 128-byte-swizzled K-major layout (row pitch 128 B, chunk ^ (row & 7)), which the 8 lanes of a
          // row write without bank conflicts.
          constexpr int kPasses = kRowsPerWarp / 4;
          float pv[kPasses][3][8];
          float wv[kPasses][3];
#pragma unroll
          for (int it = 0; it < kPasses; ++it) {
            const NnRow nr = s_nn[warp * kRowsPerWarp + it * 4 + rq];
            const float *kf = P.known_feat + (size_t)scene * P.m * P.known_stride + k0 + q8 * 8;
            ldg_row32(kf + (size_t)nr.i1 * P.known_stride, pv[it][0], P.wide);
            ldg_row32(kf + (size_t)nr.i2 * P.known_stride, pv[it][1], P.wide);
            ldg_row32(kf + (size_t)nr.i3 * P.known_stride, pv[it][2], P.wide);
            wv[it][0] = nr.w1; wv[it][1] = nr.w2; wv[it][2] = nr.w3;
          }
#pragma unroll
          for (int it = 0; it < kPasses; ++it) {
            const int r = warp * kRowsPerWarp + it * 4 + rq;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              v[e] = __fmaf_rn(pv[it][2][e], wv[it][2], __fmaf_rn(pv[it][0][e], wv[it][0],
                                                                 __fmul_rn(pv[it][1][e], wv[it][1])));
            dst[r * 8 + (q8 ^ (r & 7))] = make_uint4(pack16(v[0], v[1], P.fp16), pack16(v[2], v[3], P.fp16),
                                                     pack16(v[4], v[5], P.fp16), pack16(v[6], v[7], P.fp16));
          }
        } else {
#pragma unroll
        for (int it = 0; it < kRowsPerWarp / 8; ++it) {
          const int r = warp * kRowsPerWarp + it * 8 + (lane & 7);
          const int jr = min(row0 + r, P.n - 1);
          if (skip_early) {
            // already in the X1 buffer
          } else if (P.skip16) {
            // the previous fused layer's 16-bit twin: the very bits this operand needs
            const uint16_t *src = P.skip16 + ((size_t)scene * P.n + jr) * P.skip16_stride + (k0 - P.c_known);
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              const int q = g4 + 4 * qq;
              cp_async16(smem_u32(dst + q * kRows + r), src + q * 8);
            }
          } else {
            const float *src = P.skip_feat + ((size_t)scene * P.n + jr) * P.skip_stride + (k0 - P.c_known);
#pragma unroll
            for (int qq = 0; qq < 2; ++qq) {
              const int q = g4 + 4 * qq;
              const float4 a = __ldg(reinterpret_cast<const float4 *>(src + q * 8));
              const float4 bq = __ldg(reinterpret_cast<const float4 *>(src + q * 8 + 4));
              dst[q * kRows + r] = make_uint4(pack16(a.x, a.y, P.fp16), pack16(a.z, a.w, P.fp16),
                                              pack16(bq.x, bq.y, P.fp16), pack16(bq.z, bq.w, P.fp16));
            }
          }
        }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
      }
      FP_TRACE(2 + 3 * c)
      umma::fence_proxy_async_smem();
      umma::fence_before_sync();
      __syncthreads();
      FP_TRACE(3 + 3 * c)
      if (tid == 0) {
        mbar_wait(wfull0 + 8 * s, (g >> 1) & 1);
        umma::fence_after_sync();
        const bool layer1 = c < nk1;
        const int k0c = c * kChunk;
        const bool early = layer1 && skip_early && k0c >= P.c_known;
        const uint32_t abase = early ? x_addr + (uint32_t)((k0c - P.c_known) / kChunk) * (kRows * 128)
                               : layer1 ? a_addr + s * kAVecs * 16
                                        : x_addr + (uint32_t)(c - nk1) * (kChunk / 8) * kRows * 16;
#pragma unroll
        for (int ks = 0; ks < kChunk / 16; ++ks) {
          // interpolated slices sit in the 128-byte-swizzled layout (8-row groups 1024 B apart, +32 B per
          // K step inside the atom), everything else in the canonical one
          const uint64_t ad = (layer1 && (k0c < P.c_known || early))
                                  ? umma::smem_desc_sw128(abase + (uint32_t)ks * 32, 1024)
                                  : umma::smem_desc(abase + (uint32_t)(2 * ks) * kRows * 16, kRows * 16, 128);
          const uint64_t bd = umma::smem_desc(w_addr + s * kWVecs * 16 + (uint32_t)(2 * ks) * kWidth * 16,
                                              kWidth * 16, 128);
          const uint32_t acc = layer1 ? ((c | ks) != 0) : (((c - nk1) | ks) != 0);
          umma::mma_bf16_ss(layer1 ? tmem_d1 : tmem_d2, ad, bd, idesc, acc);
        }
        umma::commit(mdone0 + 8 * s);
      }
      FP_TRACE(4 + 3 * c)
      if (c == nk1 - 1) {
        // ---- epilogue 1: X1 = relu(D1 + b1) as the 16-bit A operand of layer 2; the two threads of a
        // row convert 128 columns each
        mbar_wait(mdone0 + 8 * s, (g >> 1) & 1);
        umma::fence_after_sync();
#pragma unroll 1
        for (int c0 = half * (kWidth / kParts); c0 < (half + 1) * (kWidth / kParts); c0 += 32) {
          uint32_t v[32];
          umma::ld_32x32b_x32(tmem_d1 + lane_addr + (uint32_t)c0, v);
          umma::wait_ld();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t p[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int col = c0 + q * 8 + e * 2;
              const float lo = __uint_as_float(v[q * 8 + e * 2]) + s_b1[col];
              const float hi = __uint_as_float(v[q * 8 + e * 2 + 1]) + s_b1[col + 1];
              p[e] = pack16_relu(lo, hi, P.fp16);
            }
            x1[(c0 / 8 + q) * kRows + row] = make_uint4(p[0], p[1], p[2], p[3]);
          }
        }
      }
    }
    gchunk += nchunks;
    FP_TRACE(40)

    // ---- epilogue 2: out = relu(D2 + b2), channel-major and (optionally) point-major ----------
    {
      const uint32_t g = gchunk - 1, s = g & 1;
      mbar_wait(mdone0 + 8 * s, (g >> 1) & 1);
      umma::fence_after_sync();
      // channel-major: lanes = consecutive points, one full line per warp store.  Point-major: the 32 x 32
      // block a warp just read is turned through a 4 KB shared-memory patch (the weight stages are idle
      // now; 16-byte chunks swizzled by row & 7, conflict-free both ways) so that a quarter warp stores 128
      // contiguous bytes of one row -- a lane-per-row store would touch 32 lines per instruction.
      float *ocm = P.out_cm + (size_t)scene * kWidth * P.n + j;
      float4 *patch = reinterpret_cast<float4 *>(w_st) + warp * 256;
#pragma unroll 1
      for (int c0 = half * (kWidth / kParts); c0 < (half + 1) * (kWidth / kParts); c0 += 32) {
        uint32_t v[32];
        umma::ld_32x32b_x32(tmem_d2 + lane_addr + (uint32_t)c0, v);
        umma::wait_ld();
        float o[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) o[e] = fmaxf(__uint_as_float(v[e]) + s_b2[c0 + e], 0.f);
        if (live) {
#pragma unroll
          for (int e = 0; e < 32; ++e) ocm[(size_t)(c0 + e) * P.n] = o[e];     // lanes = consecutive points
        }
        if (P.out_pm) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            patch[lane * 8 + (e ^ (lane & 7))] = make_float4(o[4 * e], o[4 * e + 1], o[4 * e + 2], o[4 * e + 3]);
          __syncwarp();
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const int rr = p * 4 + rq;                                          // row of this warp's 32
            const int jj = row0 + quarter * 32 + rr;
            const float4 t = patch[rr * 8 + (q8 ^ (rr & 7))];
            if (jj < P.n)
              *reinterpret_cast<float4 *>(P.out_pm + ((size_t)scene * P.n + jj) * kWidth + c0 + q8 * 4) = t;
          }
          __syncwarp();
        }
      }
    }
    FP_TRACE(41)
    umma::fence_before_sync();
    __syncthreads();       // TMEM, X1 and both stages are free for the next tile
    umma::fence_after_sync();
    FP_TRACE(42)
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace

int fp_supported(int n, int m, int c_known, int c_skip, int c1, int c2) {
  if (c1 != kWidth || c2 != kWidth) return 0;
  if (c_known <= 0 || c_skip <= 0 || c_known % kChunk || c_skip % kChunk) return 0;
  if (n <= 0 || m <= 0) return 0;
  return 1;
}

int fp_forward_dispatch(int b, int n, int m, int c_known, int c_skip, const float *unknown,
                        const float *known, const float *known_feat, int known_stride,
                        const float *skip_feat, int skip_stride, int c1, int c2, const void *w,
                        const float *b1, const float *b2, float *out_cm, float *out_pm, int fp16,
                        cudaStream_t stream, const void *skip16, int skip16_stride) {
  if (!fp_supported(n, m, c_known, c_skip, c1, c2))
    return set_error(BQA_ERR_UNSUPPORTED, "fp_mlp: unsupported shape c_known=%d c_skip=%d mlp=%d,%d",
                     c_known, c_skip, c1, c2);
  if ((known_stride % 4) || (reinterpret_cast<uintptr_t>(known_feat) & 15) ||
      (skip_feat && ((skip_stride % 4) || (reinterpret_cast<uintptr_t>(skip_feat) & 15))) ||
      (reinterpret_cast<uintptr_t>(out_pm) & 15))
    return set_error(BQA_ERR_INVALID_ARG, "fp_mlp: point-major features must be 16-byte aligned rows");
  FpParams P;
  P.b = b; P.n = n; P.m = m; P.c_known = c_known; P.c_skip = c_skip;
  P.known_stride = known_stride; P.skip_stride = skip_stride;
  P.unknown = unknown; P.known = known; P.known_feat = known_feat; P.skip_feat = skip_feat;
  P.skip16 = (const uint16_t *)skip16; P.skip16_stride = skip16_stride;
  if (skip16 && ((skip16_stride % 8) || skip16_stride < c_skip || (reinterpret_cast<uintptr_t>(skip16) & 15)))
    return set_error(BQA_ERR_INVALID_ARG, "fp_mlp: skip16 rows must be 16-byte aligned and hold c_skip values");
  P.w = (const uint4 *)w; P.b1 = b1; P.b2 = b2; P.out_cm = out_cm; P.out_pm = out_pm; P.fp16 = fp16;
  P.wide = (known_stride % 8 == 0) && !(reinterpret_cast<uintptr_t>(known_feat) & 31);
  P.tiles_per_scene = ceil_div(n, kRows);
  P.num_tiles = b * P.tiles_per_scene;
  const size_t smem = 16 * (2 * (size_t)(kChunk / 8 * kRows) + 2 * (size_t)(kChunk / 8 * kWidth) +
                            (size_t)(kWidth / 8 * kRows)) + 4 * (kKnownTile * 4 + 2 * kWidth) +
                      sizeof(NnRow) * kRows + 4 * 6 * kRows * (kParts - 1) + 32 + 16 + 1024;
  BQA_CUDA(cudaFuncSetAttribute(fp_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  BQA_CUDA(cudaGetDevice(&dev));
  BQA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = min(P.num_tiles, sms);
#ifdef BQA_SA_TRACE
  P.dbg = getenv("BQA_FP_DBG") ? atoi(getenv("BQA_FP_DBG")) : 0;
#endif
  fp_mlp_kernel<<<grid, kThreads, smem, stream>>>(P);
#ifdef BQA_SA_TRACE
  {
    long long t[64];
    cudaMemcpyFromSymbol(t, g_fp_trace, sizeof(t));
    fprintf(stderr, "[bqa fp trace] n=%d m=%d grid=%d | nn %lld |", n, m, grid, t[1] - t[0]);
    for (int c = 0; c < 12; ++c)
      fprintf(stderr, " c%d: build %lld sync %lld issue %lld |", c, t[2 + 3 * c] - (c ? t[4 + 3 * (c - 1)] : t[1]),
              t[3 + 3 * c] - t[2 + 3 * c], t[4 + 3 * c] - t[3 + 3 * c]);
    fprintf(stderr, " epi2 %lld | tail %lld | total %lld\n", t[41] - t[40], t[42] - t[41], t[42] - t[0]);
  }
#endif
  count_launch();
  return check_launch("fp_mlp_kernel");
}

}  // namespace bqa
