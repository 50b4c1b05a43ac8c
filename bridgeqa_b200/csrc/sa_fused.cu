// sa_fused.cu -- fused set-abstraction layer for inference on sm_100a:
//   gather neighbours -> [1x1 conv + folded BN + ReLU] x3 on tcgen05 tensor cores -> max
//   over nsample, without ever writing the grouped (B, C+3, npoint, nsample) tensor or
//   any intermediate activation to HBM.
//
// Replaces, for eval-mode forward, the chain the reference runs as ~12 separate kernels
// with full HBM round trips: QueryAndGroup.forward (pointnet2_utils.py:317-376; two
// group_points launches, subtract, divide, cat), SharedMLP (pytorch_utils.py:11-36; 3 x
// [cuDNN conv, BN, ReLU]) and F.max_pool2d (pointnet2_modules.py:259-262).
//
// One CTA of 128 threads owns a tile of 128 "rows" (row = one (centre, sample) pair;
// 128/nsample consecutive centres of one scene) and walks tiles persistently.
//
//   layer 1   D1[128 x C1]  = A1[128 x K1] * W1^T     A1 = gathered rows (bf16, smem)
//   layer 2   D2[128 x C2]  = X1[128 x C1] * W2^T     X1 = relu(D1 + b1)  (bf16, smem)
//   layer 3   D3T[C3 x 128] = W3[C3 x C2]  * X2^T     X2 = relu(D2 + b2)  (bf16, smem)
//
// Layers 1-2 keep rows on TMEM lanes, so an epilogue thread owns a row and stores it as
// 16-byte K-major vectors for the next MMA; layer 3 is issued TRANSPOSED (weights as the
// A operand) so that TMEM lanes are output channels and the 128 rows are TMEM columns:
// the max over nsample consecutive rows becomes a register-local max for the thread that
// owns the channel (no shuffles), after which bias + ReLU are applied once
// (max(relu(x+b)) == relu(max(x)+b)).
//
// Accumulators live in TMEM (D1|D2 side by side, D3T aliases them once they are dead);
// operands are bf16 with fp32 accumulation; layer-1 K is permuted to
// [features(C), xyz_rel(3), 0-pad] (the host packs W1 the same way) and processed in
// chunks of 128 so that the weights of all three layers stay resident in shared memory.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bqa {
namespace {

constexpr int kRows = 128;       // rows per tile == threads per CTA
constexpr int kKChunk = 128;     // layer-1 K processed per pass

// two fp32 -> one packed 16-bit pair, ONE instruction (SASS F2FP...PACK_AB).  fp16 mode saturates
// at +-65504 instead of producing inf (bf16 has fp32's range and needs no clamp); the _relu form
// folds max(x, 0) into the conversion.
__device__ __forceinline__ uint32_t pack2(float lo, float hi, int fp16) {
  uint32_t d;
  if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi, int fp16) {
  uint32_t d;
  if (fp16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

struct SaParams {
  int b, n, npoint, nsample, c;     // c = feature channels (K1 = c + 3); npoint = centres in this call
  int npoint_total, j_offset;       // the call covers centres [j_offset, j_offset + npoint) of npoint_total
  int k1pad;                        // K1 rounded up to a multiple of 16
  int tail_sep;                     // the last 16 channels of K1 have their own A buffer (see kernel)
  int feat_stride;                  // floats between consecutive points of feat_pm (>= c)
  int feat_vec4;                    // rows are 16-byte aligned: float4 loads allowed
  const float *xyz, *new_xyz, *feat_pm;
  const int *idx;
  float radius;
  int normalize_xyz;
  int fp16;                         // operands: 1 = fp16 (10-bit mantissa, saturating), 0 = bf16
  const uint4 *w1p, *w2p, *w3p;     // packed bf16 images [K/8][rows] of 16-byte vectors
  const float *b1, *b2, *b3;
  float *out_cm, *out_pm;
  int num_tiles;
};

// epilogue of a row-on-lane accumulator: X[r][0..CN) = relu(D[r][*] + bias) as bf16, stored
// chunk-major ([CN/8][128] x 16 B) for the next MMA.
// CB..CE: the column range this thread converts (a pipeline of 8 warps splits the columns of a row
// between warp w and warp w + 4, which both reach TMEM lane quarter w)
template <int CN>
__device__ __forceinline__ void epilogue_rows(uint32_t tmem_d, int warp, int row, const float *s_bias,
                                              uint4 *x_buf, int fp16, int cb, int ce) {
  static_assert(CN % 64 == 0, "two 32-column TMEM loads per step");
#pragma unroll 1
  for (int c0 = cb; c0 < ce; c0 += 64) {
    // two tcgen05.ld in flight, one wait: the load latency is paid once per 64 columns
    uint32_t va[32], vb[32];
    umma::ld_32x32b_x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, va);
    umma::ld_32x32b_x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c0 + 32), vb);
    umma::wait_ld();
    uint32_t v[64];
#pragma unroll
    for (int e = 0; e < 32; ++e) { v[e] = va[e]; v[32 + e] = vb[e]; }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t p[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = c0 + q * 8 + e * 2;
        const float lo = __uint_as_float(v[q * 8 + e * 2]) + s_bias[col];
        const float hi = __uint_as_float(v[q * 8 + e * 2 + 1]) + s_bias[col + 1];
        p[e] = pack2_relu(lo, hi, fp16);                 // relu(D + b), rounded once
      }
      x_buf[(c0 / 8 + q) * kRows + row] = make_uint4(p[0], p[1], p[2], p[3]);
    }
  }
}

#ifdef BQA_SA_TRACE
__device__ unsigned long long g_sa_trace[16];
#define SA_TRACE_BEGIN const bool tr = blockIdx.x == 0 && threadIdx.x == 0; long long tc = clock64();
#define SA_TRACE(i) if (tr) { const long long now = clock64(); g_sa_trace[i] += (unsigned long long)(now - tc); tc = now; }
#else
#define SA_TRACE_BEGIN
#define SA_TRACE(i)
#endif

// named barrier over the threads of one tile pipeline (id 1 or 2; 0 is __syncthreads)
template <int THREADS>
__device__ __forceinline__ void group_sync(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(THREADS) : "memory");
}

// Two independent tile pipelines per CTA (threads 0-127 and 128-255) share one copy of the
// weights in shared memory; each has its own A/X buffer, mbarrier and TMEM columns, so the
// gather / epilogue of one tile overlaps the MMAs of the other.
// W = warps per pipeline: 4 (thread = row) or 8 (two threads per row: warp w and warp w + 4 share TMEM
// lane quarter w and split the row's gather chunks and epilogue columns, which halves the two
// longest latency chains of a tile).
template <int C1, int C2, int C3, int G, int W>
__global__ void __launch_bounds__(G * W * 32)
sa_mlp_max_kernel(const SaParams P) {
  static_assert(G == 1 || G == 2, "one or two tile pipelines per CTA");
  static_assert(W == 4 || W == 8, "4 or 8 warps per pipeline");
  constexpr int kPT = W * 32;                                         // threads per pipeline
  static_assert(C1 % 32 == 0 && C2 % 32 == 0 && C3 % 128 == 0, "channel counts");
  constexpr int kXVecs = (C1 > C2 ? C1 : C2) / 8 * kRows;             // X1 / X2 buffer
  constexpr int kAVecs = kKChunk / 8 * kRows;                         // layer-1 A chunk
  constexpr int kAXVecs = kXVecs > kAVecs ? kXVecs : kAVecs;          // they alias
  constexpr uint32_t kTmemCols = (C1 + C2 > C3 ? C1 + C2 : C3) <= 128 ? 128 : 256;   // per pipeline
  static_assert((C1 + C2 > C3 ? C1 + C2 : C3) <= 256, "TMEM budget");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kTailVecs = 2 * kRows;                                // 16 channels x 128 rows
  uint4 *ax_all = reinterpret_cast<uint4 *>(smem_raw);                // G x (A1 chunk / X1 / X2)
  uint4 *tail_all = ax_all + G * kAXVecs;                             // G x last 16 channels of A1
  uint4 *w1s = tail_all + (P.tail_sep ? G * kTailVecs : 0);           // [k1pad/8][C1]
  uint4 *w2s = w1s + (P.k1pad / 8) * C1;                              // [C1/8][C2]
  uint4 *w3s = w2s + (C1 / 8) * C2;                                   // [C2/8][C3]
  float *s_b1 = reinterpret_cast<float *>(w3s + (C2 / 8) * C3);
  float *s_b2 = s_b1 + C1;
  float *s_b3 = s_b2 + C2;
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_b3 + C3);          // one per pipeline
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_bar + 2);

  const int group = threadIdx.x / kPT;      // pipeline 0 / 1
  const int ptid = threadIdx.x % kPT;       // thread of the pipeline
  const int tid = ptid & 127;               // row of the pipeline's tile
  const int half = ptid >> 7;               // 0, or 1 for the second thread of the row (W == 8)
  const int warp = tid >> 5;                // TMEM lane quarter (warp_id % 4 of the real warp)
  uint4 *ax = ax_all + group * kAXVecs;
  uint4 *ax_tail = tail_all + group * kTailVecs;

  // ---- one-time setup: weights + biases -> smem, mbarriers, TMEM ----------------------
  for (int i = threadIdx.x; i < (P.k1pad / 8) * C1; i += G * kPT) w1s[i] = P.w1p[i];
  for (int i = threadIdx.x; i < (C1 / 8) * C2; i += G * kPT) w2s[i] = P.w2p[i];
  for (int i = threadIdx.x; i < (C2 / 8) * C3; i += G * kPT) w3s[i] = P.w3p[i];
  for (int i = threadIdx.x; i < C1; i += G * kPT) s_b1[i] = P.b1[i];
  for (int i = threadIdx.x; i < C2; i += G * kPT) s_b2[i] = P.b2[i];
  for (int i = threadIdx.x; i < C3; i += G * kPT) s_b3[i] = P.b3[i];
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    mbar_init(smem_u32(&s_bar[1]), 1);
    fence_mbar_init_cluster();
  }
  if (threadIdx.x < 32) umma::tmem_alloc(smem_u32(s_tmem), G * kTmemCols);
  umma::fence_proxy_async_smem();          // weights were written with st.shared
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem = tmem_base + group * kTmemCols;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + C1, tmem_d3 = tmem;
  const uint32_t bar = smem_u32(&s_bar[group]);

  const uint32_t ax_tail_addr = smem_u32(ax_tail);
  const uint32_t ax_addr = smem_u32(ax), w1_addr = smem_u32(w1s), w2_addr = smem_u32(w2s),
                 w3_addr = smem_u32(w3s);
  const uint32_t kIdesc1 = umma::instr_desc_16b_f32(128, C1, !P.fp16);
  const uint32_t kIdesc2 = umma::instr_desc_16b_f32(128, C2, !P.fp16);
  const uint32_t kIdesc3 = umma::instr_desc_16b_f32(128, 128, !P.fp16);
  uint32_t phase = 0;

  const int ns = P.nsample;
  const int centres_per_tile = kRows / ns;
  const int tiles_per_scene = P.npoint / centres_per_tile;

  // Per-row inputs of a tile: neighbour index i, its coordinates pp, the centre qq.  They are two
  // dependent global loads (~2.7k cycles per tile when taken at the top of the tile), so the next
  // tile's are fetched while this tile computes: i and qq at the top, pp (needs i) after layer 1.
  auto row_of = [&](int tile, int &scene, int &j) {
    scene = tile / tiles_per_scene;
    j = P.j_offset + (tile % tiles_per_scene) * centres_per_tile + tid / ns;
  };
  const int tile_step = gridDim.x * G;
  int tile = blockIdx.x * G + group;
  int nx_i = 0;
  float nx_pp[3] = {0.f, 0.f, 0.f}, nx_qq[3] = {0.f, 0.f, 0.f};
  if (tile < P.num_tiles) {
    int scene, j;
    row_of(tile, scene, j);
    nx_i = P.idx[((size_t)scene * P.npoint_total + j) * ns + (tid % ns)];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      nx_pp[d] = P.xyz[((size_t)scene * P.n + nx_i) * 3 + d];
      nx_qq[d] = P.new_xyz[((size_t)scene * P.npoint_total + j) * 3 + d];
    }
  }
  SA_TRACE_BEGIN
  for (; tile < P.num_tiles; tile += tile_step) {
    SA_TRACE(9)
    const int scene = tile / tiles_per_scene;
    const int centre0 = (tile % tiles_per_scene) * centres_per_tile;
    const int i = nx_i;
    const float *frow = P.feat_pm ? P.feat_pm + ((size_t)scene * P.n + i) * P.feat_stride : nullptr;
    float rel[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = nx_pp[d] - nx_qq[d];                   // pointnet2_utils.py:350
      if (P.normalize_xyz) v = v / P.radius;           // :351-352, true division
      rel[d] = v;
    }
    // next tile: index and centre now, neighbour coordinates once the index has arrived
    const int ntile = tile + tile_step;
    int n_scene = 0, n_j = 0;
    if (ntile < P.num_tiles) {
      row_of(ntile, n_scene, n_j);
      nx_i = __ldg(P.idx + ((size_t)n_scene * P.npoint_total + n_j) * ns + (tid % ns));
#pragma unroll
      for (int d = 0; d < 3; ++d) nx_qq[d] = __ldg(P.new_xyz + ((size_t)n_scene * P.npoint_total + n_j) * 3 + d);
    }

    SA_TRACE(0)
    // ---- layer 1: gather K-chunks of A1 and accumulate D1 ------------------------------
    // one 16-byte vector (8 channels) of this row: features, then xyz_rel, then zero padding
    auto gather_q = [&](int k0, uint4 *dst) {
      float f[8];
      if (k0 + 8 <= P.c && P.feat_vec4) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(frow + k0));
        const float4 bq = __ldg(reinterpret_cast<const float4 *>(frow + k0 + 4));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = bq.x; f[5] = bq.y; f[6] = bq.z; f[7] = bq.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = k0 + e;
          float v = 0.f;
          if (k < P.c) v = __ldg(frow + k);
          else if (k < P.c + 3) v = (k - P.c == 0) ? rel[0] : ((k - P.c == 1) ? rel[1] : rel[2]);
          f[e] = v;
        }
      }
      *dst = make_uint4(pack2(f[0], f[1], P.fp16), pack2(f[2], f[3], P.fp16),
                        pack2(f[4], f[5], P.fp16), pack2(f[6], f[7], P.fp16));
    };
    // with tail_sep the last 16 channels (k1pad % 128 == 16) live in their own buffer, so they are
    // gathered together with the last full chunk and need no MMA round trip of their own
    const int k_main = P.tail_sep ? P.k1pad - 16 : P.k1pad;
    for (int kc0 = 0; kc0 < k_main; kc0 += kKChunk) {
      const int kcn = min(kKChunk, k_main - kc0);      // multiple of 16
      const int nq = kcn / 8;
      const bool last = kc0 + kKChunk >= k_main;
      // 16-byte-aligned all-feature chunks: 4 chunks (8 independent 16-byte loads) in flight
      // (8 chunks in flight measured 1.8x SLOWER per tile: 10.8k vs 6.2k cycles at SA2)
      // (with two threads per row the rounds of 4 chunks alternate between them)
      int q = 0;
      if (P.feat_vec4) {
        for (; q + 4 <= nq && kc0 + (q + 4) * 8 <= P.c; q += 4) {
          if (W == 8 && ((q >> 2) & 1) != half) continue;
          float4 lo[4], hi[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            lo[u] = __ldg(reinterpret_cast<const float4 *>(frow + kc0 + (q + u) * 8));
            hi[u] = __ldg(reinterpret_cast<const float4 *>(frow + kc0 + (q + u) * 8 + 4));
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            ax[(q + u) * kRows + tid] =
                make_uint4(pack2(lo[u].x, lo[u].y, P.fp16), pack2(lo[u].z, lo[u].w, P.fp16),
                           pack2(hi[u].x, hi[u].y, P.fp16), pack2(hi[u].z, hi[u].w, P.fp16));
        }
      }
      for (; q < nq; ++q)
        if (W == 4 || (q & 1) == half) gather_q(kc0 + q * 8, &ax[q * kRows + tid]);
      if (last && P.tail_sep) {
        if (W == 4 || half == 0) gather_q(k_main, &ax_tail[tid]);
        if (W == 4 || half == 1) gather_q(k_main + 8, &ax_tail[kRows + tid]);
      }
      SA_TRACE(1)
      umma::fence_proxy_async_smem();
      umma::fence_before_sync();
      group_sync<W * 32>(group);
      SA_TRACE(2)
      if (ptid == 0) {
        umma::fence_after_sync();
        for (int ks = 0; ks < kcn / 16; ++ks) {
          const uint64_t ad = umma::smem_desc(ax_addr + (uint32_t)(2 * ks) * kRows * 16, kRows * 16, 128);
          const uint64_t bd = umma::smem_desc(w1_addr + (uint32_t)(kc0 / 8 + 2 * ks) * C1 * 16, C1 * 16, 128);
          umma::mma_bf16_ss(tmem_d1, ad, bd, kIdesc1, (kc0 | ks) != 0);
        }
        if (last && P.tail_sep) {
          const uint64_t ad = umma::smem_desc(ax_tail_addr, kRows * 16, 128);
          const uint64_t bd = umma::smem_desc(w1_addr + (uint32_t)(k_main / 8) * C1 * 16, C1 * 16, 128);
          umma::mma_bf16_ss(tmem_d1, ad, bd, kIdesc1, true);
        }
        umma::commit(bar);
      }
      if (last && ntile < P.num_tiles) {
        // the next tile's index arrived long ago: fetch its neighbour's coordinates under the MMAs
#pragma unroll
        for (int d = 0; d < 3; ++d) nx_pp[d] = __ldg(P.xyz + ((size_t)n_scene * P.n + nx_i) * 3 + d);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      umma::fence_after_sync();
      SA_TRACE(3)
    }

    // ---- epilogue 1 -> X1 ; layer 2 ---------------------------------------------------
    epilogue_rows<C1>(tmem_d1, warp, tid, s_b1, ax, P.fp16, W == 8 ? half * (C1 / 2) : 0,
                      W == 8 ? (half + 1) * (C1 / 2) : C1);
    SA_TRACE(4)
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    group_sync<W * 32>(group);
    if (ptid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int ks = 0; ks < C1 / 16; ++ks) {
        const uint64_t ad = umma::smem_desc(ax_addr + (uint32_t)(2 * ks) * kRows * 16, kRows * 16, 128);
        const uint64_t bd = umma::smem_desc(w2_addr + (uint32_t)(2 * ks) * C2 * 16, C2 * 16, 128);
        umma::mma_bf16_ss(tmem_d2, ad, bd, kIdesc2, ks != 0);
      }
      umma::commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();

    SA_TRACE(5)
    // ---- epilogue 2 -> X2 ; layer 3 (transposed: channels on lanes) ---------------------
    epilogue_rows<C2>(tmem_d2, warp, tid, s_b2, ax, P.fp16, W == 8 ? half * (C2 / 2) : 0,
                      W == 8 ? (half + 1) * (C2 / 2) : C2);
    SA_TRACE(6)
    umma::fence_proxy_async_smem();
    umma::fence_before_sync();
    group_sync<W * 32>(group);
    if (ptid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int mt = 0; mt < C3 / 128; ++mt) {
#pragma unroll
        for (int ks = 0; ks < C2 / 16; ++ks) {
          const uint64_t ad = umma::smem_desc(w3_addr + ((uint32_t)(2 * ks) * C3 + mt * 128) * 16, C3 * 16, 128);
          const uint64_t bd = umma::smem_desc(ax_addr + (uint32_t)(2 * ks) * kRows * 16, kRows * 16, 128);
          umma::mma_bf16_ss(tmem_d3 + mt * 128, ad, bd, kIdesc3, ks != 0);
        }
      }
      umma::commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();

    SA_TRACE(7)
    // ---- epilogue 3: max over nsample columns, bias, ReLU, store -------------------------
#pragma unroll
    for (int mt = 0; mt < C3 / 128; ++mt) {
      const int ch = mt * 128 + tid;
      const float bias = s_b3[ch];
      float run = -INFINITY;
      // (W == 8: each of the row's two threads folds one half of the 128 position columns;
      //  a group of nsample <= 64 columns never straddles the halves)
#pragma unroll 1
      for (int c0 = (W == 8 ? half * 64 : 0); c0 < (W == 8 ? half * 64 + 64 : kRows); c0 += 64) {
        uint32_t va[32], vb[32];
        umma::ld_32x32b_x32(tmem_d3 + mt * 128 + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, va);
        umma::ld_32x32b_x32(tmem_d3 + mt * 128 + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c0 + 32), vb);
        umma::wait_ld();
        uint32_t v[64];
#pragma unroll
        for (int e = 0; e < 32; ++e) { v[e] = va[e]; v[32 + e] = vb[e]; }
#pragma unroll
        for (int g = 0; g < 64; g += 16) {              // nsample is a multiple of 16
          float m16 = __uint_as_float(v[g]);
#pragma unroll
          for (int e = 1; e < 16; ++e) m16 = fmaxf(m16, __uint_as_float(v[g + e]));
          run = fmaxf(run, m16);
          const int col_end = c0 + g + 16;               // rows [.., col_end) folded so far
          if (col_end % ns == 0) {
            const int jj = P.j_offset + centre0 + col_end / ns - 1;
            const float o = fmaxf(run + bias, 0.f);
            P.out_cm[((size_t)scene * C3 + ch) * P.npoint_total + jj] = o;
            if (P.out_pm) P.out_pm[((size_t)scene * P.npoint_total + jj) * C3 + ch] = o;
            run = -INFINITY;
          }
        }
      }
    }
    SA_TRACE(8)
    umma::fence_before_sync();
    group_sync<W * 32>(group);     // TMEM (D3T aliases D1/D2) and the A/X buffer are free again
    umma::fence_after_sync();
  }

  umma::fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tmem_base, G * kTmemCols);
}

// w (c_out, c_in) f32 -> bf16 image [kpad/8][c_out][8]; when xyz_first, source column
// order [xyz(3), feat(c_in-3)] becomes packed K order [feat, xyz, 0...].
__global__ void pack_weight_kernel(int c_out, int c_in, int kpad, int xyz_first, int fp16,
                                   const float *__restrict__ w, uint16_t *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c_out * kpad) return;
  const int k = t % kpad, row = t / kpad;
  float v = 0.f;
  if (k < c_in) {
    int src = k;
    if (xyz_first) src = (k < c_in - 3) ? k + 3 : k - (c_in - 3);
    v = w[(size_t)row * c_in + src];
  }
  uint16_t bits16;
  if (fp16) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    bits16 = *reinterpret_cast<const uint16_t *>(&h);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    bits16 = *reinterpret_cast<const uint16_t *>(&h);
  }
  out[((size_t)(k / 8) * c_out + row) * 8 + (k % 8)] = bits16;
}

template <int C1, int C2, int C3>
size_t sa_smem_bytes(int k1pad, int g, bool tail_sep = false) {
  constexpr int kXVecs = (C1 > C2 ? C1 : C2) / 8 * kRows;
  constexpr int kAVecs = kKChunk / 8 * kRows;
  constexpr int kAXVecs = kXVecs > kAVecs ? kXVecs : kAVecs;
  return 16 * ((size_t)g * (kAXVecs + (tail_sep ? 2 * kRows : 0)) + (size_t)(k1pad / 8) * C1 +
               (size_t)(C1 / 8) * C2 + (size_t)(C2 / 8) * C3) + 4 * (C1 + C2 + C3) + 32;
}

template <int C1, int C2, int C3, int G>
int sa_ctas_per_sm(size_t smem) {
  // CTAs per SM: shared memory (228 KB per SM, 1 KB reserved per CTA) and TMEM (512 columns)
  constexpr int kTmemCols = G * ((C1 + C2 > C3 ? C1 + C2 : C3) <= 128 ? 128 : 256);
  const int per_sm = (int)((228 * 1024) / (smem + 1024));
  return max(1, min(per_sm, 512 / kTmemCols));
}

template <int C1, int C2, int C3, int G, int W>
int launch_sa_g(SaParams P, cudaStream_t stream) {
  size_t smem = sa_smem_bytes<C1, C2, C3>(P.k1pad, G);
  // own buffer for a 16-channel K tail when it costs neither the launch nor a resident CTA
  P.tail_sep = 0;
  if (P.k1pad > 16 && P.k1pad % kKChunk == 16) {
    const size_t with_tail = sa_smem_bytes<C1, C2, C3>(P.k1pad, G, true);
    if (with_tail <= 227 * 1024 &&
        sa_ctas_per_sm<C1, C2, C3, G>(with_tail) == sa_ctas_per_sm<C1, C2, C3, G>(smem)) {
      P.tail_sep = 1;
      smem = with_tail;
    }
  }
  if (smem > 227 * 1024)
    return set_error(BQA_ERR_UNSUPPORTED, "sa_mlp_max: needs %zu bytes of shared memory", smem);
  auto kern = sa_mlp_max_kernel<C1, C2, C3, G, W>;
  BQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  BQA_CUDA(cudaGetDevice(&dev));
  BQA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int per_sm = sa_ctas_per_sm<C1, C2, C3, G>(smem);
  const int grid = min((P.num_tiles + G - 1) / G, sms * per_sm);
#ifdef BQA_SA_TRACE
  unsigned long long zeros[16] = {0};
  cudaMemcpyToSymbol(g_sa_trace, zeros, sizeof(zeros));
#endif
  kern<<<grid, G * W * 32, smem, stream>>>(P);
#ifdef BQA_SA_TRACE
  unsigned long long t[16];
  cudaMemcpyFromSymbol(t, g_sa_trace, sizeof(t));
  const double tiles = (double)((P.num_tiles - 0 + grid * G - 1) / (grid * G));
  fprintf(stderr, "[bqa sa trace] <%d,%d,%d,G=%d,W=%d> c=%d ns=%d grid=%d tiles/pipeline=%.0f | cycles per tile: idx+xyz %.0f | "
          "gather+pack %.0f | sync %.0f | mma1 %.0f | epi1 %.0f | sync+mma2 %.0f | epi2 %.0f | sync+mma3 %.0f | epi3 %.0f | "
          "tail sync+loop %.0f\n", C1, C2, C3, G, W, P.c, P.nsample, grid, tiles, t[0] / tiles, t[1] / tiles, t[2] / tiles,
          t[3] / tiles, t[4] / tiles, t[5] / tiles, t[6] / tiles, t[7] / tiles, t[8] / tiles, t[9] / tiles);
#endif
  count_launch();
  return check_launch("sa_mlp_max_kernel");
}

// Two tile pipelines per CTA when both A/X buffers fit beside the resident weights.  Eight warps per
// pipeline (two threads per row) when only one CTA fits an SM anyway -- then the 512 (or 256)
// threads still get 128 registers each -- and a group of nsample columns fits one half tile.
// BQA_SA_WARPS=4 forces the four-warp pipelines (measurement).
template <int C1, int C2, int C3>
int launch_sa(const SaParams &P, cudaStream_t stream) {
  static const int forced = [] { const char *e = getenv("BQA_SA_WARPS"); return e ? atoi(e) : 0; }();
  const bool two = sa_smem_bytes<C1, C2, C3>(P.k1pad, 2) <= 227 * 1024;
  const size_t smem = sa_smem_bytes<C1, C2, C3>(P.k1pad, two ? 2 : 1);
  const int per_sm = two ? sa_ctas_per_sm<C1, C2, C3, 2>(smem) : sa_ctas_per_sm<C1, C2, C3, 1>(smem);
  const bool wide = forced != 4 && per_sm == 1 && P.nsample <= 64 && C1 % 128 == 0 && C2 % 128 == 0;
  if (two) return wide ? launch_sa_g<C1, C2, C3, 2, 8>(P, stream) : launch_sa_g<C1, C2, C3, 2, 4>(P, stream);
  return wide ? launch_sa_g<C1, C2, C3, 1, 8>(P, stream) : launch_sa_g<C1, C2, C3, 1, 4>(P, stream);
}

}  // namespace

int sa_supported(int nsample, int npoint, int c, int c1, int c2, int c3) {
  if (nsample < 16 || nsample > 128 || (nsample & (nsample - 1))) return 0;
  if ((long long)npoint * nsample % kRows) return 0;
  if (npoint % (kRows / nsample)) return 0;
  if (c < 0 || c + 3 > 2048) return 0;
  return (c1 == 64 && c2 == 64 && c3 == 128) || (c1 == 128 && c2 == 128 && c3 == 256) ||
         (c1 == 128 && c2 == 128 && c3 == 128);
}

int pack_weight_dispatch(int c_out, int c_in, int kpad, int xyz_first, int fp16, const float *w,
                         void *packed, cudaStream_t stream) {
  const int total = c_out * kpad;
  pack_weight_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(c_out, c_in, kpad, xyz_first, fp16, w,
                                                               (uint16_t *)packed);
  count_launch();
  return check_launch("pack_weight_kernel");
}

int sa_forward_dispatch(int b, int n, int npoint, int nsample, int c, const float *xyz,
                        const float *new_xyz, const float *feat_pm, int feat_stride, const int *idx, float radius,
                        int normalize_xyz, int c1, int c2, int c3, const void *w1p, const float *b1,
                        const void *w2p, const float *b2, const void *w3p, const float *b3,
                        float *out_cm, float *out_pm, int fp16, cudaStream_t stream, int npoint_total,
                        int j_offset) {
  if (!sa_supported(nsample, npoint, c, c1, c2, c3))
    return set_error(BQA_ERR_UNSUPPORTED,
                     "sa_mlp_max: unsupported shape nsample=%d npoint=%d c=%d mlp=%d,%d,%d", nsample,
                     npoint, c, c1, c2, c3);
  SaParams P;
  P.b = b; P.n = n; P.npoint = npoint; P.nsample = nsample; P.c = c;
  P.npoint_total = npoint_total; P.j_offset = j_offset;
  P.k1pad = (c + 3 + 15) / 16 * 16;
  P.feat_stride = feat_stride;
  P.feat_vec4 = (c % 4 == 0) && (feat_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat_pm) & 15) == 0);
  P.xyz = xyz; P.new_xyz = new_xyz; P.feat_pm = feat_pm; P.idx = idx;
  P.radius = radius; P.normalize_xyz = normalize_xyz; P.fp16 = fp16;
  P.w1p = (const uint4 *)w1p; P.w2p = (const uint4 *)w2p; P.w3p = (const uint4 *)w3p;
  P.b1 = b1; P.b2 = b2; P.b3 = b3;
  P.out_cm = out_cm; P.out_pm = out_pm;
  P.num_tiles = (int)((long long)b * npoint * nsample / kRows);
  if (c1 == 64) return launch_sa<64, 64, 128>(P, stream);
  if (c3 == 256) return launch_sa<128, 128, 256>(P, stream);
  return launch_sa<128, 128, 128>(P, stream);
}

}  // namespace bqa
