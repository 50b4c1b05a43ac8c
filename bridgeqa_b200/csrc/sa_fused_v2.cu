// sa_fused_v2.cu -- warp-specialised fused set-abstraction layer for inference on sm_100a:
//   gather neighbours -> [1x1 conv + folded BN + ReLU] x3 on tcgen05 tensor cores -> max over nsample.
//
// Same contract as sa_fused.cu (it replaces QueryAndGroup.forward pointnet2_utils.py:317-376, SharedMLP
// pytorch_utils.py:11-36 and F.max_pool2d pointnet2_modules.py:259-262 of the reference in eval mode;
// the grouped (B, C+3, npoint, nsample) tensor and every activation stay on chip), re-architected for
// tensor-pipe utilisation.  What the phase trace of sa_fused.cu showed (profiles/r2_sa_trace.txt): per
// 128-row tile a pipeline spends 2.5-8k cycles in the feature gather (one L1 line request per thread and
// 16 bytes: the LSU's line-request rate is the bound), ~350 cycles of issue -> commit -> wait latency
// around every MMA burst and ~1k cycles per epilogue, all strictly one after the other, while the MMAs
// themselves need 0.6-2.2k cycles.  Here those phases belong to different warps of ONE persistent CTA
// and run concurrently on S tiles ("slots") in flight:
//
//   producer warps (4)   gather: features are read from a 16-bit point-major twin (the previous fused
//                        layer writes it; bits identical to what sa_fused.cu rounds to in registers), 8
//                        rows x 64 bytes per warp instruction with cp.async straight into the K-major
//                        operand buffer (8 line requests per instruction instead of 32, no registers, no
//                        conversion), completion tracked by the slot's mbarrier
//                        (cp.async.mbarrier.arrive.noinc); the 16-wide K tail [left-over features,
//                        (p - q)/r, 1, 1, 0...] is built in registers.  Index, coordinates and centre of
//                        the NEXT tile are prefetched while this tile's copies are issued.
//   MMA warp (1 thread)  polls the slots round robin and issues whichever layer of whichever slot is ready:
//                        layer 1 (K passes of 128), layer 2, layer 3 (transposed, weights as the A operand,
//                        so the max over nsample is register-local in the epilogue); tcgen05.commit
//                        releases operand buffers and signals the slot's epilogue warps.
//   epilogue warps (E per slot) TMEM -> registers -> relu -> 16-bit -> K-major operand of the next layer
//                        (one cvt.rn.relu per pair, no bias add: see below) / max over nsample -> outputs.
//
// Biases ride on the tensor core: the K tail of layer 1 carries two columns of ones against [b_hi, b_lo]
// (the fp32 bias split into two 16-bit halves, so it is exact to 2^-22 in fp16 mode), layer 2 has one
// extra K step with a constant ones operand; the epilogues of layers 1-2 are then LDTM -> cvt -> STS, which
// the micro-benchmark (profiles/r2_tmem_bench.txt) shows is what bounds them (FADD + LDS of a bias were
// 40 % of their issue slots).  Layer 3's bias is per output channel = per epilogue thread: one FADD after
// the max.
//
// Operands are 16-bit (fp16 default / bf16) with fp32 accumulation in TMEM, shared-memory layout as in
// tcgen05.cuh (no-swizzle K-major, [K/8][rows] 16-byte vectors).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstddef>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bqa {
namespace {

constexpr int kRows = 128;        // rows per tile (row = one (centre, neighbour) pair)
constexpr int kPassChunks = 16;   // 16-byte K chunks (8 channels each) of layer-1 features per pass
constexpr int kProducerWarps = 4;

struct Sa2Params {
  int b, n, npoint, nsample, c;
  int k1pad;            // round16(c + 5): features, rel xyz, two bias columns
  int n_main;           // (k1pad - 16) / 8 feature chunks copied with cp.async
  int kx;               // K index of rel x = max(c, k1pad - 16)
  int passes;           // ceil(n_main / 16), at least 1
  int alias;            // the feature-chunk buffer and the X1/X2 buffer of a slot share memory
  int stride16;         // 16-bit elements between consecutive rows of feat16 (multiple of 8)
  const float *xyz, *new_xyz;
  const uint16_t *feat16;
  const int *idx;
  float inv_radius;     // fl(1 / fl(radius))
  int normalize_xyz, fp16;
  const uint4 *w1p, *w2p, *w3p;     // packed images, see pack_weight_v2_kernel
  const float *b3;
  float *out_cm, *out_pm;
  uint16_t *out_pm16;
  int num_tiles;
};

__device__ __forceinline__ uint32_t pack2h(float lo, float hi, int fp16) {
  uint32_t d;
  if (fp16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t pack2h_relu(float lo, float hi, int fp16) {
  uint32_t d;
  if (fp16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint16_t to16(float v, int fp16) {
  if (fp16) {
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
    return *reinterpret_cast<const uint16_t *>(&h);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t *>(&h);
}

// mbarrier wait with a deadline: a broken hand-shake traps (the launch fails loudly) instead of hanging
// the GPU.  ~4 s at 2 GHz.
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("[bqa sa v2] mbarrier wait timed out: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
// non-blocking test (try_wait may suspend the thread for a system-dependent time: wrong for a poll loop)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the arrival is performed by the hardware once every cp.async this thread has issued so far has landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#ifdef BQA_SA_TRACE
// developer build (tools/build_stats.sh): time stamps of block 0's roles, printed by the host after the launch
constexpr int kTraceLen = 640;
__device__ long long g_v2_trace[3][kTraceLen][2];     // role (0 epilogue of slot 0, 1 producer, 2 MMA) x (tag, clock)
__device__ int g_v2_trace_n[3];
#define V2_TRACE(role, tag)                                                                      \
  if (blockIdx.x == 0 && ((role) == 2 ? threadIdx.x == (kEpiWarps + kProducerWarps) * 32 : (threadIdx.x & 31) == 0)) { \
    if (tr_n < kTraceLen) { g_v2_trace[role][tr_n][0] = (tag); g_v2_trace[role][tr_n][1] = clock64(); g_v2_trace_n[role] = ++tr_n; } \
  }
#else
#define V2_TRACE(role, tag)
#endif

// barriers of one slot
struct SlotBars {
  uint64_t a_full;      // producers -> MMA: operand chunk(s) of one layer-1 pass have landed
  uint64_t main_free;   // MMA (commit) -> producers: the feature-chunk buffer may be overwritten
  uint64_t tail_free;   // MMA (commit) -> producers: the K-tail buffer may be overwritten
  uint64_t d_full;      // MMA (commit) -> epilogue: an accumulator is complete (3 times per tile)
  uint64_t x_full;      // epilogue -> MMA: X1 / X2 is in shared memory (2 times per tile)
  uint64_t d_free;      // epilogue -> MMA: TMEM of the slot has been read (once per tile)
};
template <int C1, int C2, int C3>
struct Sa2Layout {
  static constexpr int kXVecs = (C1 > C2 ? C1 : C2) / 8 * kRows;
  static constexpr int kTailVecs = 2 * kRows;
  static constexpr int kTmemCols = (C1 + C2 > C3 ? C1 + C2 : C3);     // per slot: D1 | D2, D3T aliases
  static_assert(kTmemCols == 128 || kTmemCols == 256, "TMEM columns per slot");
  static __host__ __device__ int main_vecs(int n_main) { return (n_main < kPassChunks ? n_main : kPassChunks) * kRows; }
  static __host__ __device__ int slot_vecs(int n_main, int alias) {
    const int mv = main_vecs(n_main);
    return kTailVecs + (alias ? (mv > kXVecs ? mv : kXVecs) : mv + kXVecs);
  }
  static __host__ __device__ size_t weight_vecs(int k1pad) {
    return (size_t)(k1pad / 8) * C1 + (size_t)(C1 / 8 + 2) * C2 + (size_t)(C2 / 8) * C3;
  }
  static size_t smem_bytes(int k1pad, int n_main, int alias, int slots) {
    return 16 * (weight_vecs(k1pad) + 2 * kRows + (size_t)slots * slot_vecs(n_main, alias)) +
           sizeof(SlotBars) * slots + 64;
  }
};

// ============================================================================================
template <int C1, int C2, int C3, int S, int E>
__global__ void __launch_bounds__((S * E + kProducerWarps + S) * 32, 1)
sa_v2_kernel(const Sa2Params P) {
  using L = Sa2Layout<C1, C2, C3>;
  static_assert(E == 4 || E == 8, "epilogue warps per slot");
  static_assert(E == 4 ? (C1 == 64 && C2 == 64) : (C1 == 128 && C2 == 128), "a thread converts 64 columns");
  static_assert(C3 % 128 == 0, "layer 3 runs in 128-channel halves");
  constexpr int kEpiWarps = S * E;
  constexpr int kThreads = (kEpiWarps + kProducerWarps + S) * 32;
  constexpr uint32_t kTmemTotal = (uint32_t)(S * L::kTmemCols);
  static_assert(kTmemTotal <= 512 && (kTmemTotal & (kTmemTotal - 1)) == 0, "TMEM budget");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4 *w1s = reinterpret_cast<uint4 *>(smem_raw);                   // [k1pad/8][C1]
  uint4 *w2s = w1s + (size_t)(P.k1pad / 8) * C1;                      // [C1/8 + 2][C2]  (last K step: bias)
  uint4 *w3s = w2s + (size_t)(C1 / 8 + 2) * C2;                       // [C2/8][C3]
  uint4 *ones = w3s + (size_t)(C2 / 8) * C3;                          // [2][128]: A operand of layer 2's bias step
  uint4 *slots = ones + 2 * kRows;
  const int slot_vecs = L::slot_vecs(P.n_main, P.alias);
  SlotBars *bars = reinterpret_cast<SlotBars *>(slots + (size_t)S * slot_vecs);
  uint64_t *w_bar = reinterpret_cast<uint64_t *>(bars + S);
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(w_bar + 1);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef BQA_SA_TRACE
  int tr_n = 0;
#endif
  const int ns = P.nsample;
  const int centres_per_tile = kRows / ns;
  const int tiles_per_scene = P.npoint / centres_per_tile;

  auto slot_tail = [&](int s) { return slots + (size_t)s * slot_vecs; };
  auto slot_main = [&](int s) { return slots + (size_t)s * slot_vecs + L::kTailVecs; };
  auto slot_x = [&](int s) {
    return slots + (size_t)s * slot_vecs + L::kTailVecs + (P.alias ? 0 : L::main_vecs(P.n_main));
  };

  // ---- one-time setup ---------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bars[s].a_full), 2 * kProducerWarps * 32);   // st.shared arrival + cp.async arrival per thread
      mbar_init(smem_u32(&bars[s].main_free), 1);
      mbar_init(smem_u32(&bars[s].tail_free), 1);
      mbar_init(smem_u32(&bars[s].d_full), 1);
      mbar_init(smem_u32(&bars[s].x_full), E * 32);
      mbar_init(smem_u32(&bars[s].d_free), E * 32);
    }
    mbar_init(smem_u32(w_bar), 1);
    fence_mbar_init_cluster();
  }
  if (warp_id == kEpiWarps + kProducerWarps) umma::tmem_alloc(smem_u32(s_tmem), kTmemTotal);
  for (int i = threadIdx.x; i < 2 * kRows; i += kThreads) {
    // chunk 0 of every row: [1, 1, 0, 0, 0, 0, 0, 0]; chunk 1: zeros
    const uint32_t one2 = P.fp16 ? 0x3c003c00u : 0x3f803f80u;
    ones[i] = i < kRows ? make_uint4(one2, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *s_tmem;

  const int my_tiles = (P.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles blockIdx.x + k*gridDim.x

  if (warp_id < kEpiWarps) {
    // ======================= EPILOGUE WARPS =======================================================
    const int s = warp_id / E;                 // slot
    const int e = warp_id % E;
    const int quarter = e & 3;                 // == warp_id % 4: the TMEM lane quarter this warp may touch
    const int half = e >> 2;                   // E == 8: second thread of the row / second half of the columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tmem = tmem_base + (uint32_t)(s * L::kTmemCols);
    const uint32_t d_full = smem_u32(&bars[s].d_full), x_full = smem_u32(&bars[s].x_full),
                   d_free = smem_u32(&bars[s].d_free);
    uint4 *xb = slot_x(s);
    uint32_t ph_d = 0;
    float bias3[C3 / 128];
#pragma unroll
    for (int mt = 0; mt < C3 / 128; ++mt) bias3[mt] = __ldg(P.b3 + mt * 128 + row);

    // relu(D) -> 16-bit -> X[col/8][row]; the thread converts 64 columns [c0, c0 + 64) of its row, 32 at a
    // time (one tcgen05.ld in flight per warp: the measured per-warp rate is the same as with two, at
    // half the registers; the other warps of the SM sub-partition cover the latency)
    auto convert64 = [&](uint32_t taddr, int c0) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t va[32];
        umma::ld_32x32b_x32(taddr + lane_addr + (uint32_t)(c0 + 32 * h), va);
        umma::wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          xb[(c0 / 8 + 4 * h + q) * kRows + row] =
              make_uint4(pack2h_relu(__uint_as_float(va[q * 8 + 0]), __uint_as_float(va[q * 8 + 1]), P.fp16),
                         pack2h_relu(__uint_as_float(va[q * 8 + 2]), __uint_as_float(va[q * 8 + 3]), P.fp16),
                         pack2h_relu(__uint_as_float(va[q * 8 + 4]), __uint_as_float(va[q * 8 + 5]), P.fp16),
                         pack2h_relu(__uint_as_float(va[q * 8 + 6]), __uint_as_float(va[q * 8 + 7]), P.fp16));
        }
      }
    };

    for (int k = s; k < my_tiles; k += S) {
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      const int scene = tile / tiles_per_scene;
      const int centre0 = (tile % tiles_per_scene) * centres_per_tile;
      // ---- layer 1 accumulator -> X1
      if (warp_id == 0) { V2_TRACE(0, 100 + k * 1000) }
      wait_bar(d_full, ph_d, 10); ph_d ^= 1;
      umma::fence_after_sync();
      if (warp_id == 0) { V2_TRACE(0, 101 + k * 1000) }
      convert64(tmem, half * 64);
      if (warp_id == 0) { V2_TRACE(0, 102 + k * 1000) }
      umma::fence_proxy_async_smem();
      umma::fence_before_sync();
      mbar_arrive(x_full);
      if (warp_id == 0) { V2_TRACE(0, 103 + k * 1000) }
      // ---- layer 2 accumulator -> X2
      wait_bar(d_full, ph_d, 11); ph_d ^= 1;
      umma::fence_after_sync();
      if (warp_id == 0) { V2_TRACE(0, 111 + k * 1000) }
      convert64(tmem + C1, half * 64);
      umma::fence_proxy_async_smem();
      umma::fence_before_sync();
      mbar_arrive(x_full);
      if (warp_id == 0) { V2_TRACE(0, 113 + k * 1000) }
      // ---- layer 3 (channels on lanes, rows on columns): max over nsample, bias, ReLU, store
      wait_bar(d_full, ph_d, 12); ph_d ^= 1;
      umma::fence_after_sync();
      if (warp_id == 0) { V2_TRACE(0, 121 + k * 1000) }
      constexpr int kColsPerThread = E == 8 ? 64 : 128;
      const int cnt = kColsPerThread / ns > 0 ? kColsPerThread / ns : 1;     // centres this thread completes per channel
      const int j_first = centre0 + (half * 64) / ns;
#pragma unroll
      for (int mt = 0; mt < C3 / 128; ++mt) {
        const int ch = mt * 128 + row;
        float run = -INFINITY;
        float o[8];                       // the last `cnt` entries are this thread's outputs, in centre order
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.f;
        int j = j_first;
#pragma unroll 1
        for (int c0 = half * 64; c0 < half * 64 + kColsPerThread; c0 += 32) {
          uint32_t va[32];
          umma::ld_32x32b_x32(tmem + (uint32_t)(mt * 128) + lane_addr + (uint32_t)c0, va);
          umma::wait_ld();
#pragma unroll
          for (int g = 0; g < 32; g += 16) {             // nsample is a multiple of 16
            float m16 = __uint_as_float(va[g]);
#pragma unroll
            for (int t = 1; t < 16; ++t) m16 = fmaxf(m16, __uint_as_float(va[g + t]));
            run = fmaxf(run, m16);
            if (((c0 + g + 16) & (ns - 1)) == 0) {        // columns [.., c0 + g + 16) folded: a centre is complete
              const float v = fmaxf(run + bias3[mt], 0.f);
              const size_t pm = ((size_t)scene * P.npoint + j) * C3 + ch;
              if (P.out_pm) P.out_pm[pm] = v;              // coalesced: consecutive lanes = consecutive channels
              if (P.out_pm16) P.out_pm16[pm] = to16(v, P.fp16);
#pragma unroll
              for (int i = 0; i < 7; ++i) o[i] = o[i + 1];
              o[7] = v;
              ++j;
              run = -INFINITY;
            }
          }
        }
        // channel-major output: this thread's `cnt` consecutive centres of channel ch as ONE vector store
        // (the row (scene, ch) is npoint floats long and j_first is a multiple of cnt: aligned)
        float *dst = P.out_cm + ((size_t)scene * C3 + ch) * P.npoint + j_first;
        if (cnt == 8) {
          reinterpret_cast<float4 *>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
          reinterpret_cast<float4 *>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else if (cnt == 4) {
          reinterpret_cast<float4 *>(dst)[0] = make_float4(o[4], o[5], o[6], o[7]);
        } else if (cnt == 2) {
          reinterpret_cast<float2 *>(dst)[0] = make_float2(o[6], o[7]);
        } else {
          dst[0] = o[7];
        }
      }
      umma::fence_before_sync();
      mbar_arrive(d_free);
      if (warp_id == 0) { V2_TRACE(0, 123 + k * 1000) }
    }
  } else if (warp_id < kEpiWarps + kProducerWarps) {
    // ======================= PRODUCER WARPS ==========================================================
    // The four warps work on the same tile, a row per lane (warp pw: rows [32 pw, 32 pw + 32)), and run a
    // software pipeline over the CTA's tiles: while tile k is written to shared memory, the row data of tile
    // k + 1 (coordinates / left-over features: loads that depend on the neighbour index) and the indices of
    // tile k + 2 are in flight.  The loop is unrolled by two with two named register sets, so no register
    // move ever has to wait for a load.
    const int pw = warp_id - kEpiWarps;
    const int row = pw * 32 + lane;
    const int n_main = P.n_main;
    const int tail_p0 = P.kx - 8 * n_main;      // position of rel x inside the 16-wide tail
    // left-over feature chunks that live in the tail (source row chunks n_main, n_main + 1)
    const int lo_chunks = P.c > 8 * n_main ? (P.c - 8 * n_main + 7) / 8 : 0;
    const uint32_t one16 = P.fp16 ? 0x3c00u : 0x3f80u;
    const int sub = row / ns, smp = row % ns;   // centre of the tile / sample of the centre this row belongs to

    struct Row { int i; float px, py, pz, cv; uint4 f0; };
    auto load_index = [&](int k) -> int {
      if (k >= my_tiles) return 0;
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      const int scene = tile / tiles_per_scene;
      const int j = (tile % tiles_per_scene) * centres_per_tile + sub;
      return __ldg(P.idx + ((size_t)scene * P.npoint + j) * ns + smp);
    };
    auto load_row = [&](int k, int i) -> Row {
      Row r;
      r.i = i;
      r.px = r.py = r.pz = r.cv = 0.f;
      r.f0 = make_uint4(0u, 0u, 0u, 0u);
      if (k >= my_tiles) return r;
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      const int scene = tile / tiles_per_scene;
      const int centre0 = (tile % tiles_per_scene) * centres_per_tile;
      const float *p = P.xyz + ((size_t)scene * P.n + i) * 3;
      r.px = __ldg(p); r.py = __ldg(p + 1); r.pz = __ldg(p + 2);
      // the tile's centres (128 / nsample <= 8 of them, 3 floats each): one coalesced load per warp,
      // shuffled to the rows when the tail is built
      if (lane < 3 * centres_per_tile) r.cv = __ldg(P.new_xyz + ((size_t)scene * P.npoint + centre0) * 3 + lane);
      if (lo_chunks > 0)
        r.f0 = __ldg(reinterpret_cast<const uint4 *>(P.feat16 + ((size_t)scene * P.n + i) * P.stride16) + n_main);
      return r;
    };

    // writes tile k (row data `cur`) into its slot and signals the MMA thread
    auto produce = [&](int k, const Row &cur) {
      const int s = k % S;
      const int u = k / S;                       // u-th use of the slot
      const int tile = (int)blockIdx.x + k * (int)gridDim.x;
      const int scene = tile / tiles_per_scene;
      const uint32_t a_full = smem_u32(&bars[s].a_full);
      if (pw == 0) { V2_TRACE(1, 200 + k * 1000) }
      for (int pass = 0; pass < P.passes; ++pass) {
        if (n_main > 0) {
          // ---- feature chunks [16 pass, ...): 8 rows x 4 chunks (64 contiguous bytes of each row) per
          // warp instruction; a quarter warp writes 128 contiguous bytes of one chunk column.
          // The buffer has been released u * passes + pass times when it may be written (a fresh
          // barrier passes a wait on parity 1).
          wait_bar(smem_u32(&bars[s].main_free), (uint32_t)(((u * P.passes + pass) & 1) ^ 1), 21);
          if (pw == 0) { V2_TRACE(1, 203 + k * 1000 + 10 * pass) }
          const int ch0 = pass * kPassChunks;
          const int ch1 = min(n_main, ch0 + kPassChunks);
          const uint32_t dst_base = smem_u32(slot_main(s)) + (uint32_t)(pw * 32) * 16u;
          const uint16_t *fbase = P.feat16 + (size_t)scene * P.n * P.stride16 + (size_t)(ch0 + (lane >> 3)) * 8;
          const uint32_t dcol = (uint32_t)(lane >> 3) * (kRows * 16u);
#pragma unroll
          for (int r8 = 0; r8 < 4; ++r8) {
            const int lrow = r8 * 8 + (lane & 7);                        // row of this warp
            const int src_i = __shfl_sync(0xffffffffu, cur.i, lrow);
            const uint16_t *srow = fbase + (size_t)src_i * P.stride16;
            const uint32_t drow = dst_base + (uint32_t)lrow * 16u + dcol;
            for (int cb = ch0 + (lane >> 3); cb < ch1; cb += 4)
              cp_async_16(drow + (uint32_t)(cb - ch0 - (lane >> 3)) * (kRows * 16u), srow + (cb - ch0 - (lane >> 3)) * 8);
          }
        }
        if (pass == 0) {
          // ---- K tail: [left-over features | rel xyz | 1 1 | 0 ...]  (16 values, two 16-byte chunks)
          wait_bar(smem_u32(&bars[s].tail_free), (uint32_t)((u & 1) ^ 1), 20);
          if (pw == 0) { V2_TRACE(1, 201 + k * 1000) }
          // (p - q) [* 1/r]: torch evaluates `grouped_xyz / radius` (pointnet2_utils.py:350-352) with a Python
          // scalar divisor on CUDA as a multiplication by fl(1 / fl(radius)) (ATen BinaryDivTrueKernel.cu)
          float rel0 = cur.px - __shfl_sync(0xffffffffu, cur.cv, 3 * sub);
          float rel1 = cur.py - __shfl_sync(0xffffffffu, cur.cv, 3 * sub + 1);
          float rel2 = cur.pz - __shfl_sync(0xffffffffu, cur.cv, 3 * sub + 2);
          if (P.normalize_xyz) { rel0 *= P.inv_radius; rel1 *= P.inv_radius; rel2 *= P.inv_radius; }
          const uint32_t r01 = pack2h(rel0, rel1, P.fp16);                                 // rel x | rel y << 16
          const uint32_t r21 = (pack2h(rel2, 0.f, P.fp16) & 0xffffu) | (one16 << 16);      // rel z | 1 << 16
          uint4 t0, t1;
          if (tail_p0 == 0) {                 // no left-over features
            t0 = make_uint4(r01, r21, one16, 0u);
            t1 = make_uint4(0u, 0u, 0u, 0u);
          } else if (tail_p0 == 4) {          // e.g. c = 132
            t0 = make_uint4(cur.f0.x, cur.f0.y, r01, r21);
            t1 = make_uint4(one16, 0u, 0u, 0u);
          } else if (tail_p0 == 7) {          // e.g. c = 7
            t0 = make_uint4(cur.f0.x, cur.f0.y, cur.f0.z, (cur.f0.w & 0xffffu) | (r01 << 16));
            t1 = make_uint4((r01 >> 16) | (r21 << 16), one16 | (one16 << 16), 0u, 0u);
          } else {                            // general: 16 halves, insert [rx ry rz 1 1] at tail_p0
            uint4 f1 = make_uint4(0u, 0u, 0u, 0u);
            if (lo_chunks > 1)
              f1 = __ldg(reinterpret_cast<const uint4 *>(P.feat16 + ((size_t)scene * P.n + cur.i) * P.stride16) + n_main + 1);
            const uint32_t fw[8] = {cur.f0.x, cur.f0.y, cur.f0.z, cur.f0.w, f1.x, f1.y, f1.z, f1.w};
            uint32_t hv[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const uint32_t fh = (t & 1) ? (fw[t >> 1] >> 16) : (fw[t >> 1] & 0xffffu);
              const int d = t - tail_p0;
              uint32_t v = fh;
              if (d >= 0) v = d == 0 ? (r01 & 0xffffu) : d == 1 ? (r01 >> 16) : d == 2 ? (r21 & 0xffffu) : d <= 4 ? one16 : 0u;
              hv[t] = v;
            }
            t0 = make_uint4(hv[0] | (hv[1] << 16), hv[2] | (hv[3] << 16), hv[4] | (hv[5] << 16), hv[6] | (hv[7] << 16));
            t1 = make_uint4(hv[8] | (hv[9] << 16), hv[10] | (hv[11] << 16), hv[12] | (hv[13] << 16), hv[14] | (hv[15] << 16));
          }
          uint4 *tail = slot_tail(s);
          tail[row] = t0;
          tail[kRows + row] = t1;
        }
        umma::fence_proxy_async_smem();          // the tail's st.shared -> async proxy
        mbar_arrive(a_full);
        cp_async_arrive_noinc(a_full);           // second arrival: when this thread's copies have landed
        if (pw == 0) { V2_TRACE(1, 204 + k * 1000 + 10 * pass) }
      }
    };

    int idx_a = load_index(0);
    Row row_a = load_row(0, idx_a);
    int idx_b = load_index(1);
    for (int k = 0; k < my_tiles; k += 2) {
      // set A holds tile k; idx_b is tile k + 1's index
      const Row row_b = load_row(k + 1, idx_b);
      idx_a = load_index(k + 2);
      produce(k, row_a);
      if (k + 1 >= my_tiles) break;
      row_a = load_row(k + 2, idx_a);
      idx_b = load_index(k + 3);
      produce(k + 1, row_b);
    }
  } else if (lane == 0 && warp_id - (kEpiWarps + kProducerWarps) < S) {
    // ======================= MMA ISSUERS (one thread per slot) ==========================================
    // Each slot has its own issuing thread (any thread may issue tcgen05.mma; the slots' accumulators and
    // operand buffers are disjoint, and a commit tracks the issuing thread's own MMAs), so the control flow
    // of a slot is a plain sequence of blocking waits -- no polling loop that serialises four slots behind
    // one thread's instruction latency (measured: ~800 cycles per step, 2.2k per tile at SA1) and no spinning
    // warp stealing issue slots from the epilogue warps of its SM sub-partition.
    const int s = warp_id - (kEpiWarps + kProducerWarps);
    const uint32_t wbar = smem_u32(w_bar);
    if (s == 0) {
      const uint32_t b1 = (uint32_t)((size_t)(P.k1pad / 8) * C1 * 16), b2 = (uint32_t)((size_t)(C1 / 8 + 2) * C2 * 16),
                     b3 = (uint32_t)((size_t)(C2 / 8) * C3 * 16);
      mbar_arrive_expect_tx(wbar, b1 + b2 + b3);
      bulk_g2s(smem_u32(w1s), P.w1p, b1, wbar);
      bulk_g2s(smem_u32(w2s), P.w2p, b2, wbar);
      bulk_g2s(smem_u32(w3s), P.w3p, b3, wbar);
    }
    const uint32_t idesc1 = umma::instr_desc_16b_f32(128, C1, !P.fp16);
    const uint32_t idesc2 = umma::instr_desc_16b_f32(128, C2, !P.fp16);
    const uint32_t idesc3 = umma::instr_desc_16b_f32(128, 128, !P.fp16);
    const uint32_t tmem = tmem_base + (uint32_t)(s * L::kTmemCols);
    // descriptors of the first K step of every operand; one K step (two 16-byte chunk columns) further is a
    // constant added to the address field (bits [0,14) hold address >> 4)
    const uint64_t d_main = umma::smem_desc(smem_u32(slot_main(s)), kRows * 16, 128);
    const uint64_t d_tail = umma::smem_desc(smem_u32(slot_tail(s)), kRows * 16, 128);
    const uint64_t d_x = umma::smem_desc(smem_u32(slot_x(s)), kRows * 16, 128);
    const uint64_t d_ones = umma::smem_desc(smem_u32(ones), kRows * 16, 128);
    const uint64_t d_w1 = umma::smem_desc(smem_u32(w1s), C1 * 16, 128);
    const uint64_t d_w2 = umma::smem_desc(smem_u32(w2s), C2 * 16, 128);
    const uint64_t d_w3 = umma::smem_desc(smem_u32(w3s), C3 * 16, 128);
    constexpr uint64_t kStepRows = (2 * kRows * 16) >> 4;       // operand with 128 rows per chunk column
    constexpr uint64_t kStepW1 = (2 * C1 * 16) >> 4, kStepW2 = (2 * C2 * 16) >> 4, kStepW3 = (2 * C3 * 16) >> 4;
    const uint32_t a_full = smem_u32(&bars[s].a_full), x_full = smem_u32(&bars[s].x_full),
                   d_full = smem_u32(&bars[s].d_full), d_free = smem_u32(&bars[s].d_free);
    uint32_t ph_a = 0, ph_x = 0, ph_f = 0;
    wait_bar(wbar, 0, 30);                          // weights resident
    for (int k = s; k < my_tiles; k += S) {
      const uint32_t main_free = smem_u32(&bars[s].main_free), tail_free = smem_u32(&bars[s].tail_free);
      // ---- layer 1: D1[128 x C1] = A[128 x K1] * W1^T, K in passes of 128 (+ the 16-wide tail)
      if (k >= S) { wait_bar(d_free, ph_f, 31); ph_f ^= 1; }       // the previous tile's D3T (aliases D1) has been read
      for (int p = 0; p < P.passes; ++p) {
        wait_bar(a_full, ph_a, 32); ph_a ^= 1;
        umma::fence_after_sync();
        V2_TRACE(2, 300 + 10 * s + p)
        const int cnt = min(P.n_main - p * kPassChunks, kPassChunks);      // feature chunks of this pass (even; may be <= 0)
        uint64_t ad = d_main, bd = d_w1 + (uint64_t)(p * (kPassChunks / 2)) * kStepW1;
        for (int ks = 0; ks < cnt / 2; ++ks) {
          umma::mma_bf16_ss(tmem, ad, bd, idesc1, (p | ks) != 0);
          ad += kStepRows; bd += kStepW1;
        }
        if (p == P.passes - 1) {
          umma::mma_bf16_ss(tmem, d_tail, d_w1 + (uint64_t)(P.n_main / 2) * kStepW1, idesc1, P.n_main > 0);
          umma::commit(tail_free);
          if (P.n_main > 0 && !P.alias) umma::commit(main_free);
          umma::commit(d_full);
        } else {
          umma::commit(main_free);               // next K pass of the same tile
        }
      }
      // ---- layer 2: D2 = X1 * W2^T + ones * [b_hi b_lo 0..]^T
      wait_bar(x_full, ph_x, 33); ph_x ^= 1;
      umma::fence_after_sync();
      V2_TRACE(2, 400 + 10 * s)
      {
        uint64_t ad = d_x, bd = d_w2;
#pragma unroll
        for (int ks = 0; ks < C1 / 16; ++ks) {
          umma::mma_bf16_ss(tmem + C1, ad, bd, idesc2, ks != 0);
          ad += kStepRows; bd += kStepW2;
        }
        umma::mma_bf16_ss(tmem + C1, d_ones, bd, idesc2, 1u);
      }
      umma::commit(d_full);
      // ---- layer 3, transposed: D3T[C3 x 128] = W3 * X2^T in 128-channel halves
      wait_bar(x_full, ph_x, 34); ph_x ^= 1;
      umma::fence_after_sync();
      V2_TRACE(2, 500 + 10 * s)
#pragma unroll
      for (int mt = 0; mt < C3 / 128; ++mt) {
        uint64_t ad = d_w3 + (uint64_t)((mt * 128 * 16) >> 4), bd = d_x;
#pragma unroll
        for (int ks = 0; ks < C2 / 16; ++ks) {
          umma::mma_bf16_ss(tmem + mt * 128, ad, bd, idesc3, ks != 0);
          ad += kStepW3; bd += kStepRows;
        }
      }
      if (P.n_main > 0 && P.alias) umma::commit(main_free);
      umma::commit(d_full);
      V2_TRACE(2, 501 + 10 * s)
    }
  }

  // ---- teardown: every MMA has been consumed by an epilogue before its warp leaves the loop -----
  umma::fence_before_sync();
  __syncthreads();
  if (warp_id == kEpiWarps + kProducerWarps) umma::tmem_dealloc(tmem_base, kTmemTotal);
}

// ---- weight images ----------------------------------------------------------------------------
// w (c_out, c_in) f32 (+ bias (c_out) f32) -> 16-bit image [kpad/8][c_out][8].
//   mode 0: K = c_in, plain (layer 3)
//   mode 1: layer 1 of an SA block.  Source columns are [xyz(3), feat(c_in - 3)] (use_xyz cat order,
//           pointnet2_utils.py:357-359); packed K order: feature f at K = f, rel xyz at kx..kx+2,
//           bias hi / lo at kx+3, kx+4 (kx = max(c, kpad - 16)), zero elsewhere
//   mode 2: layer 2: K = c_in channels, then bias hi / lo at K = c_in, c_in + 1 (kpad = c_in + 16)
__global__ void pack_weight_v2_kernel(int c_out, int c_in, int kpad, int mode, int fp16,
                                      const float *__restrict__ w, const float *__restrict__ bias,
                                      uint16_t *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= c_out * kpad) return;
  const int k = t % kpad, row = t / kpad;
  float v = 0.f;
  int bias_part = 0;      // 1: hi, 2: lo
  if (mode == 0) {
    if (k < c_in) v = w[(size_t)row * c_in + k];
  } else if (mode == 1) {
    const int c = c_in - 3;
    const int kx = max(c, kpad - 16);
    if (k < c) v = w[(size_t)row * c_in + 3 + k];
    else if (k >= kx && k < kx + 3) v = w[(size_t)row * c_in + (k - kx)];
    else if (k == kx + 3) bias_part = 1;
    else if (k == kx + 4) bias_part = 2;
  } else {
    if (k < c_in) v = w[(size_t)row * c_in + k];
    else if (k == c_in) bias_part = 1;
    else if (k == c_in + 1) bias_part = 2;
  }
  uint16_t bits;
  if (bias_part) {
    const float bv = bias ? bias[row] : 0.f;
    const uint16_t hi = to16(bv, fp16);
    float hif;
    if (fp16) hif = __half2float(*reinterpret_cast<const __half *>(&hi));
    else hif = __bfloat162float(*reinterpret_cast<const __nv_bfloat16 *>(&hi));
    bits = bias_part == 1 ? hi : to16(bv - hif, fp16);
  } else {
    bits = to16(v, fp16);
  }
  out[((size_t)(k / 8) * c_out + row) * 8 + (k % 8)] = bits;
}

// (B, C, N) f32 channel-major -> (B, N, stride) 16-bit point-major, zero padded to `stride`
__global__ void to_point_major_16_kernel(int c, int n, int stride, int fp16, const float *__restrict__ in,
                                         uint16_t *__restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float *src = in + (size_t)b * c * n;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int cc = c0 + r, nn = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (cc < c && nn < n) ? src[(size_t)cc * n + nn] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int nn = n0 + r, cc = c0 + threadIdx.x;
    if (nn < n && cc < stride) out[((size_t)b * n + nn) * stride + cc] = to16(tile[threadIdx.x][r], fp16);
  }
}

// (B, N, row_stride) f32 rows [first, first + c) -> (B, N, stride) 16-bit, zero padded (the input cloud's
// feature columns: SA1 of the backbone)
__global__ void rows_to_16_kernel(long long rows, int c, int row_stride, int first, int stride, int fp16,
                                  const float *__restrict__ in, uint16_t *__restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = stride / 8;
  if (t >= rows * per_row) return;
  const long long r = t / per_row;
  const int k0 = (int)(t % per_row) * 8;
  const float *src = in + (size_t)r * row_stride + first;
  uint32_t p[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float lo = k0 + 2 * e < c ? __ldg(src + k0 + 2 * e) : 0.f;
    const float hi = k0 + 2 * e + 1 < c ? __ldg(src + k0 + 2 * e + 1) : 0.f;
    p[e] = pack2h(lo, hi, fp16);
  }
  reinterpret_cast<uint4 *>(out)[t] = make_uint4(p[0], p[1], p[2], p[3]);
}

template <int C1, int C2, int C3, int S, int E>
int launch_v2(Sa2Params P, cudaStream_t stream) {
  using L = Sa2Layout<C1, C2, C3>;
  const size_t smem = L::smem_bytes(P.k1pad, P.n_main, P.alias, S);
  auto kern = sa_v2_kernel<C1, C2, C3, S, E>;
  BQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  BQA_CUDA(cudaGetDevice(&dev));
  BQA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = min(P.num_tiles, sms);
#ifdef BQA_SA_TRACE
  int zeros[3] = {0, 0, 0};
  cudaMemcpyToSymbol(g_v2_trace_n, zeros, sizeof(zeros));
#endif
  kern<<<grid, (S * E + kProducerWarps + S) * 32, smem, stream>>>(P);
#ifdef BQA_SA_TRACE
  {
    static long long tr[3][kTraceLen][2];
    int cnt[3];
    cudaMemcpyFromSymbol(cnt, g_v2_trace_n, sizeof(cnt));
    cudaMemcpyFromSymbol(tr, g_v2_trace, sizeof(tr));
    long long t0 = 0x7fffffffffffffffll;
    for (int r = 0; r < 3; ++r) for (int i = 0; i < cnt[r]; ++i) t0 = tr[r][i][1] < t0 ? tr[r][i][1] : t0;
    fprintf(stderr, "[bqa sa v2 trace] <%d,%d,%d,S=%d,E=%d> c=%d ns=%d grid=%d tiles=%d alias=%d passes=%d smem=%zu\n", C1, C2, C3, S, E,
            P.c, P.nsample, grid, P.num_tiles, P.alias, P.passes, smem);
    const char *names[3] = {"epi", "prod", "mma"};
    for (int r = 0; r < 3; ++r) {
      fprintf(stderr, "  %s:", names[r]);
      for (int i = 0; i < cnt[r] && i < 96; ++i) fprintf(stderr, " %lld@%lld", tr[r][i][0], tr[r][i][1] - t0);
      fprintf(stderr, "\n");
    }
  }
#endif
  count_launch();
  return check_launch("sa_v2_kernel");
}

// slots: as many tiles in flight as TMEM (512 columns) and shared memory (227 KB) allow; the operand
// buffers alias only when that buys a slot
template <int C1, int C2, int C3, int E>
int launch_v2_pick(Sa2Params P, cudaStream_t stream) {
  using L = Sa2Layout<C1, C2, C3>;
  constexpr int kMaxSlots = 512 / L::kTmemCols;       // 4 or 2
  static const int forced = [] { const char *e = getenv("BQA_SA_SLOTS"); return e ? atoi(e) : 0; }();
  const size_t cap = 227 * 1024;
  for (int slots = kMaxSlots; slots >= 1; slots >>= 1) {
    if (forced && slots > forced) continue;
    for (int alias = 0; alias <= 1; ++alias) {
      if (alias && P.n_main == 0) continue;
      if (L::smem_bytes(P.k1pad, P.n_main, alias, slots) > cap) continue;
      P.alias = alias;
      if constexpr (kMaxSlots == 4) {
        if (slots == 4) return launch_v2<C1, C2, C3, 4, E>(P, stream);
      }
      if (slots == 2) return launch_v2<C1, C2, C3, 2, E>(P, stream);
      return launch_v2<C1, C2, C3, 1, E>(P, stream);
    }
  }
  return set_error(BQA_ERR_UNSUPPORTED, "sa_v2: c=%d does not fit shared memory", P.c);
}

}  // namespace

int sa_v2_supported(int nsample, int npoint, int c, int c1, int c2, int c3) {
  if (nsample < 16 || nsample > 128 || (nsample & (nsample - 1))) return 0;
  if (npoint % (kRows / nsample)) return 0;
  if (c < 0 || c > 1024) return 0;
  if (c1 == 64 && c2 == 64 && c3 == 128) return 1;                      // E = 4: a thread owns all 128 columns
  if (nsample > 64) return 0;                                           // E = 8: a group must fit a 64-column half
  return (c1 == 128 && c2 == 128 && c3 == 256) || (c1 == 128 && c2 == 128 && c3 == 128);
}

int pack_weight_v2_dispatch(int c_out, int c_in, int kpad, int mode, int fp16, const float *w, const float *bias,
                            void *packed, cudaStream_t stream) {
  const int total = c_out * kpad;
  pack_weight_v2_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(c_out, c_in, kpad, mode, fp16, w, bias,
                                                                  (uint16_t *)packed);
  count_launch();
  return check_launch("pack_weight_v2_kernel");
}

int to_point_major_16_dispatch(int b, int c, int n, int stride, int fp16, const float *in, void *out,
                               cudaStream_t stream) {
  dim3 grid((unsigned)ceil_div(n, 32), (unsigned)ceil_div(stride, 32), (unsigned)b);
  to_point_major_16_kernel<<<grid, dim3(32, 8), 0, stream>>>(c, n, stride, fp16, in, (uint16_t *)out);
  count_launch();
  return check_launch("to_point_major_16_kernel");
}

int rows_to_16_dispatch(long long rows, int c, int row_stride, int first, int stride, int fp16, const float *in,
                        void *out, cudaStream_t stream) {
  const long long total = rows * (stride / 8);
  rows_to_16_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, stream>>>(rows, c, row_stride, first, stride, fp16,
                                                                          in, (uint16_t *)out);
  count_launch();
  return check_launch("rows_to_16_kernel");
}

int sa_v2_forward_dispatch(int b, int n, int npoint, int nsample, int c, const float *xyz, const float *new_xyz,
                           const void *feat16, int stride16, const int *idx, float radius, int normalize_xyz,
                           int c1, int c2, int c3, const void *w1p, const void *w2p, const void *w3p,
                           const float *b3, float *out_cm, float *out_pm, void *out_pm16, int fp16,
                           cudaStream_t stream) {
  if (!sa_v2_supported(nsample, npoint, c, c1, c2, c3))
    return set_error(BQA_ERR_UNSUPPORTED, "sa_v2: unsupported shape nsample=%d npoint=%d c=%d mlp=%d,%d,%d",
                     nsample, npoint, c, c1, c2, c3);
  Sa2Params P;
  P.b = b; P.n = n; P.npoint = npoint; P.nsample = nsample; P.c = c;
  P.k1pad = (c + 5 + 15) / 16 * 16;
  P.n_main = (P.k1pad - 16) / 8;
  P.kx = c > P.k1pad - 16 ? c : P.k1pad - 16;
  P.passes = P.n_main > 0 ? (P.n_main + kPassChunks - 1) / kPassChunks : 1;
  P.alias = 0;
  P.stride16 = stride16;
  P.xyz = xyz; P.new_xyz = new_xyz; P.feat16 = (const uint16_t *)feat16; P.idx = idx;
  P.inv_radius = 1.0f / radius; P.normalize_xyz = normalize_xyz; P.fp16 = fp16;
  P.w1p = (const uint4 *)w1p; P.w2p = (const uint4 *)w2p; P.w3p = (const uint4 *)w3p; P.b3 = b3;
  P.out_cm = out_cm; P.out_pm = out_pm; P.out_pm16 = (uint16_t *)out_pm16;
  P.num_tiles = (int)((long long)b * npoint * nsample / kRows);
  if (c1 == 64) return launch_v2_pick<64, 64, 128, 4>(P, stream);
  if (c3 == 256) return launch_v2_pick<128, 128, 256, 8>(P, stream);
  return launch_v2_pick<128, 128, 128, 8>(P, stream);
}

}  // namespace bqa
