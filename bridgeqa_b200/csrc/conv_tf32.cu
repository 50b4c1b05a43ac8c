// conv_tf32.cu -- the 1x1 convolutions of the SharedMLP blocks in TRAINING mode on the 5th-gen
// tensor cores (sm_100a): forward (+ the BatchNorm batch statistics in the epilogue), data gradient
// and weight gradient.  Replaces the cuDNN calls behind nn.Conv2d / nn.Conv1d of
// /root/reference/lib/pointnet2/pytorch_utils.py:104-157 (Conv2d -> _ConvBase, kernel 1x1, no
// bias when followed by BN) and their autograd backward.
//
// Arithmetic: TF32 operands (fp32 tensors read through a TFLOAT32 tensor map, i.e. rounded to
// TF32 by the TMA unit), fp32 accumulation in TMEM -- the precision class of the reference's convs
// under torch's default cudnn.allow_tf32 = True.
//
// Layouts are the reference's: x (B, Cin, P) and y (B, Cout, P) with the P = npoint * nsample (or
// n) positions contiguous, W (Cout, Cin).  Nothing is transposed in memory:
//   forward   D[Cout x 128 pos] = W[Cout x Cin] . X[Cin x 128 pos]
//             A = W chunk, K-major (Cin contiguous);  B = X chunk, MN-major (positions contiguous):
//             tcgen05 takes both through shared-memory descriptors, so the channel-major activation
//             tensor is the B operand as it lies in HBM.  TMEM lanes are output channels, TMEM
//             columns are positions: an epilogue thread owns one channel, its 32-column TMEM load is
//             128 contiguous bytes of y, and the per-channel sum / sum of squares BatchNorm needs are
//             register-local sums (no shuffles, no second pass over y).
//   dgrad     the same kernel with W^T (Cin x Cout) as the weight and dy as the input
//   wgrad     dW[Cout x Cin] = sum over positions of dy[Cout x pos] . x[Cin x pos]^T:  both operands
//             K-major (K = positions), the CTA owns a range of positions and a 128/256 x <=256 block of
//             dW in TMEM, partial results are added with red.global.add.f32 (split-K over the grid).
// Data movement: cp.async.bulk.tensor (TMA, SASS UTMALDG) with 128-byte swizzle into a 4-stage ring,
// one producer thread, one MMA-issuing thread, four epilogue warps, two TMEM accumulator stages so the
// epilogue of tile t runs under the loads and MMAs of tile t + 1.  The kernels are HBM-bound
// (arithmetic intensity Cin*Cout/(2(Cin+Cout)) flop/byte = 16-64): the roofline is the copy bandwidth.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tcgen05.cuh"

namespace bqa {
namespace {

constexpr int kTileN = 128;      // positions per forward tile
constexpr int kKC = 32;          // K per pipeline stage: 32 fp32 = one 128-byte swizzle row
constexpr int kMaxStages = 12;   // ring depth is chosen per launch: as many stages as shared memory holds
constexpr int kThreads = 192;    // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue

// ---- tensor maps (host) -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 tensor (d2, d1, d0) with d0 contiguous, read as TF32; box (b0 <= 32, b1, 1); 128-byte swizzle
// (16-byte chunks, or 32-byte chunks for `atom32`: the only swizzle an MN-major TF32 operand may
// have); out-of-bounds elements read as zero
int make_map(CUtensorMap *tm, const void *base, long long d0, long long d1, long long d2, long long stride1,
             long long stride2, int b0, int b1, bool atom32 = false, bool plain_f32 = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(BQA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)stride1 * 4, (cuuint64_t)stride2 * 4};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(tm, plain_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 3, const_cast<void *>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(BQA_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): dims %lld %lld %lld strides %lld %lld box %d %d",
                     (int)r, d0, d1, d2, stride1, stride2, b0, b1);
  return BQA_OK;
}

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// a broken hand-shake traps (the launch fails loudly) instead of hanging the GPU: ~4 s at 2 GHz
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {
      printf("[bqa conv] mbarrier wait timed out: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap *tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// shared-memory matrix descriptor, version 1; layout type 2 = 128-byte swizzle of 16-byte chunks,
// 1 = 128-byte swizzle of 32-byte chunks (MN-major TF32)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                               uint64_t layout_type = 2) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (layout_type << 61);
}
// instruction descriptor for kind::tf32: D f32, A / B tf32 (format 2), A K-major, B MN- or K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int m, int n, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Bars {
  uint64_t full[kMaxStages], empty[kMaxStages], acc_full[2], acc_empty[2];
};

// ---- forward / dgrad ----------------------------------------------------------------------------
struct ConvParams {
  int b, cin, cout, p;           // sizes; p = positions per scene
  int rows;                      // output channels of one CTA slab, padded: 128 or 256
  int nchunks;                   // ceil(cin / 32)
  int tiles_per_scene, num_tiles;
  int nstages;                   // ring depth
  float *y;                      // (b, cout, p)
  const float *shift;            // per-channel shift of the statistics (or NULL: 0)
  double *sums;                  // [2 * cout]: sum (y - shift), sum (y - shift)^2; or NULL
};

__global__ void __launch_bounds__(kThreads, 1)
conv1x1_tf32_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                    const __grid_constant__ CUtensorMap tm_y, const ConvParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128-byte swizzle atoms
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t w_bytes = (uint32_t)P.rows * 128u;            // W chunk: rows x 32 tf32
  const uint32_t x_bytes = kTileN * kKC * 4;                   // X chunk: 4 blocks of [32 ch][32 pos]
  const uint32_t stage_bytes = w_bytes + x_bytes;
  const uint32_t nst = (uint32_t)P.nstages;
  __shared__ Bars bars;
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int halves = P.rows / 128;
  const int slab = blockIdx.y;                                  // 256 output channels per slab
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(smem_u32(&bars.full[s]), 1); mbar_init(smem_u32(&bars.empty[s]), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(smem_u32(&bars.acc_full[a]), 1); mbar_init(smem_u32(&bars.acc_empty[a]), 128); }
    fence_mbar_init_cluster();
    prefetch_map(&tm_x);
    prefetch_map(&tm_w);
    prefetch_map(&tm_y);
  }
  if (warp == 1) umma::tmem_alloc(smem_u32(&s_tmem), 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ---- producer: one thread feeds the ring with TMA ------------------------------------
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int scene = tile / P.tiles_per_scene;
        const int p0 = (tile % P.tiles_per_scene) * kTileN;
        for (int c = 0; c < P.nchunks; ++c, ++g) {
          const uint32_t s = g % nst, u = g / nst;
          wait_bar(smem_u32(&bars.empty[s]), (u & 1) ^ 1, 1);   // a fresh barrier passes a wait on parity 1
          const uint32_t full = smem_u32(&bars.full[s]);
          mbar_arrive_expect_tx(full, stage_bytes);
          const uint32_t dst = smem_base + s * stage_bytes;
          tma_load_3d(dst, &tm_w, c * kKC, slab * 256, 0, full);
#pragma unroll
          for (int q = 0; q < kTileN / 32; ++q)
            tma_load_3d(dst + w_bytes + q * (kKC * 128), &tm_x, p0 + 32 * q, c * kKC, scene, full);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, kTileN, 1);
      // A (W chunk): K-major, 16-byte-chunk swizzle, 8-row groups 1024 B apart.  B (X chunk): MN-major,
      // 32-byte-chunk swizzle (atoms of 4 channel rows x 128 B): 32-position blocks 4096 B apart (LBO),
      // 4-channel groups 512 B apart (SBO)
      const uint64_t ad0 = desc_sw128(smem_base, 16, 1024);
      const uint64_t bd0 = desc_sw128(smem_base + w_bytes, kKC * 128, 512, 1);
      uint32_t g = 0, it = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t a = it & 1, ua = it >> 1;
        wait_bar(smem_u32(&bars.acc_empty[a]), (ua & 1) ^ 1, 2);
        umma::fence_after_sync();
        const uint32_t d0 = tmem + a * 256;
        for (int c = 0; c < P.nchunks; ++c, ++g) {
          const uint32_t s = g % nst, u = g / nst;
          wait_bar(smem_u32(&bars.full[s]), u & 1, 3);
          umma::fence_after_sync();
          // descriptors of this stage = stage-0 descriptors + (s * stage_bytes) >> 4 in the address field
          const uint64_t ad_s = ad0 + (uint64_t)(s * (stage_bytes >> 4));
          const uint64_t bd_s = bd0 + (uint64_t)(s * (stage_bytes >> 4));
#pragma unroll
          for (int k = 0; k < kKC / 8; ++k) {
            // A: 128 channels x 8 tf32 = 32 B of every 128-byte row (+2 per k-step); second half 16 KB on.
            // B: 8 channels x 128 positions = 8 rows of every 32-position block (+1024 B per k-step)
            mma_tf32(d0, ad_s + (uint64_t)(2 * k), bd_s + (uint64_t)(64 * k), idesc, (c | k) != 0);
            if (halves == 2)
              mma_tf32(d0 + 128, ad_s + (uint64_t)(2 * k + (128 * 128 >> 4)), bd_s + (uint64_t)(64 * k), idesc, (c | k) != 0);
          }
          umma::commit(smem_u32(&bars.empty[s]));              // stage reusable once these MMAs have read it
        }
        umma::commit(smem_u32(&bars.acc_full[a]));
      }
    }
  } else {
    // ---- epilogue: TMEM -> y (+ statistics) -------------------------------------------------------
    const int quarter = warp & 3;                               // TMEM lane quarter this warp may read
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    double s1[2] = {0.0, 0.0}, s2[2] = {0.0, 0.0};
    float shift[2] = {0.f, 0.f};
    // two 4 KB staging tiles per epilogue warp, behind the ring
    const uint32_t stage_out = smem_base + nst * stage_bytes + (uint32_t)(warp - 2) * 8192u;
    int buf = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int co = slab * 256 + h * 128 + quarter * 32 + lane;
      if (h < halves && P.shift && co < P.cout) shift[h] = P.shift[co];
    }
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t a = it & 1, ua = it >> 1;
      const int scene = tile / P.tiles_per_scene;
      const int p0 = (tile % P.tiles_per_scene) * kTileN;
      wait_bar(smem_u32(&bars.acc_full[a]), ua & 1, 4);
      umma::fence_after_sync();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h >= halves) break;
        const int co = slab * 256 + h * 128 + quarter * 32 + lane;
        float t1 = 0.f, t2 = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < kTileN; c0 += 32) {
          uint32_t v[32];
          umma::ld_32x32b_x32(tmem + lane_addr + a * 256 + h * 128 + c0, v);
          umma::wait_ld();
          // y: this warp's 32 channels x 32 positions go through a swizzled staging tile (row = channel,
          // 128 B) and one TMA store, which also clips channels >= cout and positions >= p
          {
            if (lane == 0) bulk_wait_read<1>();                 // the store that last read this buffer is done
            __syncwarp();
            const uint32_t tile_s = stage_out + (uint32_t)buf * 4096u;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                           ::"r"(tile_s + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) * 16)), "r"(v[4 * q]),
                             "r"(v[4 * q + 1]), "r"(v[4 * q + 2]), "r"(v[4 * q + 3])
                           : "memory");
            umma::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tm_y, tile_s, p0 + c0, slab * 256 + h * 128 + quarter * 32, scene);
              bulk_commit();
            }
            buf ^= 1;
          }
          if (co < P.cout) {
            if (p0 + c0 + 32 <= P.p) {
              // four independent chains per sum (a 32-long dependent chain would cost ~130 cycles)
              float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const float d = __uint_as_float(v[e]) - shift[h];
                a1[e & 3] += d;
                a2[e & 3] = fmaf(d, d, a2[e & 3]);
              }
              t1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
              t2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
            } else {
              for (int e = 0; e < 32; ++e) {
                if (p0 + c0 + e < P.p) {
                  const float d = __uint_as_float(v[e]) - shift[h];
                  t1 += d;
                  t2 = fmaf(d, d, t2);
                }
              }
            }
          }
        }
        s1[h] += (double)t1;
        s2[h] += (double)t2;
      }
      umma::fence_before_sync();
      mbar_arrive(smem_u32(&bars.acc_empty[a]));
    }
    if (lane == 0) bulk_wait_read<0>();                         // staging tiles are read before the CTA exits
    if (P.sums) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int co = slab * 256 + h * 128 + quarter * 32 + lane;
        if (h < halves && co < P.cout) {
          atomicAdd(&P.sums[co], s1[h]);
          atomicAdd(&P.sums[P.cout + co], s2[h]);
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

// ---- weight gradient ---------------------------------------------------------------------------
struct WgradParams {
  int b, cin, cout, p;
  int rows;                      // dy rows per CTA, padded: 128 or 256
  int ncols;                     // x channels of one CTA slab, padded to a multiple of 16, <= 256
  int steps_per_scene;           // ceil(p / 32)
  long long total_steps;         // b * steps_per_scene
  int splits;                    // gridDim.x
  int nstages;                   // ring depth
  float *dw;                     // (cout, cin), accumulated with red.add
};

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                  const WgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = (uint32_t)P.rows * 128u;            // dy step: rows x 32 positions
  const uint32_t b_bytes = (uint32_t)P.ncols * 128u;           // x step: ncols channels x 32 positions
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t nst = (uint32_t)P.nstages;
  __shared__ Bars bars;
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int halves = P.rows / 128;
  const int slab = blockIdx.y;                                  // 256 input channels per slab
  const uint32_t smem_base = smem_u32(smem);
  const long long per = (P.total_steps + P.splits - 1) / P.splits;
  const long long st0 = (long long)blockIdx.x * per;
  const long long st1 = st0 + per < P.total_steps ? st0 + per : P.total_steps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(smem_u32(&bars.full[s]), 1); mbar_init(smem_u32(&bars.empty[s]), 1); }
    mbar_init(smem_u32(&bars.acc_full[0]), 1);
    fence_mbar_init_cluster();
    prefetch_map(&tm_dy);
    prefetch_map(&tm_x);
  }
  if (warp == 1) umma::tmem_alloc(smem_u32(&s_tmem), 512);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = s_tmem;
  if (st0 >= st1) {                                             // no positions for this CTA
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
    return;
  }

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (long long st = st0; st < st1; ++st, ++g) {
        const int scene = (int)(st / P.steps_per_scene);
        const int p0 = (int)(st % P.steps_per_scene) * 32;
        const uint32_t s = g % nst, u = g / nst;
        wait_bar(smem_u32(&bars.empty[s]), (u & 1) ^ 1, 11);
        const uint32_t full = smem_u32(&bars.full[s]);
        mbar_arrive_expect_tx(full, stage_bytes);
        const uint32_t dst = smem_base + s * stage_bytes;
        tma_load_3d(dst, &tm_dy, p0, 0, scene, full);
        tma_load_3d(dst + a_bytes, &tm_x, p0, slab * 256, scene, full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(128, P.ncols, 0);
      const uint64_t ad0 = desc_sw128(smem_base, 16, 1024);              // dy rows: K-major, 8-row groups 1024 B apart
      const uint64_t bd0 = desc_sw128(smem_base + a_bytes, 16, 1024);    // x rows: the same
      uint32_t g = 0;
      for (long long st = st0; st < st1; ++st, ++g) {
        const uint32_t s = g % nst, u = g / nst;
        wait_bar(smem_u32(&bars.full[s]), u & 1, 13);
        umma::fence_after_sync();
        const uint64_t ad_s = ad0 + (uint64_t)(s * (stage_bytes >> 4));
        const uint64_t bd_s = bd0 + (uint64_t)(s * (stage_bytes >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {                           // 8 positions (32 B of every row) per MMA
          mma_tf32(tmem, ad_s + (uint64_t)(2 * k), bd_s + (uint64_t)(2 * k), idesc, (g | (uint32_t)k) != 0);
          if (halves == 2)
            mma_tf32(tmem + 256, ad_s + (uint64_t)(2 * k + (128 * 128 >> 4)), bd_s + (uint64_t)(2 * k), idesc,
                     (g | (uint32_t)k) != 0);
        }
        umma::commit(smem_u32(&bars.empty[s]));
      }
      umma::commit(smem_u32(&bars.acc_full[0]));
    }
  } else {
    const int quarter = warp & 3;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    wait_bar(smem_u32(&bars.acc_full[0]), 0, 14);
    umma::fence_after_sync();
    for (int h = 0; h < halves; ++h) {
      const int co = h * 128 + quarter * 32 + lane;
      float *row = P.dw + (size_t)(co < P.cout ? co : 0) * P.cin + slab * 256;
      for (int c0 = 0; c0 < P.ncols; c0 += 32) {
        uint32_t v[32];
        umma::ld_32x32b_x32(tmem + lane_addr + h * 256 + c0, v);
        umma::wait_ld();
        if (co < P.cout) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c0 + e < P.ncols && slab * 256 + c0 + e < P.cin) atomicAdd(row + c0 + e, __uint_as_float(v[e]));
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

int sm_count() {
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  return sms;
}

}  // namespace

// x (b, cin, p), w (cout, ldw) with ldw >= cin a multiple of 4 (rows 16-byte aligned), y (b, cout, p);
// p a multiple of 4.  sums (optional): [2 * cout] doubles, ACCUMULATED into (the caller zeroes them).
bool conv1x1_tf32_supported(int b, int cin, int cout, long long p, int ldw) {
  return b >= 1 && cin >= 1 && cout >= 1 && p >= 1 && (p % 4) == 0 && p < (1ll << 31) && ldw >= cin && (ldw % 4) == 0;
}

int conv1x1_tf32_forward(int b, int cin, int cout, int p, const float *x, const float *w, int ldw, float *y,
                         const float *shift, double *sums, cudaStream_t stream) {
  if (!conv1x1_tf32_supported(b, cin, cout, p, ldw))
    return set_error(BQA_ERR_UNSUPPORTED, "conv1x1_tf32: unsupported shape b=%d cin=%d cout=%d p=%d ldw=%d", b, cin, cout, p, ldw);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(y) & 15))
    return set_error(BQA_ERR_INVALID_ARG, "conv1x1_tf32: tensors must be 16-byte aligned");
  ConvParams P;
  P.b = b; P.cin = cin; P.cout = cout; P.p = p;
  P.rows = cout > 128 ? 256 : 128;
  P.nchunks = ceil_div(cin, kKC);
  P.tiles_per_scene = ceil_div(p, kTileN);
  P.num_tiles = b * P.tiles_per_scene;
  P.y = y; P.shift = shift; P.sums = sums;
  CUtensorMap tm_x, tm_w;
  if (int rc = make_map(&tm_x, x, p, cin, b, p, (long long)cin * p, 32, kKC, true)) return rc;
  if (int rc = make_map(&tm_w, w, cin, cout, 1, ldw, (long long)cout * ldw, kKC, P.rows)) return rc;
  CUtensorMap tm_y;
  if (int rc = make_map(&tm_y, y, p, cout, b, p, (long long)cout * p, 32, 32, false, true)) return rc;
  const size_t stage = (size_t)P.rows * 128 + kTileN * kKC * 4;
  P.nstages = (int)((size_t)(226 * 1024 - 4 * 8192 - 1024) / stage);
  if (P.nstages > kMaxStages) P.nstages = kMaxStages;
  const size_t smem = P.nstages * stage + 4 * 8192 + 1024;
  BQA_CUDA(cudaFuncSetAttribute(conv1x1_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int slabs = ceil_div(cout, 256);
  dim3 grid((unsigned)min(P.num_tiles, max(1, sm_count() / slabs)), (unsigned)slabs);
  conv1x1_tf32_kernel<<<grid, kThreads, smem, stream>>>(tm_x, tm_w, tm_y, P);
  count_launch();
  return check_launch("conv1x1_tf32_kernel");
}

// dw (cout, cin) += sum_{b,p} dy[b,:,p] x[b,:,p]^T   (the caller zeroes dw); cout <= 256
int conv1x1_tf32_wgrad(int b, int cin, int cout, int p, const float *x, const float *dy, float *dw,
                       cudaStream_t stream) {
  if (!conv1x1_tf32_supported(b, cin, cout, p, cin + (4 - cin % 4) % 4) || cout > 256)
    return set_error(BQA_ERR_UNSUPPORTED, "conv1x1_tf32 wgrad: unsupported shape b=%d cin=%d cout=%d p=%d", b, cin, cout, p);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(dy) & 15))
    return set_error(BQA_ERR_INVALID_ARG, "conv1x1_tf32 wgrad: tensors must be 16-byte aligned");
  WgradParams P;
  P.b = b; P.cin = cin; P.cout = cout; P.p = p;
  P.rows = cout > 128 ? 256 : 128;
  const int slabs = ceil_div(cin, 256);
  // every slab uses the same padded column count (the last one reads zeros beyond cin)
  P.ncols = cin >= 256 ? 256 : ceil_div(cin, 16) * 16;
  P.steps_per_scene = ceil_div(p, 32);
  P.total_steps = (long long)b * P.steps_per_scene;
  long long want = P.total_steps / 16;                          // at least 16 steps per CTA
  if (want < 1) want = 1;
  P.splits = (int)(want < sm_count() / slabs ? want : max(1, sm_count() / slabs));
  P.dw = dw;
  CUtensorMap tm_dy, tm_x;
  if (int rc = make_map(&tm_dy, dy, p, cout, b, p, (long long)cout * p, 32, P.rows)) return rc;
  if (int rc = make_map(&tm_x, x, p, cin, b, p, (long long)cin * p, 32, P.ncols)) return rc;
  const size_t stage = (size_t)(P.rows + P.ncols) * 128;
  P.nstages = (int)((size_t)(226 * 1024 - 1024) / stage);
  if (P.nstages > kMaxStages) P.nstages = kMaxStages;
  const size_t smem = P.nstages * stage + 1024;
  BQA_CUDA(cudaFuncSetAttribute(wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)P.splits, (unsigned)slabs);
  wgrad_tf32_kernel<<<grid, kThreads, smem, stream>>>(tm_dy, tm_x, P);
  count_launch();
  return check_launch("wgrad_tf32_kernel");
}

}  // namespace bqa
