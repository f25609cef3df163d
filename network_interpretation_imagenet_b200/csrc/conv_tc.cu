// conv_tc.cu — stage 2 hot kernel: implicit-GEMM convolution on the 5th-gen tensor cores.
//
//   out[m, co] = act( sum_k A[m, k] * Wt[co, k] + bias[co] + residual[m, co] ),  bf16 x bf16 -> fp32
//
// with m = (n, p, q) flattened output pixels, k = (r, s, c).  A is never materialised:
//   * 1x1 stride-1 convs read the NHWC activation matrix [N*H*W, C] with a tiled TMA map;
//   * every other geometry (3x3, strided) uses a TMA *im2col* map: the copy engine walks 128
//     consecutive output pixels (w fastest, then h, then n) of the padded bounding box and
//     fetches 64 channels of filter tap (r, s) for each, zero-filling the padding.
// Both land as a 128 x 64 bf16, 128B-swizzled, K-major tile = one tcgen05.mma operand.
// Weights are KRSC ([Cout][R*S*Cin], BN folded) and arrive through a second tiled map.
//
// CTA = 192 threads, warp-specialised:  warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer
// (warp 1 also owns the TMEM allocation), warps 2..5 = epilogue (TMEM -> registers -> bias /
// residual / ReLU -> bf16 -> global).  The accumulator (128 lanes x BLOCK_N fp32 columns) lives
// in TMEM; smem holds a STAGES-deep ring of (A, B) tiles guarded by full/empty mbarriers.
// Two CTAs are co-resident per SM (<= 113 KB smem, <= 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.
//
// This replaces the cuDNN calls behind `model(masked_img_tensor)`
// (generate_gp_training_data_imagenet.py:246, bayesian_active_learning_imagenet.py:192).
#include "common.cuh"
#include "layers.cuh"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>

namespace nib {

static constexpr int TC_BLOCK_M = 128;
static constexpr int TC_BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
static constexpr int TC_UMMA_K = 16;
static constexpr int TC_THREADS = 192;

struct TcKernelParams {
  const float* bias;
  const __nv_bfloat16* res;
  void* out;
  int M;            // valid output rows
  int Cout;
  int out_cstride, out_coff;
  int res_cstride, res_coff;
  int relu;
  int out_f32;      // 1: store fp32 (GEMM self-test), 0: bf16
  int num_k_blocks;
  int cblocks;      // Cin / 64
  int S;            // filter width (tap -> (r, s))
  int im2col;       // 0: tiled 2D A map, 1: im2col 4D A map, 2: stem rows (overlapping-window 4D tiled map)
  int tile_rows;    // valid output rows per M tile (128; Q for the stem mode: one output row per tile)
  int a_bytes;      // bytes the A load delivers per stage (tile_rows * 128)
  int P, Q, stride, pad;
  int in_coff;
  int n_tiles;      // ceil(Cout / BLOCK_N)
  int m_tiles;      // ceil(M / tile_rows)
  int pf_dist;      // v2: L2-prefetch the operands of the tile this many iterations ahead (0 = off)
  unsigned int* err_flag;
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a descriptor bug must fail the launch, not hang the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* err_flag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
      if (err_flag) atomicExch(err_flag, (unsigned)code);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a future tile's operands: with one CTA per SM the smem ring holds < 1 tile of a K=256 layer, so a
// DRAM-latency load (~3 us measured under load) throttles the ring to ring_bytes / latency.  Prefetching tiles that
// are `pf_dist` iterations ahead into L2 turns those into L2-hit loads without spending shared memory.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_im2col_4d(const CUtensorMap* tm, int c, int w, int h, int n, uint16_t off_w,
                                                       uint16_t off_h) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.im2col [%0, {%1, %2, %3, %4}], {%5, %6};"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled smem tile (rows of 128 B, 8-row groups 1024 B apart).
// cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

template <int BLOCK_N>
__host__ __device__ constexpr uint32_t make_idesc_bf16() {
  // cute::UMMA::InstrDescriptor: c_format F32=1 @4, a_format BF16=1 @7, b_format BF16=1 @10,
  // a_major/b_major K=0 @15/@16, N>>3 @17, M>>4 @24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(TC_BLOCK_M >> 4) << 24);
}

template <int BLOCK_N, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;  // 16 KB
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFFSET + (2 * STAGES + 1) * 8 + 16 + 1024 /*alignment slack*/;
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TcKernelParams p) {
  using SM = TcSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + SM::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int m0 = m_tile * p.tile_rows;
  const int n0 = n_tile * BLOCK_N;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)(BLOCK_N < 32 ? 32 : BLOCK_N))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int w0 = 0, h0 = 0, img = 0;
      if (p.im2col == 1) {
        const int pq = p.P * p.Q;
        img = m0 / pq;
        const int rem = m0 - img * pq;
        const int pp = rem / p.Q, qq = rem - pp * p.Q;
        w0 = qq * p.stride - p.pad;
        h0 = pp * p.stride - p.pad;
      } else if (p.im2col == 2) {   // one output row (img, pp) per tile; input rows pp*stride + r of the haloed image
        img = m_tile / p.P;
        h0 = (m_tile - img * p.P) * p.stride;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
        const uint32_t a_dst = smem_base + stage * SM::STAGE_BYTES;
        const uint32_t b_dst = a_dst + SM::A_BYTES;
        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(p.a_bytes + SM::B_BYTES));
        const int tap = kb / p.cblocks;
        const int c0 = (kb - tap * p.cblocks) * TC_BLOCK_K + p.in_coff;
        if (p.im2col == 1) {
          const int r = tap / p.S, s = tap - r * p.S;
          tma_load_im2col_4d(a_dst, &tmA, full_bar(stage), c0, w0, h0, img, (uint16_t)s, (uint16_t)r);
        } else if (p.im2col == 2) {
          tma_load_4d(a_dst, &tmA, full_bar(stage), 0, 0, h0 + kb, img);   // k-block kb = filter row r
        } else {
          tma_load_2d(a_dst, &tmA, full_bar(stage), c0, m0);
        }
        tma_load_2d(b_dst, &tmB, full_bar(stage), kb * TC_BLOCK_K, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16<BLOCK_N>();
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(full_bar(stage), phase, p.err_flag, 2);
        tc_fence_after();
        const uint32_t a_addr = smem_base + stage * SM::STAGE_BYTES;
        const uint32_t b_addr = a_addr + SM::A_BYTES;
        const uint64_t adesc = make_smem_desc_sw128(a_addr);
        const uint64_t bdesc = make_smem_desc_sw128(b_addr);
#pragma unroll
        for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
        }
        umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    mbar_wait(tmem_full_bar, 0, p.err_flag, 3);
    tc_fence_after();
    const int lrow = quarter * 32 + lane;
    const int row = m0 + lrow;
    const bool row_ok = row < p.M && lrow < p.tile_rows;
#pragma unroll 1
    for (int cc = 0; cc < BLOCK_N; cc += 32) {
      uint32_t v[32];
      __syncwarp();  // .sync.aligned TMEM loads need the full warp converged
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)cc, v);
      tmem_ld_wait();
      const int col0 = n0 + cc;
      if (row_ok && col0 < p.Cout) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
          f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
        }
      }
      if (p.res != nullptr) {
        const __nv_bfloat16* rp = p.res + (size_t)row * p.res_cstride + p.res_coff + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 raw = *reinterpret_cast<const uint4*>(rp + j);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 r2 = __bfloat1622float2(h2[t]);
            f[j + 2 * t] += r2.x;
            f[j + 2 * t + 1] += r2.y;
          }
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      if (p.out_f32) {
        float* op = reinterpret_cast<float*>(p.out) + (size_t)row * p.out_cstride + p.out_coff + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      } else {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.out_cstride + p.out_coff + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 o;
          __nv_bfloat162 h;
          h = __floats2bfloat162_rn(f[j], f[j + 1]);     o.x = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2bfloat162_rn(f[j + 2], f[j + 3]); o.y = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2bfloat162_rn(f[j + 4], f[j + 5]); o.z = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2bfloat162_rn(f[j + 6], f[j + 7]); o.w = *reinterpret_cast<uint32_t*>(&h);
          *reinterpret_cast<uint4*>(op + j) = o;
        }
      }
      }  // row_ok
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)(BLOCK_N < 32 ? 32 : BLOCK_N))
                 : "memory");
  }
}


// =====================================================================================================
// v2: persistent kernel.  One CTA per SM loops over output tiles (n-tile fastest so the CTAs that share an
// A tile run together and hit L2).  Differences from v1 that matter for the memory-heavy layers (1x1
// expansions with residual, K = 64..256, where the epilogue — not the MMA — is the critical path):
//   * accumulators are double-buffered in TMEM (2 x BLOCK_N columns): the MMA issuer starts tile i+1 while
//     the epilogue warps drain tile i;
//   * the residual tile is prefetched by the TMA producer into (double-buffered) swizzled smem while the
//     main loop runs, instead of 64-byte strided global loads per lane;
//   * the output goes registers -> swizzled smem box (128 rows x 64 channels) -> one TMA store per box:
//     full 128 B lines instead of 32 sectors per store request (profiles/r01_ncu_conv_tc_v1_layer3.csv).
static constexpr int TC2_EPI_WGS = 2;                         // epilogue warpgroups (one per TMEM accumulator buffer)
static constexpr int TC2_THREADS = 64 + 128 * TC2_EPI_WGS;    // producer warp + MMA warp + epilogue warps
static constexpr int TC2_BRES_KBLOCKS = 9;                    // resident-weights mode: up to 9 k-blocks (3x3 x 64 ch)

// BRES: the whole weight matrix of the layer (Cout == BLOCK_N, <= 9 k-blocks) is loaded once per CTA and stays in
// smem for every tile; the ring then carries A tiles only.  For the 64-channel layers (stem, 56x56 3x3) the
// per-tile weight re-fetch was the dominant L2 -> SM traffic.
template <int BLOCK_N, int STAGES, bool HAS_RES, bool BRES>
struct Tc2Smem {
  static constexpr int A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * TC_BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + (BRES ? 0 : B_BYTES);
  static constexpr int NBOX = BLOCK_N / 64;
  static constexpr int BOX_BYTES = TC_BLOCK_M * 128;
  static constexpr int BRES_OFFSET = STAGES * STAGE_BYTES;
  static constexpr int BRES_BYTES = BRES ? TC2_BRES_KBLOCKS * B_BYTES : 0;
  static constexpr int RES_OFFSET = BRES_OFFSET + BRES_BYTES;
  static constexpr int RES_BYTES = HAS_RES ? 2 * NBOX * BOX_BYTES : 0;
  static constexpr int OUT_OFFSET = RES_OFFSET + RES_BYTES;
  static constexpr int OUT_BYTES = TC2_EPI_WGS * BOX_BYTES;   // one store-staging box per epilogue warpgroup
  static constexpr int BAR_OFFSET = OUT_OFFSET + OUT_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 9;
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(wg + 1) : "memory"); }
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int BLOCK_N, int STAGES, bool HAS_RES, bool BRES>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                const TcKernelParams p) {
  using SM = Tc2Smem<BLOCK_N, STAGES, HAS_RES, BRES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + SM::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  auto res_full_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 4 + b); };
  auto res_empty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 6 + b); };
  const uint32_t bres_full_bar = bar_base + 8u * (2 * STAGES + 8);
  const uint32_t tmem_slot = bar_base + 8u * SM::NUM_BARS;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  auto res_box = [&](int buf, int b) { return smem_base + SM::RES_OFFSET + (uint32_t)(buf * SM::NBOX + b) * SM::BOX_BYTES; };
  auto out_box = [&](int buf) { return smem_base + SM::OUT_OFFSET + (uint32_t)buf * SM::BOX_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  constexpr bool has_res = HAS_RES;
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmOut);
    if (has_res) prefetch_tmap(&tmRes);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), 4);
      mbar_init(res_full_bar(b), 1);
      mbar_init(res_empty_bar(b), 4);
    }
    mbar_init(bres_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (BRES) {   // the layer's whole weight matrix, once
        mbar_arrive_expect_tx(bres_full_bar, (uint32_t)(p.num_k_blocks * SM::B_BYTES));
        for (int kb = 0; kb < p.num_k_blocks; ++kb)
          tma_load_2d(smem_base + SM::BRES_OFFSET + kb * SM::B_BYTES, &tmB, bres_full_bar, kb * TC_BLOCK_K, 0);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int m0 = m_tile * p.tile_rows;
        const int n0 = n_tile * BLOCK_N;
        if (p.pf_dist > 0) {
          const int ft = tile + p.pf_dist * (int)gridDim.x;
          if (ft < total_tiles) {
            const int fn = ft % p.n_tiles, fm = ft / p.n_tiles;
            const int fm0 = fm * p.tile_rows;
            if (has_res) {
#pragma unroll
              for (int b = 0; b < SM::NBOX; ++b) tma_prefetch_2d(&tmRes, p.res_coff + fn * BLOCK_N + 64 * b, fm0);
            }
            // the A tile is shared by the n_tiles CTAs working on the same rows: one of them prefetches it
            if (fn == 0 || p.n_tiles > (int)gridDim.x) {
              if (p.im2col == 1) {
                const int pq = p.P * p.Q;
                const int fi = fm0 / pq, rem = fm0 - fi * pq;
                const int pp = rem / p.Q, qq = rem - pp * p.Q;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                  const int tap = kb / p.cblocks;
                  const int r = tap / p.S, sx = tap - r * p.S;
                  tma_prefetch_im2col_4d(&tmA, (kb - tap * p.cblocks) * TC_BLOCK_K + p.in_coff, qq * p.stride - p.pad,
                                         pp * p.stride - p.pad, fi, (uint16_t)sx, (uint16_t)r);
                }
              } else if (p.im2col == 2) {
                const int fi = fm / p.P;
                const int fh = (fm - fi * p.P) * p.stride;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) tma_prefetch_4d(&tmA, 0, 0, fh + kb, fi);
              } else {
                for (int kb = 0; kb < p.num_k_blocks; ++kb) tma_prefetch_2d(&tmA, kb * TC_BLOCK_K + p.in_coff, fm0);
              }
            }
          }
        }
        int w0 = 0, h0 = 0, img = 0;
        if (p.im2col == 1) {
          const int pq = p.P * p.Q;
          img = m0 / pq;
          const int rem = m0 - img * pq;
          const int pp = rem / p.Q, qq = rem - pp * p.Q;
          w0 = qq * p.stride - p.pad;
          h0 = pp * p.stride - p.pad;
        } else if (p.im2col == 2) {
          img = m_tile / p.P;
          h0 = (m_tile - img * p.P) * p.stride;
        }
        if (has_res) {
          const int rb = it & 1;
          mbar_wait(res_empty_bar(rb), (uint32_t)(((it >> 1) & 1) ^ 1), p.err_flag, 4);
          mbar_arrive_expect_tx(res_full_bar(rb), (uint32_t)(SM::NBOX * SM::BOX_BYTES));
#pragma unroll
          for (int b = 0; b < SM::NBOX; ++b)
            tma_load_2d(res_box(rb, b), &tmRes, res_full_bar(rb), p.res_coff + n0 + 64 * b, m0);
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
          const uint32_t a_dst = smem_base + stage * SM::STAGE_BYTES;
          const uint32_t b_dst = a_dst + SM::A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(p.a_bytes + (BRES ? 0 : SM::B_BYTES)));
          const int tap = kb / p.cblocks;
          const int c0 = (kb - tap * p.cblocks) * TC_BLOCK_K + p.in_coff;
          if (p.im2col == 1) {
            const int r = tap / p.S, s = tap - r * p.S;
            tma_load_im2col_4d(a_dst, &tmA, full_bar(stage), c0, w0, h0, img, (uint16_t)s, (uint16_t)r);
          } else if (p.im2col == 2) {
            tma_load_4d(a_dst, &tmA, full_bar(stage), 0, 0, h0 + kb, img);
          } else {
            tma_load_2d(a_dst, &tmA, full_bar(stage), c0, m0);
          }
          if (!BRES) tma_load_2d(b_dst, &tmB, full_bar(stage), kb * TC_BLOCK_K, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_bf16<BLOCK_N>();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (BRES) mbar_wait(bres_full_bar, 0, p.err_flag, 7);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int ab = it & 1;
        mbar_wait(tmem_empty_bar(ab), (uint32_t)(((it >> 1) & 1) ^ 1), p.err_flag, 5);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(ab * BLOCK_N);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 2);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * SM::STAGE_BYTES;
          const uint32_t b_addr = BRES ? smem_base + SM::BRES_OFFSET + kb * SM::B_BYTES : a_addr + SM::A_BYTES;
          const uint64_t adesc = make_smem_desc_sw128(a_addr);
          const uint64_t bdesc = make_smem_desc_sw128(b_addr);
#pragma unroll
          for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k)
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full_bar(ab));
      }
    }
  } else {
    // ===== epilogue: warpgroup wg (warps 2+4wg .. 5+4wg) drains the tiles whose accumulator is TMEM buffer wg =====
    const int wg = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int lrow = quarter * 32 + lane;
    const uint32_t row_off = (uint32_t)lrow * 128u;
    const uint32_t sw = (uint32_t)(lrow & 7);
    const bool leader = (warp == 2 + 4 * wg && lane == 0);
    for (int tile = blockIdx.x + wg * (int)gridDim.x, it = wg; tile < total_tiles; tile += 2 * (int)gridDim.x, it += 2) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int m0 = m_tile * p.tile_rows;
      const int n0 = n_tile * BLOCK_N;
      const int ab = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      mbar_wait(tmem_full_bar(ab), ph, p.err_flag, 3);
      tc_fence_after();
      if (has_res) mbar_wait(res_full_bar(ab), ph, p.err_flag, 6);
#pragma unroll 1
      for (int b = 0; b < SM::NBOX; ++b) {
        if (leader) bulk_wait_read<0>();   // this warpgroup's previous store has finished reading its staging box
        epi_bar_sync(wg);
        const uint32_t obase = out_box(wg) + row_off;
        const uint32_t rbase = res_box(ab, b) + row_off;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BLOCK_N + b * 64 + half * 32), v);
          tmem_ld_wait();
          if (b == SM::NBOX - 1 && half == 1) {   // accumulator fully read: hand the TMEM buffer back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar(ab));
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          const int col0 = n0 + b * 64 + half * 32;
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              f[j] += bb.x; f[j + 1] += bb.y; f[j + 2] += bb.z; f[j + 3] += bb.w;
            }
          }
          if (has_res) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 raw = lds_v4(rbase + ((((uint32_t)(half * 4 + c)) ^ sw) << 4));
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 r2 = __bfloat1622float2(h2[t]);
                f[c * 8 + 2 * t] += r2.x;
                f[c * 8 + 2 * t + 1] += r2.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint4 o;
            __nv_bfloat162 h;
            h = __floats2bfloat162_rn(f[c * 8 + 0], f[c * 8 + 1]); o.x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(f[c * 8 + 2], f[c * 8 + 3]); o.y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(f[c * 8 + 4], f[c * 8 + 5]); o.z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(f[c * 8 + 6], f[c * 8 + 7]); o.w = *reinterpret_cast<uint32_t*>(&h);
            sts_v4(obase + ((((uint32_t)(half * 4 + c)) ^ sw) << 4), o);
          }
        }
        if (has_res && b == SM::NBOX - 1) {
          __syncwarp();
          if (lane == 0) mbar_arrive(res_empty_bar(ab));
        }
        fence_async_smem();   // generic-proxy smem writes -> visible to the TMA (async proxy)
        epi_bar_sync(wg);
        if (leader) {
          tma_store_2d(&tmOut, out_box(wg), p.out_coff + n0 + b * 64, m0);
          bulk_commit();
        }
      }
    }
    if (leader) bulk_wait_read<0>();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encodeTiled = nullptr;
static PFN_encodeIm2col g_encodeIm2col = nullptr;

static int load_driver_fns() {
  if (g_encodeTiled && g_encodeIm2col) return NIB_OK;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  NIB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled not available"); return NIB_ECUDA; }
  g_encodeTiled = (PFN_encodeTiled)fn;
  fn = nullptr;
  NIB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeIm2col not available"); return NIB_ECUDA; }
  g_encodeIm2col = (PFN_encodeIm2col)fn;
  return NIB_OK;
}

struct TcConvPlan {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  int v2;
  int block_n, stages;
  int im2col;
  int tile_rows;
  int num_k_blocks, cblocks;
  unsigned int* err_flag;
};

static unsigned int* g_err_flag = nullptr;

static int encode_2d_bf16(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_bytes,
                          uint32_t box_inner, uint32_t box_outer) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu) row_bytes=%llu box=(%u,%u) ptr=%p", (int)r,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_bytes, box_inner,
              box_outer, ptr);
    return NIB_ECUDA;
  }
  return NIB_OK;
}

// 7x7 stride-2 pad-3 stem on a haloed, 8-channel-padded NHWC input (torchvision conv1): for filter row r the 7 taps x 8
// channels of output pixel q are 56 contiguous elements starting at pixel 2q of input row 2p+r, so a tiled map whose
// pixel dimension advances 2 pixels (32 B) per index delivers a [Q x 64] K-major tile per filter row (8 zero-weight
// filler elements).  K = 7 x 64.
bool tc_conv_is_stem(const ConvParams& p) {
  return p.R == 7 && p.S == 7 && p.stride == 2 && p.pad == 3 && p.Cin <= 8 && p.in_cstride == 8 && p.in_coff == 0 &&
         p.in_halo == 3 && p.Cout % 32 == 0 && p.Q <= TC_BLOCK_M && p.out_halo == 0 && p.out_cstride % 8 == 0 &&
         p.out_coff % 8 == 0 && p.pre_scale == nullptr && p.res == nullptr && p.w_alt != nullptr &&
         (p.Win + 2 * p.in_halo) >= 2 * (p.Q - 1) + 8;
}

bool tc_conv_supported(const ConvParams& p) {
  if (tc_conv_is_stem(p)) return true;
  if (p.Cin % TC_BLOCK_K != 0) return false;
  if (p.in_cstride % 8 != 0 || p.in_coff % 8 != 0) return false;   // 16 B TMA alignment
  if (p.Cout % 32 != 0) return false;
  if (p.out_cstride % 8 != 0 || p.out_coff % 8 != 0) return false;
  if (p.in_halo != 0 || p.out_halo != 0) return false;
  if (p.pre_scale != nullptr) return false;                        // pre-activation needs a register path
  if (p.res != nullptr && (p.res_C != p.Cout || p.res_cstride % 8 != 0 || p.res_coff % 8 != 0)) return false;
  if (p.R != p.S) return false;
  if (p.pad > 127 || p.R > 16) return false;
  return true;
}

static int pick_block_n(int Cout, bool has_res = true, int ksize = 1) {
  const char* e = getenv("NIB_TC_BLOCK_N");
  if (e) {
    int v = atoi(e);
    if ((v == 32 || v == 64 || v == 128 || v == 256) && Cout % v == 0) return v;
  }
  // tuning knobs for the sweep: block-N of residual-free 3x3 / 1x1 layers
  const char* e3 = getenv(ksize > 1 ? "NIB_TC_BLOCK_N_3X3" : "NIB_TC_BLOCK_N_1X1");
  if (e3 && !has_res) {
    int v = atoi(e3);
    if ((v == 64 || v == 128 || v == 256) && Cout % v == 0) return v;
  }
  if (!has_res && Cout % 256 == 0) return 256;   // 128x256 tiles: 85 FLOP per byte of smem fill (L2 -> SM is the limiter)
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  return 32;
}

// v2 (persistent, TMA-store epilogue) needs 64-column output boxes: BLOCK_N in {64,128}, Cout % BLOCK_N == 0.
static int finish_plan(TcConvPlan* plan, const ConvParams& p, int max_batch) {
  plan->v2 = 0;
  const char* e = getenv("NIB_TC_V1");
  if (e && atoi(e) != 0) return NIB_OK;
  const bool bn_ok = plan->block_n == 64 || plan->block_n == 128 || (plan->block_n == 256 && p.res == nullptr);
  if (!bn_ok || p.Cout % plan->block_n != 0) return NIB_OK;
  const uint64_t rows = (uint64_t)max_batch * p.P * p.Q;
  int rc = encode_2d_bf16(&plan->tmOut, p.out, (uint64_t)p.out_cstride, rows, (uint64_t)p.out_cstride * 2, 64,
                          (uint32_t)plan->tile_rows);
  if (rc != NIB_OK) return rc;
  if (p.res != nullptr) {
    rc = encode_2d_bf16(&plan->tmRes, p.res, (uint64_t)p.res_cstride, rows, (uint64_t)p.res_cstride * 2, 64, TC_BLOCK_M);
    if (rc != NIB_OK) return rc;
  } else {
    plan->tmRes = plan->tmOut;
  }
  plan->v2 = 1;
  return NIB_OK;
}

int tc_conv_plan_create(const ConvParams& p, int max_batch, TcConvPlan** out) {
  int rc = load_driver_fns();
  if (rc != NIB_OK) return rc;
  if (!g_err_flag) {
    NIB_CUDA(cudaMalloc(&g_err_flag, sizeof(unsigned int)));
    NIB_CUDA(cudaMemset(g_err_flag, 0, sizeof(unsigned int)));
  }
  TcConvPlan* plan = new TcConvPlan();
  memset(plan, 0, sizeof(*plan));
  plan->err_flag = g_err_flag;
  plan->block_n = pick_block_n(p.Cout, p.res != nullptr, p.R);
  plan->tile_rows = TC_BLOCK_M;
  if (tc_conv_is_stem(p)) {
    plan->im2col = 2;
    plan->cblocks = 1;
    plan->num_k_blocks = 7;
    plan->tile_rows = p.Q;
    rc = encode_2d_bf16(&plan->tmB, p.w_alt, 7 * 64, (uint64_t)p.Cout, 7 * 64 * 2, TC_BLOCK_K, plan->block_n);
    if (rc != NIB_OK) { delete plan; return rc; }
    const int Hp = p.Hin + 6, Wp = p.Win + 6;
    cuuint64_t dims[4] = {64, (cuuint64_t)p.Q, (cuuint64_t)Hp, (cuuint64_t)max_batch};
    cuuint64_t strides[3] = {32, (cuuint64_t)Wp * 16, (cuuint64_t)Hp * Wp * 16};   // 2 pixels, one row, one image
    cuuint32_t box[4] = {64, (cuuint32_t)p.Q, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_encodeTiled(&plan->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.in), dims, strides,
                               box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled (stem, overlapping windows) failed (%d)", (int)r);
      delete plan;
      return NIB_ECUDA;
    }
    rc = finish_plan(plan, p, max_batch);
    if (rc != NIB_OK) { delete plan; return rc; }
    *out = plan;
    return NIB_OK;
  }
  plan->cblocks = p.Cin / TC_BLOCK_K;
  plan->num_k_blocks = p.R * p.S * plan->cblocks;
  plan->im2col = !(p.R == 1 && p.S == 1 && p.stride == 1 && p.pad == 0);
  const int K = p.R * p.S * p.Cin;
  // B: weights [Cout][K]
  rc = encode_2d_bf16(&plan->tmB, p.w, (uint64_t)K, (uint64_t)p.Cout, (uint64_t)K * 2, TC_BLOCK_K, plan->block_n);
  if (rc != NIB_OK) { delete plan; return rc; }
  if (!plan->im2col) {
    const uint64_t rows = (uint64_t)max_batch * p.Hin * p.Win;
    rc = encode_2d_bf16(&plan->tmA, p.in, (uint64_t)p.in_cstride, rows, (uint64_t)p.in_cstride * 2, TC_BLOCK_K,
                        TC_BLOCK_M);
    if (rc != NIB_OK) { delete plan; return rc; }
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)p.in_cstride, (cuuint64_t)p.Win, (cuuint64_t)p.Hin, (cuuint64_t)max_batch};
    cuuint64_t strides[3] = {(cuuint64_t)p.in_cstride * 2, (cuuint64_t)p.Win * p.in_cstride * 2,
                             (cuuint64_t)p.Hin * p.Win * p.in_cstride * 2};
    int lower[2] = {-p.pad, -p.pad};
    int upper[2] = {p.pad - (p.S - 1), p.pad - (p.R - 1)};
    cuuint32_t es[4] = {1, (cuuint32_t)p.stride, (cuuint32_t)p.stride, 1};
    CUresult r = g_encodeIm2col(&plan->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.in), dims,
                                strides, lower, upper, TC_BLOCK_K, TC_BLOCK_M, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeIm2col failed (%d): C=%d W=%d H=%d N=%d R=%d stride=%d pad=%d", (int)r,
                p.in_cstride, p.Win, p.Hin, max_batch, p.R, p.stride, p.pad);
      delete plan;
      return NIB_ECUDA;
    }
    // Driver quirk mirrored from CUTLASS (cute/atom/copy_traits_sm90_im2col.hpp): for tensors smaller
    // than 128 KiB, drivers <= 13.1 set a descriptor bit that must be cleared.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    const uint64_t bytes = (uint64_t)max_batch * p.Hin * p.Win * p.in_cstride * 2;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(&plan->tmA)[1] &= ~(1ull << 21);
  }
  rc = finish_plan(plan, p, max_batch);
  if (rc != NIB_OK) { delete plan; return rc; }
  *out = plan;
  return NIB_OK;
}

void tc_conv_plan_destroy(TcConvPlan* plan) { delete plan; }
int tc_conv_plan_block_n(const TcConvPlan* plan) { return plan->block_n; }

template <int BLOCK_N, int STAGES>
static int launch_tc(const TcConvPlan* plan, const TcKernelParams& kp, int tiles, cudaStream_t st) {
  using SM = TcSmem<BLOCK_N, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    NIB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SM::TOTAL));
    attr_set = true;
  }
  conv_tc_kernel<BLOCK_N, STAGES><<<tiles, TC_THREADS, SM::TOTAL, st>>>(plan->tmA, plan->tmB, kp);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

template <int BLOCK_N, int STAGES, bool HAS_RES, bool BRES = false>
static int launch_tc2(const TcConvPlan* plan, const TcKernelParams& kp, int tiles, cudaStream_t st) {
  using SM = Tc2Smem<BLOCK_N, STAGES, HAS_RES, BRES>;
  static_assert(SM::TOTAL <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
  static bool attr_set = false;
  if (!attr_set) {
    NIB_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BLOCK_N, STAGES, HAS_RES, BRES>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    attr_set = true;
  }
  const int grid = tiles < num_sms() ? tiles : num_sms();
  conv_tc2_kernel<BLOCK_N, STAGES, HAS_RES, BRES><<<grid, TC2_THREADS, SM::TOTAL, st>>>(plan->tmA, plan->tmB, plan->tmOut,
                                                                                         plan->tmRes, kp);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

static int tc_dispatch(const TcConvPlan* plan, const TcKernelParams& kp, int tiles, cudaStream_t st) {
  if (plan->v2 && !kp.out_f32) {
    // stage depth is what hides the L2 -> smem latency (Little's law: ~1.5 us x per-SM fill rate); the residual
    // double buffer costs 2 x BLOCK_N/64 x 16 KB, so layers without a residual get the deeper ring.
    const bool res = kp.res != nullptr;
    static const bool no_bres = getenv("NIB_TC_NO_BRES") != nullptr;
    if (plan->block_n == 64 && !res && kp.n_tiles == 1 && kp.num_k_blocks <= TC2_BRES_KBLOCKS && !no_bres)
      return launch_tc2<64, 6, false, true>(plan, kp, tiles, st);   // resident weights: stem, 56x56 64-channel layers
    if (plan->block_n == 64) return res ? launch_tc2<64, 5, true>(plan, kp, tiles, st) : launch_tc2<64, 7, false>(plan, kp, tiles, st);
    if (plan->block_n == 128) return res ? launch_tc2<128, 3, true>(plan, kp, tiles, st) : launch_tc2<128, 5, false>(plan, kp, tiles, st);
    if (plan->block_n == 256 && !res) return launch_tc2<256, 4, false>(plan, kp, tiles, st);
  }
  switch (plan->block_n) {
    case 32:  return launch_tc<32, 4>(plan, kp, tiles, st);
    case 64:  return launch_tc<64, 4>(plan, kp, tiles, st);
    case 128: return launch_tc<128, 3>(plan, kp, tiles, st);
    case 256: return launch_tc<256, 4>(plan, kp, tiles, st);
  }
  set_error("tc_dispatch: bad block_n %d", plan->block_n);
  return NIB_EINVAL;
}

int tc_conv_launch(const TcConvPlan* plan, const ConvParams& p, cudaStream_t st) {
  TcKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.bias = p.bias;
  kp.res = reinterpret_cast<const __nv_bfloat16*>(p.res);
  kp.out = p.out;
  kp.M = p.M;
  kp.Cout = p.Cout;
  kp.out_cstride = p.out_cstride;
  kp.out_coff = p.out_coff;
  kp.res_cstride = p.res_cstride;
  kp.res_coff = p.res_coff;
  kp.relu = p.relu;
  kp.out_f32 = 0;
  kp.num_k_blocks = plan->num_k_blocks;
  kp.cblocks = plan->cblocks;
  kp.S = p.S;
  kp.im2col = plan->im2col;
  kp.P = p.P;
  kp.Q = p.Q;
  kp.stride = p.stride;
  kp.pad = p.pad;
  kp.in_coff = p.in_coff;
  kp.n_tiles = ceil_div(p.Cout, plan->block_n);
  kp.err_flag = plan->err_flag;
  kp.tile_rows = plan->tile_rows;
  kp.a_bytes = plan->tile_rows * TC_BLOCK_K * 2;
  kp.m_tiles = ceil_div(p.M, plan->tile_rows);
  {
    static int pf = -1;
    if (pf < 0) { const char* e = getenv("NIB_TC_PF"); pf = e ? atoi(e) : 2; }
    kp.pf_dist = pf;
  }
  const int tiles = kp.m_tiles * kp.n_tiles;
  return tc_dispatch(plan, kp, tiles, st);
}

}  // namespace nib

// Standalone GEMM hook: C[M,N] (fp32) = A[M,K] (bf16) * B[N,K]^T (bf16).  Exercises exactly the
// descriptors, swizzle and pipeline of the convolution kernel (tiled A map).
extern "C" int nib_tc_gemm_bf16(const void* d_A, const void* d_B, float* d_C, int M, int N, int K, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(d_A && d_B && d_C, "nib_tc_gemm_bf16: null pointer");
  NIB_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0 && N % 32 == 0, "nib_tc_gemm_bf16: need K%%64==0, N%%32==0");
  int rc = load_driver_fns();
  if (rc != NIB_OK) return rc;
  if (!g_err_flag) {
    NIB_CUDA(cudaMalloc(&g_err_flag, sizeof(unsigned int)));
    NIB_CUDA(cudaMemset(g_err_flag, 0, sizeof(unsigned int)));
  }
  TcConvPlan plan;
  memset(&plan, 0, sizeof(plan));
  plan.err_flag = g_err_flag;
  plan.block_n = pick_block_n(N);
  plan.cblocks = K / TC_BLOCK_K;
  plan.num_k_blocks = plan.cblocks;
  plan.im2col = 0;
  rc = encode_2d_bf16(&plan.tmB, d_B, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, TC_BLOCK_K, plan.block_n);
  if (rc != NIB_OK) return rc;
  rc = encode_2d_bf16(&plan.tmA, d_A, (uint64_t)K, (uint64_t)M, (uint64_t)K * 2, TC_BLOCK_K, TC_BLOCK_M);
  if (rc != NIB_OK) return rc;
  TcKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  kp.out = d_C;
  kp.M = M;
  kp.Cout = N;
  kp.out_cstride = N;
  kp.out_f32 = 1;
  kp.num_k_blocks = plan.num_k_blocks;
  kp.cblocks = plan.cblocks;
  kp.S = 1;
  kp.n_tiles = ceil_div(N, plan.block_n);
  kp.tile_rows = TC_BLOCK_M;
  kp.a_bytes = TC_BLOCK_M * TC_BLOCK_K * 2;
  kp.m_tiles = ceil_div(M, TC_BLOCK_M);
  kp.err_flag = g_err_flag;
  const int tiles = ceil_div(M, TC_BLOCK_M) * kp.n_tiles;
  return tc_dispatch(&plan, kp, tiles, (cudaStream_t)stream);
}
