// conv_x3.cu — fp32-grade implicit-GEMM convolution on the tensor cores: split-bf16 ("bf16x3") products, fp32 accumulate.
//
// Role: the re-score path of the bf16 tie policy (score.cu) and the fast variant of the fp32 parity mode.  Activations and
// outputs are fp32 NHWC exactly as in the CUDA-core mode (conv_simt.cu); each fp32 operand is split on the fly into
//     x = hi + lo,   hi = bf16(x),   lo = bf16(x - hi)            (16 mantissa bits between them)
// and the product is accumulated in fp32 as  hi_a*hi_w + lo_a*hi_w + hi_a*lo_w  (the lo*lo term, ~2^-18 relative, is
// dropped).  Three mma.sync.m16n8k16.bf16 per output tile and k-step instead of one: ~1e-5 relative to an fp32 FMA chain,
// at several times the CUDA-core kernel's rate.  Weights arrive pre-split (two bf16 KRSC arrays, net.cu).
//
//   out[m, co] = act( sum_{r,s,c} pre(in[n, p*st-pad+r, q*st-pad+s, c]) * w[co, r, s, c] + bias[co] + residual[m, co] )
//
// CTA: 256 threads, tile 128 pixels x BN couts x 32 k; warp tile (128 / WARPS_M) x 32.  Operands are staged in shared
// memory as bf16 rows of 32 k padded to 40 elements (80 B: the 8 rows of an ldmatrix phase fall into distinct banks).
// Weights stream through a 4-stage cp.async ring (three slabs in flight); the next activation slab's global loads are
// issued before the MMAs of the current one and split into hi / lo on their way into shared memory.  Needs Cin % 16 == 0 (a 16-element k run
// stays inside one filter tap); the stems (Cin = 1 / 3) stay on conv_simt_kernel.
// Same reference call sites as conv_simt.cu: `model(masked_img_tensor)`, generate_gp_training_data_imagenet.py:246.
#include "common.cuh"
#include "layers.cuh"

namespace nib {

static constexpr int X3_BM = 128, X3_BK = 32, X3_LD = 40;
static constexpr int X3_STAGES = 4;   // weight slabs in flight (cp.async ring): the weights of a 100-layer network do not
                                      // stay in L2 between forwards, and at the tie policy's batch sizes (tens of images,
                                      // < 1 wave of CTAs) one DRAM round trip per slab was the whole kernel time

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (hi, lo) bf16 pair of two consecutive fp32 values, packed low element first
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

template <int BN>
__global__ void __launch_bounds__(256, 2)
conv_x3_kernel(ConvParams p, const __nv_bfloat16* __restrict__ w_hi, const __nv_bfloat16* __restrict__ w_lo) {
  constexpr int WARPS_M = BN == 128 ? 2 : 4;
  constexpr int WM = X3_BM / WARPS_M;          // 64 or 32 rows per warp
  constexpr int MT = WM / 16;                  // m16 tiles per warp
  constexpr int NT = 4;                        // n8 tiles per warp (32 columns)
  constexpr int B_ELEMS = BN * X3_BK / 256;    // weight elements per thread, slab and array (16 or 8)
  constexpr int B_TPR = X3_BK / B_ELEMS;       // threads per weight row (2 or 4)
  constexpr int A_ELEMS = X3_BM * X3_LD, B_STAGE = BN * X3_LD;   // elements per array
  extern __shared__ __align__(16) __nv_bfloat16 x3_smem[];
  __nv_bfloat16* Ah = x3_smem;
  __nv_bfloat16* Al = x3_smem + A_ELEMS;
  __nv_bfloat16* Bring = x3_smem + 2 * A_ELEMS;                   // [stage][hi | lo][BN * X3_LD]

  const float* __restrict__ in = reinterpret_cast<const float*>(p.in);
  float* __restrict__ out = reinterpret_cast<float*>(p.out);
  const float* __restrict__ res = reinterpret_cast<const float*>(p.res);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * X3_BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.R * p.S * p.Cin;
  if (p.dyn_n != nullptr) {   // device-side live batch (tie policy): whole blocks beyond it leave before the first barrier
    const long long live = (long long)max(*p.dyn_n, 0) * p.P * p.Q;
    if (live < p.M) p.M = (int)live;
    if (m0 >= p.M) return;
  }

  // A-load role: pixel tid % 128, 16 consecutive k starting at (tid / 128) * 16 of the slab
  const int a_pix = tid & (X3_BM - 1);
  const int a_k0 = (tid >> 7) * 16;
  const int am = m0 + a_pix;
  const bool a_valid = am < p.M;
  int an = 0, ap = 0, aq = 0;
  if (a_valid) {
    an = am / (p.P * p.Q);
    const int rem = am - an * p.P * p.Q;
    ap = rem / p.Q;
    aq = rem - ap * p.Q;
  }
  const int ih0 = ap * p.stride - p.pad, iw0 = aq * p.stride - p.pad;
  const int Hp = p.Hin + 2 * p.in_halo, Wp = p.Win + 2 * p.in_halo;
  const float* in_img = in + (size_t)an * Hp * Wp * p.in_cstride + p.in_coff;

  // B-load role: weight row tid / B_TPR, B_ELEMS consecutive k
  const int b_row = tid / B_TPR;
  const int b_co = n0 + b_row;
  const int b_k0 = (tid % B_TPR) * B_ELEMS;

  const int wm = warp % WARPS_M, wn = warp / WARPS_M;
  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

  float av[16];
  const uint32_t bring_s = (uint32_t)__cvta_generic_to_shared(Bring);
  // this thread's share of one weight slab (hi and lo), asynchronously into ring stage `stage`; zero-filled past Cout / K
  auto issue_b = [&](int stage, int kb) {
#pragma unroll
    for (int q = 0; q < B_ELEMS / 8; ++q) {
      const int kk = kb + b_k0 + 8 * q;
      const bool ok = b_co < p.Cout && kk < K;
      const size_t off = ok ? (size_t)b_co * K + kk : 0;
      const uint32_t dst = bring_s + (uint32_t)((stage * 2 * B_STAGE + b_row * X3_LD + b_k0 + 8 * q) * 2);
      cp_async16(dst, w_hi + off, ok ? 16 : 0);
      cp_async16(dst + (uint32_t)(B_STAGE * 2), w_lo + off, ok ? 16 : 0);
    }
  };
  auto load_slab = [&](int kb) {
#pragma unroll
    for (int j = 0; j < 16; ++j) av[j] = 0.f;
    const int k = kb + a_k0;
    if (a_valid && k < K) {
      const int tap = k / p.Cin, c = k - tap * p.Cin;
      const int r = tap / p.S, s = tap - r * p.S;
      const int ih = ih0 + r, iw = iw0 + s;
      if (ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win) {
        const float* src = in_img + ((size_t)(ih + p.in_halo) * Wp + (iw + p.in_halo)) * p.in_cstride + c;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = *reinterpret_cast<const float4*>(src + 4 * q);
          av[4 * q] = f.x; av[4 * q + 1] = f.y; av[4 * q + 2] = f.z; av[4 * q + 3] = f.w;
        }
        if (p.pre_scale != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j) av[j] = fmaxf(fmaf(av[j], p.pre_scale[c + j], p.pre_shift[c + j]), 0.f);
        }
      }
    }
  };

  const uint32_t ah_s = (uint32_t)__cvta_generic_to_shared(Ah), al_s = (uint32_t)__cvta_generic_to_shared(Al);
  // ldmatrix lane addresses (bytes) inside a 16 x 16 A tile / a 16(n) x 16(k) pair of B tiles
  const uint32_t a_lane = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * X3_LD + (lane >> 4) * 8) * 2u;
  const uint32_t b_lane = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * X3_LD + ((lane >> 3) & 1) * 8) * 2u;

#pragma unroll
  for (int s2 = 0; s2 < X3_STAGES - 1; ++s2) {
    if (s2 * X3_BK < K) issue_b(s2, s2 * X3_BK);
    cp_async_commit();
  }
  load_slab(0);
  int stage = 0;
  for (int kb = 0; kb < K; kb += X3_BK) {
    cp_async_wait<X3_STAGES - 2>();   // this thread's part of the current weight slab has landed ...
    __syncthreads();                  // ... and everyone's; the previous slab (activations and ring stage) has been consumed
    {
      uint32_t h[8], l[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split2(av[2 * j], av[2 * j + 1], h[j], l[j]);
      uint4* dh = reinterpret_cast<uint4*>(Ah + a_pix * X3_LD + a_k0);
      uint4* dl = reinterpret_cast<uint4*>(Al + a_pix * X3_LD + a_k0);
      dh[0] = make_uint4(h[0], h[1], h[2], h[3]); dh[1] = make_uint4(h[4], h[5], h[6], h[7]);
      dl[0] = make_uint4(l[0], l[1], l[2], l[3]); dl[1] = make_uint4(l[4], l[5], l[6], l[7]);
      // refill the ring stage the previous iteration multiplied from
      const int kn = kb + (X3_STAGES - 1) * X3_BK;
      if (kn < K) issue_b((stage + X3_STAGES - 1) % X3_STAGES, kn);
      cp_async_commit();
    }
    __syncthreads();
    if (kb + X3_BK < K) load_slab(kb + X3_BK);   // in flight while this slab is multiplied
    const uint32_t bh_s = bring_s + (uint32_t)(stage * 2 * B_STAGE * 2), bl_s = bh_s + (uint32_t)(B_STAGE * 2);
#pragma unroll
    for (int ks = 0; ks < X3_BK; ks += 16) {
      uint32_t fbh[NT][2], fbl[NT][2];
#pragma unroll
      for (int np = 0; np < NT / 2; ++np) {
        const uint32_t off = (uint32_t)((wn * 32 + np * 16) * X3_LD + ks) * 2u + b_lane;
        uint32_t r[4];
        ldsm_x4(r, bh_s + off);
        fbh[2 * np][0] = r[0]; fbh[2 * np][1] = r[1]; fbh[2 * np + 1][0] = r[2]; fbh[2 * np + 1][1] = r[3];
        ldsm_x4(r, bl_s + off);
        fbl[2 * np][0] = r[0]; fbl[2 * np][1] = r[1]; fbl[2 * np + 1][0] = r[2]; fbl[2 * np + 1][1] = r[3];
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t off = (uint32_t)((wm * WM + mt * 16) * X3_LD + ks) * 2u + a_lane;
        uint32_t fah[4], fal[4];
        ldsm_x4(fah, ah_s + off);
        ldsm_x4(fal, al_s + off);
        // small terms first; each pass walks the four accumulators so dependent MMAs are four issues apart
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[mt][nt], fal, fbh[nt][0], fbh[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[mt][nt], fah, fbl[nt][0], fbl[nt][1]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[mt][nt], fah, fbh[nt][0], fbh[nt][1]);
      }
    }
    stage = (stage + 1) % X3_STAGES;
  }
  cp_async_wait<0>();

  // ---- epilogue: accumulator fragment (row = lane / 4 (+8), columns 2 * (lane % 4) + {0, 1}) ----
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm * WM + mt * 16 + (lane >> 2) + half * 8;
      if (m >= p.M) continue;
      const size_t out_row = (size_t)m * p.out_cstride + p.out_coff;
      const size_t res_row = (size_t)m * p.res_cstride + p.res_coff;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = n0 + wn * 32 + nt * 8 + (lane & 3) * 2 + e;
          if (co >= p.Cout) continue;
          float v = acc[mt][nt][half * 2 + e];
          if (p.bias != nullptr) v += p.bias[co];
          if (res != nullptr && co < p.res_C) v += res[res_row + co];
          if (p.relu) v = fmaxf(v, 0.f);
          out[out_row + co] = v;
        }
    }
}

bool conv_x3_supported(const ConvParams& p) {
  return p.Cin % 16 == 0 && p.in_cstride % 4 == 0 && p.in_coff % 4 == 0 && p.out_halo == 0;
}

int launch_conv_x3(const ConvParams& p, const void* w_hi, const void* w_lo, cudaStream_t st) {
  const __nv_bfloat16* wh = reinterpret_cast<const __nv_bfloat16*>(w_hi);
  const __nv_bfloat16* wl = reinterpret_cast<const __nv_bfloat16*>(w_lo);
  constexpr int smem128 = (2 * X3_BM * X3_LD + X3_STAGES * 2 * 128 * X3_LD) * 2;
  constexpr int smem64 = (2 * X3_BM * X3_LD + X3_STAGES * 2 * 64 * X3_LD) * 2;
  static bool attr = false;
  if (!attr) {
    NIB_CUDA(cudaFuncSetAttribute(conv_x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem128));
    NIB_CUDA(cudaFuncSetAttribute(conv_x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64));
    attr = true;
  }
  if (p.Cout >= 128) {
    dim3 grid(ceil_div(p.M, X3_BM), ceil_div(p.Cout, 128));
    conv_x3_kernel<128><<<grid, 256, smem128, st>>>(p, wh, wl);
  } else {
    dim3 grid(ceil_div(p.M, X3_BM), ceil_div(p.Cout, 64));
    conv_x3_kernel<64><<<grid, 256, smem64, st>>>(p, wh, wl);
  }
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- fully connected layer of the bf16 networks on the tensor cores ---------------------------------------------------
// logits[n][k] = sum_c feat[n][c] * W[k][c] + b[k] with bf16 features (exact operands) and the fp32 weights split as above:
// two mma.sync passes (feat * W_lo, feat * W_hi), fp32 accumulate — the fp32 weights' 16 leading mantissa bits, so top-1 is
// still decided at fp32 grade (the CUDA-core SGEMM this replaces ran at 10 TFLOP/s: 155 us of a 384-image forward).
// CTA: 128 threads, 32 samples x 64 classes, 64-deep slabs through a 3-stage cp.async ring; warp tile 32 x 16.
static constexpr int FCX_BM = 32, FCX_BN = 64, FCX_BK = 64, FCX_LD = 72, FCX_STAGES = 3;
static constexpr int FCX_STAGE_ELEMS = (FCX_BM + 2 * FCX_BN) * FCX_LD;
static constexpr int FCX_SMEM = FCX_STAGES * FCX_STAGE_ELEMS * 2;

__global__ void __launch_bounds__(128)
fc_x2_kernel(const __nv_bfloat16* __restrict__ feat, int feat_stride, const __nv_bfloat16* __restrict__ w_hi,
             const __nv_bfloat16* __restrict__ w_lo, const float* __restrict__ bias, int N, int Cin, int Cout,
             float* __restrict__ logits, const int* __restrict__ dyn_n) {
  extern __shared__ __align__(16) __nv_bfloat16 fcx_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.y * FCX_BM, k0 = blockIdx.x * FCX_BN;
  if (dyn_n != nullptr) N = min(max(*dyn_n, 0), N);
  if (n0 >= N) return;
  const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(fcx_smem);
  // loader roles per slab: 16-byte pieces; A has 32 rows x 8 pieces, each B array 64 rows x 8 pieces
  auto issue = [&](int stage, int c0) {
    const uint32_t base = smem_s + (uint32_t)(stage * FCX_STAGE_ELEMS * 2);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + q * 128, row = idx >> 3, pc = idx & 7;
      const bool ok = n0 + row < N;
      cp_async16(base + (uint32_t)((row * FCX_LD + pc * 8) * 2), feat + (ok ? (size_t)(n0 + row) * feat_stride + c0 + pc * 8 : 0),
                 ok ? 16 : 0);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 128, row = idx >> 3, pc = idx & 7;
      const bool ok = k0 + row < Cout;
      const size_t off = ok ? (size_t)(k0 + row) * Cin + c0 + pc * 8 : 0;
      const uint32_t dst = base + (uint32_t)(((FCX_BM + row) * FCX_LD + pc * 8) * 2);
      cp_async16(dst, w_hi + off, ok ? 16 : 0);
      cp_async16(dst + (uint32_t)(FCX_BN * FCX_LD * 2), w_lo + off, ok ? 16 : 0);
    }
  };
  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;
  const uint32_t a_lane = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * FCX_LD + (lane >> 4) * 8) * 2u;
  const uint32_t b_lane = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * FCX_LD + ((lane >> 3) & 1) * 8) * 2u;
  const int nslab = Cin / FCX_BK;
#pragma unroll
  for (int s2 = 0; s2 < FCX_STAGES - 1; ++s2) {
    if (s2 < nslab) issue(s2, s2 * FCX_BK);
    cp_async_commit();
  }
  int stage = 0;
  for (int sl = 0; sl < nslab; ++sl) {
    cp_async_wait<FCX_STAGES - 2>();
    __syncthreads();                         // slab sl has landed for everyone; slab sl - 1's stage is free again
    if (sl + FCX_STAGES - 1 < nslab) issue((stage + FCX_STAGES - 1) % FCX_STAGES, (sl + FCX_STAGES - 1) * FCX_BK);
    cp_async_commit();
    const uint32_t a_s = smem_s + (uint32_t)(stage * FCX_STAGE_ELEMS * 2);
    const uint32_t bh_s = a_s + (uint32_t)(FCX_BM * FCX_LD * 2), bl_s = bh_s + (uint32_t)(FCX_BN * FCX_LD * 2);
#pragma unroll
    for (int ks = 0; ks < FCX_BK; ks += 16) {
      uint32_t fbh[4], fbl[4], fa[2][4];
      const uint32_t boff = (uint32_t)((warp * 16) * FCX_LD + ks) * 2u + b_lane;
      ldsm_x4(fbh, bh_s + boff);
      ldsm_x4(fbl, bl_s + boff);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) ldsm_x4(fa[mt], a_s + (uint32_t)((mt * 16) * FCX_LD + ks) * 2u + a_lane);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_bf16(acc[mt][nt], fa[mt], fbl[2 * nt], fbl[2 * nt + 1]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_bf16(acc[mt][nt], fa[mt], fbh[2 * nt], fbh[2 * nt + 1]);
    }
    stage = (stage + 1) % FCX_STAGES;
  }
  cp_async_wait<0>();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + mt * 16 + (lane >> 2) + half * 8;
      if (n >= N) continue;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = k0 + warp * 16 + nt * 8 + (lane & 3) * 2 + e;
          if (k < Cout) logits[(size_t)n * Cout + k] = acc[mt][nt][half * 2 + e] + (bias ? bias[k] : 0.f);
        }
    }
}

bool fc_x2_supported(int Cin, int feat_stride) { return Cin % FCX_BK == 0 && feat_stride % 8 == 0; }

int launch_fc_x2(const void* feat, int feat_stride, const void* w_hi, const void* w_lo, const float* b, int N, int Cin,
                 int Cout, float* logits, const int* dyn_n, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    NIB_CUDA(cudaFuncSetAttribute(fc_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FCX_SMEM));
    attr = true;
  }
  dim3 grid(ceil_div(Cout, FCX_BN), ceil_div(N, FCX_BM));
  fc_x2_kernel<<<grid, 128, FCX_SMEM, st>>>((const __nv_bfloat16*)feat, feat_stride, (const __nv_bfloat16*)w_hi,
                                             (const __nv_bfloat16*)w_lo, b, N, Cin, Cout, logits, dyn_n);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

}  // namespace nib
