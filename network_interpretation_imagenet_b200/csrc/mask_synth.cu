// mask_synth.cu — stage 1: superpixel mask synthesis, blend and normalise in one pass.
//
// Reference arithmetic (bit-exact targets):
//   KEEP_MUL       generate_gp_training_data_imagenet.py:234-240
//                  bayesian_active_learning_imagenet.py:182-187
//   REMOVE_MINMAX  generate_gp_training_data_mnist.py:218-242, generate_gp_training_data_cifar.py:310-321,
//                  utils.py:92-94
//
// HBM-bound: the only per-mask traffic is the output (C*H*W*sizeof(out) bytes per mask).  One CTA
// keeps a strip of pixels (labels + C channel values) in registers and loops over a slice of the
// masks, so the image and label map are read from L2 once per CTA, never once per mask.
// Stores are 128-bit and streaming (st.global.cs): the masked batch is consumed by the next
// kernel from HBM/L2, not re-read here.
#include "common.cuh"
#include <math.h>

namespace nib {

static constexpr int kMaxC = 4;

__device__ __forceinline__ void st_cs_v4(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cs_v2(void* p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_cs_u32(void* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Per-mask (min, max-after-subtract) of org*mask for REMOVE_MINMAX, from the per-segment table.
// fl(v*255) is monotone in v, so the extreme of a kept segment is fl(seg_extreme*255); removed
// pixels contribute v*0 = 0 (v >= 0 after the a1 rescale).  max(m - mn) = fl(max(m) - mn) by
// monotonicity of fl(x - c).
// Computed inside mask_synth_kernel for the CTA's own slice of masks (shared memory): no scratch buffer is shared between
// launches, so forwards running on different streams cannot see each other's statistics.
__device__ __forceinline__ float2 mask_stats(const float* __restrict__ seg_minmax, const uint64_t* __restrict__ sel,
                                             int words, int n, int S) {
  float mn = INFINITY, mx = -INFINITY;
  bool any_removed = false, any = false;
  for (int s = 0; s < S; ++s) {
    float smin = seg_minmax[2 * s], smax = seg_minmax[2 * s + 1];
    if (!(smin <= smax)) continue;  // empty segment (label absent from the map)
    any = true;
    bool removed = (sel[(size_t)n * words + (s >> 6)] >> (s & 63)) & 1ull;
    if (removed) {
      any_removed = true;
    } else {
      mn = fminf(mn, __fmul_rn(smin, 255.0f));
      mx = fmaxf(mx, __fmul_rn(smax, 255.0f));
    }
  }
  if (any_removed) {
    mn = fminf(mn, 0.0f);
    mx = fmaxf(mx, 0.0f);
  }
  if (!any) { mn = 0.f; mx = 0.f; }
  return make_float2(mn, __fsub_rn(mx, mn));
}

template <int MODE>
__device__ __forceinline__ float blend(float x, bool bit, float mn, float mxs) {
  if (MODE == NIB_MASK_KEEP_MUL) {
    // fp32 * uint8 -> fp32 in numpy: the product with an exact 0.0f/1.0f keeps the sign of x on zeros.
    return __fmul_rn(x, bit ? 1.0f : 0.0f);
  } else {
    float t = __fmul_rn(x, bit ? 0.0f : 255.0f);
    t = __fsub_rn(t, mn);
    t = __fdiv_rn(t, mxs);
    t = __fmul_rn(t, 255.0f);
    return __fmul_rn(t, (float)(1.0 / 255.0));  // np.multiply(f32, 1.0/255.0): weak scalar -> f32
  }
}

// VEC pixels per thread along W (VEC=4 needs W % 4 == 0).
template <typename OutT, int LAYOUT, int MODE, int VEC, typename LabT>
__global__ void __launch_bounds__(256)
mask_synth_kernel(const float* __restrict__ img, const LabT* __restrict__ labels,
                  const uint64_t* __restrict__ sel, int words, int N, int C, int H, int W,
                  const float* __restrict__ seg_minmax, int S, OutT* __restrict__ out, int c_stride, int pad_h,
                  int pad_w, uint8_t* __restrict__ pixel_mask, int masks_per_cta, const int* __restrict__ dyn_n) {
  if (dyn_n != nullptr) N = min(N, max(*dyn_n, 0));   // device-side live mask count (tie policy re-score batch)
  constexpr int kStatChunk = 256;   // == blockDim.x: one thread computes the (min, max - min) of one mask of the chunk
  __shared__ float2 s_stats[MODE == NIB_MASK_REMOVE_MINMAX ? kStatChunk : 1];
  const int HW = H * W;
  const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const bool active = p0 < HW;
  if (MODE != NIB_MASK_REMOVE_MINMAX && !active) return;   // (the REMOVE_MINMAX path has block-wide barriers below)
  const int h = p0 / W, w = p0 - h * W;

  int lab[VEC];
  float x[kMaxC][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) lab[v] = (p0 + v < HW) ? (int)labels[p0 + v] : 0;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c)
#pragma unroll
    for (int v = 0; v < VEC; ++v) x[c][v] = (c < C && p0 + v < HW) ? img[(size_t)c * HW + p0 + v] : 0.f;

  const int n0 = blockIdx.y * masks_per_cta;
  const int n1 = min(N, n0 + masks_per_cta);
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;

  for (int n = n0; n < n1; ++n) {
    if (MODE == NIB_MASK_REMOVE_MINMAX) {
      const int k = (n - n0) % kStatChunk;
      if (k == 0) {
        __syncthreads();   // the previous chunk's readers are done
        const int nn = n + (int)threadIdx.x;
        if (nn < n1) s_stats[threadIdx.x] = mask_stats(seg_minmax, sel, words, nn, S);
        __syncthreads();
      }
      if (!active) continue;
    }
    bool bit[VEC];
    if (words == 1) {
      const uint64_t z = __ldg(sel + n);
#pragma unroll
      for (int v = 0; v < VEC; ++v) bit[v] = (z >> lab[v]) & 1ull;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        bit[v] = (__ldg(sel + (size_t)n * words + (lab[v] >> 6)) >> (lab[v] & 63)) & 1ull;
    }
    float mn = 0.f, mxs = 1.f;
    if (MODE == NIB_MASK_REMOVE_MINMAX) {
      const float2 st = s_stats[(n - n0) % kStatChunk];
      mn = st.x;
      mxs = st.y;
    }
    float y[kMaxC][VEC];
#pragma unroll
    for (int c = 0; c < kMaxC; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) y[c][v] = blend<MODE>(x[c][v], bit[v], mn, mxs);

    if (pixel_mask != nullptr) {
      uint8_t mv[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        mv[v] = (MODE == NIB_MASK_KEEP_MUL) ? (bit[v] ? 1 : 0) : (bit[v] ? 0 : 255);
      uint8_t* pm = pixel_mask + (size_t)n * HW + p0;
      if (VEC == 4) {
        st_cs_u32(pm, (uint32_t)mv[0] | ((uint32_t)mv[1] << 8) | ((uint32_t)mv[2] << 16) |
                          ((uint32_t)mv[3 % VEC] << 24));
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (p0 + v < HW) pm[v] = mv[v];
      }
    }

    if (LAYOUT == NIB_NCHW) {
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) {
        if (c >= C) break;
        OutT* o = out + ((size_t)n * C + c) * HW + p0;
        if (VEC == 4) {
          if (sizeof(OutT) == 4) {
            st_cs_v4(o, make_uint4(__float_as_uint(y[c][0]), __float_as_uint(y[c][1 % VEC]),
                                   __float_as_uint(y[c][2 % VEC]), __float_as_uint(y[c][3 % VEC])));
          } else {
            st_cs_v2(o, make_uint2(pack_bf16x2(y[c][0], y[c][1 % VEC]),
                                   pack_bf16x2(y[c][2 % VEC], y[c][3 % VEC])));
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            if (p0 + v < HW) Elem<OutT>::st(o + v, y[c][v]);
        }
      }
    } else {  // NHWC with channel padding and halo offset
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (p0 + v >= HW) break;
        OutT* o = out + (((size_t)n * Hp + (h + pad_h)) * Wp + (w + v + pad_w)) * c_stride;
        if (sizeof(OutT) == 2 && c_stride == 8) {
          st_cs_v4(o, make_uint4(pack_bf16x2(y[0][v], y[1][v]), pack_bf16x2(y[2][v], y[3][v]), 0u, 0u));
        } else if (sizeof(OutT) == 2 && c_stride == 4) {
          st_cs_v2(o, make_uint2(pack_bf16x2(y[0][v], y[1][v]), pack_bf16x2(y[2][v], y[3][v])));
        } else if (sizeof(OutT) == 4 && c_stride == 4) {
          st_cs_v4(o, make_uint4(__float_as_uint(y[0][v]), __float_as_uint(y[1][v]),
                                 __float_as_uint(y[2][v]), __float_as_uint(y[3][v])));
        } else {
          for (int c = 0; c < c_stride; ++c) {
            float val = 0.f;
#pragma unroll
            for (int cc = 0; cc < kMaxC; ++cc)
              if (cc == c) val = y[cc][v];
            Elem<OutT>::st(o + c, (c < C) ? val : 0.f);
          }
        }
      }
    }
  }
}

// ---- the classifier's own input layout: NHWC, 4-channel bf16 pixels (8 B) inside a zero halo ---------------------------
// The tcgen05 stem reads [N][H + 2p][W + 2p][4] bf16.  Writing it pixel by pixel means 8-byte stores (a quarter of each
// 32 B sector per instruction); here a thread owns TWO 16-byte units = four consecutive padded pixels of the interior rows
// (row pitch (W + 2p) * 8 B must be a multiple of 16), halo columns included (they are written as +0, which is what they
// must hold anyway), so every store is a full, aligned 128-bit store.  The top and bottom halo rows are not touched: the
// buffer owner zeroes them once (net.cu) or halo_zero_kernel does.
template <int MODE, typename LabT>
__global__ void __launch_bounds__(256)
mask_synth_nhwc4_kernel(const float* __restrict__ img, const LabT* __restrict__ labels, const uint64_t* __restrict__ sel,
                        int words, int N, int C, int H, int W, const float* __restrict__ seg_minmax, int S,
                        __nv_bfloat16* __restrict__ out, int pad, uint8_t* __restrict__ pixel_mask, int masks_per_cta,
                        const int* __restrict__ dyn_n) {
  if (dyn_n != nullptr) N = min(N, max(*dyn_n, 0));
  constexpr int kStatChunk = 256;
  __shared__ float2 s_stats[MODE == NIB_MASK_REMOVE_MINMAX ? kStatChunk : 1];
  const int Wp = W + 2 * pad, Hp = H + 2 * pad;
  const int HW = H * W;
  const long long total_px = (long long)H * Wp;                 // padded pixels of the interior rows
  const long long q0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const bool active = q0 < total_px;
  if (MODE != NIB_MASK_REMOVE_MINMAX && !active) return;
  int lab[4];
  bool inside[4];
  int pix[4];
  float x[kMaxC][4];
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const long long q = q0 + v;
    const int row = (int)(q / Wp), px = (int)(q - (long long)row * Wp);
    inside[v] = active && q < total_px && px >= pad && px < pad + W;
    pix[v] = inside[v] ? row * W + (px - pad) : 0;
    lab[v] = inside[v] ? (int)labels[pix[v]] : 0;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) x[c][v] = (c < C && inside[v]) ? img[(size_t)c * HW + pix[v]] : 0.f;
  }
  const int n0 = blockIdx.y * masks_per_cta;
  const int n1 = min(N, n0 + masks_per_cta);
  for (int n = n0; n < n1; ++n) {
    if (MODE == NIB_MASK_REMOVE_MINMAX) {
      const int k = (n - n0) % kStatChunk;
      if (k == 0) {
        __syncthreads();
        const int nn = n + (int)threadIdx.x;
        if (nn < n1) s_stats[threadIdx.x] = mask_stats(seg_minmax, sel, words, nn, S);
        __syncthreads();
      }
      if (!active) continue;
    }
    float mn = 0.f, mxs = 1.f;
    if (MODE == NIB_MASK_REMOVE_MINMAX) {
      const float2 st = s_stats[(n - n0) % kStatChunk];
      mn = st.x;
      mxs = st.y;
    }
    uint32_t o[8];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const bool bit = (__ldg(sel + (size_t)n * words + (lab[v] >> 6)) >> (lab[v] & 63)) & 1ull;
      float y[kMaxC];
#pragma unroll
      for (int c = 0; c < kMaxC; ++c) y[c] = (inside[v] && c < C) ? blend<MODE>(x[c][v], bit, mn, mxs) : 0.f;
      o[2 * v] = pack_bf16x2(y[0], y[1]);
      o[2 * v + 1] = pack_bf16x2(y[2], y[3]);
      if (pixel_mask != nullptr && inside[v])
        pixel_mask[(size_t)n * HW + pix[v]] = (MODE == NIB_MASK_KEEP_MUL) ? (bit ? 1 : 0) : (bit ? 0 : 255);
    }
    __nv_bfloat16* dst = out + ((size_t)n * Hp + pad) * Wp * 4 + q0 * 4;
    st_cs_v4(dst, make_uint4(o[0], o[1], o[2], o[3]));
    if (q0 + 2 < total_px) st_cs_v4(dst + 8, make_uint4(o[4], o[5], o[6], o[7]));
  }
}

// ---- display images of the PNG side channel ("next" row 3) ------------------------------------------------------
// KEEP_MUL (bayesian_active_learning_imagenet.py:199-205, generate_gp_training_data_imagenet.py:250-256 commented):
//   show = masked.transpose(1,2,0); show -= show.min(); show /= show.max(); show *= 255; show.astype(uint8)   (fp32, truncation)
// REMOVE_MINMAX (generate_gp_training_data_mnist.py:225-236 `pic`, cifar :316-320): the [0,255] float image the reference
//   hands to cv2.imwrite, which rounds to nearest when it converts to 8 bits.
// Per-mask min / max come from the per-segment table like the classifier input's statistics.  Output [N][H][W][C] uint8.
template <int MODE, typename LabT>
__global__ void __launch_bounds__(256)
mask_display_kernel(const float* __restrict__ img, const LabT* __restrict__ labels, const uint64_t* __restrict__ sel,
                    int words, int N, int C, int H, int W, const float* __restrict__ seg_minmax, int S,
                    uint8_t* __restrict__ out, int masks_per_cta) {
  constexpr int kStatChunk = 256;
  __shared__ float2 s_stats[kStatChunk];
  const int HW = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < HW;
  const int lab = active ? (int)labels[p] : 0;
  float x[kMaxC];
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) x[c] = (c < C && active) ? img[(size_t)c * HW + p] : 0.f;
  const int n0 = blockIdx.y * masks_per_cta;
  const int n1 = min(N, n0 + masks_per_cta);
  for (int n = n0; n < n1; ++n) {
    const int k = (n - n0) % kStatChunk;
    if (k == 0) {
      __syncthreads();
      const int nn = n + (int)threadIdx.x;
      if (nn < n1) {
        if (MODE == NIB_MASK_REMOVE_MINMAX) {
          s_stats[threadIdx.x] = mask_stats(seg_minmax, sel, words, nn, S);
        } else {   // min / max of x * mask: extremes of the kept segments, and 0 as soon as one present segment is dropped
          float mn = INFINITY, mx = -INFINITY;
          bool any_dropped = false, any = false;
          for (int s2 = 0; s2 < S; ++s2) {
            const float smin = seg_minmax[2 * s2], smax = seg_minmax[2 * s2 + 1];
            if (!(smin <= smax)) continue;
            any = true;
            if ((sel[(size_t)nn * words + (s2 >> 6)] >> (s2 & 63)) & 1ull) { mn = fminf(mn, smin); mx = fmaxf(mx, smax); }
            else any_dropped = true;
          }
          if (any_dropped) { mn = fminf(mn, 0.f); mx = fmaxf(mx, 0.f); }
          if (!any) { mn = 0.f; mx = 0.f; }
          s_stats[threadIdx.x] = make_float2(mn, __fsub_rn(mx, mn));
        }
      }
      __syncthreads();
    }
    if (!active) continue;
    const bool bit = (sel[(size_t)n * words + (lab >> 6)] >> (lab & 63)) & 1ull;
    const float2 st = s_stats[k];
    uint8_t* o = out + ((size_t)n * HW + p) * C;
#pragma unroll
    for (int c = 0; c < kMaxC; ++c) {
      if (c >= C) break;
      float t = (MODE == NIB_MASK_KEEP_MUL) ? __fmul_rn(x[c], bit ? 1.0f : 0.0f) : __fmul_rn(x[c], bit ? 0.0f : 255.0f);
      t = __fsub_rn(t, st.x);
      t = __fdiv_rn(t, st.y);
      t = __fmul_rn(t, 255.0f);
      if (MODE == NIB_MASK_REMOVE_MINMAX) t = rintf(fminf(fmaxf(t, 0.f), 255.f));   // cv2 saturate_cast<uchar>
      o[c] = (t == t) ? (uint8_t)t : (uint8_t)0;
    }
  }
}

// zero the halo ring of an NHWC padded batch
template <typename OutT>
__global__ void halo_zero_kernel(OutT* out, int N, int H, int W, int c_stride, int pad_h, int pad_w) {
  const int Hp = H + 2 * pad_h, Wp = W + 2 * pad_w;
  const int halo = Hp * Wp - H * W;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * halo;
  if (idx >= total) return;
  int n = (int)(idx / halo), r = (int)(idx - (long long)n * halo);
  int hp, wp;
  const int top = pad_h * Wp;
  if (r < top) {
    hp = r / Wp; wp = r - hp * Wp;
  } else if (r < 2 * top) {
    int q = r - top;
    hp = H + pad_h + q / Wp; wp = q % Wp;
  } else {
    int q = r - 2 * top;            // side columns of the H interior rows
    int row = q / (2 * pad_w), col = q - row * (2 * pad_w);
    hp = pad_h + row;
    wp = col < pad_w ? col : (W + pad_w + (col - pad_w));
  }
  OutT* o = out + (((size_t)n * Hp + hp) * Wp + wp) * c_stride;
  for (int c = 0; c < c_stride; ++c) Elem<OutT>::st(o + c, 0.f);
}

// ---- per-segment min/max --------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
  int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int k) {
  return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

__global__ void segmm_init_kernel(int* tab, int S) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S) {
    tab[2 * s] = float_to_ordered(INFINITY);
    tab[2 * s + 1] = float_to_ordered(-INFINITY);
  }
}
template <typename LabT>
__global__ void segmm_accum_kernel(const float* __restrict__ img, const LabT* __restrict__ labels, int C,
                                   int HW, int S, int* tab) {
  extern __shared__ int sm[];  // [2*S]
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    sm[2 * i] = float_to_ordered(INFINITY);
    sm[2 * i + 1] = float_to_ordered(-INFINITY);
  }
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    int lab = (int)labels[p];
    if (lab >= S) continue;
    float mn = INFINITY, mx = -INFINITY;
    for (int c = 0; c < C; ++c) {
      float v = img[(size_t)c * HW + p];
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    atomicMin(&sm[2 * lab], float_to_ordered(mn));
    atomicMax(&sm[2 * lab + 1], float_to_ordered(mx));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) {
    atomicMin(&tab[2 * i], sm[2 * i]);
    atomicMax(&tab[2 * i + 1], sm[2 * i + 1]);
  }
}
__global__ void segmm_decode_kernel(int* tab, int S) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * S) reinterpret_cast<float*>(tab)[i] = ordered_to_float(tab[i]);
}

// ---- a1: per-image min-max rescale to [0,255] ------------------------------------------------
__global__ void prep_minmax_kernel(float* org, int C, int H, int W, uint8_t* u8) {
  __shared__ float s_mn[32], s_mx[32];
  __shared__ float g_mn, g_mx;
  const int total = C * H * W;
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    float v = org[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = INFINITY, b = -INFINITY;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { a = fminf(a, s_mn[i]); b = fmaxf(b, s_mx[i]); }
    g_mn = a;
    g_mx = __fsub_rn(b, a);  // max of (x - min) == fl(max - min)
  }
  __syncthreads();
  const float a = g_mn, b = g_mx;
  const int HW = H * W;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    float v = __fsub_rn(org[i], a);
    v = __fdiv_rn(v, b);
    v = __fmul_rn(v, 255.0f);
    org[i] = v;
    if (u8) {
      int c = i / HW, p = i - c * HW;
      u8[(size_t)p * C + c] = (uint8_t)v;  // astype(np.uint8) truncates
    }
  }
}

// ---- heat map --------------------------------------------------------------------------------
__global__ void heat_weights_kernel(const uint64_t* __restrict__ sel, int words, const float* __restrict__ y,
                                    int N, int S, double* __restrict__ wseg, int* __restrict__ cover) {
  // one block per segment; exact for integer-valued labels (gp_regression.py:82-94 adds ints).  cover[s] = number of
  // masks that select segment s (a pixel is a key of the reference's dict_pixel iff some mask covers it)
  const int s = blockIdx.x;
  double acc = 0.0;
  int cnt = 0;
  for (int n = threadIdx.x; n < N; n += blockDim.x)
    if ((sel[(size_t)n * words + (s >> 6)] >> (s & 63)) & 1ull) { acc += (double)y[n]; ++cnt; }
  __shared__ double sh[32];
  __shared__ int shc[32];
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = acc; shc[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    int c = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { t += sh[i]; c += shc[i]; }
    wseg[s] = t;
    if (cover) cover[s] = c;
  }
}
template <typename LabT>
__global__ void heat_scatter_kernel(const LabT* __restrict__ labels, int HW, int S,
                                    const double* __restrict__ wseg, float* __restrict__ heat) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < HW) {
    int lab = (int)labels[p];
    heat[p] = lab < S ? (float)wseg[lab] : 0.f;
  }
}

template <typename OutT, int LAYOUT, int MODE, typename LabT>
static int launch_mask(const nib_mask_args* a, bool skip_halo, const int* dyn_n, cudaStream_t st) {
  const int HW = a->H * a->W;
  if (LAYOUT == NIB_NHWC && sizeof(OutT) == 2 && a->c_stride == 4 && a->pad_h == a->pad_w && a->pad_w > 0 &&
      ((a->W + 2 * a->pad_w) % 2) == 0) {
    // the network input layout: full 128-bit stores over the padded interior rows
    const long long total_px = (long long)a->H * (a->W + 2 * a->pad_w);
    const int gx = (int)ceil_div_ll(ceil_div_ll(total_px, 4), 256);
    int gy = max(1, min(a->N, ceil_div(num_sms() * 16, gx)));
    const int mpc = ceil_div(a->N, gy);
    gy = ceil_div(a->N, mpc);
    mask_synth_nhwc4_kernel<MODE, LabT><<<dim3(gx, gy), 256, 0, st>>>(
        a->d_img, (const LabT*)a->d_labels, a->d_sel, a->sel_words, a->N, a->C, a->H, a->W, a->d_seg_minmax, a->S,
        (__nv_bfloat16*)a->d_out, a->pad_w, a->d_pixel_mask, mpc, dyn_n);
    NIB_LAUNCH_CHECK();
    if (!skip_halo) {   // only the top / bottom halo rows remain (the side columns were written above); the generic
                        // kernel covers them, re-zeroing the sides as well
      const int Hp = a->H + 2 * a->pad_h, Wp = a->W + 2 * a->pad_w;
      long long total = (long long)a->N * (Hp * Wp - HW);
      halo_zero_kernel<OutT><<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(
          (OutT*)a->d_out, a->N, a->H, a->W, a->c_stride, a->pad_h, a->pad_w);
      NIB_LAUNCH_CHECK();
    }
    return NIB_OK;
  }
  const bool vec4 = (a->W % 4 == 0);
  const int threads = 256;
  const int per = vec4 ? 4 : 1;
  const int gx = ceil_div(ceil_div(HW, per), threads);
  // enough CTAs for >= 4 waves of 148 SMs x 8 resident CTAs when N allows
  int target_ctas = num_sms() * 16;
  int gy = max(1, min(a->N, ceil_div(target_ctas, gx)));
  int mpc = ceil_div(a->N, gy);
  gy = ceil_div(a->N, mpc);
  dim3 grid(gx, gy);
  if (vec4)
    mask_synth_kernel<OutT, LAYOUT, MODE, 4, LabT><<<grid, threads, 0, st>>>(
        a->d_img, (const LabT*)a->d_labels, a->d_sel, a->sel_words, a->N, a->C, a->H, a->W, a->d_seg_minmax, a->S,
        (OutT*)a->d_out, a->c_stride, a->pad_h, a->pad_w, a->d_pixel_mask, mpc, dyn_n);
  else
    mask_synth_kernel<OutT, LAYOUT, MODE, 1, LabT><<<grid, threads, 0, st>>>(
        a->d_img, (const LabT*)a->d_labels, a->d_sel, a->sel_words, a->N, a->C, a->H, a->W, a->d_seg_minmax, a->S,
        (OutT*)a->d_out, a->c_stride, a->pad_h, a->pad_w, a->d_pixel_mask, mpc, dyn_n);
  NIB_LAUNCH_CHECK();
  if (LAYOUT == NIB_NHWC && (a->pad_h > 0 || a->pad_w > 0) && !skip_halo) {
    const int Hp = a->H + 2 * a->pad_h, Wp = a->W + 2 * a->pad_w;
    long long total = (long long)a->N * (Hp * Wp - HW);
    halo_zero_kernel<OutT><<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(
        (OutT*)a->d_out, a->N, a->H, a->W, a->c_stride, a->pad_h, a->pad_w);
    NIB_LAUNCH_CHECK();
  }
  return NIB_OK;
}

template <typename OutT, int LAYOUT, typename LabT>
static int dispatch_mode(const nib_mask_args* a, bool skip_halo, const int* dyn_n, cudaStream_t st) {
  if (a->mode == NIB_MASK_KEEP_MUL) return launch_mask<OutT, LAYOUT, NIB_MASK_KEEP_MUL, LabT>(a, skip_halo, dyn_n, st);
  return launch_mask<OutT, LAYOUT, NIB_MASK_REMOVE_MINMAX, LabT>(a, skip_halo, dyn_n, st);
}
template <typename OutT, typename LabT>
static int dispatch_layout(const nib_mask_args* a, bool skip_halo, const int* dyn_n, cudaStream_t st) {
  if (a->layout == NIB_NCHW) return dispatch_mode<OutT, NIB_NCHW, LabT>(a, skip_halo, dyn_n, st);
  return dispatch_mode<OutT, NIB_NHWC, LabT>(a, skip_halo, dyn_n, st);
}

// skip_halo: the caller guarantees the halo ring of d_out already holds +0 (the network's own input buffer is zeroed
// when it is allocated and nothing else ever writes its halo), so it is not re-written on every forward.
int mask_synth_impl(const nib_mask_args* a, cudaStream_t st, bool skip_halo, const int* dyn_n) {
  NIB_REQUIRE(a != nullptr, "nib_mask_synth: null args");
  if (a->N == 0) return NIB_OK;  // an empty batch is legal (and its tensors have null data pointers)
  NIB_REQUIRE(a->d_img && a->d_labels && a->d_sel && a->d_out, "nib_mask_synth: null device pointer");
  NIB_REQUIRE(a->N >= 0 && a->C >= 1 && a->C <= kMaxC, "nib_mask_synth: C=%d unsupported (1..%d)", a->C, kMaxC);
  NIB_REQUIRE(a->H > 0 && a->W > 0 && a->S > 0, "nib_mask_synth: bad geometry H=%d W=%d S=%d", a->H, a->W, a->S);
  NIB_REQUIRE(a->label_bytes == 1 || a->label_bytes == 2, "nib_mask_synth: label_bytes must be 1 or 2");
  NIB_REQUIRE(a->label_bytes == 2 || a->S <= 256, "nib_mask_synth: S=%d needs uint16 labels", a->S);
  NIB_REQUIRE(a->sel_words * 64 >= a->S, "nib_mask_synth: sel_words=%d too small for S=%d", a->sel_words, a->S);
  NIB_REQUIRE(a->mode == NIB_MASK_KEEP_MUL || a->mode == NIB_MASK_REMOVE_MINMAX, "nib_mask_synth: bad mode %d", a->mode);
  NIB_REQUIRE(a->out_dtype == NIB_F32 || a->out_dtype == NIB_BF16, "nib_mask_synth: bad out_dtype");
  NIB_REQUIRE(a->layout == NIB_NCHW || a->layout == NIB_NHWC, "nib_mask_synth: bad layout");
  if (a->layout == NIB_NHWC)
    NIB_REQUIRE(a->c_stride >= a->C && a->pad_h >= 0 && a->pad_w >= 0, "nib_mask_synth: bad NHWC c_stride/pad");
  if (a->N == 0) return NIB_OK;
  if (a->mode == NIB_MASK_REMOVE_MINMAX)
    NIB_REQUIRE(a->d_seg_minmax != nullptr, "nib_mask_synth: REMOVE_MINMAX needs d_seg_minmax (nib_segment_minmax)");
  if (a->out_dtype == NIB_F32) {
    if (a->label_bytes == 1) return dispatch_layout<float, uint8_t>(a, skip_halo, dyn_n, st);
    return dispatch_layout<float, uint16_t>(a, skip_halo, dyn_n, st);
  } else {
    if (a->label_bytes == 1) return dispatch_layout<__nv_bfloat16, uint8_t>(a, skip_halo, dyn_n, st);
    return dispatch_layout<__nv_bfloat16, uint16_t>(a, skip_halo, dyn_n, st);
  }
}

}  // namespace nib

extern "C" {

int nib_mask_synth(const nib_mask_args* args, void* stream) {
  NIB_DEVICE_OR_FAIL();
  return nib::mask_synth_impl(args, (cudaStream_t)stream, false, nullptr);
}

int nib_segment_minmax(const float* d_img, const void* d_labels, int label_bytes, int C, int H, int W,
                       int S, float* d_seg_minmax, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_img && d_labels && d_seg_minmax, "nib_segment_minmax: null pointer");
  NIB_REQUIRE(S > 0 && S <= 4096 && C > 0 && H > 0 && W > 0, "nib_segment_minmax: bad geometry");
  NIB_REQUIRE(label_bytes == 1 || label_bytes == 2, "nib_segment_minmax: label_bytes must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  int* tab = reinterpret_cast<int*>(d_seg_minmax);
  nib::segmm_init_kernel<<<nib::ceil_div(S, 128), 128, 0, st>>>(tab, S);
  const int HW = H * W;
  int blocks = min(nib::ceil_div(HW, 256), 64);
  size_t smem = sizeof(int) * 2 * S;
  if (label_bytes == 1)
    nib::segmm_accum_kernel<uint8_t><<<blocks, 256, smem, st>>>(d_img, (const uint8_t*)d_labels, C, HW, S, tab);
  else
    nib::segmm_accum_kernel<uint16_t><<<blocks, 256, smem, st>>>(d_img, (const uint16_t*)d_labels, C, HW, S, tab);
  nib::segmm_decode_kernel<<<nib::ceil_div(2 * S, 128), 128, 0, st>>>(tab, S);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_prep_minmax_u8(float* d_org, int C, int H, int W, uint8_t* d_u8, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_org && C > 0 && H > 0 && W > 0, "nib_prep_minmax_u8: bad arguments");
  nib::prep_minmax_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_org, C, H, W, d_u8);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_mask_display_u8(const nib_mask_args* a, uint8_t* d_out, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(a && d_out, "nib_mask_display_u8: null pointer");
  if (a->N == 0) return NIB_OK;
  NIB_REQUIRE(a->d_img && a->d_labels && a->d_sel && a->d_seg_minmax, "nib_mask_display_u8: null device pointer (d_seg_minmax is required)");
  NIB_REQUIRE(a->C >= 1 && a->C <= kMaxC && a->H > 0 && a->W > 0 && a->S > 0, "nib_mask_display_u8: bad geometry");
  NIB_REQUIRE(a->label_bytes == 1 || a->label_bytes == 2, "nib_mask_display_u8: label_bytes must be 1 or 2");
  NIB_REQUIRE(a->mode == NIB_MASK_KEEP_MUL || a->mode == NIB_MASK_REMOVE_MINMAX, "nib_mask_display_u8: bad mode");
  const int HW = a->H * a->W;
  const int gx = ceil_div(HW, 256);
  int gy = max(1, min(a->N, ceil_div(num_sms() * 16, gx)));
  const int mpc = ceil_div(a->N, gy);
  gy = ceil_div(a->N, mpc);
  dim3 grid(gx, gy);
  cudaStream_t st = (cudaStream_t)stream;
#define NIB_DISPLAY(MODE, LABT)                                                                                         \
  mask_display_kernel<MODE, LABT><<<grid, 256, 0, st>>>(a->d_img, (const LABT*)a->d_labels, a->d_sel, a->sel_words, a->N, \
                                                        a->C, a->H, a->W, a->d_seg_minmax, a->S, d_out, mpc)
  if (a->mode == NIB_MASK_KEEP_MUL) {
    if (a->label_bytes == 1) NIB_DISPLAY(NIB_MASK_KEEP_MUL, uint8_t); else NIB_DISPLAY(NIB_MASK_KEEP_MUL, uint16_t);
  } else {
    if (a->label_bytes == 1) NIB_DISPLAY(NIB_MASK_REMOVE_MINMAX, uint8_t); else NIB_DISPLAY(NIB_MASK_REMOVE_MINMAX, uint16_t);
  }
#undef NIB_DISPLAY
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_segment_weights(const uint64_t* d_sel, int sel_words, const float* d_y, int N, int S, double* d_wseg,
                        int32_t* d_cover, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_sel && d_y && d_wseg && N > 0, "nib_segment_weights: bad arguments");
  NIB_REQUIRE(S > 0 && S <= 4096 && sel_words * 64 >= S, "nib_segment_weights: bad S/sel_words");
  nib::heat_weights_kernel<<<S, 256, 0, (cudaStream_t)stream>>>(d_sel, sel_words, d_y, N, S, d_wseg, d_cover);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_heatmap(const void* d_labels, int label_bytes, int H, int W, int S, const uint64_t* d_sel,
                int sel_words, const float* d_y, int N, float* d_heat, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_labels && d_sel && d_y && d_heat, "nib_heatmap: null pointer");
  NIB_REQUIRE(S > 0 && S <= 4096 && sel_words * 64 >= S, "nib_heatmap: bad S/sel_words");
  NIB_REQUIRE(label_bytes == 1 || label_bytes == 2, "nib_heatmap: label_bytes must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  double* wseg = nullptr;   // per-stream scratch (common.cuh): concurrent heat maps on different streams do not share it
  int rcs = nib::stream_scratch(nib::SCRATCH_HEAT_WSEG, st, sizeof(double) * 4096, sizeof(double) * 4096, reinterpret_cast<void**>(&wseg));
  if (rcs != NIB_OK) return rcs;
  nib::heat_weights_kernel<<<S, 256, 0, st>>>(d_sel, sel_words, d_y, N, S, wseg, nullptr);
  const int HW = H * W;
  if (label_bytes == 1)
    nib::heat_scatter_kernel<uint8_t><<<nib::ceil_div(HW, 256), 256, 0, st>>>((const uint8_t*)d_labels, HW, S, wseg, d_heat);
  else
    nib::heat_scatter_kernel<uint16_t><<<nib::ceil_div(HW, 256), 256, 0, st>>>((const uint16_t*)d_labels, HW, S, wseg, d_heat);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

}  // extern "C"
