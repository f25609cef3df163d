// conv_simt.cu — generic implicit-GEMM convolution on CUDA cores (fp32 accumulate), plus the
// small layer kernels (pooling, fully-connected, layout transforms).
//
// Role: (1) the fp32 parity mode (logits <= 1e-4 vs the reference's fp32 PyTorch forward,
// generate_gp_training_data_imagenet.py:246), (2) layers the tcgen05 path does not cover
// (Cin not a multiple of 64: the 3-channel stems, CIFAR ResNet-56 16/32-channel stages,
// MNIST 1-channel input).  All activations NHWC; weights KRSC ([Cout][R][S][Cin]).
//
//   out[m, co] = act( sum_{r,s,c} pre(in[n, p*st-pad+r, q*st-pad+s, c]) * w[co, r, s, c] + bias[co]
//                     + residual[m, co] )
#include "common.cuh"
#include "layers.cuh"

namespace nib {

static constexpr int BM = 128, BK = 16;

// BN x TN: 64 x 4 (8 x 4 outputs per thread) for narrow layers, 128 x 8 (8 x 8 outputs per thread: 64 FMAs per four
// 128-bit shared loads) for Cout >= 128.  The next K slab's global loads are issued before the FMAs of the current one
// (register prefetch), so their latency hides behind the arithmetic.  Every output accumulates its products in ascending
// k with fmaf: results do not depend on the tile shape.
template <typename T, int BN, int TN>
__global__ void __launch_bounds__(256)
conv_simt_kernel(ConvParams p) {
  static_assert(BN == 16 * TN, "16 thread columns of TN outputs");
  constexpr int BPT = BN * BK / 256;          // weight elements each thread stages per slab (4 or 8), consecutive in k
  constexpr int B_TPR = BK / BPT;             // threads per weight row (4 or 2)
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];

  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  const T* __restrict__ wgt = reinterpret_cast<const T*>(p.w);
  T* __restrict__ out = reinterpret_cast<T*>(p.out);
  const T* __restrict__ res = reinterpret_cast<const T*>(p.res);

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int K = p.R * p.S * p.Cin;
  if (p.dyn_n != nullptr) {   // device-side live batch: whole blocks beyond it leave before the first barrier
    const long long live = (long long)max(*p.dyn_n, 0) * p.P * p.Q;
    if (live < p.M) p.M = (int)live;
    if (m0 >= p.M) return;
  }

  // A-load role: one pixel per thread (tid % 128), 8 consecutive k (tid / 128)
  const int a_pix = tid & (BM - 1);
  const int a_k0 = (tid >> 7) * 8;
  const int am = m0 + a_pix;
  const bool a_valid = am < p.M;
  int an = 0, ap = 0, aq = 0;
  if (a_valid) {
    an = am / (p.P * p.Q);
    int rem = am - an * p.P * p.Q;
    ap = rem / p.Q;
    aq = rem - ap * p.Q;
  }
  const int ih0 = ap * p.stride - p.pad, iw0 = aq * p.stride - p.pad;
  const int Hp = p.Hin + 2 * p.in_halo, Wp = p.Win + 2 * p.in_halo;
  const T* in_img = in + (size_t)an * Hp * Wp * p.in_cstride + p.in_coff;
  const bool cin8 = (p.Cin % 8 == 0) && (p.in_cstride % 8 == 0) && (p.in_coff % 8 == 0);

  // B-load role: cout = tid / B_TPR, BPT consecutive k
  const int b_row = tid / B_TPR;
  const int b_co = n0 + b_row;
  const int b_k0 = (tid % B_TPR) * BPT;
  const bool kvec = (K % 4 == 0);

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float av[8], bv[BPT];
  auto load_slab = [&](int kb) {
    // ---- gather A ----
#pragma unroll
    for (int j = 0; j < 8; ++j) av[j] = 0.f;
    if (a_valid) {
      const int k = kb + a_k0;
      if (cin8) {
        if (k < K) {
          const int tap = k / p.Cin, c = k - tap * p.Cin;
          const int r = tap / p.S, s = tap - r * p.S;
          const int ih = ih0 + r, iw = iw0 + s;
          if (ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win) {
            const T* src = in_img + ((size_t)(ih + p.in_halo) * Wp + (iw + p.in_halo)) * p.in_cstride + c;
            if (sizeof(T) == 2) {
              uint4 raw = *reinterpret_cast<const uint4*>(src);
              const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
              for (int j = 0; j < 8; ++j) av[j] = __bfloat162float(h[j]);
            } else {
              float4 f0 = *reinterpret_cast<const float4*>(src);
              float4 f1 = *reinterpret_cast<const float4*>(src + 4);
              av[0] = f0.x; av[1] = f0.y; av[2] = f0.z; av[3] = f0.w;
              av[4] = f1.x; av[5] = f1.y; av[6] = f1.z; av[7] = f1.w;
            }
            if (p.pre_scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                av[j] = fmaxf(fmaf(av[j], p.pre_scale[c + j], p.pre_shift[c + j]), 0.f);
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int kk = k + j;
          if (kk < K) {
            const int tap = kk / p.Cin, c = kk - tap * p.Cin;
            const int r = tap / p.S, s = tap - r * p.S;
            const int ih = ih0 + r, iw = iw0 + s;
            if (ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win) {
              float v = Elem<T>::ld(in_img + ((size_t)(ih + p.in_halo) * Wp + (iw + p.in_halo)) * p.in_cstride + c);
              if (p.pre_scale != nullptr) v = fmaxf(fmaf(v, p.pre_scale[c], p.pre_shift[c]), 0.f);
              av[j] = v;
            }
          }
        }
      }
    }
    // ---- load B ----
#pragma unroll
    for (int j = 0; j < BPT; ++j) bv[j] = 0.f;
    if (b_co < p.Cout) {
      const int k = kb + b_k0;
      const T* src = wgt + (size_t)b_co * K + k;
#pragma unroll
      for (int q = 0; q < BPT; q += 4) {
        if (kvec && k + q + 3 < K) {
          if (sizeof(T) == 2) {
            uint2 raw = *reinterpret_cast<const uint2*>(src + q);
            const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[q + j] = __bfloat162float(h[j]);
          } else {
            float4 f = *reinterpret_cast<const float4*>(src + q);
            bv[q] = f.x; bv[q + 1] = f.y; bv[q + 2] = f.z; bv[q + 3] = f.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k + q + j < K) bv[q + j] = Elem<T>::ld(src + q + j);
        }
      }
    }
  };

  load_slab(0);
  for (int kb = 0; kb < K; kb += BK) {
    __syncthreads();  // previous slab fully consumed
#pragma unroll
    for (int j = 0; j < 8; ++j) As[a_k0 + j][a_pix] = av[j];
#pragma unroll
    for (int j = 0; j < BPT; ++j) Bs[b_k0 + j][b_row] = bv[j];
    __syncthreads();
    if (kb + BK < K) load_slab(kb + BK);   // in flight while this slab is multiplied
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[TN];
#pragma unroll
      for (int q = 0; q < TN; q += 4) {
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + q]);
        bb[q] = b.x; bb[q + 1] = b.y; bb[q + 2] = b.z; bb[q + 3] = b.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= p.M) continue;
    size_t out_row, res_row = 0;
    if (p.out_halo == 0) {
      out_row = (size_t)m * p.out_cstride + p.out_coff;
    } else {
      int n = m / (p.P * p.Q), rem = m - n * p.P * p.Q;
      int pp = rem / p.Q, qq = rem - pp * p.Q;
      out_row = (((size_t)n * (p.P + 2 * p.out_halo) + pp + p.out_halo) * (p.Q + 2 * p.out_halo) + qq + p.out_halo) *
                    p.out_cstride + p.out_coff;
    }
    if (res != nullptr) res_row = (size_t)m * p.res_cstride + p.res_coff;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co >= p.Cout) continue;
      float v = acc[i][j];
      if (p.bias != nullptr) v += p.bias[co];
      if (res != nullptr && co < p.res_C) v += Elem<T>::ld(res + res_row + co);
      if (p.relu) v = fmaxf(v, 0.f);
      Elem<T>::st(out + out_row + co, v);
    }
  }
}

int launch_conv_simt(const ConvParams& p, bool bf16, cudaStream_t st) {
  if (p.Cout >= 128) {
    dim3 grid(ceil_div(p.M, BM), ceil_div(p.Cout, 128));
    if (bf16)
      conv_simt_kernel<__nv_bfloat16, 128, 8><<<grid, 256, 0, st>>>(p);
    else
      conv_simt_kernel<float, 128, 8><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(ceil_div(p.M, BM), ceil_div(p.Cout, 64));
    if (bf16)
      conv_simt_kernel<__nv_bfloat16, 64, 4><<<grid, 256, 0, st>>>(p);
    else
      conv_simt_kernel<float, 64, 4><<<grid, 256, 0, st>>>(p);
  }
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- pooling ---------------------------------------------------------------------------------
template <typename T>
__global__ void pool_kernel(PoolParams p) {
  const T* __restrict__ in = reinterpret_cast<const T*>(p.in);
  T* __restrict__ out = reinterpret_cast<T*>(p.out);
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Nl = p.dyn_n ? min(max(*p.dyn_n, 0), p.N) : p.N;
  long long total = (long long)Nl * p.P * p.Q * p.C;
  if (idx >= total) return;
  int c = (int)(idx % p.C);
  long long t = idx / p.C;
  int q = (int)(t % p.Q); t /= p.Q;
  int pp = (int)(t % p.P);
  int n = (int)(t / p.P);
  float acc = p.kind == NIB_POOL_MAX ? -INFINITY : 0.f;
  for (int r = 0; r < p.k; ++r) {
    int ih = pp * p.stride - p.pad + r;
    if (ih < 0 || ih >= p.Hin) continue;
    for (int s = 0; s < p.k; ++s) {
      int iw = q * p.stride - p.pad + s;
      if (iw < 0 || iw >= p.Win) continue;
      float v = Elem<T>::ld(in + (((size_t)n * p.Hin + ih) * p.Win + iw) * p.in_cstride + p.in_coff + c);
      if (p.pre_scale != nullptr) v = fmaxf(fmaf(v, p.pre_scale[c], p.pre_shift[c]), 0.f);
      acc = p.kind == NIB_POOL_MAX ? fmaxf(acc, v) : acc + v;
    }
  }
  if (p.kind == NIB_POOL_AVG) acc = acc / (float)(p.k * p.k);
  Elem<T>::st(out + (((size_t)n * p.P + pp) * p.Q + q) * p.out_cstride + p.out_coff + c, acc);
}

// 4 fp32 channels (one 128-bit load/store) per thread; no pre-activation.  The fp32 / split networks' max-pool and global
// average pool (the scalar kernel above spent 260 us on the 80-image re-score batch of the tie policy).
__global__ void __launch_bounds__(256)
pool_f32x4_kernel(PoolParams p) {
  const float* __restrict__ in = reinterpret_cast<const float*>(p.in);
  float* __restrict__ out = reinterpret_cast<float*>(p.out);
  const int C4 = p.C >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Nl = p.dyn_n ? min(max(*p.dyn_n, 0), p.N) : p.N;
  long long total = (long long)Nl * p.P * p.Q * C4;
  if (idx >= total) return;
  const int c = (int)(idx % C4) * 4;
  long long t = idx / C4;
  const int q = (int)(t % p.Q); t /= p.Q;
  const int pp = (int)(t % p.P);
  const int n = (int)(t / p.P);
  const float init = p.kind == NIB_POOL_MAX ? -INFINITY : 0.f;
  float4 acc = make_float4(init, init, init, init);
  for (int r = 0; r < p.k; ++r) {
    const int ih = pp * p.stride - p.pad + r;
    if (ih < 0 || ih >= p.Hin) continue;
    for (int s = 0; s < p.k; ++s) {
      const int iw = q * p.stride - p.pad + s;
      if (iw < 0 || iw >= p.Win) continue;
      const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * p.Hin + ih) * p.Win + iw) * p.in_cstride + p.in_coff + c);
      if (p.kind == NIB_POOL_MAX) {
        acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
      } else {
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  if (p.kind == NIB_POOL_AVG) {
    const float d = (float)(p.k * p.k);
    acc.x = acc.x / d; acc.y = acc.y / d; acc.z = acc.z / d; acc.w = acc.w / d;
  }
  *reinterpret_cast<float4*>(out + (((size_t)n * p.P + pp) * p.Q + q) * p.out_cstride + p.out_coff + c) = acc;
}

// 8 bf16 channels (one 128-bit load/store) per thread; no pre-activation.
__global__ void __launch_bounds__(256)
pool_bf16x8_kernel(PoolParams p) {
  const __nv_bfloat16* __restrict__ in = reinterpret_cast<const __nv_bfloat16*>(p.in);
  __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(p.out);
  const int C8 = p.C >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Nl = p.dyn_n ? min(max(*p.dyn_n, 0), p.N) : p.N;
  long long total = (long long)Nl * p.P * p.Q * C8;
  if (idx >= total) return;
  const int c = (int)(idx % C8) * 8;
  long long t = idx / C8;
  const int q = (int)(t % p.Q); t /= p.Q;
  const int pp = (int)(t % p.P);
  const int n = (int)(t / p.P);
  if (p.kind == NIB_POOL_MAX && p.k == 3 && p.pad <= 1) {
    // 3 x 3 max (the ResNet / DenseNet stem pool): window coordinates clamped into the image instead of skipped — a clamped
    // tap re-reads a pixel the window already holds, so the maximum is unchanged — which makes the nine 128-bit loads
    // independent and lets them all be in flight at once (the skip-branches serialised them: 2.9 TB/s)
    uint4 raw[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = min(max(pp * p.stride - p.pad + r, 0), p.Hin - 1);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = min(max(q * p.stride - p.pad + s, 0), p.Win - 1);
        raw[r * 3 + s] = __ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * p.Hin + ih) * p.Win + iw) * p.in_cstride + p.in_coff + c));
      }
    }
    uint4 o = raw[0];
    __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int t = 1; t < 9; ++t) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[t]);
#pragma unroll
      for (int j = 0; j < 4; ++j) oh[j] = __hmax2(oh[j], h2[j]);
    }
    *reinterpret_cast<uint4*>(out + (((size_t)n * p.P + pp) * p.Q + q) * p.out_cstride + p.out_coff + c) = o;
    return;
  }
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = p.kind == NIB_POOL_MAX ? -INFINITY : 0.f;
  for (int r = 0; r < p.k; ++r) {
    const int ih = pp * p.stride - p.pad + r;
    if (ih < 0 || ih >= p.Hin) continue;
    for (int s = 0; s < p.k; ++s) {
      const int iw = q * p.stride - p.pad + s;
      if (iw < 0 || iw >= p.Win) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(in + (((size_t)n * p.Hin + ih) * p.Win + iw) * p.in_cstride + p.in_coff + c);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        if (p.kind == NIB_POOL_MAX) {
          acc[2 * j] = fmaxf(acc[2 * j], f.x);
          acc[2 * j + 1] = fmaxf(acc[2 * j + 1], f.y);
        } else {
          acc[2 * j] += f.x;
          acc[2 * j + 1] += f.y;
        }
      }
    }
  }
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float a = acc[2 * j], b = acc[2 * j + 1];
    if (p.kind == NIB_POOL_AVG) { a = a / (float)(p.k * p.k); b = b / (float)(p.k * p.k); }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    ow[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(out + (((size_t)n * p.P + pp) * p.Q + q) * p.out_cstride + p.out_coff + c) = o;
}

int launch_pool(const PoolParams& p, bool bf16, cudaStream_t st) {
  if (bf16 && p.pre_scale == nullptr && p.C % 8 == 0 && p.in_cstride % 8 == 0 && p.in_coff % 8 == 0 &&
      p.out_cstride % 8 == 0 && p.out_coff % 8 == 0) {
    long long total8 = (long long)p.N * p.P * p.Q * (p.C / 8);
    pool_bf16x8_kernel<<<(unsigned)ceil_div_ll(total8, 256), 256, 0, st>>>(p);
    NIB_LAUNCH_CHECK();
    return NIB_OK;
  }
  if (!bf16 && p.pre_scale == nullptr && p.C % 4 == 0 && p.in_cstride % 4 == 0 && p.in_coff % 4 == 0 &&
      p.out_cstride % 4 == 0 && p.out_coff % 4 == 0) {
    long long total4 = (long long)p.N * p.P * p.Q * (p.C / 4);
    pool_f32x4_kernel<<<(unsigned)ceil_div_ll(total4, 256), 256, 0, st>>>(p);
    NIB_LAUNCH_CHECK();
    return NIB_OK;
  }
  long long total = (long long)p.N * p.P * p.Q * p.C;
  unsigned blocks = (unsigned)ceil_div_ll(total, 256);
  if (bf16)
    pool_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(p);
  else
    pool_kernel<float><<<blocks, 256, 0, st>>>(p);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- pre-activation pack (DenseNet BN-ReLU-conv on the tensor-core path) ------------------------
// y[m][c] = relu(x[m][in_coff + c] * scale[c] + shift[c]) for c < Cin, 0 for Cin <= c < Cpad, written compactly as
// [M][Cpad] bf16 so the tcgen05 conv can read it through TMA (models/densenet.py:16-21 applies norm1/relu1 per consumer
// to the shared concat buffer, so it cannot be folded into the producer).  8 channels (16 B) per thread.
__global__ void __launch_bounds__(256)
bnrelu_pack_kernel(const __nv_bfloat16* __restrict__ x, int in_cstride, int in_coff, int Cin, int Cpad,
                   const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ y,
                   long long M) {
  const int groups = Cpad >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * groups) return;
  const long long m = idx / groups;
  const int c0 = (int)(idx - m * groups) << 3;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (c0 < Cin) {
    const uint4 raw = *reinterpret_cast<const uint4*>(x + (size_t)m * in_cstride + in_coff + c0);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
    const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
    const float2 a = __bfloat1622float2(v[0]), b = __bfloat1622float2(v[1]), c = __bfloat1622float2(v[2]), d = __bfloat1622float2(v[3]);
    __nv_bfloat162 r;
    r = __floats2bfloat162_rn(fmaxf(fmaf(a.x, s0.x, h0.x), 0.f), fmaxf(fmaf(a.y, s0.y, h0.y), 0.f)); o.x = *reinterpret_cast<uint32_t*>(&r);
    r = __floats2bfloat162_rn(fmaxf(fmaf(b.x, s0.z, h0.z), 0.f), fmaxf(fmaf(b.y, s0.w, h0.w), 0.f)); o.y = *reinterpret_cast<uint32_t*>(&r);
    r = __floats2bfloat162_rn(fmaxf(fmaf(c.x, s1.x, h1.x), 0.f), fmaxf(fmaf(c.y, s1.y, h1.y), 0.f)); o.z = *reinterpret_cast<uint32_t*>(&r);
    r = __floats2bfloat162_rn(fmaxf(fmaf(d.x, s1.z, h1.z), 0.f), fmaxf(fmaf(d.y, s1.w, h1.w), 0.f)); o.w = *reinterpret_cast<uint32_t*>(&r);
  }
  *reinterpret_cast<uint4*>(y + (size_t)m * Cpad + c0) = o;
}

int launch_bnrelu_pack(const void* x, int in_cstride, int in_coff, int Cin, int Cpad, const float* scale,
                       const float* shift, void* y, long long M, cudaStream_t st) {
  const long long total = M * (Cpad >> 3);
  bnrelu_pack_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, st>>>(
      (const __nv_bfloat16*)x, in_cstride, in_coff, Cin, Cpad, scale, shift, (__nv_bfloat16*)y, M);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- fully connected (fp32 weights, fp32 math; features in the net dtype) ---------------------
// logits[n][k] = sum_c feat[n][c] * W[k][c] + b[k].  Register-tiled SGEMM: 32 samples x 64 classes per block,
// 2 x 4 outputs per thread, 32-wide K chunks staged (transposed) in smem; global loads are issued for chunk i+1
// before the FMAs of chunk i.  fp32 weights and fp32 math in both precisions: top-1 is decided here.
static constexpr int FC_BM = 32, FC_BN = 64, FC_BK = 32;
template <typename T>
__global__ void __launch_bounds__(256)
fc_kernel(const T* __restrict__ feat, int feat_stride, const float* __restrict__ w, const float* __restrict__ b,
          int N, int Cin, int Cout, float* __restrict__ logits, const int* __restrict__ dyn_n) {
  __shared__ float Fs[FC_BK][FC_BM + 2];
  __shared__ __align__(16) float Ws[FC_BK][FC_BN + 4];
  const int tid = threadIdx.x;
  const int n0 = blockIdx.y * FC_BM, k0 = blockIdx.x * FC_BN;
  if (dyn_n != nullptr) {
    N = min(max(*dyn_n, 0), N);
    if (n0 >= N) return;
  }
  const int ty = tid >> 4, tx = tid & 15;   // samples 2ty, 2ty+1; classes 4tx..4tx+3
  // loader roles: feature element (row fr, k fc*4..+3); weight elements (row wr and wr+32, k wc*4..+3)
  const int fr = tid >> 3, fc = tid & 7;
  const int wr = tid >> 3, wc = tid & 7;
  float acc[2][4] = {};
  float fv[4], wv0[4], wv1[4];
  auto load_chunk = [&](int c0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + fc * 4 + j;
      fv[j] = (n0 + fr < N && c < Cin) ? Elem<T>::ld(feat + (size_t)(n0 + fr) * feat_stride + c) : 0.f;
      wv0[j] = (k0 + wr < Cout && c < Cin) ? __ldg(w + (size_t)(k0 + wr) * Cin + c) : 0.f;
      wv1[j] = (k0 + wr + 32 < Cout && c < Cin) ? __ldg(w + (size_t)(k0 + wr + 32) * Cin + c) : 0.f;
    }
  };
  load_chunk(0);
  for (int c0 = 0; c0 < Cin; c0 += FC_BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Fs[fc * 4 + j][fr] = fv[j];
      Ws[wc * 4 + j][wr] = wv0[j];
      Ws[wc * 4 + j][wr + 32] = wv1[j];
    }
    __syncthreads();
    if (c0 + FC_BK < Cin) load_chunk(c0 + FC_BK);
#pragma unroll
    for (int c = 0; c < FC_BK; ++c) {
      const float a0 = Fs[c][2 * ty], a1 = Fs[c][2 * ty + 1];
      const float4 bv = *reinterpret_cast<const float4*>(&Ws[c][4 * tx]);
      acc[0][0] = fmaf(a0, bv.x, acc[0][0]); acc[0][1] = fmaf(a0, bv.y, acc[0][1]);
      acc[0][2] = fmaf(a0, bv.z, acc[0][2]); acc[0][3] = fmaf(a0, bv.w, acc[0][3]);
      acc[1][0] = fmaf(a1, bv.x, acc[1][0]); acc[1][1] = fmaf(a1, bv.y, acc[1][1]);
      acc[1][2] = fmaf(a1, bv.z, acc[1][2]); acc[1][3] = fmaf(a1, bv.w, acc[1][3]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = n0 + 2 * ty + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + 4 * tx + j;
      if (k < Cout) logits[(size_t)n * Cout + k] = acc[i][j] + (b ? b[k] : 0.f);
    }
  }
}

int launch_fc(const void* feat, int feat_stride, bool bf16, const float* w, const float* b, int N, int Cin,
              int Cout, float* logits, const int* dyn_n, cudaStream_t st) {
  dim3 grid(ceil_div(Cout, FC_BN), ceil_div(N, FC_BM));
  if (bf16)
    fc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)feat, feat_stride, w, b, N, Cin, Cout, logits, dyn_n);
  else
    fc_kernel<float><<<grid, 256, 0, st>>>((const float*)feat, feat_stride, w, b, N, Cin, Cout, logits, dyn_n);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- layout transforms -----------------------------------------------------------------------
// NCHW fp32 -> NHWC T with channel padding (zeros) and halo offset (halo assumed already zero)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int H, int W, T* __restrict__ out,
                                    int cs, int halo) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * H * W;
  if (idx >= total) return;
  int w = (int)(idx % W);
  long long t = idx / W;
  int h = (int)(t % H);
  int n = (int)(t / H);
  T* o = out + (((size_t)n * (H + 2 * halo) + h + halo) * (W + 2 * halo) + w + halo) * cs;
  for (int c = 0; c < cs; ++c) {
    float v = c < C ? x[(((size_t)n * C + c) * H + h) * W + w] : 0.f;
    Elem<T>::st(o + c, v);
  }
}
int launch_nchw_to_nhwc(const float* x, int N, int C, int H, int W, void* out, int cs, int halo, bool bf16,
                        cudaStream_t st) {
  unsigned blocks = (unsigned)ceil_div_ll((long long)N * H * W, 256);
  if (bf16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(x, N, C, H, W, (__nv_bfloat16*)out, cs, halo);
  else
    nchw_to_nhwc_kernel<float><<<blocks, 256, 0, st>>>(x, N, C, H, W, (float*)out, cs, halo);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int N, int C, int H, int W, int cs, int halo,
                                    float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)N * C * H * W;
  if (idx >= total) return;
  int w = (int)(idx % W);
  long long t = idx / W;
  int h = (int)(t % H); t /= H;
  int c = (int)(t % C);
  int n = (int)(t / C);
  out[idx] = Elem<T>::ld(in + (((size_t)n * (H + 2 * halo) + h + halo) * (W + 2 * halo) + w + halo) * cs + c);
}
int launch_nhwc_to_nchw(const void* in, int N, int C, int H, int W, int cs, int halo, bool bf16, float* out,
                        cudaStream_t st) {
  unsigned blocks = (unsigned)ceil_div_ll((long long)N * C * H * W, 256);
  if (bf16)
    nhwc_to_nchw_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, N, C, H, W, cs, halo, out);
  else
    nhwc_to_nchw_kernel<float><<<blocks, 256, 0, st>>>((const float*)in, N, C, H, W, cs, halo, out);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- fp32 <-> split bf16 -------------------------------------------------------------------------
// Boundaries of a NIB_PREC_SPLIT network (conv_tc.cu split mode): the stem and the pooled features are fp32, the body's
// tensors carry each value as two bf16 halves in the channel dimension.  4 channels per thread.
__global__ void __launch_bounds__(256)
split_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long M, int C, const int* __restrict__ dyn_n,
                  int rows_per_image) {
  const int g = C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (dyn_n != nullptr) M = min(M, (long long)max(*dyn_n, 0) * rows_per_image);
  if (idx >= M * g) return;
  const long long m = idx / g;
  const int c = (int)(idx - m * g) << 2;
  const float4 v = *reinterpret_cast<const float4*>(x + m * C + c);
  const float f[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = __float2bfloat16_rn(f[j]);
    lo[j] = __float2bfloat16_rn(f[j] - __bfloat162float(hi[j]));
  }
  *reinterpret_cast<uint2*>(y + m * 2 * C + c) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(y + m * 2 * C + C + c) = *reinterpret_cast<const uint2*>(lo);
}
__global__ void __launch_bounds__(256)
split_merge_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long M, int C, const int* __restrict__ dyn_n,
                   int rows_per_image) {
  const int g = C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (dyn_n != nullptr) M = min(M, (long long)max(*dyn_n, 0) * rows_per_image);
  if (idx >= M * g) return;
  const long long m = idx / g;
  const int c = (int)(idx - m * g) << 2;
  const uint2 h = *reinterpret_cast<const uint2*>(x + m * 2 * C + c), l = *reinterpret_cast<const uint2*>(x + m * 2 * C + C + c);
  const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h);
  const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l);
  float4 o;
  o.x = __bfloat162float(hb[0]) + __bfloat162float(lb[0]); o.y = __bfloat162float(hb[1]) + __bfloat162float(lb[1]);
  o.z = __bfloat162float(hb[2]) + __bfloat162float(lb[2]); o.w = __bfloat162float(hb[3]) + __bfloat162float(lb[3]);
  *reinterpret_cast<float4*>(y + m * C + c) = o;
}
int launch_split_convert(const void* in, void* out, long long M, int C, bool to_split, const int* dyn_n, int rows_per_image,
                         cudaStream_t st) {
  NIB_REQUIRE(C % 4 == 0, "split convert: C = %d is not a multiple of 4", C);
  const long long total = M * (C >> 2);
  const unsigned blocks = (unsigned)ceil_div_ll(total, 256);
  if (to_split)
    split_pack_kernel<<<blocks, 256, 0, st>>>((const float*)in, (__nv_bfloat16*)out, M, C, dyn_n, rows_per_image);
  else
    split_merge_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)in, (float*)out, M, C, dyn_n, rows_per_image);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

}  // namespace nib
