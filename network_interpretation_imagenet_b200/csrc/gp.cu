// gp.cu — stage 3: exact Gaussian-process surrogate over binary masks + Expected Improvement, fp64.
//
// Arithmetic restated from scikit-learn 1.9.0 `GaussianProcessRegressor` as configured at
// BayesianOptimization.py:154-159 (RBF kernel, alpha=1e-5, normalize_y=True):
//   fit      sklearn/gaussian_process/_gpr.py:233-368   (Gram, +alpha*I, cholesky, cho_solve)
//   predict  _gpr.py:446-496                             (K* alpha; V = L^-1 K*^T; 1 - sum V^2, clip)
//   LML/grad _gpr.py:588-655
//   RBF      sklearn/gaussian_process/kernels.py:1561-1569 (exp(-0.5 * sqeuclidean(X / l)))
// and the acquisition from BayesianOptimization.py:37-54.
//
// For binary X, sqeuclidean(z_i, z_j) = popcount(z_i xor z_j), so the Gram matrix is a (S+1)-entry
// LUT over Hamming distances: integer ALU + an n^2*8-byte HBM write.  Cholesky/TRSM are blocked
// (NB = 64) around one fp64 SIMT GEMM kernel (128x128 tiles, 8x8 register micro-tiles).
#include "common.cuh"
#include <math.h>

namespace nib {

static constexpr int NB = 64;

// ---- Gram ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gram_binary_kernel(const uint64_t* __restrict__ Za, int na, const uint64_t* __restrict__ Zb, int nb, int words,
                   double inv_l2_half /* 0.5 / l^2 */, double jitter, int same, double* __restrict__ K, int ldk) {
  extern __shared__ double lut[];  // [64*words + 1]
  const int nl = 64 * words + 1;
  for (int h = threadIdx.x + threadIdx.y * blockDim.x; h < nl; h += blockDim.x * blockDim.y)
    lut[h] = exp(-(double)h * inv_l2_half);
  __syncthreads();
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int i0 = blockIdx.y * 32;
  if (j >= nb) return;
  if (words == 1) {
    const uint64_t zj = Zb[j];
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
      const int i = i0 + ii;
      if (i >= na) break;
      const int h = __popcll(Za[i] ^ zj);
      double v = lut[h];
      if (same && i == j) v += jitter;
      K[(size_t)i * ldk + j] = v;
    }
  } else {
    for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
      const int i = i0 + ii;
      if (i >= na) break;
      int h = 0;
      for (int w = 0; w < words; ++w) h += __popcll(Za[(size_t)i * words + w] ^ Zb[(size_t)j * words + w]);
      double v = lut[h];
      if (same && i == j) v += jitter;
      K[(size_t)i * ldk + j] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
gram_rbf_kernel(const double* __restrict__ Xa, int na, const double* __restrict__ Xb, int nb, int d, double inv_l,
                double jitter, int same, double* __restrict__ K, int ldk) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int i0 = blockIdx.y * 32;
  if (j >= nb) return;
  for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
    const int i = i0 + ii;
    if (i >= na) break;
    double s = 0.0;
    for (int k = 0; k < d; ++k) {
      // kernels.py:1566: dists on X / length_scale
      const double diff = Xa[(size_t)i * d + k] * inv_l - Xb[(size_t)j * d + k] * inv_l;
      s += diff * diff;
    }
    double v = exp(-0.5 * s);
    if (same && i == j) v = 1.0 + jitter;  // np.fill_diagonal(K, 1) then K[diag] += alpha
    K[(size_t)i * ldk + j] = v;
  }
}

// ---- generic fp64 GEMM update:  C[i][j] -= sum_t A(i,t) * B(t,j) ------------------------------
// A(i,t) = A[i*sai + t*sat], B(t,j) = B[t*sbt + j*sbj], C row-major ldc.  lower_only skips tiles
// strictly above the diagonal (SYRK trailing update of the Cholesky).
struct DgemmArgs {
  const double* A; long long sai, sat;
  const double* B; long long sbt, sbj;
  double* C; long long ldc;
  int M, N, K;
  int lower_only;
};

static constexpr int GT = 128, GK = 8;

__global__ void __launch_bounds__(256)
dgemm_sub_kernel(DgemmArgs g) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (g.lower_only && bj > bi) return;
  __shared__ __align__(16) double As[GK][GT];
  __shared__ __align__(16) double Bs[GK][GT];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int i0 = bi * GT, j0 = bj * GT;
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;

  // loader mapping: 1024 elements each for A and B per chunk, 4 per thread.
  // choose the fastest-varying thread index along the unit-stride direction.
  const bool a_t_fast = (g.sat == 1);
  const bool b_j_fast = (g.sbj == 1);

  for (int k0 = 0; k0 < g.K; k0 += GK) {
    double av[4], bv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int ii, tt;
      if (a_t_fast) { tt = idx & (GK - 1); ii = idx >> 3; } else { ii = idx & (GT - 1); tt = idx >> 7; }
      const int gi = i0 + ii, gt = k0 + tt;
      av[e] = (gi < g.M && gt < g.K) ? g.A[gi * g.sai + gt * g.sat] : 0.0;
      int jj, t2;
      if (b_j_fast) { jj = idx & (GT - 1); t2 = idx >> 7; } else { t2 = idx & (GK - 1); jj = idx >> 3; }
      const int gj = j0 + jj, gt2 = k0 + t2;
      bv[e] = (gj < g.N && gt2 < g.K) ? g.B[gt2 * g.sbt + gj * g.sbj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int ii, tt;
      if (a_t_fast) { tt = idx & (GK - 1); ii = idx >> 3; } else { ii = idx & (GT - 1); tt = idx >> 7; }
      As[tt][ii] = av[e];
      int jj, t2;
      if (b_j_fast) { jj = idx & (GT - 1); t2 = idx >> 7; } else { t2 = idx & (GK - 1); jj = idx >> 3; }
      Bs[t2][jj] = bv[e];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < GK; ++t) {
      double a[8], b[8];
#pragma unroll
      for (int u = 0; u < 8; u += 2) {
        const double2 a2 = *reinterpret_cast<const double2*>(&As[t][ty * 8 + u]);
        a[u] = a2.x; a[u + 1] = a2.y;
        const double2 b2 = *reinterpret_cast<const double2*>(&Bs[t][tx * 8 + u]);
        b[u] = b2.x; b[u + 1] = b2.y;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = fma(a[u], b[v], acc[u][v]);
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int i = i0 + ty * 8 + u;
    if (i >= g.M) continue;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int j = j0 + tx * 8 + v;
      if (j >= g.N) continue;
      if (g.lower_only && j > i) continue;
      g.C[i * g.ldc + j] -= acc[u][v];
    }
  }
}

static int dgemm_sub(const DgemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return NIB_OK;
  dim3 grid(ceil_div(g.N, GT), ceil_div(g.M, GT));
  dgemm_sub_kernel<<<grid, 256, 0, st>>>(g);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

// ---- Cholesky --------------------------------------------------------------------------------
// factor the nb x nb diagonal block at (k0,k0) in shared memory
__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* __restrict__ A, int ld, int k0, int nb, int* __restrict__ info) {
  __shared__ double s[NB][NB + 1];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < nb * nb; idx += 256) {
    int i = idx / nb, j = idx - i * nb;
    s[i][j] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
  }
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    const double d = s[j][j];
    if (!(d > 0.0)) {  // also catches NaN
      if (tid == 0 && *info == 0) *info = k0 + j + 1;
      return;
    }
    const double r = sqrt(d);
    __syncthreads();
    if (tid == 0) s[j][j] = r;
    for (int i = j + 1 + tid; i < nb; i += 256) s[i][j] = s[i][j] / r;
    __syncthreads();
    // trailing update of the lower triangle: (i, k) with j < k <= i
    const int rem = nb - j - 1;
    for (int idx = tid; idx < rem * rem; idx += 256) {
      const int ii = idx / rem, kk = idx - ii * rem;
      if (kk <= ii) {
        const int i = j + 1 + ii, k = j + 1 + kk;
        s[i][k] -= s[i][j] * s[k][j];
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nb * nb; idx += 256) {
    int i = idx / nb, j = idx - i * nb;
    if (j <= i) A[(size_t)(k0 + i) * ld + k0 + j] = s[i][j];
  }
}

// rows below the diagonal block: solve X * Lkk^T = A[i, k0:k0+nb], one row per thread
__global__ void __launch_bounds__(128)
trsm_panel_kernel(double* __restrict__ A, int ld, int k0, int nb, int n, const int* __restrict__ info) {
  __shared__ double L[NB][NB + 1];
  if (*info != 0) return;
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    int i = idx / nb, j = idx - i * nb;
    L[i][j] = (j <= i) ? A[(size_t)(k0 + i) * ld + k0 + j] : 0.0;
  }
  __syncthreads();
  const int i = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[NB];
  double* row = A + (size_t)i * ld + k0;
#pragma unroll
  for (int j = 0; j < NB; ++j) x[j] = (j < nb) ? row[j] : 0.0;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (j < nb) {
      double sacc = x[j];
#pragma unroll
      for (int t = 0; t < j; ++t) sacc = fma(-x[t], L[j][t], sacc);
      x[j] = sacc / L[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j)
    if (j < nb) row[j] = x[j];
}

// ---- triangular solves on many right-hand sides ----------------------------------------------
// forward:  rows [k0,k0+nb) of B <- Lkk^{-1} * rows;  one column per thread
__global__ void __launch_bounds__(128)
trsm_diag_fwd_kernel(const double* __restrict__ Lm, int ldl, double* __restrict__ B, int ldb, int nrhs, int k0, int nb) {
  __shared__ double L[NB][NB + 1];
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    int i = idx / nb, j = idx - i * nb;
    L[i][j] = (j <= i) ? Lm[(size_t)(k0 + i) * ldl + k0 + j] : 0.0;
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nrhs) return;
  double x[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    if (r < nb) {
      double sacc = B[(size_t)(k0 + r) * ldb + c];
#pragma unroll
      for (int t = 0; t < r; ++t) sacc = fma(-L[r][t], x[t], sacc);
      x[r] = sacc / L[r][r];
      B[(size_t)(k0 + r) * ldb + c] = x[r];
    }
  }
}
// backward: rows [k0,k0+nb) of B <- Lkk^{-T} * rows
__global__ void __launch_bounds__(128)
trsm_diag_bwd_kernel(const double* __restrict__ Lm, int ldl, double* __restrict__ B, int ldb, int nrhs, int k0, int nb) {
  __shared__ double L[NB][NB + 1];
  for (int idx = threadIdx.x; idx < nb * nb; idx += blockDim.x) {
    int i = idx / nb, j = idx - i * nb;
    L[i][j] = (j <= i) ? Lm[(size_t)(k0 + i) * ldl + k0 + j] : 0.0;
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nrhs) return;
  double x[NB];
#pragma unroll
  for (int rr = 0; rr < NB; ++rr) {
    const int r = NB - 1 - rr;
    if (r < nb) {
      double sacc = B[(size_t)(k0 + r) * ldb + c];
#pragma unroll
      for (int t = r + 1; t < NB; ++t)
        if (t < nb) sacc = fma(-L[t][r], x[t], sacc);
      x[r] = sacc / L[r][r];
      B[(size_t)(k0 + r) * ldb + c] = x[r];
    }
  }
}

static int trsm_impl(const double* L, int n, int ldl, double* B, int nrhs, int ldb, int trans, cudaStream_t st) {
  if (n <= 0 || nrhs <= 0) return NIB_OK;
  const int cb = ceil_div(nrhs, 128);
  if (!trans) {
    for (int k0 = 0; k0 < n; k0 += NB) {
      const int nb = min(NB, n - k0);
      trsm_diag_fwd_kernel<<<cb, 128, 0, st>>>(L, ldl, B, ldb, nrhs, k0, nb);
      NIB_LAUNCH_CHECK();
      const int below = n - (k0 + nb);
      if (below > 0) {
        DgemmArgs g;
        g.A = L + (size_t)(k0 + nb) * ldl + k0; g.sai = ldl; g.sat = 1;   // L[i][k0+t]
        g.B = B + (size_t)k0 * ldb; g.sbt = ldb; g.sbj = 1;                // X[t][j]
        g.C = B + (size_t)(k0 + nb) * ldb; g.ldc = ldb;
        g.M = below; g.N = nrhs; g.K = nb; g.lower_only = 0;
        int rc = dgemm_sub(g, st);
        if (rc != NIB_OK) return rc;
      }
    }
  } else {
    const int nblk = ceil_div(n, NB);
    for (int b = nblk - 1; b >= 0; --b) {
      const int k0 = b * NB;
      const int nb = min(NB, n - k0);
      trsm_diag_bwd_kernel<<<cb, 128, 0, st>>>(L, ldl, B, ldb, nrhs, k0, nb);
      NIB_LAUNCH_CHECK();
      if (k0 > 0) {
        DgemmArgs g;
        g.A = L + (size_t)k0 * ldl; g.sai = 1; g.sat = ldl;                // A(i,t) = L[k0+t][i]
        g.B = B + (size_t)k0 * ldb; g.sbt = ldb; g.sbj = 1;
        g.C = B; g.ldc = ldb;
        g.M = k0; g.N = nrhs; g.K = nb; g.lower_only = 0;
        int rc = dgemm_sub(g, st);
        if (rc != NIB_OK) return rc;
      }
    }
  }
  return NIB_OK;
}

// ---- posterior pieces ------------------------------------------------------------------------
__global__ void transpose_kernel(const double* __restrict__ in, int rows, int cols, int ldi, double* __restrict__ out,
                                 int ldo) {
  __shared__ double tile[32][33];
  int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
    int y = y0 + yy;
    if (x < cols && y < rows) tile[yy][threadIdx.x] = in[(size_t)y * ldi + x];
  }
  __syncthreads();
  int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
  for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
    int oy = oy0 + yy;
    if (ox < rows && oy < cols) out[(size_t)oy * ldo + ox] = tile[threadIdx.x][yy];
  }
}

__global__ void __launch_bounds__(256)
gemv_mean_kernel(const double* __restrict__ Ks, int m, int n, int ldks, const double* __restrict__ alpha, double y_mean,
                 double y_std, double* __restrict__ mu) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= m) return;
  double acc = 0.0;
  for (int i = lane; i < n; i += 32) acc = fma(Ks[(size_t)q * ldks + i], alpha[i], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) mu[q] = y_std * acc + y_mean;
}

__global__ void __launch_bounds__(256)
colsumsq_var_kernel(const double* __restrict__ V, int n, int m, double prior_var, double y_std, double* __restrict__ var,
                    double* __restrict__ sd) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    const double v = V[(size_t)i * m + q];
    acc = fma(v, v, acc);
  }
  double r = prior_var - acc;
  if (r < 0.0) r = 0.0;  // _gpr.py:485-492
  r = r * y_std * y_std;
  if (var) var[q] = r;
  if (sd) sd[q] = sqrt(r);
}

// ---- LML -------------------------------------------------------------------------------------
__global__ void lml_kernel(const double* __restrict__ L, int n, int ldl, const double* __restrict__ y,
                           const double* __restrict__ alpha, double* __restrict__ out) {
  __shared__ double s1[32], s2[32];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    a = fma(y[i], alpha[i], a);
    b += log(L[(size_t)i * ldl + i]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { ta += s1[i]; tb += s2[i]; }
    out[0] = -0.5 * ta - tb - 0.5 * (double)n * 1.8378770664093453;  // log(2*pi)
  }
}

__global__ void __launch_bounds__(256)
lml_grad_kernel(const double* __restrict__ K0, const double* __restrict__ Kinv, const double* __restrict__ alpha,
                const uint64_t* __restrict__ Z, int words, int n, int ld, double* __restrict__ partial) {
  // partial[block] = sum over this block's rows of (alpha_i alpha_j - Kinv_ij) * K0_ij * hamming_ij
  const int i = blockIdx.x;
  double acc = 0.0;
  const double ai = alpha[i];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int h = 0;
    for (int w = 0; w < words; ++w) h += __popcll(Z[(size_t)i * words + w] ^ Z[(size_t)j * words + w]);
    acc = fma((ai * alpha[j] - Kinv[(size_t)i * ld + j]) * K0[(size_t)i * ld + j], (double)h, acc);
  }
  __shared__ double s[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += s[k];
    partial[i] = t;
  }
}
__global__ void __launch_bounds__(256)
lml_grad_rbf_kernel(const double* __restrict__ K0, const double* __restrict__ Kinv, const double* __restrict__ alpha,
                    const double* __restrict__ X, int d, int n, int ld, double inv_l, double* __restrict__ partial) {
  const int i = blockIdx.x;
  double acc = 0.0;
  const double ai = alpha[i];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    double s2 = 0.0;
    for (int k = 0; k < d; ++k) {
      const double diff = X[(size_t)i * d + k] * inv_l - X[(size_t)j * d + k] * inv_l;
      s2 += diff * diff;
    }
    acc = fma((ai * alpha[j] - Kinv[(size_t)i * ld + j]) * K0[(size_t)i * ld + j], s2, acc);
  }
  __shared__ double s[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += s[k];
    partial[i] = t;
  }
}
__global__ void sum_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double s[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += v[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) t += s[k];
    out[0] = t;
  }
}

// ---- Expected improvement ----------------------------------------------------------------------
__global__ void ei_kernel(const double* __restrict__ mu, const double* __restrict__ sigma, int m, double best, double sf,
                          double* __restrict__ ei) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m) return;
  const double s = sigma[q];
  const double d = sf * (mu[q] - best);
  const double Z = d / s;  // sigma == 0 -> inf/NaN, as the reference (np.errstate(divide='ignore'))
  const double Phi = 0.5 * erfc(-Z * 0.70710678118654752440);
  const double phi = exp(-0.5 * Z * Z) * 0.39894228040143267794;
  ei[q] = d * Phi + s * phi;
}
__global__ void argmax_kernel(const double* __restrict__ v, int m, long long* __restrict__ out) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  double best = -INFINITY;
  long long bi = -1;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    const double x = v[i];
    if (x == x && (bi < 0 || x > best)) { best = x; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (oi >= 0 && (bi < 0 || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (blockDim.x >> 5); ++k) {
      if (si[k] >= 0 && (bi < 0 || sv[k] > best || (sv[k] == best && si[k] < bi))) { best = sv[k]; bi = si[k]; }
    }
    out[0] = bi;
  }
}

static double* g_scratch = nullptr;  // small device scratch for scalar reductions
static size_t g_scratch_cap = 0;
static int ensure_scratch(size_t doubles) {
  if (g_scratch_cap >= doubles) return NIB_OK;
  if (g_scratch) cudaFree(g_scratch);
  g_scratch_cap = doubles < 16384 ? 16384 : doubles;
  NIB_CUDA(cudaMalloc(&g_scratch, g_scratch_cap * sizeof(double)));
  return NIB_OK;
}

}  // namespace nib

using namespace nib;

extern "C" {

int nib_gp_gram_binary(const uint64_t* d_Za, int na, const uint64_t* d_Zb, int nb, int words, double length_scale,
                       double jitter, double* d_K, int ldk, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_Za && d_Zb && d_K, "nib_gp_gram_binary: null pointer");
  NIB_REQUIRE(na > 0 && nb > 0 && words > 0 && words <= 64 && ldk >= nb, "nib_gp_gram_binary: bad shape");
  NIB_REQUIRE(length_scale > 0.0, "nib_gp_gram_binary: length_scale must be > 0");
  dim3 grid(ceil_div(nb, 32), ceil_div(na, 32)), block(32, 8);
  const size_t smem = sizeof(double) * (64 * words + 1);
  gram_binary_kernel<<<grid, block, smem, (cudaStream_t)stream>>>(d_Za, na, d_Zb, nb, words,
                                                                  0.5 / (length_scale * length_scale), jitter,
                                                                  d_Za == d_Zb ? 1 : 0, d_K, ldk);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_gp_gram_rbf(const double* d_Xa, int na, const double* d_Xb, int nb, int d, double length_scale, double jitter,
                    int same, double* d_K, int ldk, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_Xa && d_Xb && d_K, "nib_gp_gram_rbf: null pointer");
  NIB_REQUIRE(na > 0 && nb > 0 && d > 0 && ldk >= nb && length_scale > 0.0, "nib_gp_gram_rbf: bad shape");
  dim3 grid(ceil_div(nb, 32), ceil_div(na, 32)), block(32, 8);
  gram_rbf_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_Xa, na, d_Xb, nb, d, 1.0 / length_scale, jitter, same,
                                                            d_K, ldk);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_gp_cholesky(double* d_K, int n, int ldk, int* d_info, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_K && d_info && n > 0 && ldk >= n, "nib_gp_cholesky: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  NIB_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), st));
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int nb = min(NB, n - k0);
    potrf_diag_kernel<<<1, 256, 0, st>>>(d_K, ldk, k0, nb, d_info);
    NIB_LAUNCH_CHECK();
    const int below = n - (k0 + nb);
    if (below > 0) {
      trsm_panel_kernel<<<ceil_div(below, 128), 128, 0, st>>>(d_K, ldk, k0, nb, n, d_info);
      NIB_LAUNCH_CHECK();
      DgemmArgs g;
      const double* X = d_K + (size_t)(k0 + nb) * ldk + k0;
      g.A = X; g.sai = ldk; g.sat = 1;      // X[i][t]
      g.B = X; g.sbt = 1; g.sbj = ldk;      // B(t,j) = X[j][t]
      g.C = d_K + (size_t)(k0 + nb) * ldk + (k0 + nb); g.ldc = ldk;
      g.M = below; g.N = below; g.K = nb; g.lower_only = 1;
      int rc = dgemm_sub(g, st);
      if (rc != NIB_OK) return rc;
    }
  }
  return NIB_OK;
}

int nib_gp_trsm(const double* d_L, int n, int ldl, double* d_B, int nrhs, int ldb, int trans, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_L && d_B && n > 0 && nrhs > 0 && ldl >= n && ldb >= nrhs, "nib_gp_trsm: bad arguments");
  return trsm_impl(d_L, n, ldl, d_B, nrhs, ldb, trans, (cudaStream_t)stream);
}

int nib_gp_posterior(const double* d_L, int n, int ldl, const double* d_alpha, const double* d_Ks, int m, int ldks,
                     double y_mean, double y_std, double prior_var, double* d_work, double* d_mu, double* d_var,
                     double* d_std, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_L && d_alpha && d_Ks && n > 0 && m > 0 && ldl >= n && ldks >= n, "nib_gp_posterior: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (d_mu) {
    gemv_mean_kernel<<<ceil_div(m * 32, 256), 256, 0, st>>>(d_Ks, m, n, ldks, d_alpha, y_mean, y_std, d_mu);
    NIB_LAUNCH_CHECK();
  }
  if (d_var || d_std) {
    NIB_REQUIRE(d_work != nullptr, "nib_gp_posterior: variance needs d_work (m*n doubles)");
    dim3 grid(ceil_div(n, 32), ceil_div(m, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, st>>>(d_Ks, m, n, ldks, d_work, m);  // work = Ks^T  [n][m]
    NIB_LAUNCH_CHECK();
    int rc = trsm_impl(d_L, n, ldl, d_work, m, m, 0, st);
    if (rc != NIB_OK) return rc;
    colsumsq_var_kernel<<<ceil_div(m, 256), 256, 0, st>>>(d_work, n, m, prior_var, y_std, d_var, d_std);
    NIB_LAUNCH_CHECK();
  }
  return NIB_OK;
}

int nib_gp_lml(const double* d_L, int n, int ldl, const double* d_y, const double* d_alpha, double* h_lml,
               void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_L && d_y && d_alpha && h_lml && n > 0, "nib_gp_lml: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch(16);
  if (rc != NIB_OK) return rc;
  lml_kernel<<<1, 1024, 0, st>>>(d_L, n, ldl, d_y, d_alpha, g_scratch);
  NIB_LAUNCH_CHECK();
  NIB_CUDA(cudaMemcpyAsync(h_lml, g_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
  NIB_CUDA(cudaStreamSynchronize(st));
  return NIB_OK;
}

int nib_gp_lml_grad(const double* d_K0, const double* d_Kinv, const double* d_alpha, const uint64_t* d_Z, int words,
                    int n, int ld, double length_scale, double* h_grad, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_K0 && d_Kinv && d_alpha && d_Z && h_grad && n > 0 && ld >= n, "nib_gp_lml_grad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch((size_t)n + 16);
  if (rc != NIB_OK) return rc;
  lml_grad_kernel<<<n, 256, 0, st>>>(d_K0, d_Kinv, d_alpha, d_Z, words, n, ld, g_scratch + 16);
  NIB_LAUNCH_CHECK();
  sum_kernel<<<1, 1024, 0, st>>>(g_scratch + 16, n, g_scratch);
  NIB_LAUNCH_CHECK();
  double t = 0.0;
  NIB_CUDA(cudaMemcpyAsync(&t, g_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
  NIB_CUDA(cudaStreamSynchronize(st));
  // dK/dtheta = K0 .* D2 / l^2 ;  grad = 0.5 * sum((alpha alpha^T - K^-1) .* dK/dtheta)
  *h_grad = 0.5 * t / (length_scale * length_scale);
  return NIB_OK;
}

int nib_gp_lml_grad_rbf(const double* d_K0, const double* d_Kinv, const double* d_alpha, const double* d_X, int d,
                        int n, int ld, double length_scale, double* h_grad, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_K0 && d_Kinv && d_alpha && d_X && h_grad && n > 0 && d > 0 && ld >= n, "nib_gp_lml_grad_rbf: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_scratch((size_t)n + 16);
  if (rc != NIB_OK) return rc;
  lml_grad_rbf_kernel<<<n, 256, 0, st>>>(d_K0, d_Kinv, d_alpha, d_X, d, n, ld, 1.0 / length_scale, g_scratch + 16);
  NIB_LAUNCH_CHECK();
  sum_kernel<<<1, 1024, 0, st>>>(g_scratch + 16, n, g_scratch);
  NIB_LAUNCH_CHECK();
  double t = 0.0;
  NIB_CUDA(cudaMemcpyAsync(&t, g_scratch, sizeof(double), cudaMemcpyDeviceToHost, st));
  NIB_CUDA(cudaStreamSynchronize(st));
  *h_grad = 0.5 * t;  // D2 here is already sqeuclidean(X / l): dK/dtheta = K0 .* D2
  return NIB_OK;
}

int nib_gp_ei(const double* d_mu, const double* d_sigma, int m, double best, int greater_is_better, double* d_ei,
              long long* d_argmax, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_mu && d_sigma && d_ei && m > 0, "nib_gp_ei: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ei_kernel<<<ceil_div(m, 256), 256, 0, st>>>(d_mu, d_sigma, m, best, greater_is_better ? 1.0 : -1.0, d_ei);
  NIB_LAUNCH_CHECK();
  if (d_argmax) {
    argmax_kernel<<<1, 1024, 0, st>>>(d_ei, m, d_argmax);
    NIB_LAUNCH_CHECK();
  }
  return NIB_OK;
}

}  // extern "C"
