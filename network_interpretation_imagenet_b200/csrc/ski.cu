// ski.cu — pixel-coordinate GP regression on an inducing grid and the pixel heat map it is trained on.
//
// Replaces the arithmetic of gp_regression.py:
//   :63-104   heat map  H[p] = sum_i label_i * [mask_i[p] == 255]   (an O(N * 224^2) Python dict loop there)
//   :160-176  GPRegressionModel = ExactGP + GridInterpolationKernel(RBF, grid_size=30, grid_bounds=[(0,n),(0,n)]) + outputscale
//   :244-261  posterior mean (and variance) at every pixel of the n x n grid
// gpytorch is absent from this image and the reference pins no version (pre-0.1 API) -> parity unpinned.  Restated from the
// published KISS-GP / SKI construction (Wilson & Nickisch 2015): K_XX ~= W K_UU W^T with W the local cubic-convolution
// interpolation (Keys 1981, a = -0.5; 4 grid points per dimension) onto a regular grid U.  With A = W^T W, b = W^T (y - c),
// noise s2 and Sigma = s2 K + K A K (all G x G, G = grid_size^2), the exact posterior of that model is
//   mean_U = K Sigma^-1 K b          mean(x*) = c + w*^T mean_U
//   Cov_U  = s2 K Sigma^-1 K         var(x*)  = w*^T Cov_U w* = s2 |C^-1 K w*|^2   (Sigma = C C^T)
// The dense G x G algebra runs in the fp64 kernels of gp.cu (Gram, DGEMM, Cholesky, TRSM); this file holds the sparse
// parts: accumulating A and b over the n training pixels, and interpolating mean / variance at the query pixels.
#include "common.cuh"

namespace nib {
namespace {

// Keys cubic convolution kernel, a = -0.5
__device__ __forceinline__ double keys(double s) {
  s = fabs(s);
  if (s <= 1.0) return (1.5 * s - 2.5) * s * s + 1.0;
  if (s < 2.0) return ((-0.5 * s + 2.5) * s - 4.0) * s + 2.0;
  return 0.0;
}

// 4 x 4 stencil of a point: flat grid indices (row-major over (dim0, dim1)) and weights.  g0/h: first grid coordinate and
// spacing, gs: grid points per dimension.  The grid is built with one spacing of margin on either side of the data
// bounds (host side), so the stencil of an in-bounds point never leaves it; indices are clamped all the same.
__device__ __forceinline__ void stencil(double x0, double x1, double g0, double h, int gs, int (&idx)[16], double (&w)[16]) {
  const double u0 = (x0 - g0) / h, u1 = (x1 - g0) / h;
  const int i0 = (int)floor(u0) - 1, i1 = (int)floor(u1) - 1;
  double w0[4], w1[4];
  int c0[4], c1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    w0[k] = keys(u0 - (double)(i0 + k));
    w1[k] = keys(u1 - (double)(i1 + k));
    c0[k] = min(max(i0 + k, 0), gs - 1);
    c1[k] = min(max(i1 + k, 0), gs - 1);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      idx[a * 4 + b] = c0[a] * gs + c1[b];
      w[a * 4 + b] = w0[a] * w1[b];
    }
}

// A += W^T W, b += W^T (y - c): one thread per training point, 16 + 256 fp64 atomics into L2-resident G x G / G arrays
__global__ void __launch_bounds__(128)
ski_accumulate_kernel(const double* __restrict__ X, const double* __restrict__ y, int n, double g0, double h, int gs,
                      double cmean, double* __restrict__ A, double* __restrict__ b) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int idx[16];
  double w[16];
  stencil(X[2 * p], X[2 * p + 1], g0, h, gs, idx, w);
  const double r = y[p] - cmean;
  const int G = gs * gs;
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    atomicAdd(b + idx[a], w[a] * r);
#pragma unroll
    for (int c = 0; c < 16; ++c) atomicAdd(A + (size_t)idx[a] * G + idx[c], w[a] * w[c]);
  }
}

// mean[q] = c + w^T mean_U;  var[q] = s2 * |Gm w|^2 (+ s2 with the Gaussian likelihood), Gm = C^-1 K  [G x G] row-major.
// One warp per query: lanes stride over the G rows of Gm, each row needs the 16 stencil columns.
__global__ void __launch_bounds__(256)
ski_predict_kernel(const double* __restrict__ Xq, int m, double g0, double h, int gs, double cmean,
                   const double* __restrict__ mean_u, const double* __restrict__ Gm, double s2, int add_noise,
                   double* __restrict__ mean, double* __restrict__ var) {
  const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= m) return;
  int idx[16];
  double w[16];
  stencil(Xq[2 * q], Xq[2 * q + 1], g0, h, gs, idx, w);
  const int G = gs * gs;
  if (lane == 0 && mean) {
    double acc = cmean;
#pragma unroll
    for (int a = 0; a < 16; ++a) acc = fma(w[a], mean_u[idx[a]], acc);
    mean[q] = acc;
  }
  if (var) {
    double ss = 0.0;
    for (int r = lane; r < G; r += 32) {
      const double* row = Gm + (size_t)r * G;
      double v = 0.0;
#pragma unroll
      for (int a = 0; a < 16; ++a) v = fma(w[a], row[idx[a]], v);
      ss = fma(v, v, ss);
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) var[q] = s2 * ss + (add_noise ? s2 : 0.0);
  }
}

// ---- variational GP classification on the inducing grid (gp_classification.py:139-264) -----------------------------------
// Model (gpytorch GridInducingVariationalGP + BernoulliLikelihood, pre-0.1 API, absent here -> parity unpinned; restated
// from its published construction): inducing values u ~ N(0, K_UU) on the grid, q(u) = N(m, S), latent f(x) = c + w(x)^T u
// with the same cubic interpolation weights as above, probit likelihood p(y | f) = Phi(y f).  The O(n) part of one
// optimiser step is the expected log-likelihood and its gradients:
//     mu_i = c + w_i^T m,   s2_i = w_i^T S w_i,   E_i = E_{f ~ N(mu_i, s2_i)}[ log Phi(y_i f) ]
//     dE/dm = sum_i dE_i/dmu w_i,   dE/dS = sum_i dE_i/ds2 w_i w_i^T       (dE/ds2 = 1/2 E[d^2/df^2 log Phi(y f)])
// with the expectations by Gauss-Hermite quadrature (20 nodes; the reference's version samples, which no parity test
// could pin).  One thread per training pixel; 16 + 256 fp64 atomics into the L2-resident G / G x G gradient arrays.
__constant__ double c_gh_x[20];   // Gauss-Hermite nodes / weights for the weight function exp(-x^2) (set by the host)
__constant__ double c_gh_w[20];

__device__ __forceinline__ double log_ndtr(double z) {
  // log Phi(z), stable in the lower tail: Phi(z) = erfc(-z / sqrt 2) / 2, erfcx for z << 0
  if (z > -5.0) return log(0.5 * erfc(-z * 0.70710678118654752440));
  const double t = -z * 0.70710678118654752440;
  return log(0.5 * erfcx(t)) - t * t;
}
__device__ __forceinline__ double pdf_over_cdf(double z) {
  // phi(z) / Phi(z) (inverse Mills ratio of -z), stable in the lower tail
  if (z > -5.0) return 0.39894228040143267794 * exp(-0.5 * z * z) / (0.5 * erfc(-z * 0.70710678118654752440));
  return 0.79788456080286535588 / erfcx(-z * 0.70710678118654752440);
}

__global__ void __launch_bounds__(128)
vgp_loglik_kernel(const double* __restrict__ X, const double* __restrict__ y, int n, double g0, double h, int gs,
                  double cmean, const double* __restrict__ m, const double* __restrict__ S, double* __restrict__ ell_out,
                  double* __restrict__ g_m, double* __restrict__ g_S, double* __restrict__ g_c) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0.0, dmu = 0.0;
  if (p < n) {
    int idx[16];
    double w[16];
    stencil(X[2 * p], X[2 * p + 1], g0, h, gs, idx, w);
    const int G = gs * gs;
    double mu = cmean, s2 = 0.0;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      mu = fma(w[a], m[idx[a]], mu);
      double r = 0.0;
#pragma unroll
      for (int c = 0; c < 16; ++c) r = fma(w[c], S[(size_t)idx[a] * G + idx[c]], r);
      s2 = fma(w[a], r, s2);
    }
    s2 = fmax(s2, 0.0);
    const double sd = sqrt(s2), yy = y[p];
    double ds2 = 0.0;
    for (int k = 0; k < 20; ++k) {
      const double z = yy * (mu + 1.41421356237309504880 * sd * c_gh_x[k]);
      const double r = pdf_over_cdf(z);
      const double wk = c_gh_w[k] * 0.56418958354775628695;   // / sqrt(pi)
      e = fma(wk, log_ndtr(z), e);
      dmu = fma(wk, yy * r, dmu);
      ds2 = fma(wk, -0.5 * yy * yy * r * (z + r), ds2);
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      atomicAdd(g_m + idx[a], dmu * w[a]);
#pragma unroll
      for (int c = 0; c < 16; ++c) atomicAdd(g_S + (size_t)idx[a] * G + idx[c], ds2 * w[a] * w[c]);
    }
  }
  // block-reduce the scalar terms (expected log-likelihood, gradient w.r.t. the constant mean)
  __shared__ double sh[2][4];
  for (int o = 16; o > 0; o >>= 1) {
    e += __shfl_xor_sync(0xffffffffu, e, o);
    dmu += __shfl_xor_sync(0xffffffffu, dmu, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = e; sh[1][threadIdx.x >> 5] = dmu; }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(ell_out, sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]);
    if (g_c) atomicAdd(g_c, sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3]);
  }
}

// predictive class probability E_q[Phi(f)] = Phi(mu / sqrt(1 + s2)) at the query pixels, plus the latent moments
__global__ void __launch_bounds__(128)
vgp_predict_kernel(const double* __restrict__ Xq, int mq, double g0, double h, int gs, double cmean,
                   const double* __restrict__ m, const double* __restrict__ S, double* __restrict__ prob,
                   double* __restrict__ mu_out, double* __restrict__ var_out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= mq) return;
  int idx[16];
  double w[16];
  stencil(Xq[2 * q], Xq[2 * q + 1], g0, h, gs, idx, w);
  const int G = gs * gs;
  double mu = cmean, s2 = 0.0;
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    mu = fma(w[a], m[idx[a]], mu);
    double r = 0.0;
#pragma unroll
    for (int c = 0; c < 16; ++c) r = fma(w[c], S[(size_t)idx[a] * G + idx[c]], r);
    s2 = fma(w[a], r, s2);
  }
  s2 = fmax(s2, 0.0);
  if (prob) prob[q] = 0.5 * erfc(-(mu / sqrt(1.0 + s2)) * 0.70710678118654752440);
  if (mu_out) mu_out[q] = mu;
  if (var_out) var_out[q] = s2;
}

// heat[p] += sum_i y[i] * [masks[i][p] == on]   — masks u8 [N][P] (the PNG side channel read back, 0 / 255)
__global__ void __launch_bounds__(256)
heatmap_pixels_kernel(const uint8_t* __restrict__ masks, const float* __restrict__ y, int N, int P, int on,
                      float* __restrict__ heat) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float acc = 0.f;
  for (int i = blockIdx.y; i < N; i += gridDim.y)
    if (masks[(size_t)i * P + p] == (uint8_t)on) acc += y[i];
  atomicAdd(heat + p, acc);
}

}  // namespace
}  // namespace nib

extern "C" {

int nib_ski_accumulate(const double* d_X, const double* d_y, int n, double grid0, double spacing, int grid_size,
                       double const_mean, double* d_A, double* d_b, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(d_X && d_y && d_A && d_b && n > 0 && grid_size >= 4 && spacing > 0.0, "nib_ski_accumulate: bad arguments");
  ski_accumulate_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_X, d_y, n, grid0, spacing, grid_size,
                                                                             const_mean, d_A, d_b);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_ski_predict(const double* d_Xq, int m, double grid0, double spacing, int grid_size, double const_mean,
                    const double* d_mean_u, const double* d_Gm, double noise, int add_noise, double* d_mean,
                    double* d_var, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(d_Xq && d_mean_u && m > 0 && grid_size >= 4 && spacing > 0.0, "nib_ski_predict: bad arguments");
  NIB_REQUIRE(d_var == nullptr || d_Gm != nullptr, "nib_ski_predict: variance needs d_Gm");
  ski_predict_kernel<<<ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(d_Xq, m, grid0, spacing, grid_size, const_mean,
                                                                        d_mean_u, d_Gm, noise, add_noise, d_mean, d_var);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

static int vgp_set_quadrature() {
  // 20-point Gauss-Hermite rule (physicists': weight exp(-x^2)); positive half, mirrored.  Abramowitz & Stegun 25.10.
  static bool done = false;
  if (done) return NIB_OK;
  static const double xh[10] = {0.2453407083009012, 0.7374737285453944, 1.2340762153953231, 1.7385377121165861,
                                2.2549740020892757, 2.7888060584281305, 3.3478545673832163, 3.9447640401156252,
                                4.6036824495507442, 5.3874808900112328};
  static const double wh[10] = {4.622436696006101e-01, 2.866755053628341e-01, 1.090172060200233e-01, 2.481052088746361e-02,
                                3.243773342237862e-03, 2.283386360163540e-04, 7.802556478532064e-06, 1.086069370769282e-07,
                                4.399340992273181e-10, 2.229393645534151e-13};
  double x[20], w[20];
  for (int i = 0; i < 10; ++i) { x[i] = -xh[9 - i]; w[i] = wh[9 - i]; x[10 + i] = xh[i]; w[10 + i] = wh[i]; }
  NIB_CUDA(cudaMemcpyToSymbol(nib::c_gh_x, x, sizeof(x)));
  NIB_CUDA(cudaMemcpyToSymbol(nib::c_gh_w, w, sizeof(w)));
  done = true;
  return NIB_OK;
}

int nib_vgp_loglik_grad(const double* d_X, const double* d_y, int n, double grid0, double spacing, int grid_size,
                        double const_mean, const double* d_m, const double* d_S, double* d_ell, double* d_grad_m,
                        double* d_grad_S, double* d_grad_c, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(d_X && d_y && d_m && d_S && d_ell && d_grad_m && d_grad_S && n > 0 && grid_size >= 4 && spacing > 0.0,
              "nib_vgp_loglik_grad: bad arguments");
  int rc = vgp_set_quadrature();
  if (rc != NIB_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t G = (size_t)grid_size * grid_size;
  NIB_CUDA(cudaMemsetAsync(d_ell, 0, sizeof(double), st));
  NIB_CUDA(cudaMemsetAsync(d_grad_m, 0, G * sizeof(double), st));
  NIB_CUDA(cudaMemsetAsync(d_grad_S, 0, G * G * sizeof(double), st));
  if (d_grad_c) NIB_CUDA(cudaMemsetAsync(d_grad_c, 0, sizeof(double), st));
  vgp_loglik_kernel<<<ceil_div(n, 128), 128, 0, st>>>(d_X, d_y, n, grid0, spacing, grid_size, const_mean, d_m, d_S, d_ell,
                                                      d_grad_m, d_grad_S, d_grad_c);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_vgp_predict(const double* d_Xq, int m, double grid0, double spacing, int grid_size, double const_mean,
                    const double* d_m, const double* d_S, double* d_prob, double* d_mu, double* d_var, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(d_Xq && d_m && d_S && m > 0 && grid_size >= 4 && spacing > 0.0, "nib_vgp_predict: bad arguments");
  vgp_predict_kernel<<<ceil_div(m, 128), 128, 0, (cudaStream_t)stream>>>(d_Xq, m, grid0, spacing, grid_size, const_mean,
                                                                         d_m, d_S, d_prob, d_mu, d_var);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_heatmap_pixels(const uint8_t* d_masks, const float* d_y, int N, int P, int on_value, float* d_heat, void* stream) {
  NIB_DEVICE_OR_FAIL();
  using namespace nib;
  NIB_REQUIRE(d_masks && d_y && d_heat && N > 0 && P > 0, "nib_heatmap_pixels: bad arguments");
  int slices = ceil_div(N, 64);
  if (slices > 64) slices = 64;
  dim3 grid(ceil_div(P, 256), slices);
  heatmap_pixels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_masks, d_y, N, P, on_value, d_heat);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

}  // extern "C"
