// localize.cu — "next" row 1 tail: heat map -> 8-bit map -> threshold -> largest connected region's bounding box.
//
// Replaces, on the device,
//   * `g = H - H.min(); g = g / g.max(); g *= 255; np.array(g, dtype=np.uint8)`   (float64 numpy, truncation)
//     generate_gp_training_data_imagenet.py:519-525, bayesian_active_learning_imagenet.py:349-355, gp_regression.py:96-104
//   * utils.py:96-109 `generate_boundingbox`: cv2.threshold(gray, t, 255, THRESH_BINARY) -> cv2.findContours(RETR_EXTERNAL)
//     -> the cv2.boundingRect with the largest w*h (first one on ties, in OpenCV's contour order).
// The reference then feeds the box to utils.py:114-142 `generate_IOU` (scalar arithmetic: stays on the host).
//
// The external contours of a binary image are the outlines of its 8-connected foreground components; a component nested
// in another one's hole is not "external", but its box lies inside the outer box and can never be the largest.  So:
// 8-connected component labelling (one CTA, labels in shared memory, min-label propagation with pointer jumping) + one
// bounding box per component (atomics into a per-stream scratch table) + arg-max of w*h.  OpenCV returns contours in
// reverse order of their first pixel in raster order and the reference keeps the first strict maximum, so ties go to the
// component whose first raster pixel comes LAST.
#include "common.cuh"

namespace nib {

__global__ void __launch_bounds__(1024)
heat_normalize_u8_kernel(const float* __restrict__ heat, int P, uint8_t* __restrict__ gray, double* __restrict__ stats) {
  __shared__ double s_mn[32], s_mx[32];
  __shared__ double g_mn, g_mx;
  double mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const double v = (double)heat[i];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = INFINITY, b = -INFINITY;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a = fmin(a, s_mn[i]); b = fmax(b, s_mx[i]); }
    g_mn = a;
    g_mx = __dsub_rn(b, a);          // max of (x - min) == fl(max - min): subtraction is monotone
    if (stats) { stats[0] = a; stats[1] = b; }
  }
  __syncthreads();
  const double a = g_mn, b = g_mx;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    double v = __dsub_rn((double)heat[i], a);
    v = __ddiv_rn(v, b);             // 0/0 = NaN for a constant map, as numpy; NaN -> uint8 is 0 on the reference's platforms
    v = __dmul_rn(v, 255.0);
    gray[i] = (v == v) ? (uint8_t)v : (uint8_t)0;
  }
}

// out: {x, y, w, h, number of components, pixels above threshold}; all zero when nothing exceeds the threshold
__global__ void __launch_bounds__(1024)
threshold_bbox_kernel(const uint8_t* __restrict__ gray, int H, int W, int threshold, int* __restrict__ tab, int* __restrict__ out) {
  extern __shared__ unsigned short lab[];      // 0 = background, else 1 + raster index of the component's current root
  __shared__ int changed;
  __shared__ int best_area, best_root, ncomp, nfg;
  const int P = H * W;
  for (int p = threadIdx.x; p < P; p += blockDim.x) lab[p] = gray[p] > threshold ? (unsigned short)(p + 1) : 0;
  if (threadIdx.x == 0) { best_area = 0; best_root = -1; ncomp = 0; nfg = 0; }
  __syncthreads();
  for (int iter = 0; iter < P; ++iter) {          // converges in far fewer rounds; P bounds the longest possible chain
    if (threadIdx.x == 0) changed = 0;
    __syncthreads();
    bool ch = false;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      unsigned short l = lab[p];
      if (!l) continue;
      const int y = p / W, x = p - y * W;
      unsigned short m = l;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          const unsigned short q = lab[yy * W + xx];
          if (q && q < m) m = q;
        }
      }
      // pointer jumping: follow the chain of roots a few steps (labels only ever decrease, stale reads are harmless)
      for (int j = 0; j < 4; ++j) {
        const unsigned short r = lab[m - 1];
        if (r && r < m) m = r; else break;
      }
      if (m < l) { lab[p] = m; ch = true; }
    }
    if (ch) changed = 1;
    __syncthreads();
    const int c = changed;
    __syncthreads();
    if (!c) break;
  }
  // bounding boxes: tab[4][P] = xmin, ymin, xmax, ymax per root (only root rows are initialised / read)
  int* xmin = tab; int* ymin = tab + P; int* xmax = tab + 2 * P; int* ymax = tab + 3 * P;
  for (int p = threadIdx.x; p < P; p += blockDim.x)
    if (lab[p] == p + 1) { xmin[p] = W; ymin[p] = H; xmax[p] = -1; ymax[p] = -1; }
  __syncthreads();
  int fg = 0;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const unsigned short l = lab[p];
    if (!l) continue;
    ++fg;
    const int r = l - 1, y = p / W, x = p - y * W;
    atomicMin(&xmin[r], x); atomicMax(&xmax[r], x);
    atomicMin(&ymin[r], y); atomicMax(&ymax[r], y);
  }
  atomicAdd(&nfg, fg);
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x)
    if (lab[p] == p + 1) {
      atomicAdd(&ncomp, 1);
      atomicMax(&best_area, (xmax[p] - xmin[p] + 1) * (ymax[p] - ymin[p] + 1));
    }
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x)
    if (lab[p] == p + 1 && (xmax[p] - xmin[p] + 1) * (ymax[p] - ymin[p] + 1) == best_area) atomicMax(&best_root, p);
  __syncthreads();
  if (threadIdx.x == 0) {
    const int r = best_root;
    if (r >= 0) {
      out[0] = xmin[r]; out[1] = ymin[r]; out[2] = xmax[r] - xmin[r] + 1; out[3] = ymax[r] - ymin[r] + 1;
    } else {
      out[0] = out[1] = out[2] = out[3] = 0;
    }
    out[4] = ncomp;
    out[5] = nfg;
  }
}

}  // namespace nib

extern "C" {

int nib_heat_normalize_u8(const float* d_heat, int P, uint8_t* d_gray, double* d_minmax, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_heat && d_gray && P > 0, "nib_heat_normalize_u8: bad arguments");
  nib::heat_normalize_u8_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_heat, P, d_gray, d_minmax);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

int nib_threshold_bbox(const uint8_t* d_gray, int H, int W, int threshold, int32_t* d_box, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_gray && d_box && H > 0 && W > 0, "nib_threshold_bbox: bad arguments");
  const long long P = (long long)H * W;
  NIB_REQUIRE(P <= 65535, "nib_threshold_bbox: %d x %d exceeds the 65535-pixel limit of the shared-memory labelling", H, W);
  cudaStream_t st = (cudaStream_t)stream;
  int* tab = nullptr;
  int rc = nib::stream_scratch(nib::SCRATCH_BBOX, st, (size_t)4 * P * sizeof(int), (size_t)4 * 65535 * sizeof(int),
                               reinterpret_cast<void**>(&tab));
  if (rc != NIB_OK) return rc;
  const size_t smem = (size_t)P * sizeof(unsigned short);
  static bool attr = false;
  if (!attr) {
    NIB_CUDA(cudaFuncSetAttribute(nib::threshold_bbox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65535 * 2 + 16));
    attr = true;
  }
  nib::threshold_bbox_kernel<<<1, 1024, smem, st>>>(d_gray, H, W, threshold, tab, d_box);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

}  // extern "C"
