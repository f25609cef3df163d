// segment.cu — superpixel label maps: the reference's once-per-image pre-step
//   segments = felzenszwalb(img_as_float(img), scale=100, sigma=0.5, min_size=50)
// (generate_gp_training_data_imagenet.py:183, ...mnist.py:187, ...cifar.py:293,
// bayesian_active_learning_imagenet.py:150,:263,:463).  scikit-image runs this in Cython on the host; it is a sorted-edge
// union-find (Kruskal order), inherently sequential and ~2e5 edges for a 224^2 image, so the native equivalent is host C++
// too (about a millisecond per image) — the label map is then uploaded once and every mask of the image reuses it.
// Algorithm restated from scikit-image's published implementation (see oracle/segmentation.py for the step list).
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace nib {
namespace {

// scipy.ndimage.correlate1d, mode='reflect', symmetric kernel: centre tap, then (x[l+j] + x[l-j]) * w[j] from the outermost
// pair inwards — the accumulation order of ni_filters.c, so the blur is bit-identical to ndi.gaussian_filter
void blur_axis(const std::vector<double>& in, std::vector<double>& out, int n_outer, int n, int n_inner,
               const std::vector<double>& w) {
  const int r = (int)w.size() / 2;
  auto refl = [n](int i) {
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
  };
  for (int o = 0; o < n_outer; ++o)
    for (int l = 0; l < n; ++l)
      for (int c = 0; c < n_inner; ++c) {
        const size_t base = (size_t)o * n * n_inner + c;
        double acc = in[base + (size_t)l * n_inner] * w[r];
        for (int j = -r; j < 0; ++j)
          acc += (in[base + (size_t)refl(l + j) * n_inner] + in[base + (size_t)refl(l - j) * n_inner]) * w[j + r];
        out[base + (size_t)l * n_inner] = acc;
      }
}

struct Forest {
  std::vector<int> parent, size;
  std::vector<double> cint;
  explicit Forest(int n) : parent(n), size(n, 1), cint(n, 0.0) { std::iota(parent.begin(), parent.end(), 0); }
  int find(int n) {
    int root = n;
    while (parent[root] != root) root = parent[root];
    while (parent[n] != root) { const int nx = parent[n]; parent[n] = root; n = nx; }
    return root;
  }
  int join(int a, int b) {   // the smaller root survives: labels end up numbered in raster order of first appearance
    const int keep = a < b ? a : b, drop = a < b ? b : a;
    parent[drop] = keep;
    size[keep] = size[a] + size[b];
    return keep;
  }
};

}  // namespace
}  // namespace nib

extern "C" int nib_felzenszwalb(const double* h_image, int H, int W, int C, double scale, double sigma, int min_size,
                                int32_t* h_labels, int* num_segments) {
  using namespace nib;
  NIB_REQUIRE(h_image && h_labels, "nib_felzenszwalb: null pointer");
  NIB_REQUIRE(H >= 2 && W >= 2 && C >= 1 && (long long)H * W < (1ll << 30), "nib_felzenszwalb: bad shape %dx%dx%d", H, W, C);
  NIB_REQUIRE(sigma >= 0.0 && scale >= 0.0 && min_size >= 0, "nib_felzenszwalb: negative parameter");
  const int n = H * W;
  std::vector<double> im(h_image, h_image + (size_t)n * C);
  if (sigma > 1e-15) {
    const int radius = (int)(4.0 * sigma + 0.5);      // truncate = 4.0
    std::vector<double> w(2 * radius + 1);
    double sum = 0.0;
    for (int x = -radius; x <= radius; ++x) sum += (w[x + radius] = std::exp(-0.5 / (sigma * sigma) * (double)(x * x)));
    for (double& v : w) v /= sum;
    std::vector<double> tmp(im.size());
    blur_axis(im, tmp, 1, H, W * C, w);   // axis 0
    blur_axis(tmp, im, H, W, C, w);       // axis 1
  }
  // 8-connected edges in scikit-image's block order: right, down, down-right, up-right
  struct Edge { int a, b; };
  std::vector<Edge> edges;
  std::vector<double> costs;
  edges.reserve((size_t)4 * n);
  costs.reserve((size_t)4 * n);
  auto add = [&](int ya, int xa, int yb, int xb) {
    const double* p = &im[((size_t)ya * W + xa) * C];
    const double* q = &im[((size_t)yb * W + xb) * C];
    double s = 0.0;
    for (int c = 0; c < C; ++c) { const double d = p[c] - q[c]; s += d * d; }
    edges.push_back({ya * W + xa, yb * W + xb});
    costs.push_back(std::sqrt(s));
  };
  for (int y = 0; y < H; ++y) for (int x = 1; x < W; ++x) add(y, x, y, x - 1);
  for (int y = 1; y < H; ++y) for (int x = 0; x < W; ++x) add(y, x, y - 1, x);
  for (int y = 1; y < H; ++y) for (int x = 1; x < W; ++x) add(y, x, y - 1, x - 1);
  for (int y = 1; y < H; ++y) for (int x = 0; x < W - 1; ++x) add(y, x, y - 1, x + 1);
  std::vector<int> order(edges.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int i, int j) { return costs[i] < costs[j]; });
  Forest f(n);
  const double k = scale / 255.0;
  for (int e : order) {
    const int s0 = f.find(edges[e].a), s1 = f.find(edges[e].b);
    if (s0 == s1) continue;
    const double i0 = f.cint[s0] + k / f.size[s0], i1 = f.cint[s1] + k / f.size[s1];
    if (costs[e] < (i0 < i1 ? i0 : i1)) f.cint[f.join(s0, s1)] = costs[e];
  }
  for (int e : order) {
    const int s0 = f.find(edges[e].a), s1 = f.find(edges[e].b);
    if (s0 == s1) continue;
    if (f.size[s0] < min_size || f.size[s1] < min_size) f.join(s0, s1);
  }
  // a root is the smallest pixel index of its component, so a raster scan meets roots in label order
  std::vector<int> label_of(n, -1);
  int S = 0;
  for (int i = 0; i < n; ++i) {
    const int r = f.find(i);
    if (label_of[r] < 0) label_of[r] = S++;
    h_labels[i] = label_of[r];
  }
  if (num_segments) *num_segments = S;
  return NIB_OK;
}
