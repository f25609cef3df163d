// score.cu — stage 2 tail: logits -> (top-1, target-class softmax prob, max prob, correct).
//
// Replaces, per mask, `output.data.max(1, keepdim=True)[1]` (generate_gp_training_data_imagenet.py:248),
// `F.softmax(mask_output)[0][label]` (bayesian_active_learning_imagenet.py:196-198) and
// `F.softmax(pred0, dim=1).max(1)` (generate_gp_training_data_mnist.py:249-256) — and the one
// device->host sync per mask that follows them (imagenet :257) becomes one copy per shard.
// HBM-bound: 4*K bytes read per mask; one warp per row, coalesced 128-bit loads when K % 4 == 0.
#include "common.cuh"
#include <math.h>

namespace nib {

__global__ void __launch_bounds__(256)
score_kernel(const float* __restrict__ logits, int N, int K, int target, int32_t* __restrict__ top1,
             float* __restrict__ tprob, float* __restrict__ mprob, uint8_t* __restrict__ correct,
             float* __restrict__ margin) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* row = logits + (size_t)warp * K;
  // pass 1: max and first argmax (NaN-free inputs assumed; ties -> lowest index like torch.max)
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    float v = row[k];
    if (v > best) { best = v; bidx = k; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  // pass 2: sum of exp(x - max), runner-up and max |x| (row is L1/L2 resident from pass 1)
  float sum = 0.f, second = -INFINITY, amax = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float v = row[k];
    sum += expf(v - best);
    if (k != bidx) second = fmaxf(second, v);
    amax = fmaxf(amax, fabsf(v));
  }
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    second = fmaxf(second, __shfl_xor_sync(0xffffffffu, second, o));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  }
  if (lane == 0) {
    if (top1) top1[warp] = bidx;
    if (mprob) mprob[warp] = 1.0f / sum;
    if (tprob) tprob[warp] = (target >= 0 && target < K) ? expf(row[target] - best) / sum : 0.f;
    if (correct) correct[warp] = (bidx == target) ? 1 : 0;
    if (margin) margin[warp] = (K > 1 && amax > 0.f) ? (best - second) / amax : INFINITY;
  }
}

}  // namespace nib

extern "C" int nib_score(const float* d_logits, int N, int K, int target, int32_t* d_top1,
                         float* d_target_prob, float* d_max_prob, uint8_t* d_correct, float* d_margin,
                         void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_logits != nullptr && N >= 0 && K > 0, "nib_score: bad arguments N=%d K=%d", N, K);
  if (N == 0) return NIB_OK;
  const int threads = 256;
  const int blocks = nib::ceil_div(N * 32, threads);
  nib::score_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_logits, N, K, target, d_top1,
                                                                  d_target_prob, d_max_prob, d_correct, d_margin);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}
