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
             float* __restrict__ margin, float2* __restrict__ table) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  const float* row = logits + (size_t)warp * K;
  // pass 1: max and first argmax (ties -> lowest index like torch.max; a row of NaN / -inf only reports index 0)
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
  if (bidx == 0x7fffffff) bidx = 0;
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
    // (target prob, top-1 as float): the layout of the all-gather send buffer; class ids < 2^24 are exact in fp32
    if (table) table[warp] = make_float2((target >= 0 && target < K) ? expf(row[target] - best) / sum : 0.f, (float)bidx);
  }
}

}  // namespace nib

extern "C" int nib_score_table(const float* d_logits, int N, int K, int target, int32_t* d_top1,
                               float* d_target_prob, float* d_max_prob, uint8_t* d_correct, float* d_margin,
                               float* d_table, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_logits != nullptr && N >= 0 && K > 0, "nib_score: bad arguments N=%d K=%d", N, K);
  if (N == 0) return NIB_OK;
  const int threads = 256;
  const int blocks = nib::ceil_div(N * 32, threads);
  nib::score_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(d_logits, N, K, target, d_top1, d_target_prob,
                                                                  d_max_prob, d_correct, d_margin,
                                                                  reinterpret_cast<float2*>(d_table));
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

extern "C" int nib_score(const float* d_logits, int N, int K, int target, int32_t* d_top1,
                         float* d_target_prob, float* d_max_prob, uint8_t* d_correct, float* d_margin,
                         void* stream) {
  return nib_score_table(d_logits, N, K, target, d_top1, d_target_prob, d_max_prob, d_correct, d_margin, nullptr, stream);
}

// ---- tie policy of the bf16 path (device side, no host synchronisation) -----------------------------------------
// bf16 logits agree with the reference's fp32 logits to a stated relative tolerance, so a row whose top-2 margin is
// inside that band may flip its arg-max (generate_gp_training_data_imagenet.py:248-257 compares the arg-max with the
// target).  Those rows are re-scored by an fp32 copy of the classifier.  Everything happens on the stream:
//   nib_score / nib_score_table  ->  margin per row (+ the (prob, top1) table the all-gather sends)
//   nib_tie_compact              ->  ascending indices of the rows with margin < threshold (first `cap` of them), their
//                                    selection words gathered into a cap-row table, the total count on the device
//   fp32 forward of the cap-row table with the device-side count as the live batch size (nib_net_set_dynamic_batch)
//   nib_tie_scatter              ->  refined (prob, top1, ...) written back over the bf16 results
namespace nib {

// One block; rows are taken in chunks of blockDim.x with a running base, so the output order is ascending and the set
// of refined rows is deterministic (the first `cap` in index order) even when more than `cap` rows qualify.
__global__ void __launch_bounds__(1024)
tie_compact_kernel(const float* __restrict__ margin, int N, float threshold, const uint64_t* __restrict__ sel, int words,
                   int cap, int32_t* __restrict__ idx, uint64_t* __restrict__ sel_out, int32_t* __restrict__ count,
                   long long* __restrict__ totals) {
  __shared__ int warp_sums[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int r0 = 0; r0 < N; r0 += blockDim.x) {
    const int r = r0 + (int)threadIdx.x;
    const bool hit = r < N && margin[r] < threshold;     // NaN margins never qualify (comparison false)
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const int s = warp_sums[w];
      if (w < warp) before += s;
      total += s;
    }
    const int base = base_s;
    if (hit) {
      const int slot = base + before + __popc(ballot & ((1u << lane) - 1u));
      if (slot < cap) {
        idx[slot] = r;
        for (int w = 0; w < words; ++w) sel_out[(size_t)slot * words + w] = sel[(size_t)r * words + w];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + total;
    __syncthreads();
  }
  const int found = base_s;
  for (int s = (found < cap ? found : cap) + (int)threadIdx.x; s < cap; s += blockDim.x) {
    idx[s] = -1;
    for (int w = 0; w < words; ++w) sel_out[(size_t)s * words + w] = 0ull;
  }
  if (threadIdx.x == 0) {
    *count = found;
    if (totals) {   // running (qualified, did-not-fit) row counts over all calls, read by the host when it wants them
      totals[0] += found;
      totals[1] += found > cap ? found - cap : 0;
    }
  }
}

__global__ void tie_scatter_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ count, int cap,
                                   const int32_t* __restrict__ r_top1, const float* __restrict__ r_tprob,
                                   const float* __restrict__ r_mprob, const uint8_t* __restrict__ r_correct,
                                   int32_t* __restrict__ top1, float* __restrict__ tprob, float* __restrict__ mprob,
                                   uint8_t* __restrict__ correct, float2* __restrict__ table) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int live = min(*count, cap);
  if (s >= live) return;
  const int r = idx[s];
  if (r < 0) return;
  if (top1 && r_top1) top1[r] = r_top1[s];
  if (tprob && r_tprob) tprob[r] = r_tprob[s];
  if (mprob && r_mprob) mprob[r] = r_mprob[s];
  if (correct && r_correct) correct[r] = r_correct[s];
  if (table && r_tprob && r_top1) table[r] = make_float2(r_tprob[s], (float)r_top1[s]);
}

}  // namespace nib

extern "C" int nib_tie_compact(const float* d_margin, int N, float threshold, const uint64_t* d_sel, int words, int cap,
                               int32_t* d_idx, uint64_t* d_sel_out, int32_t* d_count, long long* d_totals, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_margin && d_sel && d_idx && d_sel_out && d_count, "nib_tie_compact: null pointer");
  NIB_REQUIRE(N >= 0 && words > 0 && cap > 0, "nib_tie_compact: bad arguments N=%d words=%d cap=%d", N, words, cap);
  nib::tie_compact_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(d_margin, N, threshold, d_sel, words, cap, d_idx,
                                                               d_sel_out, d_count, d_totals);
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}

extern "C" int nib_tie_scatter(const int32_t* d_idx, const int32_t* d_count, int cap, const int32_t* r_top1,
                               const float* r_target_prob, const float* r_max_prob, const uint8_t* r_correct,
                               int32_t* d_top1, float* d_target_prob, float* d_max_prob, uint8_t* d_correct,
                               float* d_table, void* stream) {
  NIB_DEVICE_OR_FAIL();
  NIB_REQUIRE(d_idx && d_count && cap > 0, "nib_tie_scatter: bad arguments");
  nib::tie_scatter_kernel<<<nib::ceil_div(cap, 128), 128, 0, (cudaStream_t)stream>>>(
      d_idx, d_count, cap, r_top1, r_target_prob, r_max_prob, r_correct, d_top1, d_target_prob, d_max_prob, d_correct,
      reinterpret_cast<float2*>(d_table));
  NIB_LAUNCH_CHECK();
  return NIB_OK;
}
