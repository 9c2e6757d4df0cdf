// Device-side post-processing of boundary scores: only matters when hundreds of thousands of
// clips are scored per call (BASELINE.json configs[3]); the results must equal the reference's Python bit for bit.
//
//   labels      pred_label = argmax(logits) via topk(1) (test_video_segment_point.py:201-203): 1 iff logit1 > logit0
//   cut points  convert_clip_label2cut_point (eval_utils/eval_utils.py:3-18): every maximal run of 1-labels that is
//               FOLLOWED by a 0 yields round((i_begin*2*max_offset + (i_end-1)*2*max_offset + T - 1) / 2) with
//               Python's round (half to even); a trailing run is dropped
//   hits        calculate_pr (eval_utils.py:21-92): for every point of list A, is there a point of list B at distance
//               0 / <= 3 / <= 5 — the six counts the host divides into recall and precision
//   AUC / AP    sklearn.metrics.roc_curve + auc and average_precision_score per video
//               (test_video_segment_point.py:253-255, 299-301) without a sort: with TP_i / FP_i = the positives /
//               negatives scored >= clip i,  AP = 1/P * sum over positives i of TP_i / (TP_i + FP_i)  and
//               AUC = sum over positives of (negatives scored lower + half the negatives tied) / (P * N)
//               — the step-wise sums sklearn takes over distinct thresholds, regrouped per positive clip
// One CTA per video; videos are contiguous clip ranges [offsets[v], offsets[v+1]).
#include "kernels.cuh"
#include "launch.cuh"

namespace vcg {

namespace {

__global__ void __launch_bounds__(256) cut_points_kernel(const float* __restrict__ logits, const int32_t* __restrict__ offsets,
                                                         int clip_frames, int max_offset, int cap,
                                                         int32_t* __restrict__ labels_out, int32_t* __restrict__ cuts,
                                                         int32_t* __restrict__ counts) {
  pdl_enter();
  const int v = blockIdx.x;
  const int lo = offsets[v], n = offsets[v + 1] - lo;
  __shared__ int s_base;
  __shared__ int s_warp[8];
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto label = [&](int i) { return logits[2L * (lo + i) + 1] > logits[2L * (lo + i)] ? 1 : 0; };
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    int lab = 0, is_end = 0;
    if (i < n) {
      lab = label(i);
      if (labels_out) labels_out[lo + i] = lab;
      is_end = (lab == 0 && i > 0 && label(i - 1) == 1);   // clip i closes a run of 1s
    }
    // order-preserving compaction of the run ends
    const unsigned bal = __ballot_sync(0xffffffffu, is_end);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (is_end) {
      int b = i - 1;
      while (b > 0 && label(b - 1) == 1) --b;               // first clip of the run
      const int stride = 2 * max_offset;
      const int S = b * stride + (i - 1) * stride + clip_frames - 1;   // begin_sec + end_sec - 1 >= 0
      const int k = S >> 1;
      const int cut = (S & 1) ? ((k & 1) ? k + 1 : k) : k;             // round half to even
      const int idx = before + __popc(bal & ((1u << lane) - 1));
      if (idx < cap) cuts[static_cast<long>(v) * cap + idx] = cut;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 8; ++w) t += s_warp[w];
      s_base += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[v] = s_base;     // may exceed cap: the caller re-runs with a larger buffer
}

// hits[v] = {A->B exact, <=3, <=5, B->A exact, <=3, <=5}; A = ground truth, B = predictions of video v
__global__ void __launch_bounds__(128) pr_hits_kernel(const int32_t* __restrict__ gt, const int32_t* __restrict__ gt_off,
                                                      const int32_t* __restrict__ pred, const int32_t* __restrict__ pred_off,
                                                      int32_t* __restrict__ hits) {
  pdl_enter();
  const int v = blockIdx.x;
  const int32_t* a = gt + gt_off[v];
  const int na = gt_off[v + 1] - gt_off[v];
  const int32_t* b = pred + pred_off[v];
  const int nb = pred_off[v + 1] - pred_off[v];
  __shared__ int s_hits[6];
  if (threadIdx.x < 6) s_hits[threadIdx.x] = 0;
  __syncthreads();
  for (int dir = 0; dir < 2; ++dir) {
    const int32_t* x = dir ? b : a;
    const int32_t* y = dir ? a : b;
    const int nx = dir ? nb : na, ny = dir ? na : nb;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) {
      int best = 1 << 30;
      for (int j = 0; j < ny; ++j) best = min(best, abs(x[i] - y[j]));
      if (best == 0) atomicAdd(&s_hits[dir * 3], 1);
      if (best <= 3) atomicAdd(&s_hits[dir * 3 + 1], 1);
      if (best <= 5) atomicAdd(&s_hits[dir * 3 + 2], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) hits[v * 6 + threadIdx.x] = s_hits[threadIdx.x];
}

// One warp per positive clip (round-robin), lanes stride over the video's clips.  Integer counts are exact; the double
// sums are taken in a fixed order (clip order per warp, then warp order), so results are run-to-run identical.
__global__ void __launch_bounds__(256) auc_ap_kernel(const float* __restrict__ scores, const int32_t* __restrict__ labels,
                                                     const int32_t* __restrict__ offsets, double* __restrict__ auc,
                                                     double* __restrict__ ap) {
  pdl_enter();
  const int v = blockIdx.x;
  const int lo = offsets[v], n = offsets[v + 1] - lo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __shared__ double s_ap[8];
  __shared__ long long s_u2[8];
  __shared__ int s_pos[8];
  double ap_sum = 0.0;
  long long u2 = 0;      // 2 * (negatives below) + (negatives tied), summed over this warp's positives
  int n_pos = 0;
  for (int i = warp; i < n; i += nw) {
    if (labels[lo + i] != 1) continue;          // warp-uniform
    const float si = scores[lo + i];
    int tp = 0, fp = 0, eq = 0, lt = 0;
    for (int j = lane; j < n; j += 32) {
      const float sj = scores[lo + j];
      const bool pos = labels[lo + j] == 1;
      tp += (pos && sj >= si);
      fp += (!pos && sj >= si);
      eq += (!pos && sj == si);
      lt += (!pos && sj < si);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      tp += __shfl_xor_sync(0xffffffffu, tp, d);
      fp += __shfl_xor_sync(0xffffffffu, fp, d);
      eq += __shfl_xor_sync(0xffffffffu, eq, d);
      lt += __shfl_xor_sync(0xffffffffu, lt, d);
    }
    ap_sum += static_cast<double>(tp) / static_cast<double>(tp + fp);
    u2 += 2LL * lt + eq;
    ++n_pos;
  }
  if (lane == 0) { s_ap[warp] = ap_sum; s_u2[warp] = u2; s_pos[warp] = n_pos; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    long long u = 0;
    int P = 0;
    for (int w = 0; w < nw; ++w) { a += s_ap[w]; u += s_u2[w]; P += s_pos[w]; }
    const int N = n - P;
    // sklearn: no positives -> AP 0.0 (recall defined as 1 everywhere); a single class -> ROC AUC undefined (nan)
    ap[v] = P > 0 ? a / static_cast<double>(P) : 0.0;
    auc[v] = (P > 0 && N > 0) ? static_cast<double>(u) / (2.0 * static_cast<double>(P) * static_cast<double>(N))
                              : __longlong_as_double(0x7ff8000000000000LL);
  }
}

}  // namespace

void launch_auc_ap(const float* scores, const int32_t* labels, const int32_t* offsets, int n_videos, double* auc, double* ap,
                   cudaStream_t s) {
  if (n_videos == 0) return;
  launch_pdl(auc_ap_kernel, n_videos, 256, 0, s, scores, labels, offsets, auc, ap);
}

void launch_cut_points(const float* logits, const int32_t* offsets, int n_videos, int clip_frames, int max_offset, int cap,
                       int32_t* labels_out, int32_t* cuts, int32_t* counts, cudaStream_t s) {
  if (n_videos == 0) return;
  launch_pdl(cut_points_kernel, n_videos, 256, 0, s, logits, offsets, clip_frames, max_offset, cap, labels_out, cuts, counts);
}

void launch_pr_hits(const int32_t* gt, const int32_t* gt_off, const int32_t* pred, const int32_t* pred_off, int n_videos,
                    int32_t* hits, cudaStream_t s) {
  if (n_videos == 0) return;
  launch_pdl(pr_hits_kernel, n_videos, 128, 0, s, gt, gt_off, pred, pred_off, hits);
}

}  // namespace vcg
