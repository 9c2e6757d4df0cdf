// Device-side post-processing of boundary scores: only matters when hundreds of thousands of
// clips are scored per call (BASELINE.json configs[3]); the results must equal the reference's Python bit for bit.
//
//   labels      pred_label = argmax(logits) via topk(1) (test_video_segment_point.py:201-203): 1 iff logit1 > logit0
//   cut points  convert_clip_label2cut_point (eval_utils/eval_utils.py:3-18): every maximal run of 1-labels that is
//               FOLLOWED by a 0 yields round((i_begin*2*max_offset + (i_end-1)*2*max_offset + T - 1) / 2) with
//               Python's round (half to even); a trailing run is dropped
//   hits        calculate_pr (eval_utils.py:21-92): for every point of list A, is there a point of list B at distance
//               0 / <= 3 / <= 5 — the six counts the host divides into recall and precision
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

}  // namespace

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
