"""Mirror of the reference's eval_utils/eval_utils.py: clip labels -> chapter cut points, and precision/recall.

Same names, arguments and results as video_chapter_generation/eval_utils/eval_utils.py:3-92 (pinned by
tests/golden/cut_points.npz, generated with the reference functions).  Pure host-side Python: O(clips) work.
"""


def convert_clip_label2cut_point(clip_label_array, clip_frame_num, max_offset):
    """Every maximal run of 1-labels that is followed by a 0 yields one cut point: the rounded (half-to-even, as
    Python's round) midpoint of the seconds covered by the run's clips; a trailing run is dropped."""
    stride = max_offset * 2
    cut_points = []
    run_start = None
    for i, label in enumerate(clip_label_array):
        if label == 1 and run_start is None:
            run_start = i
        elif label == 0 and run_start is not None:
            begin_sec = run_start * stride
            end_sec = (i - 1) * stride + clip_frame_num
            cut_points.append(round((begin_sec + end_sec - 1) / 2))
            run_start = None
    return cut_points


def _hit_counts(points, others):
    exact = within3 = within5 = 0
    for p in points:
        exact += any(p == o for o in others)
        within3 += any(abs(p - o) <= 3 for o in others)
        within5 += any(abs(p - o) <= 5 for o in others)
    return exact, within3, within5


def calculate_pr(gt_cut_points, pred_cut_points):
    """-> (recall, recall@3s, recall@5s, precision, precision@3s, precision@5s); precisions are None when nothing
    was predicted; an empty ground truth raises ZeroDivisionError exactly like the reference."""
    n_gt = len(gt_cut_points)
    r0, r3, r5 = _hit_counts(gt_cut_points, pred_cut_points)
    recall, recall_3, recall_5 = r0 / n_gt, r3 / n_gt, r5 / n_gt
    precision = precision_3 = precision_5 = None
    if len(pred_cut_points) > 0:
        n_pred = len(pred_cut_points)
        p0, p3, p5 = _hit_counts(pred_cut_points, gt_cut_points)
        precision, precision_3, precision_5 = p0 / n_pred, p3 / n_pred, p5 / n_pred
    return recall, recall_3, recall_5, precision, precision_3, precision_5
