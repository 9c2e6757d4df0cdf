"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the two scikit-learn metrics the reference's evaluation
loop calls per video (test_video_segment_point.py:253-257, 299-303):

    fpr, tpr, _ = metrics.roc_curve(gt_labels, pred_scores, pos_label=1);  metrics.auc(fpr, tpr)
    metrics.average_precision_score(gt_labels, pred_scores)

scikit-learn is a third-party dependency of the reference, UNPINNED in its requirements.txt:16; the published algorithm
(sklearn/metrics/_ranking.py: _binary_clf_curve, roc_curve, auc, precision_recall_curve, average_precision_score) is
restated here and pinned by tests/test_oracle.py against the scikit-learn installed in this image (1.9.0).  Also restates
the reference loop's grouping of clips into videos, including its double count of every video's first clip (:284-292).
Only tests may import this.
"""
import numpy as np


def _binary_clf_curve(y_true, y_score):
    """Cumulative true / false positives at every DISTINCT score, scores descending (stable sort)."""
    y_true = (np.asarray(y_true) == 1)
    y_score = np.asarray(y_score, dtype=np.float64)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[order], y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    return fps, tps


def roc_auc(y_true, y_score):
    """auc(*roc_curve(...)[:2]): trapezoid rule over (fpr, tpr) with (0, 0) prepended; nan when one class is absent.
    (roc_curve's drop_intermediate only removes collinear points, which leaves the area unchanged.)"""
    fps, tps = _binary_clf_curve(y_true, y_score)
    fps, tps = np.r_[0.0, fps], np.r_[0.0, tps]
    if fps[-1] <= 0 or tps[-1] <= 0:
        return float("nan")
    fpr, tpr = fps / fps[-1], tps / tps[-1]
    return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0))


def average_precision(y_true, y_score):
    """-sum(diff(recall) * precision[:-1]) over precision_recall_curve's points (recall descending, final (1, 0) point
    appended); without positives recall is defined as 1 everywhere and the score is 0."""
    fps, tps = _binary_clf_curve(y_true, y_score)
    precision = tps / (tps + fps)
    recall = np.ones_like(tps) if tps[-1] == 0 else tps / tps[-1]
    precision, recall = np.r_[precision[::-1], 1.0], np.r_[recall[::-1], 0.0]
    return float(-np.sum(np.diff(recall) * precision[:-1]))


def reference_video_groups(vids):
    """Clip indices per video exactly as the loop at test_video_segment_point.py:250-296 accumulates them: a new list is
    started WITH the first clip of a video and the same clip is appended again right after, so it counts twice."""
    groups = []
    prev = object()
    for i, v in enumerate(vids):
        if v != prev:
            groups.append([i])
            prev = v
        groups[-1].append(i)
    return groups
