"""Device-side post-processing (C ABI: vcg_op_cut_points / vcg_op_pr_hits / vcg_op_auc_ap).  Same results as the reference's Python
(eval_utils/eval_utils.py:3-92, test_video_segment_point.py:201-203) without moving per-clip data to the host: one CTA
per video.  Meant for the many-videos configuration (BASELINE.json configs[3]); for a single video the Python mirror
in eval_utils/eval_utils.py is just as good."""
import torch

from . import binding as _b


def _stream():
    return torch.cuda.current_stream().cuda_stream


def cut_points_device(logits, video_offsets, clip_frames, max_offset=2):
    """logits [N,2] fp32 CUDA, video_offsets int32 [V+1] -> (labels [N] int32 CUDA, list of per-video cut-point lists)."""
    if not logits.is_cuda:
        raise RuntimeError("vcg_b200 post-processing needs CUDA tensors: there is no CPU fallback")
    logits = logits.float().contiguous()
    off = video_offsets.to(device=logits.device, dtype=torch.int32).contiguous()
    V = off.numel() - 1
    labels = torch.empty(logits.shape[0], dtype=torch.int32, device=logits.device)
    counts = torch.empty(V, dtype=torch.int32, device=logits.device)
    lib = _b.load_library()
    cap = 64
    while True:
        cuts = torch.empty(V, cap, dtype=torch.int32, device=logits.device)
        _b.check(lib.vcg_op_cut_points(logits.data_ptr(), off.data_ptr(), V, clip_frames, max_offset, cap,
                                       labels.data_ptr(), cuts.data_ptr(), counts.data_ptr(), _stream()))
        c = counts.cpu()
        if V == 0 or int(c.max()) <= cap:
            break
        cap = int(c.max())
    cuts = cuts.cpu()
    return labels, [cuts[v, :int(c[v])].tolist() for v in range(V)]


def pr_hits_device(gt_lists, pred_lists, device="cuda"):
    """Per video (recall, recall@3, recall@5, precision, precision@3, precision@5) exactly like calculate_pr."""
    def pack(lists):
        off = [0]
        for l in lists:
            off.append(off[-1] + len(l))
        flat = [x for l in lists for x in l] or [0]
        return (torch.tensor(flat, dtype=torch.int32, device=device), torch.tensor(off, dtype=torch.int32, device=device))
    gt, gt_off = pack(gt_lists)
    pr, pr_off = pack(pred_lists)
    V = len(gt_lists)
    hits = torch.empty(V, 6, dtype=torch.int32, device=device)
    lib = _b.load_library()
    _b.check(lib.vcg_op_pr_hits(gt.data_ptr(), gt_off.data_ptr(), pr.data_ptr(), pr_off.data_ptr(), V, hits.data_ptr(),
                                _stream()))
    hits = hits.cpu().tolist()
    out = []
    for v in range(V):
        n_gt, n_pr = len(gt_lists[v]), len(pred_lists[v])
        rec = tuple(h / n_gt for h in hits[v][:3])          # ZeroDivisionError on empty ground truth, like the reference
        prec = tuple(h / n_pr for h in hits[v][3:]) if n_pr > 0 else (None, None, None)
        out.append(rec + prec)
    return out


def auc_ap_device(scores, labels, video_offsets):
    """Per-video ROC AUC and average precision of the clip scores (softmax prob of class 1) against the 0/1 clip labels,
    as sklearn.metrics computes them in test_video_segment_point.py:253-257 -> (auc [V], ap [V]) float64 host tensors.
    scores [N] fp32 CUDA, labels [N] int, video_offsets [V+1]; a video with a single class has auc = nan."""
    if not scores.is_cuda:
        raise RuntimeError("vcg_b200 post-processing needs CUDA tensors: there is no CPU fallback")
    scores = scores.float().contiguous()
    labels = labels.to(device=scores.device, dtype=torch.int32).contiguous()
    off = video_offsets.to(device=scores.device, dtype=torch.int32).contiguous()
    V = off.numel() - 1
    auc = torch.empty(V, dtype=torch.float64, device=scores.device)
    ap = torch.empty(V, dtype=torch.float64, device=scores.device)
    if V > 0:
        lib = _b.load_library()
        _b.check(lib.vcg_op_auc_ap(scores.data_ptr(), labels.data_ptr(), off.data_ptr(), V, auc.data_ptr(), ap.data_ptr(),
                                   _stream()))
    return auc.cpu(), ap.cpu()


def reference_video_groups(vids):
    """The clip index lists the reference's evaluation loop actually scores per video (test_video_segment_point.py:
    238-296): clips grouped by consecutive ``vid``, and — because the loop re-initialises its lists WITH the first clip
    of a video and then appends that clip again (:284-292) — the first clip of every video counted twice.
    -> (index tensor int64, offsets int32 [V+1]) for gathering scores / labels before auc_ap_device /
    cut_points_device, so that the numbers equal the reference script's."""
    idx, off = [], [0]
    prev = object()
    for i, v in enumerate(vids):
        if v != prev:
            if idx:
                off.append(len(idx))
            idx.append(i)                     # the duplicated first clip
            prev = v
        idx.append(i)
    if idx:
        off.append(len(idx))
    return torch.tensor(idx, dtype=torch.int64), torch.tensor(off, dtype=torch.int32)
