"""Device-side post-processing (C ABI: vcg_op_cut_points / vcg_op_pr_hits).  Same results as the reference's Python
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
