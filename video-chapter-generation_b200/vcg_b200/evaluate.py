"""The evaluation of test_video_segment_point.py (:168-391) as one call on the B200 path.

``evaluate_flat_clips(engine, dataset)`` scores every clip of an InferYoutubeClipDataset video by video through
Engine.score_clips_u8_host (uint8 frames, each decoded once), then computes on the device what the reference's Python
loop computes per video — labels, chapter cut points, the six hit counts of calculate_pr, ROC AUC and average precision
(vcg_op_cut_points / vcg_op_pr_hits / vcg_op_auc_ap) — and averages exactly like :345-358.  The reference loop counts the
first clip of every video twice (:284-292); ``reference_grouping=True`` (default) reproduces that so the numbers are the
script's, ``False`` scores every clip once.
"""
import json
import os
import random

import torch

from . import postprocess as pp


def score_dataset(engine, dataset, pin=True):
    """-> (logits [N,2], probs [N,2]) host tensors for all clips of a flat-clip dataset, one engine call per video."""
    vids = [info["vid"] for info in dataset.all_clip_infos]
    logits, probs, lo = [], [], 0
    while lo < len(vids):
        hi = lo
        while hi < len(vids) and vids[hi] == vids[lo]:
            hi += 1
        frames, clip_start, ids, mask, _ = dataset.clips_u8(lo, hi)
        if pin:
            frames, clip_start, ids, mask = (t.pin_memory() for t in (frames, clip_start, ids, mask))
        lg, pr = engine.score_clips_u8_host(frames, clip_start, ids, mask)
        logits.append(lg.clone())
        probs.append(pr.clone())
        lo = hi
    return torch.cat(logits), torch.cat(probs)


def _mean(xs):
    return sum(xs) / len(xs)


def evaluate_flat_clips(engine, dataset, reference_grouping=True, logits=None, probs=None, random_baseline=True):
    """-> dict with the script's summary numbers: "mAP", "auc", "recall", "recall@3", "recall@5", "precision", ...,
    "f-score", ..., the "*_rand" random-guess lines, per-video lists and "vid2cut_points" (the content of
    vid2cut_points.json).  Pass ``logits`` / ``probs`` to evaluate scores computed elsewhere."""
    infos = dataset.all_clip_infos
    T, max_offset = dataset.clip_frame_num, dataset.max_offset
    if logits is None:
        logits, probs = score_dataset(engine, dataset)
    dev = torch.device("cuda", torch.cuda.current_device())
    logits, probs = logits.to(dev), probs.to(dev)
    vids = [info["vid"] for info in infos]
    if reference_grouping:
        idx, off = pp.reference_video_groups(vids)
    else:
        idx = torch.arange(len(vids))
        bounds = [0] + [i for i in range(1, len(vids)) if vids[i] != vids[i - 1]] + [len(vids)]
        off = torch.tensor(bounds, dtype=torch.int32)
    gidx = idx.to(dev)
    gt = torch.tensor([int(info["clip_label"]) for info in infos], device=dev)
    gt_logits = torch.stack([1.0 - gt.float(), gt.float()], dim=1)
    _, pred_cuts = pp.cut_points_device(logits[gidx], off, T, max_offset)
    _, gt_cuts = pp.cut_points_device(gt_logits[gidx], off, T, max_offset)
    auc, ap = pp.auc_ap_device(probs[:, 1][gidx], gt[gidx], off)
    names = [vids[int(idx[int(off[v])])] for v in range(off.numel() - 1)]
    # calculate_pr divides by len(gt): like the reference, a video without ground-truth cut points is an error there;
    # here it is left out of the recall / precision averages
    scored = [v for v in range(len(names)) if len(gt_cuts[v]) > 0]
    pr = pp.pr_hits_device([gt_cuts[v] for v in scored], [pred_cuts[v] for v in scored], device=dev)
    out = {"videos": names, "auc_list": auc.tolist(), "map_list": ap.tolist(),
           "vid2cut_points": {names[v]: {"second_gt_cut_points": gt_cuts[v], "second_pred_cut_points": pred_cuts[v]}
                              for v in range(len(names))}}
    out["mAP"], out["auc"] = _mean(out["map_list"]), _mean(out["auc_list"])
    _summarise(out, pr, "")
    if random_baseline:      # :262, :277-284 — Python's global random stream, as in the script
        from eval_utils.eval_utils import calculate_pr
        rand = []
        for v in scored:
            last = infos[int(idx[int(off[v + 1]) - 1])]
            guess = [random.randint(0, last["clip_start_end"][1] - 1) for _ in range(len(last["cut_points"]))]
            rand.append(calculate_pr(gt_cuts[v], guess))
        _summarise(out, rand, "_rand")
    return out


def _summarise(out, pr, suffix):
    rec = [[p[k] for p in pr] for k in range(3)]
    prec = [[p[k] for p in pr if p[k] is not None] for k in range(3, 6)]
    for k, tag in enumerate(("", "@3", "@5")):
        r = _mean(rec[k]) if rec[k] else float("nan")
        p = _mean(prec[k]) if prec[k] else float("nan")
        out[f"recall{suffix}{tag}"], out[f"precision{suffix}{tag}"] = r, p
        out[f"f-score{suffix}{tag}"] = 2 * r * p / (r + p) if (r + p) > 0 else float("nan")


def write_results(out, result_file, vid2cut_points_file=None):
    """The script's result txt (:381-391) and vid2cut_points.json (:343-344)."""
    if vid2cut_points_file:
        os.makedirs(os.path.dirname(vid2cut_points_file) or ".", exist_ok=True)
        with open(vid2cut_points_file, "w") as f:
            json.dump(out["vid2cut_points"], f)
    os.makedirs(os.path.dirname(result_file) or ".", exist_ok=True)
    with open(result_file, "w") as f:
        f.write(f"mAP {out['mAP']}\n")
        for name in ("recall", "precision", "f-score"):
            f.write(f"{name} {out[name]}, {name}@3 {out[name + '@3']}, {name}@5 {out[name + '@5']}\n")
        if "recall_rand" in out:
            f.write("\n")
            for name in ("recall", "precision", "f-score"):
                f.write(f"{name}_rand {out[name + '_rand']}, {name}_rand@3 {out[name + '_rand@3']}, "
                        f"{name}_rand@5 {out[name + '_rand@5']}\n")


def infer_videos(engine, video_dataset, vids, pin=True):
    """The per-video harness of video_segment/test_video_segment_point_per_video.py (:104-175) on the B200 path: for
    every vid choose it, score all its clips through Engine.score_clips_u8_host (frames of the video decoded once),
    turn ground-truth and predicted labels into cut points, and accumulate the reference's throughput metric
    ``video infer fps = total_frames / sum(forward time)`` — here the forward time is measured with CUDA events around the
    host-buffer call (copies included, decode excluded, as in the reference where decode happens in the DataLoader).
    -> {"videos": {vid: {"duration", "gt_cut_points", "pred_cut_points", "pred_labels", "infer_seconds"}},
        "total_frames", "total_infer_seconds", "video_infer_fps"}"""
    from eval_utils.eval_utils import convert_clip_label2cut_point
    out = {"videos": {}, "total_frames": 0, "total_infer_seconds": 0.0}
    T, max_offset = video_dataset.clip_frame_num, video_dataset.max_offset
    for vid in vids:
        video_dataset.manual_choose_vid(vid)
        duration = video_dataset.get_duration()
        frames, clip_start, ids, mask, labels = video_dataset.video_u8()
        if pin:
            frames, clip_start, ids, mask = (t.pin_memory() for t in (frames, clip_start, ids, mask))
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        logits, _ = engine.score_clips_u8_host(frames, clip_start, ids, mask)
        stop.record()
        stop.synchronize()
        seconds = start.elapsed_time(stop) / 1e3
        pred = logits.topk(1, 1, True, True)[1].view(-1).tolist()
        out["videos"][vid] = {"duration": duration, "pred_labels": pred, "infer_seconds": seconds,
                              "gt_cut_points": convert_clip_label2cut_point(labels.tolist(), T, max_offset),
                              "pred_cut_points": convert_clip_label2cut_point(pred, T, max_offset)}
        out["total_frames"] += duration
        out["total_infer_seconds"] += seconds
    out["video_infer_fps"] = out["total_frames"] / out["total_infer_seconds"] if out["total_infer_seconds"] > 0 else 0.0
    return out
