"""The evaluation of test_video_segment_point.py end to end on a synthetic dataset, two ways that must agree:

  A  the caller's own flow through the mirrored API: InferYoutubeClipDataset -> DataLoader (fp32 CHW clips) ->
     TwoStream.forward -> topk / prob[:, 1] -> the per-video loop (:250-333) with sklearn metrics and eval_utils
  B  the B200 flow: clips_u8 per video (every frame decoded once, uint8) -> Engine.score_clips_u8_host ->
     vcg_op_cut_points / vcg_op_pr_hits / vcg_op_auc_ap on the device

Labels and chapter timestamps (cut points) must be identical; P/R identical; AUC / AP equal to 1e-9.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_point_evaluation_reference_flow_vs_device_flow(tmp_path):
    sk = pytest.importorskip("sklearn.metrics")
    from torch.utils.data import DataLoader
    from torchvision import transforms
    from transformers import BertTokenizer
    from data.infer_youtube_video_dataset import InferYoutubeClipDataset
    from eval_utils.eval_utils import calculate_pr, convert_clip_label2cut_point
    from oracle import metrics_oracle as mo
    from oracle import synthetic_dataset as syn
    from vcg_b200 import postprocess as pp
    from test_parity_gpu import build_model

    p = syn.build(str(tmp_path))
    tok = BertTokenizer(vocab_file=p["vocab"], do_lower_case=True)
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    T, L = syn.T, 20
    ds = InferYoutubeClipDataset(p["img_dir"], p["clips_json"], tok, T, L, transform=tf)
    model, sd = build_model(T, "mlp", "fp32")

    def run_loader():
        logits, probs = [], []
        for img, ids, mask, _ in DataLoader(ds, batch_size=5, shuffle=False):
            lg, pr = model(img.float().to(0), ids.to(0), mask.to(0))
            logits.append(lg)
            probs.append(pr)
        return torch.cat(logits), torch.cat(probs)

    # random weights put every clip on one side: centre the decision so that both labels (and cut points) occur
    logits, _ = run_loader()
    shift = float((logits[:, 1] - logits[:, 0]).median())
    sd["fusion_head.head.bias"] = sd["fusion_head.head.bias"] + torch.tensor([shift, 0.0])
    model, _ = build_model(T, "mlp", "fp32", sd=sd)

    # ---- A: the caller's flow
    logits, probs = run_loader()
    pred_label = logits.data.topk(1, 1, True, True)[1].squeeze(1).cpu().numpy()
    pred_score = probs[:, 1].cpu().numpy()
    assert 0 < pred_label.sum() < len(pred_label)
    vids = [info["vid"] for info in ds.all_clip_infos]
    gt_label = np.array([info["clip_label"] for info in ds.all_clip_infos])
    flow_a = {}
    for g in mo.reference_video_groups(vids):             # first clip of each video twice, like the reference loop
        vid = vids[g[0]]
        fpr, tpr, _ = sk.roc_curve(gt_label[g], pred_score[g], pos_label=1)
        gt_cuts = convert_clip_label2cut_point(gt_label[g].tolist(), T, ds.max_offset)
        pred_cuts = convert_clip_label2cut_point(pred_label[g].tolist(), T, ds.max_offset)
        flow_a[vid] = {"auc": sk.auc(fpr, tpr), "ap": sk.average_precision_score(gt_label[g], pred_score[g]),
                       "gt_cuts": gt_cuts, "pred_cuts": pred_cuts,
                       "pr": calculate_pr(gt_cuts, pred_cuts) if gt_cuts else None}

    # ---- B: uint8 frames, host entry point, device post-processing
    eng = model.engine
    lg_b, pr_b, lo = [], [], 0
    for vid in dict.fromkeys(vids):
        n = vids.count(vid)
        frames, clip_start, ids, mask, _ = ds.clips_u8(lo, lo + n)
        lg, pr = eng.score_clips_u8_host(frames.pin_memory(), clip_start.pin_memory(), ids.pin_memory(), mask.pin_memory())
        lg_b.append(lg.clone())
        pr_b.append(pr.clone())
        lo += n
    lg_b, pr_b = torch.cat(lg_b).cuda(), torch.cat(pr_b).cuda()
    assert float((lg_b - logits).abs().max() / logits.abs().max()) <= 1e-4
    idx, off = pp.reference_video_groups(vids)
    labels_dev, pred_cuts_dev = pp.cut_points_device(lg_b[idx.cuda()], off, T, ds.max_offset)
    gt_logits = torch.stack([1.0 - torch.from_numpy(gt_label).float(), torch.from_numpy(gt_label).float()], 1).cuda()
    _, gt_cuts_dev = pp.cut_points_device(gt_logits[idx.cuda()], off, T, ds.max_offset)
    auc, ap = pp.auc_ap_device(pr_b[:, 1][idx.cuda()], torch.from_numpy(gt_label).cuda()[idx.cuda()], off)
    assert labels_dev.cpu().tolist() == pred_label[idx.numpy()].tolist()
    scored = [v for v, vid in enumerate(dict.fromkeys(vids)) if flow_a[vid]["gt_cuts"]]
    pr_dev = pp.pr_hits_device([gt_cuts_dev[v] for v in scored], [pred_cuts_dev[v] for v in scored])
    for v, vid in enumerate(dict.fromkeys(vids)):
        a = flow_a[vid]
        print(vid, a, float(auc[v]), float(ap[v]))
        assert pred_cuts_dev[v] == a["pred_cuts"] and gt_cuts_dev[v] == a["gt_cuts"]       # identical timestamps
        assert abs(float(auc[v]) - a["auc"]) <= 1e-9 and abs(float(ap[v]) - a["ap"]) <= 1e-9
        if v in scored:
            assert pr_dev[scored.index(v)] == a["pr"]

    # ---- the same through vcg_b200.evaluate (one call), against the script's averaging (:345-358) of flow A
    from vcg_b200 import evaluate as ev
    res = ev.evaluate_flat_clips(eng, ds)
    order = list(dict.fromkeys(vids))
    assert res["videos"] == order
    assert abs(res["mAP"] - np.mean([flow_a[v]["ap"] for v in order])) <= 1e-9
    prs = [flow_a[v]["pr"] for v in order if flow_a[v]["pr"] is not None]
    want_recall = np.mean([p_[0] for p_ in prs])
    want_prec3 = np.mean([p_[4] for p_ in prs if p_[4] is not None])
    assert abs(res["recall"] - want_recall) <= 1e-12 and abs(res["precision@3"] - want_prec3) <= 1e-12
    assert res["vid2cut_points"]["vidA"]["second_pred_cut_points"] == flow_a["vidA"]["pred_cuts"]
    ev.write_results(res, str(tmp_path / "out" / "result.txt"), str(tmp_path / "out" / "vid2cut_points.json"))
    text = (tmp_path / "out" / "result.txt").read_text()
    assert text.startswith(f"mAP {res['mAP']}\nrecall ") and "f-score_rand" in text
    once = ev.evaluate_flat_clips(eng, ds, reference_grouping=False, random_baseline=False)
    assert len(once["auc_list"]) == 2 and "recall_rand" not in once

    # ---- the per-video harness (video_segment/test_video_segment_point_per_video.py:104-175)
    from data.infer_youtube_video_dataset import InferYoutubeVideoDataset
    vds = InferYoutubeVideoDataset(p["img_dir"], p["data_file"], p["vid_file"], tok, T, L, transform=tf)
    per_video = ev.infer_videos(eng, vds, list(syn.VIDEOS))
    assert per_video["total_frames"] == sum(n for n, _ in syn.VIDEOS.values()) and per_video["video_infer_fps"] > 0
    for vid in syn.VIDEOS:          # the same clips as the flat file -> the same labels; cut points without the duplicate
        rows = [i for i, v in enumerate(vids) if v == vid]
        assert per_video["videos"][vid]["pred_labels"] == pred_label[rows].tolist()
        assert per_video["videos"][vid]["pred_cut_points"] == convert_clip_label2cut_point(pred_label[rows].tolist(), T, 2)
        assert per_video["videos"][vid]["gt_cut_points"] == convert_clip_label2cut_point(gt_label[rows].tolist(), T, 2)


def test_command_line_counterpart(tmp_path):
    """video_segment/test_video_segment_point_b200.py — the caller's switches, model construction, evaluation and result
    files — on the synthetic dataset: --data_mode all (uint8 path, device metrics) and text (DataLoader path)."""
    import json
    from oracle import synthetic_dataset as syn
    from oracle import weights as W
    from video_segment import test_video_segment_point_b200 as cli
    p = syn.build(str(tmp_path / "data"))
    for mode, sd in (("all", W.make_state_dict(syn.T, "mlp", seed=123)), ("text", W.make_unimodal_state_dict("bert", syn.T, seed=123))):
        ckpt = str(tmp_path / f"ckpt_{mode}.pth")
        torch.save({"epoch": 1, "best_result": 0.0, "model_state_dict": sd}, ckpt)
        out = cli.main(["--gpu", "0", "--data_mode", mode, "--clip_frame_num", str(syn.T), "--clips_json", p["clips_json"],
                        "--img_dir", p["img_dir"], "--ckpt", ckpt, "--vocab", p["vocab"], "--max_text_len", "20",
                        "--precision", "fp32", "--num_workers", "0", "--batch_size", "5",
                        "--result_file", str(tmp_path / mode / "result.txt"),
                        "--vid2cut_points_file", str(tmp_path / mode / "cuts.json")])
        assert (tmp_path / mode / "result.txt").read_text().startswith(f"mAP {out['mAP']}")
        cuts = json.loads((tmp_path / mode / "cuts.json").read_text())
        assert set(cuts) == set(syn.VIDEOS) and 0.0 <= out["mAP"] <= 1.0


def test_convert2vision_emb_round_trip(tmp_path):
    """convert2vision_emb.py's job on the uint8 path: the files it writes hold what forward(..., return_emb=True) returns,
    and feeding them back as precomputed embeddings (Identity vision model, BASELINE.json configs[1]) reproduces the logits."""
    from torchvision import transforms
    from transformers import BertTokenizer
    from data.infer_youtube_video_dataset import InferYoutubeClipDataset
    from oracle import synthetic_dataset as syn
    from vcg_b200 import vision_emb_io as vio
    from test_parity_gpu import build_model, rel
    p = syn.build(str(tmp_path / "data"))
    tok = BertTokenizer(vocab_file=p["vocab"], do_lower_case=True)
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    T, L = syn.T, 20
    ds = InferYoutubeClipDataset(p["img_dir"], p["clips_json"], tok, T, L, transform=tf)
    model, sd = build_model(T, "mlp", "fp32")
    n = vio.convert_flat_clips(model, ds, str(tmp_path / "emb"))
    assert n == len(ds)
    rows = [0, 3, 11]                                                # vidA x2, vidB
    items = [ds[i] for i in rows]
    img = torch.stack([it[0] for it in items]).cuda()
    ids = torch.stack([it[1] for it in items]).cuda()
    mask = torch.stack([it[2] for it in items]).cuda()
    logits, _, vis, _ = model(img, ids, mask, return_emb=True)
    for k, i in enumerate(rows):
        info = ds.all_clip_infos[i]
        back = vio.load_vision_embs(str(tmp_path / "emb"), info["vid"], [info["clip_start_end"][0]], T)
        assert back.shape == (1, T, 2048, 1, 1)
        assert rel(back.view(T, 2048), vis[k]) <= 1e-5
        light, _ = build_model(T, "mlp", "fp32", sd=sd, vision=False)
        lg, _ = light(back.cuda(), ids[k:k + 1], mask[k:k + 1])
        assert rel(lg, logits[k:k + 1]) <= 1e-4
