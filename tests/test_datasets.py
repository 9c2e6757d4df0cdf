"""The inference datasets of the named callers (video-chapter-generation_b200/data/infer_youtube_video_dataset.py)
against what the UNMODIFIED reference classes returned on the same synthetic dataset
(tests/golden/dataset_synthetic.npz, written by oracle/make_golden_dataset.py), plus the uint8 hand-over to the engine."""
import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def synthetic(tmp_path_factory):
    from oracle import synthetic_dataset as syn
    from torchvision import transforms
    from transformers import BertTokenizer
    root = str(tmp_path_factory.mktemp("dataset"))
    p = syn.build(root)
    tok = BertTokenizer(vocab_file=p["vocab"], do_lower_case=True)
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return p, tok, tf


def _same_images(items_img, want):
    from oracle import synthetic_dataset as syn
    got = np.stack([syn.summarise(x) for x in items_img])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)


def test_video_dataset_matches_reference(golden_dir, synthetic):
    from data.infer_youtube_video_dataset import InferYoutubeVideoDataset
    from oracle import synthetic_dataset as syn
    g = np.load(f"{golden_dir}/dataset_synthetic.npz")
    p, tok, tf = synthetic
    ds = InferYoutubeVideoDataset(p["img_dir"], p["data_file"], p["vid_file"], tok, syn.T, 20, transform=tf)
    with pytest.raises(RuntimeError):
        len(ds)
    with pytest.raises(RuntimeError):
        ds.manual_choose_vid("nope")
    for vid in syn.VIDEOS:
        ds.manual_choose_vid(vid)
        assert [len(ds), ds.get_duration()] == g[f"video_{vid}_len"].tolist()
        assert ds.cut_points == g[f"video_{vid}_cut_points"].tolist() == ds.real_cut_points
        assert ds.descriptions == g[f"video_{vid}_descriptions"].tolist()
        items = [ds[i] for i in range(len(ds))]
        assert np.array_equal(torch.stack([it[1] for it in items]).numpy(), g[f"video_{vid}_ids"])
        assert np.array_equal(torch.stack([it[2] for it in items]).numpy(), g[f"video_{vid}_mask"])
        assert [it[3] for it in items] == g[f"video_{vid}_label"].tolist()
        assert items[0][0].shape == (syn.T, 3, 224, 224) and items[0][0].dtype == torch.float32
        _same_images([it[0] for it in items], g[f"video_{vid}_img"])
        # the uint8 hand-over: every file once, the +1 / +3 file offset folded into the clip start rows
        frames, clip_start, ids, mask, labels = ds.video_u8()
        assert frames.dtype == torch.uint8 and frames.shape[0] <= ds.get_duration()
        assert np.array_equal(ids.numpy(), g[f"video_{vid}_ids"]) and labels.tolist() == g[f"video_{vid}_label"].tolist()
        mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
        for i in (0, 1, len(ds) - 1):
            clip = frames[clip_start[i]:clip_start[i] + syn.T].float().div(255).sub(mean).div(std).permute(0, 3, 1, 2)
            assert torch.allclose(clip, items[i][0], atol=1e-6)
    ds.mode = "text"
    assert ds[0][0] == 0


def test_flat_clip_dataset_matches_reference(golden_dir, synthetic):
    from data.infer_youtube_video_dataset import InferYoutubeClipDataset
    from oracle import synthetic_dataset as syn
    g = np.load(f"{golden_dir}/dataset_synthetic.npz")
    p, tok, tf = synthetic
    for paths in (p["clips_json"], [p["clips_json"]]):
        ds = InferYoutubeClipDataset(p["img_dir"], paths, tok, syn.T, 20, transform=tf)
        assert len(ds) == len(g["clip_label"]) and ds.max_offset == 2
    items = [ds[i] for i in range(len(ds))]
    assert np.array_equal(torch.stack([it[1] for it in items]).numpy(), g["clip_ids"])
    assert np.array_equal(torch.stack([it[2] for it in items]).numpy(), g["clip_mask"])
    assert [it[3] for it in items] == g["clip_label"].tolist()
    _same_images([it[0] for it in items], g["clip_img"])
    assert np.array_equal(items[3][0].numpy(), g["clip_img_first"])
    # a DataLoader batch looks like the caller's (test_video_segment_point.py:154-161, 172-176)
    from torch.utils.data import DataLoader
    img, ids, mask, label = next(iter(DataLoader(ds, batch_size=4, shuffle=False)))
    assert img.shape == (4, syn.T, 3, 224, 224) and ids.shape == (4, 20) and label.tolist() == g["clip_label"][:4].tolist()
    # uint8 hand-over per video
    n_a = sum(1 for info in ds.all_clip_infos if info["vid"] == "vidA")
    frames, clip_start, ids, mask, labels = ds.clips_u8(0, n_a)
    assert frames.shape[0] == len({q for info in ds.all_clip_infos[:n_a] for q in info["image_paths"]})
    assert np.array_equal(ids.numpy(), g["clip_ids"][:n_a]) and np.array_equal(mask.numpy(), g["clip_mask"][:n_a])
    mean, std = torch.tensor([0.485, 0.456, 0.406]), torch.tensor([0.229, 0.224, 0.225])
    clip = frames[clip_start[3]:clip_start[3] + syn.T].float().div(255).sub(mean).div(std).permute(0, 3, 1, 2)
    assert torch.allclose(clip, torch.from_numpy(g["clip_img_first"]), atol=1e-6)


@pytest.mark.parametrize("w", [1, 2])
def test_window_dataset_matches_reference(golden_dir, synthetic, w):
    from data.infer_youtube_video_dataset import InferWindowClipDataset
    from oracle import synthetic_dataset as syn
    g = np.load(f"{golden_dir}/dataset_synthetic.npz")
    p, tok, tf = synthetic
    ds = InferWindowClipDataset(p["img_dir"], p["clips_json"], tok, syn.T, 20, window_size=w, transform=tf)
    assert len(ds) == len(g[f"window{w}_label"])
    assert [ds.get_clip_info(i)[1] for i in range(len(ds))] == g[f"window{w}_indices"].tolist()
    items = [ds[i] for i in range(len(ds))]
    assert np.array_equal(torch.stack([it[1] for it in items]).numpy(), g[f"window{w}_ids"])
    assert np.array_equal(torch.stack([it[2] for it in items]).numpy(), g[f"window{w}_mask"])
    assert [int(it[3]) for it in items] == g[f"window{w}_label"].tolist()
    for key in ("clip_start_frame", "total_frames", "target_clip_idx", "total_num_clips"):
        assert np.array_equal(torch.stack([it[4][key] for it in items]).numpy(), g[f"window{w}_{key}"]), key
    from oracle import synthetic_dataset as syn2
    got = np.stack([np.stack([syn2.summarise(c) for c in it[0]]) for it in items])
    assert np.allclose(got, g[f"window{w}_img"], rtol=1e-9, atol=1e-9)
    assert items[0][0].shape == (2 * w + 1, syn.T, 3, 224, 224)


def test_timestamp_helpers():
    from data._timestamps import extract_first_timestamp, extract_timestamp
    assert extract_timestamp("intro 1:02:03 x") == ("1:02:03", 3723, 6, 13)
    assert extract_timestamp("12:34 go") == ("12:34", 754, 0, 5)
    assert extract_timestamp("no stamp") == ("", -1, -1, -1)
    assert extract_first_timestamp("00:25 second part 0:31") == (25, " second part ")
    assert extract_first_timestamp("7:15 a 3:10 b") == (190, " a  b")
