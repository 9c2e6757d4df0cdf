"""TEST INFRASTRUCTURE — golden fixtures of the reference's inference datasets.

Run in the build container only (reads /root/reference):   python oracle/make_golden_dataset.py

Builds the synthetic dataset of oracle/synthetic_dataset.py in a temporary directory, iterates the UNMODIFIED reference
classes InferYoutubeVideoDataset / InferYoutubeClipDataset / InferWindowClipDataset (data/infer_youtube_video_dataset.py)
over it with the callers' transform and a BertTokenizer, and stores what they return in tests/golden/dataset_*.npz.
matplotlib (imported by the reference module, unused on this path, absent from the image) is stubbed.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/video_chapter_generation"
GOLDEN = os.path.join(ROOT, "tests", "golden")

from oracle import synthetic_dataset as syn  # noqa: E402


def main():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    from data import infer_youtube_video_dataset as ref          # the reference module
    from torchvision import transforms
    from transformers import BertTokenizer
    tf = transforms.Compose([transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    with tempfile.TemporaryDirectory() as root:
        p = syn.build(root)
        tok = BertTokenizer(vocab_file=p["vocab"], do_lower_case=True)
        T, L = syn.T, 20
        out = {}
        # one chosen video
        ds = ref.InferYoutubeVideoDataset(p["img_dir"], p["data_file"], p["vid_file"], tok, T, L, transform=tf)
        for vid in syn.VIDEOS:
            ds.manual_choose_vid(vid)
            items = [ds[i] for i in range(len(ds))]
            out[f"video_{vid}_len"] = np.array([len(ds), ds.get_duration()])
            out[f"video_{vid}_cut_points"] = np.array(ds.cut_points)
            out[f"video_{vid}_descriptions"] = np.array(ds.descriptions)
            out[f"video_{vid}_ids"] = torch.stack([it[1] for it in items]).numpy()
            out[f"video_{vid}_mask"] = torch.stack([it[2] for it in items]).numpy()
            out[f"video_{vid}_label"] = np.array([it[3] for it in items])
            out[f"video_{vid}_img"] = np.stack([syn.summarise(it[0]) for it in items])
        # flat clips
        ds = ref.InferYoutubeClipDataset(p["img_dir"], p["clips_json"], tok, T, L, transform=tf)
        items = [ds[i] for i in range(len(ds))]
        out["clip_ids"] = torch.stack([it[1] for it in items]).numpy()
        out["clip_mask"] = torch.stack([it[2] for it in items]).numpy()
        out["clip_label"] = np.array([it[3] for it in items])
        out["clip_img"] = np.stack([syn.summarise(it[0]) for it in items])
        out["clip_img_first"] = items[3][0].numpy()              # one full clip [T,3,224,224]
        # windows
        for w in (1, 2):
            ds = ref.InferWindowClipDataset(p["img_dir"], p["clips_json"], tok, T, L, window_size=w, transform=tf)
            items = [ds[i] for i in range(len(ds))]
            out[f"window{w}_ids"] = torch.stack([it[1] for it in items]).numpy()
            out[f"window{w}_mask"] = torch.stack([it[2] for it in items]).numpy()
            out[f"window{w}_label"] = np.array([int(it[3]) for it in items])
            out[f"window{w}_img"] = np.stack([np.stack([syn.summarise(c) for c in it[0]]) for it in items])
            for key in ("clip_start_frame", "total_frames", "target_clip_idx", "total_num_clips"):
                out[f"window{w}_{key}"] = torch.stack([it[4][key] for it in items]).numpy()
            out[f"window{w}_indices"] = np.array([ds.get_clip_info(i)[1] for i in range(len(ds))])
        np.savez_compressed(os.path.join(GOLDEN, "dataset_synthetic.npz"), **out)
        print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
