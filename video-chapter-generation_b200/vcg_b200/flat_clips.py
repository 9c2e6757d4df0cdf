"""Flat-clip JSON -> the uint8 scoring pipeline.

The reference evaluates from "flat clip" JSON files written by
video_chapter_youtube_dataset/flat_video2clip_for_quick_infer.py:112-119 — a list of
``{"image_paths": [T jpg paths], "text_clip": str, "clip_label": 0/1, "clip_start_end": [s, e], "cut_points": [...],
"vid": str}`` — and its InferYoutubeClipDataset (data/infer_youtube_video_dataset.py:218-313) decodes EVERY clip's T
JPEGs, normalises them on the CPU to fp32 CHW and ships 9.6 MB per clip to the GPU.  Neighbouring clips share 3/4 of
their frames, so this reader decodes every distinct frame of a video ONCE, keeps it as uint8 HWC (150 KB), and hands the
engine a frame table + per-clip start indices: ``Engine.score_clips_u8_host`` then does ToTensor + Normalize on the
device (and, for a regular clip grid, runs the ResNet stem once per distinct frame).

Tokenisation is the reference's (:267-285): "[CLS] " + text, ``tokenizer.tokenize``, truncate, pad with [PAD], mask 1/0;
any object with ``tokenize`` and ``convert_tokens_to_ids`` works (transformers' BertTokenizer in the reference).
Host-side I/O only: no arithmetic of the scoring path happens here.
"""
import json
from collections import OrderedDict

import numpy as np
import torch


def load_flat_clips(json_paths):
    """-> OrderedDict vid -> list of clip-info dicts, in file order (a video's clips are contiguous in the files)."""
    if isinstance(json_paths, str):
        json_paths = [json_paths]
    infos = []
    for path in json_paths:
        with open(path, "r", encoding="utf-8") as f:
            infos.extend(json.load(f))
    videos = OrderedDict()
    for info in infos:
        videos.setdefault(info["vid"], []).append(info)
    return videos


def tokenize_clip(tokenizer, text_clip, max_text_len):
    """(ids [L] int64, attention_mask [L] int64) exactly like infer_youtube_video_dataset.py:267-285."""
    tokens = tokenizer.tokenize("[CLS] " + text_clip)[:max_text_len]
    mask = [1] * len(tokens) + [0] * (max_text_len - len(tokens))
    tokens = tokens + ["[PAD]"] * (max_text_len - len(tokens))
    ids = tokenizer.convert_tokens_to_ids(tokens)
    return torch.tensor(ids, dtype=torch.int64), torch.tensor(mask, dtype=torch.int64)


def _decode(path, size=224):
    from PIL import Image
    with Image.open(path) as img:
        img = img.convert("RGB")
        if img.size != (size, size):     # frames are extracted at 224x224 (extract_video_to_frames.py:28); be lenient
            img = img.resize((size, size), Image.BILINEAR)
        return np.asarray(img, dtype=np.uint8)


class FlatClipVideo:
    """One video of a flat-clip file, ready for ``Engine.score_clips_u8_host`` / ``score_clips_u8``:

    frames      uint8 [n_unique, 224, 224, 3]  every distinct JPEG of the video's clips, decoded once
    clip_start  int32 [n_clips]                index of each clip's first frame in ``frames`` (its T frames are consecutive)
    text_ids, attention_mask  int64 [n_clips, L]
    labels      int64 [n_clips];  cut_points: the video's ground-truth chapter starts;  clip_start_end: [n_clips, 2]
    """

    def __init__(self, clip_infos, tokenizer, clip_frame_num, max_text_len, decode=_decode, pin=True):
        assert len(clip_infos) > 0
        self.vid = clip_infos[0]["vid"]
        T = clip_frame_num
        index = OrderedDict()                      # image path -> row of the frame table, in first-use order
        starts = []
        for info in clip_infos:
            paths = info["image_paths"]
            if len(paths) != T:
                raise ValueError(f"clip of video {self.vid} has {len(paths)} frames, expected {T}")
            rows = [index.setdefault(p, len(index)) for p in paths]
            if any(rows[i + 1] != rows[i] + 1 for i in range(T - 1)):
                raise ValueError(f"video {self.vid}: a clip's frames are not consecutive in first-use order; "
                                 "use one FlatClipVideo per contiguous segment")
            starts.append(rows[0])
        self.frames = torch.from_numpy(np.stack([decode(p) for p in index]))
        self.clip_start = torch.tensor(starts, dtype=torch.int32)
        toks = [tokenize_clip(tokenizer, info["text_clip"], max_text_len) for info in clip_infos]
        self.text_ids = torch.stack([t[0] for t in toks])
        self.attention_mask = torch.stack([t[1] for t in toks])
        self.labels = torch.tensor([int(info["clip_label"]) for info in clip_infos], dtype=torch.int64)
        self.clip_start_end = torch.tensor([list(info["clip_start_end"]) for info in clip_infos], dtype=torch.int64)
        self.cut_points = list(clip_infos[-1]["cut_points"])
        if pin and torch.cuda.is_available():
            self.frames, self.clip_start = self.frames.pin_memory(), self.clip_start.pin_memory()
            self.text_ids, self.attention_mask = self.text_ids.pin_memory(), self.attention_mask.pin_memory()

    def __len__(self):
        return self.clip_start.numel()

    def score(self, engine):
        """-> (logits, probs) host tensors [n_clips, 2] through the host-buffer entry point of the engine."""
        return engine.score_clips_u8_host(self.frames, self.clip_start, self.text_ids, self.attention_mask)


def iter_videos(json_paths, tokenizer, clip_frame_num, max_text_len, decode=_decode, pin=True):
    """Yields a FlatClipVideo per video of the flat-clip file(s)."""
    for clip_infos in load_flat_clips(json_paths).values():
        yield FlatClipVideo(clip_infos, tokenizer, clip_frame_num, max_text_len, decode=decode, pin=pin)
