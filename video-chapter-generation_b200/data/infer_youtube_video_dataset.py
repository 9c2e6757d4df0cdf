"""Inference datasets of the named callers, same classes / constructor arguments / item tuples as the reference's
data/infer_youtube_video_dataset.py, so test_video_segment_point.py (:148-152), test_whole_pipeline_per_video.py and
test_video_segment_update.py construct and iterate them unchanged:

  InferYoutubeVideoDataset   one chosen video, clips range(0, n_frames - T, 4)                     (reference :31-215)
  InferYoutubeClipDataset    flat-clip JSON of flat_video2clip_for_quick_infer.py                  (reference :218-313)
  InferWindowClipDataset     windows of 2w+1 clips around every clip, zero padding at the borders  (reference :429-576)

Items are what the reference returns (fp32 CHW clips through the caller's transform, int64 ids / masks, labels), pinned
against the unmodified reference classes by oracle/make_golden_dataset.py -> tests/golden/dataset_*.npz.

What is new here is the second way out of each dataset, for the B200 engine: ``video_u8()`` / ``clips_u8()`` hand over
every DISTINCT frame once as uint8 HWC plus per-clip start rows (the 4x overlap of neighbouring clips is not decoded,
normalised or copied four times), which is what Engine.score_clips_u8_host consumes.  The reference's experimental
dataset variants (InferYoutubeAllClipDataset, InferWindowClipDatasetv2 — broken in the reference, :730 —,
InferWindowClipIDDataset) are not mirrored: no caller in scope uses them.
"""
import glob
import json
import os
import random

import numpy as np
import torch
from PIL import Image

from data._timestamps import extract_first_timestamp, parse_csv_to_list

X_PAD = 0
Y_PAD = -1


def _encode_text(tokenizer, text_clip, max_text_len):
    """"[CLS] " + text -> wordpieces, cut at max_text_len, [PAD]-filled; mask 1 for real tokens (no [SEP])."""
    tokens = tokenizer.tokenize("[CLS] " + text_clip)[:max_text_len]
    n_real = len(tokens)
    tokens = tokens + ["[PAD]"] * (max_text_len - n_real)
    ids = torch.tensor(tokenizer.convert_tokens_to_ids(tokens), dtype=torch.int64)
    mask = torch.zeros(max_text_len, dtype=torch.int64)
    mask[:n_real] = 1
    return ids, mask


def _load_rgb(path):
    with Image.open(path) as img:
        return img.convert("RGB")


def _frame_file(image_dir, sec, clip_start_sec, image_num, clip_frame_num):
    """File of the frame shown at second ``sec`` of a clip starting at ``clip_start_sec``: frames are 1-based and, away
    from both ends of the video, two files late (the reference compensates an ffmpeg extraction misalignment this way)."""
    near_edge = clip_start_sec <= 2 or clip_start_sec >= image_num - clip_frame_num - 2
    return os.path.join(image_dir, "%05d.jpg" % (sec + (1 if near_edge else 3)))


def _read_json_list(json_paths):
    if isinstance(json_paths, str):
        json_paths = [json_paths]
    infos = []
    for p in json_paths:
        with open(p, "r", encoding="utf-8") as f:
            infos.extend(json.load(f))
    return infos


def _frame_table(paths_per_clip, size=224):
    """Distinct files of the clips -> (uint8 [n,size,size,3] in first-use order, int64 row index [n_clips, T])."""
    rows, order = {}, []
    index = []
    for paths in paths_per_clip:
        r = []
        for p in paths:
            if p not in rows:
                rows[p] = len(order)
                order.append(p)
            r.append(rows[p])
        index.append(r)
    frames = np.empty((len(order), size, size, 3), dtype=np.uint8)
    for k, p in enumerate(order):
        img = _load_rgb(p)
        if img.size != (size, size):
            img = img.resize((size, size), Image.BILINEAR)
        frames[k] = np.asarray(img, dtype=np.uint8)
    return torch.from_numpy(frames), torch.tensor(index, dtype=torch.int64)


class InferYoutubeVideoDataset:
    """Clips of ONE video in temporal order (1 frame per second); choose the video before iterating."""

    def __init__(self, img_dir, data_file, vid_file, tokenizer, clip_frame_num, max_text_len, mode="all", transform=None,
                 target_transform=None):
        self.max_offset = 2
        self.tokenizer = tokenizer
        self.clip_frame_num = clip_frame_num
        self.max_text_len = max_text_len
        self.mode = mode                       # "text", "image" or "all"
        self.half_clip_frame_num = clip_frame_num // 2
        self.img_dir = img_dir
        vids, titles, durations, timestamps = parse_csv_to_list(data_file)
        self.vid2title = dict(zip(vids, titles))
        self.vid2timestamps = dict(zip(vids, timestamps))
        self.vid2durations = dict(zip(vids, durations))
        with open(vid_file, "r") as f:
            self.vids = [line.strip() for line in f.readlines()]
        self.asr_files = {}
        for path in glob.glob(os.path.dirname(data_file) + "/*/subtitle_*.json"):
            self.asr_files[os.path.basename(path).split(".")[0][len("subtitle_"):]] = path
        self.infer_vid = None
        self.transform = transform
        self.target_transform = target_transform

    # -- choosing the video
    def manual_choose_vid(self, vid):
        if vid not in self.vids:
            raise RuntimeError(f"The vid {vid} is not existed in dataset")
        self.infer_vid = vid
        self._load_gt_data()

    def random_choose_vid(self):
        self.infer_vid = random.sample(self.vids, 1)[0]
        self._load_gt_data()

    def _image_dir(self):
        return os.path.join(self.img_dir, self.infer_vid)

    def _image_num(self):
        return len(glob.glob(self._image_dir() + "/*.jpg"))

    def _load_gt_data(self):
        image_num = self._image_num()
        with open(self.asr_files[self.infer_vid], "r") as f:
            self.subtitles = json.load(f)
        self.cut_points, self.real_cut_points, self.descriptions = [], [], []
        for stamp in self.vid2timestamps[self.infer_vid]:
            sec, description = extract_first_timestamp(stamp)
            if sec < 4 or sec > image_num - 4:          # chapters too close to either end are not scored
                continue
            self.cut_points.append(sec)
            self.real_cut_points.append(sec)
            self.descriptions.append(description)

    # -- the clip grid
    def _clip_starts(self, image_num=None):
        image_num = self._image_num() if image_num is None else image_num
        return range(0, image_num - self.clip_frame_num, 2 * self.max_offset)

    def __len__(self):
        if self.infer_vid is None:
            raise RuntimeError("You should run choose_vid before iterate this dataset")
        return len(self._clip_starts())

    def get_duration(self):
        return self._image_num()

    def _label(self, start, end):
        """1 when the clip overlaps the +-T/2 neighbourhood of a chapter start with IoU >= (T - 2) / (T + 2)."""
        need = (self.clip_frame_num - self.max_offset) / (self.clip_frame_num + self.max_offset)
        label = 0
        for cp in self.cut_points:
            lo, hi = cp - self.half_clip_frame_num, cp + self.half_clip_frame_num
            inter = min(end, hi) - max(start, lo)
            union = max(end, hi) - min(start, lo)
            if inter / union >= need:
                label = 1
        return label

    def _text(self, start, end):
        """Subtitles that begin strictly inside (start - 1, end + 1), joined by blanks."""
        text = ""
        for sub in self.subtitles:
            if start - 1 < sub["start"] < end + 1:
                text = sub["text"] if len(text) == 0 else text + " " + sub["text"]
        return text

    def __getitem__(self, i):
        image_num = self._image_num()
        start = self._clip_starts(image_num)[i]
        end = start + self.clip_frame_num
        text_ids, attention_mask = _encode_text(self.tokenizer, self._text(start, end), self.max_text_len)
        if self.mode == "text":
            img_clip = 0
        else:
            img_clip = torch.stack([
                self.transform(_load_rgb(_frame_file(self._image_dir(), sec, start, image_num, self.clip_frame_num)))
                for sec in range(start, end)], dim=0)
        return img_clip, text_ids, attention_mask, self._label(start, end)

    def video_u8(self):
        """The chosen video for Engine.score_clips_u8_host: (frames uint8 [n,224,224,3] — each file decoded once —,
        clip_start int32 [N] rows of each clip's first frame, text_ids [N,L], attention_mask [N,L], labels [N])."""
        image_num = self._image_num()
        starts = list(self._clip_starts(image_num))
        T = self.clip_frame_num
        paths = [[_frame_file(self._image_dir(), sec, s, image_num, T) for sec in range(s, s + T)] for s in starts]
        frames, index = _frame_table(paths)
        assert bool((index[:, 1:] == index[:, :-1] + 1).all()), "a clip's frames are consecutive files"
        enc = [_encode_text(self.tokenizer, self._text(s, s + T), self.max_text_len) for s in starts]
        return (frames, index[:, 0].to(torch.int32).contiguous(), torch.stack([e[0] for e in enc]),
                torch.stack([e[1] for e in enc]), torch.tensor([self._label(s, s + T) for s in starts], dtype=torch.int64))


class InferYoutubeClipDataset:
    """All clips of all test videos from the flat-clip JSON(s)."""

    def __init__(self, img_dir, json_paths, tokenizer, clip_frame_num, max_text_len, mode="all", transform=None,
                 target_transform=None):
        self.max_offset = 2
        self.tokenizer = tokenizer
        self.clip_frame_num = clip_frame_num
        self.max_text_len = max_text_len
        self.mode = mode
        self.half_clip_frame_num = clip_frame_num // 2
        self.img_dir = img_dir
        self.all_clip_infos = _read_json_list(json_paths)
        self.transform = transform
        self.target_transform = target_transform

    def __len__(self):
        return len(self.all_clip_infos)

    def __getitem__(self, i):
        info = self.all_clip_infos[i]
        text_ids, attention_mask = _encode_text(self.tokenizer, info["text_clip"], self.max_text_len)
        if self.mode == "text":
            img_clip = 0
        else:
            img_clip = torch.stack([self.transform(_load_rgb(p)) for p in info["image_paths"]], dim=0)
        return img_clip, text_ids, attention_mask, info["clip_label"]

    def clips_u8(self, lo=0, hi=None):
        """Clips [lo, hi) for Engine.score_clips_u8_host: (frames uint8 [n,224,224,3], clip_start int32, text_ids,
        attention_mask, labels); frames shared by neighbouring clips are decoded once."""
        infos = self.all_clip_infos[lo:hi]
        frames, index = _frame_table([info["image_paths"] for info in infos])
        if not bool((index[:, 1:] == index[:, :-1] + 1).all()):
            raise ValueError("a clip's frames must be consecutive in first-use order; split the range per video")
        enc = [_encode_text(self.tokenizer, info["text_clip"], self.max_text_len) for info in infos]
        return (frames, index[:, 0].to(torch.int32).contiguous(), torch.stack([e[0] for e in enc]),
                torch.stack([e[1] for e in enc]), torch.tensor([int(info["clip_label"]) for info in infos], dtype=torch.int64))


class InferWindowClipDataset:
    """Every clip with its window: the clips ``skip * k`` positions away in the same video, k = -w..w, with
    skip = T // (2 * max_offset) so that window members do not overlap; positions outside the video are zero clips
    with zero ids and ZERO mask."""

    def __init__(self, img_dir, json_paths, tokenizer, clip_frame_num, max_text_len, window_size=2, mode="all",
                 transform=None):
        self.fps = 1
        self.max_offset = 2 * self.fps
        self.tokenizer = tokenizer
        self.clip_frame_num = clip_frame_num
        self.max_text_len = max_text_len
        self.window_size = window_size
        self.mode = mode
        self.img_dir = img_dir
        self.transform = transform
        self.all_clip_infos = _read_json_list(json_paths)
        self.vid2clips = {}
        for idx, info in enumerate(self.all_clip_infos):
            self.vid2clips.setdefault(info["vid"], []).append(idx)
        self._pos_in_video = {}
        for members in self.vid2clips.values():
            for pos, idx in enumerate(members):
                self._pos_in_video[idx] = pos

    def __len__(self):
        return len(self.all_clip_infos)

    def get_clip_info(self, idx):
        info = self.all_clip_infos[idx]
        members = self.vid2clips[info["vid"]]
        here = self._pos_in_video[idx]
        skip = self.clip_frame_num // (2 * self.max_offset)
        window = []
        for k in range(-self.window_size, self.window_size + 1):
            pos = here + k * skip
            window.append(members[pos] if 0 <= pos < len(members) else -1)
        return info, window

    def __getitem__(self, i):
        info, window = self.get_clip_info(i)
        image_dir = os.path.join(self.img_dir, info["vid"])
        image_num = len(glob.glob(image_dir + "/*.jpg"))
        T, L = self.clip_frame_num, self.max_text_len
        images, ids, masks, start_frames = [], [], [], []
        for idx in window:
            if idx == -1:
                if self.mode != "text":
                    images.append(torch.zeros((T, 3, 224, 224)))
                ids.append(torch.zeros(L, dtype=torch.long))
                masks.append(torch.zeros(L, dtype=torch.long))
                start_frames.append(-1)
                continue
            member = self.all_clip_infos[idx]
            start, end = member["clip_start_end"]
            start_frames.append(start)
            if self.mode != "text":
                frames = []
                for sec in range(start, end):
                    img = _load_rgb(_frame_file(image_dir, sec, start, image_num, T))
                    frames.append(self.transform(img) if self.transform else img)
                images.append(torch.stack(frames))
            t, m = _encode_text(self.tokenizer, member["text_clip"], L)
            ids.append(t)
            masks.append(m)
        img_clips = torch.tensor(0) if self.mode == "text" else torch.stack(images)
        clips_info = {"clip_start_frame": torch.tensor(start_frames), "total_frames": torch.tensor(image_num),
                      "target_clip_idx": torch.tensor(window[self.window_size]),
                      "total_num_clips": torch.tensor(len(self.vid2clips[info["vid"]]))}
        return img_clips, torch.stack(ids), torch.stack(masks), torch.tensor(info["clip_label"]), clips_info
