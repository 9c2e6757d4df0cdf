"""On-disk format of convert2vision_emb.py (reference :190-198): one ``vision_emb_{start}_{end}.npy`` per clip, fp32
[T, 2048], under ``<save_dir>/<vid>/``.  The consumers look clips up by ``vision_emb_{st}_{st+T}.npy``
(data/youtube_chapter_title_dataset.py:233).  Plain host-side file I/O: the embeddings themselves come from
``TwoStream.forward(..., return_emb=True)`` (libvcg_b200.so)."""
import os

import numpy as np
import torch


def save_vision_embs(save_dir, clip_infos, vision_emb):
    """clip_infos: per clip a dict with "vid" and "clip_start_end" = (start_t, end_t) (the dataset's all_clip_infos
    entries, infer_youtube_video_dataset.py); vision_emb: [B, T, 2048] tensor or array.  Returns the written paths."""
    if isinstance(vision_emb, torch.Tensor):
        vision_emb = vision_emb.detach().float().cpu().numpy()
    assert len(clip_infos) == vision_emb.shape[0]
    paths = []
    for info, emb in zip(clip_infos, vision_emb):
        start_t, end_t = info["clip_start_end"]
        d = os.path.join(save_dir, info["vid"])
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, f"vision_emb_{start_t}_{end_t}.npy")
        np.save(path, emb.astype(np.float32, copy=False))
        paths.append(path)
    return paths


def load_vision_embs(save_dir, vid, starts, clip_frames=16):
    """-> fp32 tensor [len(starts), T, 2048, 1, 1]: the img_clip argument of a TwoStream whose vision_model is Identity
    (precomputed-embedding configuration, BASELINE.json configs[1])."""
    embs = [np.load(os.path.join(save_dir, vid, f"vision_emb_{st}_{st + clip_frames}.npy")) for st in starts]
    t = torch.from_numpy(np.stack(embs).astype(np.float32))
    assert t.shape[1:] == (clip_frames, 2048), t.shape
    return t.view(len(starts), clip_frames, 2048, 1, 1)
