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


def convert_flat_clips(model, dataset, save_dir, precision=None):
    """convert2vision_emb.py's loop (:150-206) on the B200 path: the ResNet-50-TSM embeddings [T, 2048] of every clip of
    an InferYoutubeClipDataset, written as ``<save_dir>/<vid>/vision_emb_{start}_{end}.npy``.

    ``model`` is the two-stream module holding the weights (model.fusion.two_stream.TwoStream); the clips of a video go
    through ONE backbone-only engine pass from uint8 frames (Engine.embed_u8: every frame decoded and normalised once)
    instead of fp32 batches through forward(..., return_emb=True).  Returns the number of files written."""
    from vcg_b200.engine import Engine
    dev = next(model.parameters()).device
    _, shift_div = model._vision_kind()
    eng = Engine(model.segment_size, "mlp", precision or model.precision, True, dataset.max_text_len, model.vision_chunk,
                 model.hidden_size, shift_div, device=dev, modality="embed")
    eng.load_state_dict(model.state_dict())
    infos = dataset.all_clip_infos
    written, lo = 0, 0
    try:
        while lo < len(infos):
            hi = lo
            while hi < len(infos) and infos[hi]["vid"] == infos[lo]["vid"]:
                hi += 1
            frames, clip_start, ids, mask, _ = dataset.clips_u8(lo, hi)
            vis, _ = eng.embed_u8(frames.to(dev), ids.to(dev), mask.to(dev), clip_start.to(dev))
            written += len(save_vision_embs(save_dir, infos[lo:hi], vis))
            lo = hi
    finally:
        eng.close()
    return written
