"""TEST INFRASTRUCTURE — a tiny synthetic copy of the reference's on-disk dataset layout (seeded, a few hundred KB):

    <root>/frames/<vid>/00001.jpg ...         1 fps frames, 224x224 JPEG
    <root>/meta/chapters.csv                  videoId,title,duration,timestamp ("%^&*"-joined "M:SS title" strings)
    <root>/meta/subs/subtitle_<vid>.json      [{"text", "start"}, ...]
    <root>/meta/test_vids.txt                 one vid per line
    <root>/meta/test_clips.json               flat clips (flat_video2clip_for_quick_infer.py:112-119 schema)
    <root>/vocab.txt                          a WordPiece vocabulary for transformers.BertTokenizer

Used by oracle/make_golden_dataset.py (which runs the UNMODIFIED reference datasets on it) and by tests/test_datasets.py
(which runs this repo's datasets on a rebuilt copy and compares with the stored reference outputs)."""
import json
import os

import numpy as np
from PIL import Image

WORDS = ["the", "video", "chapter", "begins", "here", "we", "talk", "about", "cooking", "music", "travel", "code",
         "and", "then", "next", "topic", "is", "fun", "lesson", "one", "two", "three", "end", "thanks"]
VIDEOS = {"vidA": (45, ["0:00 intro", "0:13 part one of vidA", "00:25 second part 0:31", "0:43 outro"]),
          "vidB": (30, ["0:00 start", "0:09 middle", "0:17 finale"])}
T = 8


def build(root, seed=5):
    rng = np.random.RandomState(seed)
    frames_dir, meta = os.path.join(root, "frames"), os.path.join(root, "meta")
    os.makedirs(os.path.join(meta, "subs"), exist_ok=True)
    with open(os.path.join(root, "vocab.txt"), "w") as f:
        f.write("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS + ["##s", "##ing"]) + "\n")
    rows = ["videoId,title,duration,timestamp"]
    flat = []
    for vid, (n, stamps) in VIDEOS.items():
        os.makedirs(os.path.join(frames_dir, vid), exist_ok=True)
        for k in range(n):
            small = rng.randint(0, 256, size=(14, 14, 3), dtype=np.uint8)
            img = Image.fromarray(small).resize((224, 224), Image.NEAREST)
            img.save(os.path.join(frames_dir, vid, "%05d.jpg" % (k + 1)), quality=92)
        subs = []
        for s in range(0, n, 3):
            words = [WORDS[j] for j in rng.randint(0, len(WORDS), size=rng.randint(2, 9))]
            subs.append({"text": " ".join(words) + ("s" if s % 2 else ""), "start": float(s) + 0.5 * (s % 2)})
        with open(os.path.join(meta, "subs", f"subtitle_{vid}.json"), "w") as f:
            json.dump(subs, f)
        rows.append(f"{vid},title of {vid},{n},{'%^&*'.join(stamps)}")
        # flat clips the way flat_video2clip_for_quick_infer.py lays them out (fps 1)
        cuts = []
        for st in stamps:
            m, s = st.split(" ")[0].split(":")
            secs = [int(m) * 60 + int(s)]
            if "0:31" in st:
                secs.append(31)
            sec = min(secs)
            if 4 <= sec <= n - 4:
                cuts.append(sec)
        for start in range(0, n - T, 4):
            end = start + T
            label = 0
            for cp in cuts:
                inter = min(end, cp + T // 2) - max(start, cp - T // 2)
                union = max(end, cp + T // 2) - min(start, cp - T // 2)
                if inter / union >= (T - 2) / (T + 2):
                    label = 1
            text = " ".join(sub["text"] for sub in subs if start - 1 < sub["start"] < end + 1)
            off = 1 if (start <= 2 or start >= n - T - 2) else 3
            flat.append({"image_paths": [os.path.join(frames_dir, vid, "%05d.jpg" % (sec + off)) for sec in range(start, end)],
                         "text_clip": text, "clip_label": label, "clip_start_end": [start, end], "cut_points": cuts,
                         "vid": vid})
    with open(os.path.join(meta, "chapters.csv"), "w") as f:
        f.write("\n".join(rows) + "\n")
    with open(os.path.join(meta, "test_vids.txt"), "w") as f:
        f.write("\n".join(VIDEOS) + "\n")
    with open(os.path.join(meta, "test_clips.json"), "w") as f:
        json.dump(flat, f)
    return {"img_dir": frames_dir, "data_file": os.path.join(meta, "chapters.csv"),
            "vid_file": os.path.join(meta, "test_vids.txt"), "clips_json": os.path.join(meta, "test_clips.json"),
            "vocab": os.path.join(root, "vocab.txt")}


def summarise(img_clip):
    """Layout-independent fingerprint of an fp32 image tensor: sum, abs-sum and 8 fixed samples."""
    flat = img_clip.reshape(-1).double()
    pick = flat[:: max(1, flat.numel() // 8)][:8]
    return np.concatenate([[flat.sum().item(), flat.abs().sum().item()], pick.numpy()])
