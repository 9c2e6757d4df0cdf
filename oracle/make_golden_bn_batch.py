"""TEST INFRASTRUCTURE — golden fixture for the batch-statistics BatchNorm mode (reference caller #1).

Run in the build container only (it reads /root/reference):

    python oracle/make_golden_bn_batch.py

Builds the UNMODIFIED reference two-stream model like make_golden.py does, puts it in .eval() and then does what
test_video_segment_point.py:116-122 does to every nn.BatchNorm2d (no running statistics, track_running_stats False), so
that each BatchNorm of the ResNet-50 normalises with the statistics of the B*T frames of the call.  The reference's
outputs become tests/golden/bn_batch_*.npz; the oracle restatement (two_stream_oracle.batch_stat_bn) is pinned on them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import two_stream_oracle as orc  # noqa: E402
from oracle import weights as W  # noqa: E402
from oracle.make_golden import GOLDEN, load_reference, rel  # noqa: E402

# (name, head_type, T, L, B)
CASES = [
    ("mlp_T16_L100_B3", "mlp", 16, 100, 3),
    ("attn_T8_L32_B2", "attn", 8, 32, 2),
]


def main():
    two_stream, bert_hugface, resnet50_tsm, _ = load_reference()
    torch.set_grad_enabled(False)
    for name, head_type, T, L, B in CASES:
        print(f"== bn_batch_{name}")
        sd = W.make_state_dict(clip_frames=T, head_type=head_type, seed=123)
        lang = bert_hugface.BertHugface(pretrain_stage=False)
        vision = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
        model = two_stream.TwoStream(lang.base_model, vision.base_model, lang.embed_size, vision.feature_dim, T, 128)
        model.build_chapter_head(output_size=2, head_type=head_type)
        model.load_state_dict(sd, strict=True)
        model = model.eval()

        ids, mask = W.make_text(B, L, seed=123)
        frames = W.make_frames_u8(4 * (B - 1) + T, seed=123)
        starts = [4 * b for b in range(B)]
        img = orc.gather_clips(orc.preprocess_u8(frames), starts, T)

        ref_eval = model(img, ids, mask, return_emb=True)            # standard eval, for the contrast below
        n_bn = 0
        for m in model.modules():                                    # caller #1's treatment of the BatchNorm layers
            if isinstance(m, torch.nn.BatchNorm2d):
                m.track_running_stats, m.running_mean, m.running_var = False, None, None
                n_bn += 1
        assert n_bn == 53, n_bn
        logits, probs, vis_emb, lang_emb = model(img, ids, mask, return_emb=True)

        with orc.batch_stat_bn():
            o = orc.two_stream_forward(sd, img, ids, mask, T, 128, head_type, 8)
        errs = {"logits": rel(o[0], logits), "probs": rel(o[1], probs), "vision_emb": rel(o[2], vis_emb),
                "lang_emb": rel(o[3], lang_emb)}
        print("   oracle vs reference (rel):", {k: f"{v:.2e}" for k, v in errs.items()})
        assert max(errs.values()) <= 1e-5, errs
        contrast = rel(vis_emb, ref_eval[2])
        print(f"   batch-stat vs standard-eval vision_emb: rel {contrast:.3f} (the two modes are different functions)")
        assert contrast > 1e-2
        # one clip alone gives different statistics: the clips of a call are coupled
        alone = model(img[:1], ids[:1], mask[:1], return_emb=True)
        print(f"   clip 0 alone vs in the batch, logits rel {rel(alone[0], logits[:1]):.3f}")
        np.savez_compressed(
            os.path.join(GOLDEN, f"bn_batch_{name}.npz"),
            logits=logits.numpy(), probs=probs.numpy(), vision_emb=vis_emb.numpy(), lang_emb=lang_emb.numpy(),
            labels=np.array(orc.predict_labels(logits)), text_ids=ids.numpy(), attention_mask=mask.numpy(),
            clip_starts=np.array(starts), logits_clip0_alone=alone[0].numpy(),
            meta=np.array([T, L, B, 123, 8, 128]))
        print("   logits", logits.tolist())


if __name__ == "__main__":
    main()
