"""Command-line counterpart of the reference's test_video_segment_point.py on the B200 path.

Same switches (--gpu --data_mode --model_type --clip_frame_num --batch_size --head_type --data_type; :35-45) and the same
model construction (:69-100), evaluation and result files (:228-391).  The paths the reference hard-codes (:57-66) are
flags here.  --data_mode all scores uint8 frames video by video (every frame decoded once) and computes the metrics on
the device (vcg_b200.evaluate); text / image go through the DataLoader like the reference and the same metrics code.

    python video_segment/test_video_segment_point_b200.py --clips_json test_clips_clip_frame_num_16.json \
        --img_dir youtube_video_frame_dataset --ckpt checkpoint.pth [--vocab vocab.txt] [--precision bf16|fp32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from common_utils import set_random_seed  # noqa: E402
from data.infer_youtube_video_dataset import InferYoutubeClipDataset  # noqa: E402
from model.fusion import two_stream  # noqa: E402
from model.lang import bert_hugface  # noqa: E402
from model.vision import resnet50, resnet50_tsm  # noqa: E402
from vcg_b200 import evaluate  # noqa: E402


def build_model(args):
    T = args.clip_frame_num
    lang_model = bert_hugface.BertHugface(pretrain_stage=False)
    if args.data_mode == "image":
        if args.model_type == "r50tsm":
            vision_model = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
        elif args.model_type == "r50":
            vision_model = resnet50.Resnet50(segments_size=T, pretrain_stage=False)
        else:
            raise RuntimeError(f"Unknown model_type {args.model_type}")
    else:
        vision_model = resnet50_tsm.Resnet50TSM(segments_size=T, shift_div=8, pretrain_stage=False)
    if args.data_mode == "text":
        model = lang_model
        model.build_chapter_head()
    elif args.data_mode == "image":
        model = vision_model
        model.build_chapter_head()
    elif args.data_mode == "all":
        model = two_stream.TwoStream(lang_model.base_model, vision_model.base_model, lang_model.embed_size,
                                     vision_model.feature_dim, T, 128)
        model.build_chapter_head(output_size=2, head_type=args.head_type)
    else:
        raise RuntimeError(f"Unknown data mode {args.data_mode}")
    if args.ckpt:
        checkpoint = torch.load(args.ckpt, map_location="cpu")
        model.load_state_dict(checkpoint["model_state_dict"] if "model_state_dict" in checkpoint else checkpoint)
    model = model.to(args.gpu).eval()
    model.precision = args.precision
    return model


def loader_scores(model, dataset, args):
    """The reference's batch loop (:168-207) for the single-modality models."""
    from torch.utils.data import DataLoader
    logits, probs = [], []
    for img_clip, text_ids, attention_mask, _ in DataLoader(dataset, shuffle=False, batch_size=args.batch_size,
                                                            num_workers=args.num_workers):
        if args.data_mode == "text":
            lg, pr = model(text_ids.to(args.gpu), attention_mask.to(args.gpu))
        else:
            lg, pr = model(img_clip.float().to(args.gpu))
        logits.append(lg)
        probs.append(pr)
    return torch.cat(logits), torch.cat(probs)


def main(argv=None):
    parser = argparse.ArgumentParser(description="video chapter model (B200)")
    parser.add_argument("--gpu", default=0, type=int)
    parser.add_argument("--data_mode", default="all", type=str, help="text (text only), image (image only) or all")
    parser.add_argument("--model_type", default="two_stream", type=str, help="bert, r50tsm, r50, two_stream")
    parser.add_argument("--clip_frame_num", default=16, type=int)
    parser.add_argument("--epoch", default=3000, type=int)
    parser.add_argument("--batch_size", default=16, type=int)
    parser.add_argument("--lr_decay_type", default="cosine", type=str)
    parser.add_argument("--head_type", default="mlp", type=str, help="only work on two_stream model")
    parser.add_argument("--data_type", default="all", type=str, help="all, easy, hard, ambiguous")
    # what the reference hard-codes
    parser.add_argument("--clips_json", required=True, nargs="+")
    parser.add_argument("--img_dir", required=True)
    parser.add_argument("--ckpt", default=None)
    parser.add_argument("--vocab", default=None, help="local WordPiece vocab.txt (default: bert-base-uncased)")
    parser.add_argument("--max_text_len", default=100, type=int)
    parser.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    parser.add_argument("--num_workers", default=4, type=int)
    parser.add_argument("--result_file", default="./test_results/chapter_localization/MVCG_.txt")
    parser.add_argument("--vid2cut_points_file", default="./test_results/chapter_localization/MVCG_vid2cut_points.json")
    args = parser.parse_args(argv)

    set_random_seed.use_fix_random_seed()
    torch.cuda.set_device(args.gpu)
    torch.set_grad_enabled(False)
    from transformers import BertTokenizer
    tokenizer = (BertTokenizer(vocab_file=args.vocab, do_lower_case=True) if args.vocab
                 else BertTokenizer.from_pretrained("bert-base-uncased"))
    transform = None
    if args.data_mode != "all":
        from torchvision import transforms
        transform = transforms.Compose([transforms.ToTensor(),
                                        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    dataset = InferYoutubeClipDataset(args.img_dir, args.clips_json if len(args.clips_json) > 1 else args.clips_json[0],
                                      tokenizer, args.clip_frame_num, args.max_text_len, mode=args.data_mode,
                                      transform=transform)
    model = build_model(args)
    if args.data_mode == "all":
        engine = model.get_engine(torch.device("cuda", args.gpu), args.max_text_len)
        out = evaluate.evaluate_flat_clips(engine, dataset)
    else:
        logits, probs = loader_scores(model, dataset, args)
        out = evaluate.evaluate_flat_clips(None, dataset, logits=logits, probs=probs)
    evaluate.write_results(out, args.result_file, args.vid2cut_points_file)
    print(f"mAP {out['mAP']}")
    for name in ("recall", "precision", "f-score"):
        print(f"{name} {out[name]}, {name}@3 {out[name + '@3']}, {name}@5 {out[name + '@5']}")
    return out


if __name__ == "__main__":
    main()
