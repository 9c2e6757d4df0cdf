#!/usr/bin/env python
"""bench.py — candidate clips scored per second by the B200-native chapter-boundary scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

The measured path is the WHOLE two-stream scorer (two_stream.py:172-194 with its caller-side pre-processing):
uint8 HWC frames -> normalise/cast -> ResNet-50 + TSM -> BERT-base -> ChapterHead (mlp) -> softmax, T=16 frames,
L=100 tokens, bf16, synthetic data, random-init weights (vcg_b200/synthetic.py, seed 123).

N = 1  BASELINE.json configs[2]: one step = one synthetic 1-hour video at 1 fps (3600 frames -> 896 candidate clips).
  value  : clips/s with the frames / tokens resident in HBM (CUDA events, barrier + synchronize both sides)
  e2e    : the same through the host-buffer C-ABI call (vcg_score_clips_u8_host): pinned host frames / ids / masks,
           H2D (541 MB per step) + D2H of the scores inside the timed region
  extra  : the full-length variant (every attention mask all ones), configs[1] (precomputed vision embeddings, batch
           256: the text stream + head alone) with both mask distributions, per-layer / per-kernel profile
N > 1  BASELINE.json configs[3]: 1024 synthetic 10-minute videos (600 frames -> 146 clips each) clip-sharded over the
  ranks, video-major.  One step = one scoring round of 6 videos (876 clips) PER RANK (weak scaling) followed by one NCCL
  all-gather of the round's [876*N, 2] boundary logits; the 1024-video job is ceil(1024 / 6N) such rounds.
  extra.configs3_full_job: the whole 149 504-clip job once (every rank scores ceil(1024/N) videos, one all-gather of
  all logits at the end), wall seconds (max over ranks) and clips/s.
--impl reference: the reference's CPU path (the oracle restatement of the reference forward: /root/reference does not
exist on the GPU box) on the host cores, rank 0 only; every step is a bounded sample (CPU_SAMPLE_CLIPS clips) of the
same workload, same K and W.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "video-chapter-generation_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T, L, HIDDEN = 16, 100, 128
HOUR_FRAMES = 3600                 # configs[2]: 1-hour video at 1 fps
VIDEO_FRAMES = 600                 # configs[3]: 10-minute videos
VIDEOS_PER_ROUND = 6               # per rank and step at N > 1 (6 x 146 = 876 clips)
JOB_VIDEOS = 1024                  # configs[3]
TEXT_BATCH = 256                   # configs[1]
VISION_CHUNK = int(os.environ.get("VCG_VISION_CHUNK", "128"))   # clips per vision pass (24 GB workspace; 128 measured 1.2 % faster than 64)
CPU_SAMPLE_CLIPS = 4
METRIC = "candidate clips scored/sec"


def flops_per_clip(t, l, vision=True):
    """algorithmic FLOPs of one clip (2*MAC over conv/matmul)."""
    f = 169869312.0 * l + 36864.0 * l * l + 1179648.0 + 2.0 * (768 * 128 + t * 2048 * 128 + (t + 1) * 128 * 2)
    if vision:
        f += 8.174272512e9 * t
    return f


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle sampling DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "power_w": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, pw, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                pw.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            # samples under load = those drawing more than half of the maximum power seen
            lim = 0.5 * max(pw)
            busy = sorted(s for s, p in zip(sm, pw) if p >= lim) or sorted(sm)
            out["sm_mhz"] = busy[len(busy) // 2]
            out["power_w"] = max(pw)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def timed(fn, steps, warmup, dist_on, drain=None):
    """W warm-up steps, then K timed steps bracketed by barrier + synchronize; returns (max-over-ranks seconds,
    this rank's seconds).  drain(): completes whatever fn() left in flight (asynchronous all-gathers); it runs INSIDE
    the timed region."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if drain:
        drain()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if drain:
        drain()
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    own = e0.elapsed_time(e1) / 1e3
    sec = own
    if dist_on:
        t = torch.tensor([own], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec, own


# ------------------------------------------------------------------------------------------------ workloads
def workload_config(n_gpus):
    base = {"clip_frames": T, "tokens": L, "head_type": "mlp", "token_lengths": "U{10..100} (SURVEY.md 8d)",
            "vision_pass_clips": VISION_CHUNK,
            "l2": "no explicit flush: every step streams >= 540 MB of uint8 frames and several GB of activations, "
                  "far beyond the 126 MB L2",
            "reference_sample": f"--impl reference / cpu_baseline time {CPU_SAMPLE_CLIPS} clips of this workload per step"}
    if n_gpus == 1:
        base.update({"workload": "configs[2]: whole per-video pipeline (uint8 frame preprocessing + ResNet-50-TSM vision "
                                 "emb + BERT-base text + ChapterHead) over a synthetic 1-hour video at 1 fps; one step = "
                                 "the whole video (3600 frames -> 896 clips)",
                     "clips_per_step_per_gpu": len(range(0, HOUR_FRAMES - T, 4)), "parallelism": "single GPU"})
    else:
        base.update({"workload": "configs[3]: 1024 synthetic 10-minute videos (600 frames -> 146 clips each) clip-sharded, "
                                 f"video-major; one step = one round of {VIDEOS_PER_ROUND} videos per rank (whole pipeline "
                                 "from uint8 frames) + one NCCL all-gather of the round's boundary logits",
                     "clips_per_step_per_gpu": VIDEOS_PER_ROUND * len(range(0, VIDEO_FRAMES - T, 4)),
                     "parallelism": f"clip-sharded x{n_gpus}, no data-path collective, one all-gather of [876*{n_gpus},2] "
                                    "logits per step (overlapped with the next step)"})
    return base


def video_major_starts(n_videos, frames_per_video):
    """Clip starts of n_videos videos stored back to back in one frame buffer (infer_youtube_video_dataset.py:117)."""
    out = []
    for v in range(n_videos):
        out += [v * frames_per_video + s for s in range(0, frames_per_video - T, 4)]
    return out


def cpu_port_clips_per_s(n_clips, steps=1, warmup=0, threads=None):
    """The reference's CPU path (oracle restatement of the WHOLE forward: preprocess + ResNet-50-TSM + BERT + head) on a
    bounded sample of the workload: n_clips consecutive clips of the synthetic video per step."""
    from oracle import two_stream_oracle as orc
    from vcg_b200 import synthetic as W
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = W.make_state_dict(T, "mlp", seed=123)
    starts = [4 * b for b in range(n_clips)]
    frames = W.make_frames_u8(starts[-1] + T, seed=5)
    ids, mask = W.make_text(n_clips, L, seed=5)

    def one():
        img = orc.gather_clips(orc.preprocess_u8(frames), starts, T)     # the caller-side ToTensor + Normalize
        return orc.two_stream_forward(sd, img, ids, mask, T, HIDDEN, "mlp", 8)

    with torch.no_grad():
        for _ in range(warmup):
            one()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = time.perf_counter() - t0
    return n_clips * steps / dt, dt / steps, threads


def cpu_sample_text(threads):
    return (f"{CPU_SAMPLE_CLIPS} consecutive clips (stride 4 frames) of the same workload per step: uint8 frames -> "
            f"preprocess -> ResNet-50-TSM + BERT + head, oracle restatement of the reference forward, torch fp32 CPU, "
            f"{threads} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cps, sec_per_step, threads = cpu_port_clips_per_s(CPU_SAMPLE_CLIPS, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                             "sample": cpu_sample_text(threads)},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def summarise_profile(prof, key):
    agg = {}
    for r in prof:
        k = agg.setdefault(r[key], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        for f in ("ms", "flops", "bytes", "launches"):
            k[f] += r[f]
    tot = sum(k["ms"] for k in agg.values())
    return {name: {"ms": round(k["ms"], 3), "share": round(k["ms"] / tot, 4), "launches": k["launches"],
                   "tflops": round(k["flops"] / k["ms"] / 1e9, 1) if k["flops"] else None,
                   "gbs": round(k["bytes"] / k["ms"] / 1e6, 1) if k["bytes"] else None}
            for name, k in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}, agg, tot


def text_only_extra(args, peaks):
    """configs[1]: the model on PRECOMPUTED vision embeddings, 256 clips per step (text stream + head only)."""
    from vcg_b200 import synthetic as W
    from vcg_b200.engine import Engine
    sd = W.make_state_dict(T, "mlp", seed=123, include_vision=False)
    eng = Engine(T, "mlp", "bf16", vision=False, max_tokens=128, max_batch=TEXT_BATCH)
    eng.load_state_dict(sd)
    del sd
    out = {"workload": "configs[1]: precomputed vision embeddings [256,16,2048] + text, batch 256, bf16; one measurement = "
                       "400 forward passes"}
    n_pass = 400
    for name, full in (("lengths_U10_100", False), ("full_length", True)):
        emb, ids, mask = W.make_precomputed_inputs(TEXT_BATCH, T, L, seed=1000, full_length=full)
        emb_h, ids_h, mask_h = emb.pin_memory(), ids.pin_memory(), mask.pin_memory()
        emb_d, ids_d, mask_d = emb_h.cuda(), ids_h.cuda(), mask_h.cuda()

        def passes():
            for _ in range(n_pass):
                eng.forward(None, ids_d, mask_d, vision_emb=emb_d)
        sec, _ = timed(passes, 1, 1, False)
        host_out = (torch.empty(TEXT_BATCH, 2).pin_memory(), torch.empty(TEXT_BATCH, 2).pin_memory())

        def passes_host():
            for _ in range(n_pass // 4):
                eng.forward_host(emb_h, ids_h, mask_h, out=host_out)
        sec_h, _ = timed(passes_host, 1, 1, False)
        eng.profile_begin()
        for _ in range(4):
            eng.forward(None, ids_d, mask_d, vision_emb=emb_d)
        prof = eng.profile_end()
        ksum, agg, tot = summarise_profile(prof, "kernel")
        flops = sum(k["flops"] for k in agg.values()) / 4
        cps = TEXT_BATCH * n_pass / sec
        out[name] = {"value": cps, "unit": "clips/s", "ms_per_pass": sec / n_pass * 1e3, "timed_region_s": sec,
                     "e2e": TEXT_BATCH * (n_pass // 4) / sec_h,
                     "tflops_executed": flops / (sec / n_pass) / 1e12,
                     "frac_of_sustained_bf16_peak": flops / (sec / n_pass) / 1e12 / peaks["tf_sustained"],
                     "mean_tokens": float(mask.sum()) / TEXT_BATCH,
                     "kernels": {n: {"share": k["share"], "tflops": k["tflops"], "launches_per_pass": k["launches"] // 4}
                                 for n, k in ksum.items()}}
    eng.close()
    return out


def emit(line):
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version there) are kept off it."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)          # anything else that writes to fd 1 goes to stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the extra measurements (configs[1], full job, ...)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if dist_on:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    from vcg_b200 import synthetic as W
    from vcg_b200 import distributed as vd
    from vcg_b200.engine import Engine
    peaks = measured_peaks()

    sd = W.make_state_dict(T, "mlp", seed=123)
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=VISION_CHUNK)
    eng.load_state_dict(sd)
    del sd

    # ---- this rank's step: 3600 uint8 frames, either one 1-hour video (N = 1) or 6 ten-minute videos (N > 1)
    if dist_on:
        starts = video_major_starts(VIDEOS_PER_ROUND, VIDEO_FRAMES)
    else:
        starts = W.clip_starts(HOUR_FRAMES, T)
    B = len(starts)
    g = torch.Generator().manual_seed(5 + rank)
    frames_h = torch.randint(0, 256, (HOUR_FRAMES, 224, 224, 3), generator=g, dtype=torch.uint8).pin_memory()
    ids_h, mask_h = W.make_text(B, L, seed=5 + rank)
    ids_h, mask_h = ids_h.pin_memory(), mask_h.pin_memory()
    starts_h = torch.tensor(starts, dtype=torch.int32).pin_memory()
    frames_d, ids_d, mask_d, starts_d = frames_h.cuda(), ids_h.cuda(), mask_h.cuda(), starts_h.cuda()
    full_mask_d = torch.ones_like(mask_d)
    full_ids_d = torch.where(ids_d == 0, torch.full_like(ids_d, 2000), ids_d)

    # the all-gather of a step's logits runs on NCCL's stream underneath the next step's kernels; at most two are in
    # flight, and every one has completed before the timed region ends (drain)
    pending = []

    def gather_async(logits_dev):
        pending.append(vd.allgather_scores_async(logits_dev, B * world))
        if len(pending) > 2:
            pending.pop(0)[1].wait()

    def drain():
        while pending:
            pending.pop(0)[1].wait()

    def step_device():
        logits, _ = eng.score_clips_u8(frames_d, starts_d, ids_d, mask_d, clip_start_host=starts_h)
        if dist_on:
            gather_async(logits)

    def step_device_full():
        eng.score_clips_u8(frames_d, starts_d, full_ids_d, full_mask_d, clip_start_host=starts_h)

    host_out = (torch.empty(B, 2).pin_memory(), torch.empty(B, 2).pin_memory())

    def step_host():
        logits, _ = eng.score_clips_u8_host(frames_h, starts_h, ids_h, mask_h, out=host_out)
        if dist_on:
            gather_async(logits.cuda(non_blocking=True))

    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count
    sec, own = timed(step_device, args.steps, args.warmup, dist_on, drain if dist_on else None)
    launches = (eng.launch_count - l0) * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop()
    sec_e2e, _ = timed(step_host, args.steps, min(args.warmup, 3), dist_on, drain if dist_on else None)

    # per-kernel CUDA-event profile of one step (events on the launching stream, separate pass so that the timed
    # region above carries no event overhead)
    eng.profile_begin()
    step_device()
    if dist_on:
        drain()
    prof = eng.profile_end()
    ksum, kagg, tot_ms = summarise_profile(prof, "kernel")
    lsum, _, _ = summarise_profile(prof, "layer")
    # the dominant kernel = the (kernel, layer) pair with the largest share of the step: launches of one pair do the same
    # work per row, so "algorithmic work per launch / launch duration" is well defined for it.  The fused bottleneck tails of
    # layer1 / layer2 (conv23h / conv23) and the 1x1 convs of layer1 are HBM-bound: their roofline is
    # bytes; every other GEMM kernel's is bf16 tensor FLOPs.
    pairs = {}
    for r in prof:
        k = pairs.setdefault((r["kernel"], r["layer"]), {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        for f in ("ms", "flops", "bytes", "launches"):
            k[f] += r[f]
    (dom_kernel, dom_layer), dom = max(pairs.items(), key=lambda kv: kv[1]["ms"])
    dom_name = f"{dom_kernel}|{dom_layer}"
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath))
    tr = traffic_tab.get(dom_name)
    traffic = None
    if isinstance(tr, dict):      # DRAM bytes of one ncu-captured launch, scaled to this run's mean launch size by algorithmic bytes
        per_launch = (dom["bytes"] if dom["bytes"] else dom["flops"]) / dom["launches"]
        traffic = tr["dram_bytes"] * per_launch / tr["algorithmic_per_launch"]
    elif tr is not None:
        traffic = tr
    # a pair is HBM-bound when its algorithmic intensity (FLOPs per HBM byte) is below the machine balance
    balance = peaks["tf_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)

    def is_hbm(v):
        return v["bytes"] > 0 and (v["flops"] == 0 or v["flops"] / v["bytes"] < balance)
    hbm_bound = is_hbm(dom)
    if hbm_bound:
        achieved = dom["bytes"] / dom["ms"] / 1e6   # GB/s
        roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                    "peak_source": peaks["source"] + ", HBM copy bandwidth",
                    "share_of_step": dom["ms"] / tot_ms, "launches_per_step": dom["launches"],
                    "tflops": dom["flops"] / dom["ms"] / 1e9,
                    "intensity_flop_per_byte": dom["flops"] / dom["bytes"], "machine_balance": balance,
                    "how": "algorithmic bytes (activations in + weights + residual + output + shifted copy, bf16) of every launch of this "
                           "kernel in one step / its CUDA-event time; traffic = dram bytes of one ncu launch (profiles/traffic.json) "
                           "scaled to the mean launch size"}
    else:
        achieved = dom["flops"] / dom["ms"] / 1e9   # TFLOP/s
        roofline = {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peaks["tf_sustained"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "share_of_step": dom["ms"] / tot_ms, "launches_per_step": dom["launches"],
                    "how": "algorithmic 2*M*N*K FLOPs of every launch of this kernel in one step / its CUDA-event time"}
    # the largest tensor-bound pair next to it, for the tensor-pipe fraction the metric asks for
    tens = {k: v for k, v in pairs.items() if v["flops"] > 0 and not is_hbm(v)}
    (tk, tl), tv = max(tens.items(), key=lambda kv: kv[1]["ms"])
    roofline_tensor = {"kernel": f"{tk}|{tl}", "intensity_flop_per_byte": (tv["flops"] / tv["bytes"]) if tv["bytes"] else None, "achieved": tv["flops"] / tv["ms"] / 1e9, "peak": peaks["tf_sustained"],
                       "unit": "TFLOP/s", "frac": tv["flops"] / tv["ms"] / 1e9 / peaks["tf_sustained"],
                       "share_of_step": tv["ms"] / tot_ms, "launches_per_step": tv["launches"]}

    value = B * world * args.steps / sec
    e2e_value = B * world * args.steps / sec_e2e
    fl = flops_per_clip(T, L, vision=True)
    executed = sum(k["flops"] for k in kagg.values())
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(n_gpus),
        "timed_region_s": sec,
        "e2e": {"value": e2e_value, "unit": "clips/s",
                "h2d_bytes_per_step": int(frames_h.numel() + ids_h.numel() * 16 + starts_h.numel() * 4),
                "d2h_bytes_per_step": B * 2 * 4 * 2},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline, "roofline_tensor": roofline_tensor,
        # executed = FLOPs the kernels really did (the text stream drops masked tokens, which is exact; the vision
        # stream runs the stem once per distinct frame); algorithmic = SURVEY.md 8d's per-clip figure at full length
        "whole_path": {"tflops_algorithmic": value / n_gpus * fl / 1e12,
                       "frac_of_sustained_bf16_peak": value / n_gpus * fl / 1e12 / peaks["tf_sustained"],
                       "frac_of_burst_bf16_peak": value / n_gpus * fl / 1e12 / peaks["tf_burst"],
                       "flops_per_clip_algorithmic": fl,
                       "tflops_executed": executed / (sec / args.steps) / 1e12,
                       "note": "algorithmic = 148.155 GF per clip (SURVEY.md 8d, L=100 padded, per-clip stem); executed "
                               "counts what ran (packed tokens, stem once per distinct frame)"},
        "kernels": ksum, "layers": lsum,
    }
    if dist_on:
        # per-rank diagnostics: own step time, clocks, dominant-kernel rate
        mine = {"rank": rank, "ms_per_step": own / args.steps * 1e3, "sm_mhz": clocks["sm_mhz"], "power_w": clocks["power_w"],
                "reasons": clocks["reasons"], "dominant_achieved": achieved, "dominant_unit": roofline["unit"]}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        ms = sorted(r["ms_per_step"] for r in allr)
        line["per_rank"] = {"ms_per_step_min": ms[0], "ms_per_step_median": ms[len(ms) // 2], "ms_per_step_max": ms[-1],
                            "ranks": allr}
    extra = {}
    if not args.no_extra:
        try:
            sec_full, _ = timed(step_device_full, max(2, args.steps // 4), 1, dist_on)
            extra["full_length_masks"] = {"value": B * world * max(2, args.steps // 4) / sec_full, "unit": "clips/s",
                                          "note": "same step with every attention mask all ones (100 tokens per clip)"}
        except Exception as ex:
            extra["full_length_masks"] = {"error": str(ex)}
        if dist_on:
            try:
                extra["configs3_full_job"] = full_job(eng, vd, frames_d, starts, ids_d, mask_d, rank, world)
            except Exception as ex:
                extra["configs3_full_job"] = {"error": str(ex)}
    eng.close()
    del frames_d
    torch.cuda.empty_cache()
    if rank == 0 and n_gpus == 1:
        cps, _, threads = cpu_port_clips_per_s(CPU_SAMPLE_CLIPS, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                                "sample": cpu_sample_text(threads) + ", 2 timed steps after 1 warm-up"}
        if not args.no_extra:
            try:
                extra["configs1_text_stream"] = text_only_extra(args, peaks)
            except Exception as ex:  # the headline line must still be printed
                extra["configs1_text_stream"] = {"error": str(ex)}
    if extra:
        line["extra"] = extra
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def full_job(eng, vd, frames_d, starts, ids_d, mask_d, rank, world):
    """configs[3] once, strong scaling: 1024 videos sharded video-major over the ranks (ceil(1024/W) videos each; the
    synthetic frame store holds 6 videos, re-used round after round), one all-gather of all 149 504 x 2 logits at the end."""
    import torch.distributed as dist
    clips_per_video = len(range(0, VIDEO_FRAMES - T, 4))
    lo, hi = vd.shard_range(JOB_VIDEOS, rank, world)
    n_local = hi - lo
    per = vd.shard_size(JOB_VIDEOS, world) * clips_per_video
    local = torch.zeros(per, 2, device="cuda")
    starts_t = torch.tensor(starts, dtype=torch.int32)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    done = 0
    while done < n_local:
        nv = min(VIDEOS_PER_ROUND, n_local - done)
        nb = nv * clips_per_video
        sh = starts_t[:nb].contiguous()
        out = (local[done * clips_per_video: done * clips_per_video + nb], torch.empty(nb, 2, device="cuda"))
        eng.score_clips_u8(frames_d, sh.cuda(), ids_d[:nb], mask_d[:nb], out=out, clip_start_host=sh)
        done += nv
    allv = vd.allgather_scores(local, per * world)
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    n_clips = JOB_VIDEOS * clips_per_video
    return {"videos": JOB_VIDEOS, "clips": n_clips, "seconds": float(dt.item()), "value": n_clips / float(dt.item()),
            "unit": "clips/s", "scaling": "strong", "gathered_rows": int(allv.shape[0]),
            "note": "wall clock between barriers, max over ranks; includes the final all-gather of every logit"}


if __name__ == "__main__":
    main()
