#!/usr/bin/env python
"""bench.py — candidate clips scored per second by the B200-native chapter-boundary scorer.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): the two-stream point model (BERT-base text stream + fusion head, mlp head,
T=16, L=100) on PRECOMPUTED vision embeddings, 256 clips per step per GPU, bf16, synthetic data, random-init weights
(vcg_b200/synthetic.py, seed 123).  One step = one pass of TwoStream.forward over one batch of 256 clips.
  value  : clips/s with inputs resident in HBM (CUDA events, max over ranks, barrier + synchronize both sides)
  e2e    : the same through the host-buffer C-ABI call (vcg_forward_host): pinned host inputs, H2D + D2H inside
  extra  : at N=1 also the whole pipeline of configs[2] (uint8 frames -> preprocess -> ResNet-50-TSM + BERT + head
           over a synthetic 1-hour video, 896 clips) device-resident and end-to-end from host buffers
N > 1: every rank scores its own 256 clips (weak scaling) and one NCCL all-gather of the [256,2] logits per step puts
all scores on every rank (clips are independent: no other collective).
--impl reference: the reference's CPU path (oracle restatement of the reference forward; /root/reference does not
exist on the GPU box) timed on the host cores, rank 0 only, each step a bounded sample of the same workload.
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

T, L, BATCH, HIDDEN = 16, 100, 256, 128
CPU_SAMPLE_CLIPS = 16
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
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
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
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            busy = sm[len(sm) // 2:]           # upper half = samples under load
            out["sm_mhz"] = busy[len(busy) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def make_inputs(batch, seed):
    from vcg_b200 import synthetic as W
    ids, mask = W.make_text(batch, L, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    # precomputed vision embeddings are post-ReLU average-pooled features: non-negative, O(1)
    emb = torch.rand(batch, T, 2048, generator=g) * 2.0
    return emb, ids, mask


def timed(fn, steps, warmup, dist_on, drain=None):
    """W warm-up steps, then K timed steps bracketed by barrier + synchronize; returns max-over-ranks seconds.
    drain(): completes whatever fn() left in flight (asynchronous all-gathers); it runs INSIDE the timed region."""
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
    sec = e0.elapsed_time(e1) / 1e3
    if dist_on:
        t = torch.tensor([sec], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


def cpu_port_clips_per_s(n_clips, steps=1, warmup=0, threads=None):
    """The reference's CPU path (oracle restatement) on a bounded sample of the workload."""
    from oracle import two_stream_oracle as orc
    from vcg_b200 import synthetic as W
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = W.make_state_dict(T, "mlp", seed=123, include_vision=False)
    emb, ids, mask = make_inputs(n_clips, seed=123)
    with torch.no_grad():
        for _ in range(warmup):
            orc.two_stream_forward(sd, None, ids, mask, T, HIDDEN, "mlp", 8, vision_emb=emb)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.two_stream_forward(sd, None, ids, mask, T, HIDDEN, "mlp", 8, vision_emb=emb)
        dt = time.perf_counter() - t0
    return n_clips * steps / dt, dt / steps, threads


def workload_config(n_gpus):
    return {"workload": "configs[1]: two-stream point model on precomputed vision embeddings "
                        "(BERT-base text stream + ChapterHead mlp), T=16 frames, L=100 tokens, batch 256 clips/GPU",
            "clips_per_step_per_gpu": BATCH, "clip_frames": T, "tokens": L, "head_type": "mlp",
            "parallelism": f"clip-sharded x{n_gpus}, one NCCL all-gather of the [256,2] logits per step, overlapped with the next step" if n_gpus > 1 else "single GPU",
            "l2": "no explicit flush: per-step working set (220 MB bf16 weights + >500 MB activations) exceeds the 126 MB L2"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cps, sec_per_step, threads = cpu_port_clips_per_s(CPU_SAMPLE_CLIPS, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                             "sample": f"{CPU_SAMPLE_CLIPS} clips per step of the same workload (oracle restatement of "
                                       "the reference forward, torch fp32 CPU)"},
            "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def pipeline_extra(args, peaks):
    """configs[2]: whole per-video pipeline over a synthetic 1-hour video at 1 fps (3600 frames -> 896 clips)."""
    from vcg_b200 import synthetic as W
    from vcg_b200.engine import Engine
    n_frames = 3600
    starts = W.clip_starts(n_frames, T)
    B = len(starts)
    sd = W.make_state_dict(T, "mlp", seed=123)
    eng = Engine(T, "mlp", "bf16", vision=True, max_tokens=128, max_batch=int(os.environ.get("VCG_VISION_CHUNK", "64")))   # clips per vision pass (12 GB workspace)
    eng.load_state_dict(sd)
    del sd
    g = torch.Generator().manual_seed(5)
    frames_h = torch.randint(0, 256, (n_frames, 224, 224, 3), generator=g, dtype=torch.uint8).pin_memory()
    ids_h, mask_h = W.make_text(B, L, seed=5)
    ids_h, mask_h = ids_h.pin_memory(), mask_h.pin_memory()
    starts_h = torch.tensor(starts, dtype=torch.int32).pin_memory()
    frames_d, ids_d, mask_d, starts_d = frames_h.cuda(), ids_h.cuda(), mask_h.cuda(), starts_h.cuda()
    steps, warm = max(2, min(args.steps, 5)), 1
    sec = timed(lambda: eng.score_video_u8(frames_d, 0, 4, ids_d, mask_d), steps, warm, False)
    out = (torch.empty(B, 2).pin_memory(), torch.empty(B, 2).pin_memory())
    sec_e2e = timed(lambda: eng.score_clips_u8_host(frames_h, starts_h, ids_h, mask_h, out=out), steps, warm, False)
    eng.profile_begin()
    eng.score_video_u8(frames_d, 0, 4, ids_d, mask_d)
    prof = eng.profile_end()
    cps = B * steps / sec
    fl = flops_per_clip(T, L, True)
    kern = {}
    for r in prof:
        k = kern.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        for f in ("ms", "flops", "bytes", "launches"):
            k[f] += r[f]
    layers = {}
    for r in prof:
        k = layers.setdefault(r["layer"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
        for f in ("ms", "flops", "bytes", "launches"):
            k[f] += r[f]
    tot_ms = sum(k["ms"] for k in kern.values())
    summary = {name: {"ms": round(k["ms"], 3), "share": round(k["ms"] / tot_ms, 4), "launches": k["launches"],
                      "tflops": round(k["flops"] / k["ms"] / 1e9, 1) if k["flops"] else None,
                      "gbs": round(k["bytes"] / k["ms"] / 1e6, 1) if k["bytes"] else None}
               for name, k in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])}
    lsummary = {name: {"ms": round(k["ms"], 3), "share": round(k["ms"] / tot_ms, 4),
                       "tflops": round(k["flops"] / k["ms"] / 1e9, 1) if k["flops"] else None,
                       "gbs": round(k["bytes"] / k["ms"] / 1e6, 1) if k["bytes"] else None}
                for name, k in sorted(layers.items(), key=lambda kv: -kv[1]["ms"])}
    eng.close()
    return {"workload": "configs[2]: uint8 frames -> preprocess -> ResNet-50-TSM + BERT + head, synthetic 1-hour video "
                        "(3600 frames, 896 clips), bf16",
            "value": cps, "unit": "clips/s", "ms_per_video": sec / steps * 1e3,
            "tflops": cps * fl / 1e12, "frac_of_sustained_bf16_peak": cps * fl / 1e12 / peaks["tf_sustained"],
            "e2e": {"value": B * steps / sec_e2e, "unit": "clips/s", "h2d_bytes_per_step": int(frames_h.numel() + ids_h.numel() * 16 + starts_h.numel() * 4),
                    "d2h_bytes_per_step": B * 16},
            "kernels": summary, "layers": lsummary}


def emit(line):
    """The contract is ONE JSON line on stdout: libraries (NCCL prints its version there) are kept off it."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)          # anything else that writes to fd 1 goes to stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the whole-pipeline (configs[2]) measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

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

    sd = W.make_state_dict(T, "mlp", seed=123, include_vision=False)
    eng = Engine(T, "mlp", "bf16", vision=False, max_tokens=128, max_batch=BATCH)
    eng.load_state_dict(sd)
    del sd

    emb_h, ids_h, mask_h = make_inputs(BATCH, seed=1000 + rank)
    emb_h, ids_h, mask_h = emb_h.pin_memory(), ids_h.pin_memory(), mask_h.pin_memory()
    emb_d, ids_d, mask_d = emb_h.cuda(), ids_h.cuda(), mask_h.cuda()
    # the all-gather of a step's [256, 2] logits runs on NCCL's stream underneath the next step's kernels; at most four
    # are in flight, and every one has completed before the timed region ends (drain)
    pending = []

    def gather_async(logits_dev):
        pending.append(vd.allgather_scores_async(logits_dev, BATCH * world))
        if len(pending) > 4:
            pending.pop(0)[1].wait()

    def drain():
        while pending:
            pending.pop(0)[1].wait()

    def step_device():
        logits, _ = eng.forward(None, ids_d, mask_d, vision_emb=emb_d)
        if dist_on:
            gather_async(logits)

    host_out = (torch.empty(BATCH, 2).pin_memory(), torch.empty(BATCH, 2).pin_memory())

    def step_host():
        logits, _ = eng.forward_host(emb_h, ids_h, mask_h, out=host_out)
        if dist_on:
            gather_async(logits.cuda(non_blocking=True))

    sampler = ClockSampler(local_rank)
    launches0 = eng.launch_count
    if rank == 0:
        sampler.start()
    sec = timed(step_device, args.steps, args.warmup, dist_on, drain if dist_on else None)
    clocks = sampler.stop() if rank == 0 else None
    launches = (eng.launch_count - launches0) * args.steps // (args.steps + args.warmup)
    sec_e2e = timed(step_host, args.steps, args.warmup, dist_on, drain if dist_on else None)

    # per-kernel CUDA-event profile over the same K steps (events on the launching stream, separate pass so that
    # the timed region above carries no event overhead)
    eng.profile_begin()
    for _ in range(args.steps):
        step_device()
    prof = eng.profile_end()
    kern = {}
    for r in prof:
        k = kern.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "launches": 0})
        k["ms"] += r["ms"]; k["flops"] += r["flops"]; k["launches"] += r["launches"]
    tot_ms = sum(k["ms"] for k in kern.values())
    dom_name, dom = max(kern.items(), key=lambda kv: kv[1]["ms"])
    achieved = dom["flops"] / dom["ms"] / 1e9   # TFLOP/s

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom_name)

    value = BATCH * world * args.steps / sec
    e2e_value = BATCH * world * args.steps / sec_e2e
    fl = flops_per_clip(T, L, vision=False)
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(n_gpus),
        "e2e": {"value": e2e_value, "unit": "clips/s",
                "h2d_bytes_per_step": int(emb_h.numel() * 4 + ids_h.numel() * 8 + mask_h.numel() * 8),
                "d2h_bytes_per_step": BATCH * 2 * 4 * 2},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": dom_name, "achieved": achieved, "peak": peaks["tf_sustained"],
                     "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"], "traffic": traffic,
                     "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                     "share_of_step": dom["ms"] / tot_ms, "launches_per_step": dom["launches"] // args.steps,
                     "how": "executed 2*M*N*K FLOPs of every launch of this kernel (M = packed token rows) / its CUDA-event time"},
        # executed = the FLOPs the kernels really did (the text stream drops masked tokens, which is exact: synthetic
        # lengths are U{10..L}); padded = SURVEY.md 8d's per-clip figure at the full length L, for reference only
        "whole_path": {"tflops_executed": sum(k["flops"] for k in kern.values()) / args.steps / (sec / args.steps) / 1e12,
                       "frac_of_sustained_bf16_peak": sum(k["flops"] for k in kern.values()) / args.steps / (sec / args.steps) / 1e12 / peaks["tf_sustained"],
                       "flops_per_clip_padded": fl, "tflops_at_padded_flops": value / n_gpus * fl / 1e12,
                       "note": "frac uses executed FLOPs; token packing skips masked tokens exactly, so clips/s x padded FLOPs would overstate the tensor work"},
        "kernels": {n: {"share": round(k["ms"] / tot_ms, 4), "tflops": round(k["flops"] / k["ms"] / 1e9, 1) if k["flops"] else None,
                        "launches_per_step": k["launches"] // args.steps} for n, k in sorted(kern.items(), key=lambda kv: -kv[1]["ms"])},
    }
    eng.close()
    if rank == 0 and n_gpus == 1:
        cps, _, threads = cpu_port_clips_per_s(CPU_SAMPLE_CLIPS, steps=1, warmup=0)
        line["cpu_baseline"] = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
                                "sample": f"{CPU_SAMPLE_CLIPS} clips of the same workload, one pass (oracle restatement of the reference forward, torch fp32 CPU)"}
        if not args.no_extra:
            try:
                line["extra"] = pipeline_extra(args, peaks)
            except Exception as ex:  # the headline line must still be printed
                line["extra"] = {"error": str(ex)}
    if dist_on:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


if __name__ == "__main__":
    main()
