"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous clip shards + one all-gather of logits."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "video-chapter-generation_b200"))
    from vcg_b200 import distributed as vd
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = torch.arange(n_clips * 2, dtype=torch.float32).view(n_clips, 2) * 0.5 - 3.0

    def score(lo, hi):          # stands in for the per-rank GPU scorer
        return full[lo:hi].clone()

    got = vd.score_sharded(score, n_clips)
    ok = torch.equal(got, full)
    # the overlapped form used by bench.py: several gathers in flight, waited for later
    lo_, hi_ = vd.shard_range(n_clips, rank, world)
    pend = [vd.allgather_scores_async(score(lo_, hi_) + k, n_clips) for k in range(3)]
    for k, (res, work) in enumerate(pend):
        work.wait()
        ok = ok and torch.equal(res, full + k)
    lo, hi = vd.shard_range(n_clips, rank, world)
    torch.save({"ok": ok, "lo": lo, "hi": hi}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_sharded_scoring_allgather(tmp_path):
    for n_clips in (146, 7, 1):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_clips, str(tmp_path)), nprocs=2, join=True)
        r = [torch.load(tmp_path / f"r{i}.pt") for i in range(2)]
        assert all(x["ok"] for x in r)
        assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == n_clips


def test_shard_ranges_cover_everything():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "video-chapter-generation_b200"))
    from vcg_b200 import distributed as vd
    for n in (0, 1, 5, 146, 149504):
        for w in (1, 2, 4, 8):
            spans = [vd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in spans) == vd.shard_size(n, w) or n == 0
