"""Multi-GPU host logic on CPU: world_size 2 over gloo.  Environments shard across ranks with no data-path
collective; only the timing / frame-count reduction and the synchronous-PAAC gradient all-reduce communicate."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util

sys.path.insert(0, util.ROOT)
import bench  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 8
    # each rank owns a contiguous block of global env ids; ALE seeds follow the global id (atari_emulator.py:20)
    ids = np.arange(rank * n, (rank + 1) * n)
    seeds = 3 * (ids + 1)
    # per-rank results of a fake timed region
    ms, frames = 100.0 + 10 * rank, 1000.0 * (rank + 1)
    stat = torch.tensor([ms, frames], dtype=torch.float64)
    mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    grad = torch.full((1000,), float(rank + 1))
    dist.all_reduce(grad)
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds.tolist())
    # the episode statistics of bench.py's collective step: count, sum return, sum length, min, max, global steps
    st = torch.tensor([3.0 + rank, 10.0 * (rank + 1), 100.0, -5.0 + 7 * rank, 20.0 - 30 * rank, 640.0], dtype=torch.float64)
    ps, pm = bench.pack_episode_stats(st)
    dist.all_reduce(ps)
    dist.all_reduce(pm, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put((float(mx[0]), float(sm[1]), float(grad[0]), gathered, bench.unpack_episode_stats(ps, pm)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms_max, frames_all, g, seeds, stats = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms_max == 110.0 and frames_all == 3000.0 and g == 3.0
    flat = [s for part in seeds for s in part]
    assert flat == [3 * (i + 1) for i in range(16)]          # global, gap-free, no overlap
    assert frames_all / (ms_max / 1000.0) == 3000.0 / 0.110  # whole-job frames / max-over-ranks time
    # episodes 3 + 4, returns 10 + 20, lengths 100 + 100, min(-5, 2), max(20, -10), global steps 640 + 640
    assert stats == (7.0, 30.0, 200.0, -5.0, 20.0, 1280.0)


def test_stale_profile_counters_are_refused(tmp_path):
    """The ncu-derived counters bench.py quotes carry the hash of the sources they were profiled on: another hash (or no
    file) gives no numbers and says why; the committed ones belong to the committed sources."""
    import json
    from manette_b200 import build as mb_build
    doc, why = bench.load_counters(str(tmp_path / "nothing.json"))
    assert doc is None and "no " in why
    stale = tmp_path / "stale.json"
    stale.write_text(json.dumps({"source_hash": "0" * 16, "warp_inst_per_next": 1.0, "dram_bytes_per_next": 1.0, "launches": []}))
    doc, why = bench.load_counters(str(stale))
    assert doc is None and why.startswith("stale")
    good = tmp_path / "good.json"
    good.write_text(json.dumps({"source_hash": mb_build.source_hash(), "warp_inst_per_next": 2.0, "dram_bytes_per_next": 3.0, "launches": []}))
    doc, why = bench.load_counters(str(good))
    assert why is None and doc["warp_inst_per_next"] == 2.0
    for path in (bench.ROUND_COUNTERS, bench.K3_COUNTERS):
        doc, why = bench.load_counters(path)
        assert why is None, why          # profiles/ was regenerated after the last change to csrc/


def test_split_games_covers_every_env():
    groups = bench.split_games(bench.GAMES12, 16384)
    assert sum(k for _, k in groups) == 16384 and len(groups) == 12
    assert max(k for _, k in groups) - min(k for _, k in groups) <= 1
    assert bench.split_games(["pong"], 32) == [("pong", 32)]


def test_workloads_match_baseline_configs():
    w = bench.WORKLOADS
    assert w["pong_paac_n32"]["n"] == 32 and w["pong_paac_n32"]["nb_choices"] == 1
    assert w["breakout_figar10_n256"]["n"] == 256 and w["breakout_figar10_n256"]["max_rep"] == 10
    assert w["seaquest_figar10_rgb_n4096"]["rgb"] and w["seaquest_figar10_rgb_n4096"]["n"] == 4096
    assert w["ms_pacman_figar10_n16384"]["n"] == 16384
    assert sorted(w["mixed12_figar10_n16384"]["games"]) == sorted(util.GAMES12)
    assert bench.tab_repetitions(10, 11) == list(range(11))
