#!/usr/bin/env python
"""bench.py -- preprocessed env frames/s of the FiGAR10 environment hot path on N B200s.

One "step" = one macro step of every environment (Runners.update_environments + wait_updated):
each env runs 1 + tab_rep[k] next() calls (4 emulated frames + one 84x84xD plane each) with early exit
and in-step reset on terminal, then the stacked states / rewards / terminals are published.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = next() calls of all ranks / max-over-ranks device time with the
policy's choices already in HBM; `e2e` = the same through Runners with HOST arrays (H2D of the one-hot
actions/repetitions and D2H of states/rewards/terminals every step inside the timed region; the states are
written into the pinned host array by the pool's kernels as environments finish their repeats, not copied
after the step -- the bytes are the same, they cross PCIe underneath the remaining FiGAR rounds).
`--impl reference` times the CPU restatement of the reference's own worker pool (oracle/host_path.py
PortRunners over the C++ oracle emulator) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMES12 = ["asterix", "asteroids", "breakout", "enduro", "gopher", "gravitar", "montezuma_revenge", "ms_pacman",
           "pong", "seaquest", "space_invaders", "yars_revenge"]
ROMS = os.path.join(ROOT, "atari_roms")

# BASELINE.json configs
WORKLOADS = {
    "pong_paac_n32": dict(games=["pong"], n=32, rgb=False, nb_choices=1, max_rep=0),
    "breakout_figar10_n256": dict(games=["breakout"], n=256, rgb=False, nb_choices=11, max_rep=10),
    "seaquest_figar10_rgb_n4096": dict(games=["seaquest"], n=4096, rgb=True, nb_choices=11, max_rep=10),
    # C4 is the LSTM configuration: the learner's 5-deep observation history (paac.py:107-112) is kept by the pool
    "ms_pacman_figar10_n16384": dict(games=["ms_pacman"], n=16384, rgb=False, nb_choices=11, max_rep=10, history=5),
    "mixed12_figar10_n16384": dict(games=GAMES12, n=16384, rgb=False, nb_choices=11, max_rep=10),
    # not a BASELINE config: the game whose resets depend on RAM (steady-state leg, --steady-state)
    "yars_revenge_figar10_n4096": dict(games=["yars_revenge"], n=4096, rgb=False, nb_choices=11, max_rep=10),
}
DEFAULT_WORKLOAD = "ms_pacman_figar10_n16384"
# algorithmic HBM bytes of one next() for the emulation kernel: the two pooled raw frames it
# must leave in HBM (2 x 33,600) + machine state in and out (2 x (168 + 128)); K3: 2 raw frames read + one plane
ROUND_BYTES_PER_NEXT = 2 * 33600 + 2 * (168 + 128)
PEAK_WARP_INST_PER_S = 148 * 4 * 1.965e9   # 148 SMs x 4 sub-partitions x 1 warp instruction per clock
# ncu-derived per-next() counters (warp instructions issued, DRAM bytes moved) are NOT literals here: they are read
# from profiles/*_counters.json, which tools/ncu_counters.py writes from an ncu report and stamps with the hash of the
# CUDA sources the profiled library was built from.  A stamp that does not match the sources of the library this run
# loads makes the dependent figures null ("stale") instead of silently quoting numbers of another kernel.
ROUND_COUNTERS = os.path.join(ROOT, "profiles", "r2_k_round_counters.json")
K3_COUNTERS = os.path.join(ROOT, "profiles", "r2_k3_counters.json")
# networks whose flat fp32 gradient the synchronous-PAAC all-reduce carries (paac.py:233-256), per workload
ARCH = {"pong_paac_n32": "NIPS", "breakout_figar10_n256": "NIPS", "seaquest_figar10_rgb_n4096": "PWYX",
        "ms_pacman_figar10_n16384": "LSTM", "mixed12_figar10_n16384": "PWYX", "yars_revenge_figar10_n4096": "NIPS"}
REAL_ALE_RAW_FPS_PER_CORE = 6000.0   # commonly quoted, NOT verifiable here (ALE is not installable offline)


def load_counters(path):
    """(counters dict, None) if the file exists and was taken from the sources this run is built from, else
    (None, reason)."""
    from manette_b200 import build as mb_build
    if not os.path.exists(path):
        return None, "no %s" % os.path.relpath(path, ROOT)
    with open(path) as f:
        doc = json.load(f)
    have = mb_build.source_hash()
    if doc.get("source_hash") != have:
        return None, "stale: %s was profiled on sources %s, this run is built from %s" % (
            os.path.relpath(path, ROOT), doc.get("source_hash"), have)
    return doc, None


def pack_episode_stats(st):
    """K6's running statistics (count, sum of returns, sum of lengths, min return, max return, global steps) as the two
    vectors the ranks all-reduce: one with SUM, one with MAX (the minimum travels negated)."""
    import torch
    return torch.stack([st[0], st[1], st[2], st[5]]), torch.stack([-st[3], st[4]])


def unpack_episode_stats(stat_sum, stat_max):
    """(count, sum of returns, sum of lengths, min, max, global steps) of all ranks from the reduced vectors."""
    return (float(stat_sum[0]), float(stat_sum[1]), float(stat_sum[2]), -float(stat_max[0]), float(stat_max[1]), float(stat_sum[3]))


def k3_bytes(depth):
    return 2 * 210 * 160 + 84 * 84 * depth


def split_games(games, n):
    base, extra = divmod(n, len(games))
    return [(g, base + (1 if i < extra else 0)) for i, g in enumerate(games)]


def tab_repetitions(max_repetition, nb_choices):
    res = [0] * nb_choices
    res[-1] = max_repetition
    if nb_choices > 2:
        for i in range(1, nb_choices - 1):
            res[i] = int(max_repetition / (nb_choices - 1)) * i
    return res


def rom_bytes(game):
    with open(os.path.join(ROMS, game + ".bin"), "rb") as f:
        return f.read()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm (oracle = the checker, timed)
def _oracle_imports():
    for d in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "shims")):
        if d not in sys.path:
            sys.path.insert(0, d)
    import host_path
    import orc_loader
    import ref_harness
    orc_loader.build()
    return host_path, orc_loader, ref_harness


def cpu_pool_run(cfg, steps, warmup, budget_s=None, envs_per_core=4, kind="auto"):
    """The reference's worker pool on the box's host cores, W = all of them, 4 envs per worker (the reference's own
    default shape, train.py:102-103), same game / FiGAR configuration, uniform random policy.
      kind "reference": the reference's UNMODIFIED runners.py / emulator_runner.py / atari_emulator.py /
                        environment.py imported from /root/reference (oracle/ref_harness.py) over the CPU oracle
                        emulator standing in for ALE -- only where the reference tree exists (not on the GPU box);
      kind "port":      oracle/host_path.py PortRunners, the restatement of the same loop (fixtures prove it equal).
    Returns a dict."""
    host_path, orc_loader, ref_harness = _oracle_imports()
    if kind == "auto":
        kind = "reference" if ref_harness.available() else "port"
    cores = os.cpu_count() or 1
    n = cores * envs_per_core
    groups = split_games(cfg["games"], n)
    tab_rep = tab_repetitions(cfg["max_rep"], cfg["nb_choices"])
    act_counter = None
    if kind == "reference":
        ref = ref_harness.load()
        import ale_python_interface as shim
        act_counter = shim.enable_act_counter(n)
        emu_cls = ref.atari_emulator.AtariEmulator
    else:
        emu_cls = host_path.PortAtariEmulator
    emus, eid = [], 0
    for g, k in groups:
        a = ref_harness.Args(g, ROMS, rgb=cfg["rgb"], max_repetition=cfg["max_rep"], nb_choices=cfg["nb_choices"])
        for _ in range(k):
            emus.append(emu_cls(eid, a))
            eid += 1
    num_actions = max(len(e.get_legal_actions()) for e in emus)
    acts_per_env = np.array([len(e.get_legal_actions()) for e in emus])
    states = np.asarray([e.get_initial_state() for e in emus], dtype=np.uint8)
    variables = [states, np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros((n, num_actions), np.float32),
                 np.zeros((n, cfg["nb_choices"]), np.float32)]
    if kind == "reference":
        runners = ref.runners.Runners(tab_rep, ref.emulator_runner.EmulatorRunner, np.asarray(emus, dtype=object), cores,
                                      variables)   # paac.py:104 (self.emulators is an ndarray there, actor_learner.py:35)
    else:
        runners = host_path.PortRunners(tab_rep, emus, cores, variables)
    runners.start()
    sv = runners.get_shared_variables()
    rng = np.random.RandomState(1234)
    seen = [int(np.sum(np.frombuffer(act_counter, dtype=np.int64)))] if act_counter is not None else None

    def one_step():
        a = (rng.randint(0, 1 << 30, size=n) % acts_per_env)
        r = rng.randint(0, cfg["nb_choices"], size=n)
        sv[3][...] = 0
        sv[3][np.arange(n), a] = 1
        sv[4][...] = 0
        sv[4][np.arange(n), r] = 1
        runners.update_environments()
        runners.wait_updated()
        if act_counter is None:
            return int(runners.next_counts().sum())
        # next() = 4 act() calls; an episode that ended inside the step cost 16 more (get_initial_state,
        # atari_emulator.py:102-107), which are not next() calls of the step
        now = int(np.sum(np.frombuffer(act_counter, dtype=np.int64)))
        acts, seen[0] = now - seen[0], now
        return (acts - 16 * int(np.sum(sv[2] != 0))) // 4

    try:
        for _ in range(warmup):
            one_step()
        t0 = time.perf_counter()
        frames, done = 0, 0
        while done < steps:
            frames += one_step()
            done += 1
            if budget_s is not None and time.perf_counter() - t0 > budget_s:
                break
        dt = time.perf_counter() - t0
    finally:
        runners.stop()
        for r in getattr(runners, "runners", []):
            r.join(timeout=5)
    return {"value": frames / dt, "frames": frames, "seconds": dt, "steps": done, "cores": cores, "n_envs": n, "kind": kind,
            "sample": "%d envs (%d per core) x %d macro steps of the same game/FiGAR config on %d worker processes (%s)"
                      % (n, envs_per_core, done, cores,
                         "the reference's unmodified Runners/EmulatorRunner/AtariEmulator over the oracle emulator"
                         if kind == "reference" else "oracle port of the reference's worker pool")}


def cpp_pool_run(cfg, budget_s=4.0, envs_per_core=4):
    """Best-case CPU line: the oracle emulators stepped by one C++ thread per host core through the
    FiGAR loop, no Python, no preprocessing (oracle_capi.cpp orc_pool_step)."""
    import ctypes as C
    host_path, orc_loader, ref_harness = _oracle_imports()
    L = orc_loader.lib()
    L.orc_pool_step.restype = C.c_long
    L.orc_pool_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_create.restype = C.c_void_p
    cores = os.cpu_count() or 1
    n = cores * envs_per_core
    tab_rep = np.array(tab_repetitions(cfg["max_rep"], cfg["nb_choices"]), np.int32)
    handles, nact = [], []
    for gi, (g, k) in enumerate(split_games(cfg["games"], n)):
        rom = rom_bytes(g)
        for j in range(k):
            h = L.orc_create(rom, len(rom), g.encode(), C.c_uint32(3 * (len(handles) + 1)))
            L.orc_reset_game(C.c_void_p(h))
            handles.append(h)
            nact.append(L.orc_num_actions(C.c_void_p(h)))
    envs = (C.c_void_p * n)(*handles)
    nact = np.array(nact)
    rng = np.random.RandomState(4321)
    rewards, terms = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    frames, t0 = 0, time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < budget_s:
        a = (rng.randint(0, 1 << 30, size=n) % nact).astype(np.int32)
        r = tab_rep[rng.randint(0, cfg["nb_choices"], size=n)].astype(np.int32)
        frames += int(L.orc_pool_step(envs, n, a.ctypes.data, r.ctypes.data, cores, rewards.ctypes.data, terms.ctypes.data))
        steps += 1
    dt = time.perf_counter() - t0
    for h in handles:
        L.orc_destroy(C.c_void_p(h))
    return {"value": frames / dt, "unit": "frames/s", "cores": cores, "raw_frames_per_s_per_core": 4.0 * frames / dt / cores,
            "sample": "%d envs x %d macro steps, %d C++ threads, emulation + FiGAR loop only (no preprocessing, no Python)"
                      % (n, steps, cores)}


def cpu_baseline_block(cfg, r, cpp):
    """The `cpu_baseline` object of the JSON line from a cpu_pool_run() result and a cpp_pool_run() result."""
    raw_per_core = cpp["raw_frames_per_s_per_core"] if cpp else None
    return {"value": r["value"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
            "cpu_sample_envs": r["n_envs"], "cpp_pool": cpp,
            "real_ale_caveat": {
                "oracle_raw_frames_per_s_per_core": raw_per_core, "real_ale_raw_frames_per_s_per_core": REAL_ALE_RAW_FPS_PER_CORE,
                "factor": (REAL_ALE_RAW_FPS_PER_CORE / raw_per_core) if raw_per_core else None,
                "note": "the CPU emulator under this arm is the oracle (a deliberately simple restatement), not ALE: "
                        "ALE is not installable offline.  Real ALE is commonly quoted at ~6 k raw frames/s/core "
                        "(unverified here); divide a GPU/CPU ratio against this arm by `factor` to estimate the ratio "
                        "against an ALE pool on the same cores"}}


# ----------------------------------------------------------------------------- main
# stdout carries exactly ONE line, the JSON: libraries that write to file descriptor 1 (NCCL prints its version
# banner there) are pointed at stderr for the whole run and the line goes to the saved descriptor
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    capture_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=0, help="override environments per GPU")
    ap.add_argument("--envs-per-warp", type=int, default=0)
    ap.add_argument("--decorrelate", type=int, default=24,
                    help="untimed random-policy macro steps before the warm-up so the envs are spread over game states")
    ap.add_argument("--steady-state", type=int, default=0,
                    help="after the timed region: this many more macro steps (episodes end and restart inside them), "
                         "reported as per-step latency percentiles + reset-memo counters")
    ap.add_argument("--random-start", action="store_true", help="random_start pools (the reset memo is off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg = dict(WORKLOADS[args.workload])
    if args.envs:
        cfg["n"] = args.envs
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": args.workload, "games": cfg["games"], "envs_per_gpu": cfg["n"], "rgb": cfg["rgb"],
              "nb_choices": cfg["nb_choices"], "max_repetition": cfg["max_rep"], "policy": "uniform random (counter-based)",
              "decorrelate_steps": args.decorrelate, "observation_history": cfg.get("history", 0),
              "frame_unit": "1 preprocessed frame = 1 next() = 4 emulated frames + one 84x84xD plane",
              "l2": "per-step working set (frame buffers %d MB + states/ring) exceeds the 126 MB L2; no flush needed"
                    % (cfg["n"] * 67200 // (1 << 20)),
              "collective": "none at 1 GPU; at N > 1 every 5 macro steps: flat fp32 gradient all-reduce of the %s net + "
                            "episode-statistics reduction (paac.py:233-256)" % ARCH[args.workload]}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # one reference "step" = REF_MACRO macro steps of the bounded CPU sample (4 envs per host core -- the
        # reference's own pool shape, train.py:102-103 -- NOT the envs_per_gpu of `config`, which names the workload
        # both arms are quoted on; the rate is per core, see cpu_baseline.cpu_sample_envs), so that the default K = 20
        # times ~10 s of CPU work instead of well under a second
        REF_MACRO = 16
        r = cpu_pool_run(cfg, args.steps * REF_MACRO, args.warmup * REF_MACRO)
        r["sample"] += " (= %d bench steps of %d macro steps)" % (args.steps, REF_MACRO)
        cpp = cpp_pool_run(cfg, budget_s=4.0)
        line = {"impl": "reference", "metric": "preprocessed env frames/sec (FiGAR10)", "value": r["value"],
                "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * r["seconds"] / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "cpu_sample_envs": r["n_envs"], "cpu_baseline": cpu_baseline_block(cfg, r, cpp),
                "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit_line(line)
        return 0

    import torch
    import torch.distributed as dist
    import manette_b200 as mb
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    n = cfg["n"]
    tab_rep = tab_repetitions(cfg["max_rep"], cfg["nb_choices"])
    groups = [(g, rom_bytes(g), k) for g, k in split_games(cfg["games"], n)]
    pool = mb.DevicePool(groups, rgb=cfg["rgb"], tab_rep=tab_rep, device=local_rank, env_id_offset=rank * n,
                         envs_per_warp=args.envs_per_warp, history=cfg.get("history", 0), random_start=args.random_start)
    pool.reset_all()
    # the caller's per-step bookkeeping (paac.py:173-205) and n-step returns (paac.py:226-231) ride along: K6 every
    # macro step, K5 every T = max_local_steps = 5 steps
    T_LOCAL = 5
    rollout = mb.Rollout(n, T_LOCAL, pool.num_actions, tab_rep, device=local_rank)
    boot = torch.zeros(n, device=dev)
    extra_launches = [0]
    # the random policy's choices, resident in HBM before the timed region
    gen = torch.Generator(device=dev)
    gen.manual_seed(1000 + rank)
    n_act = torch.cat([torch.full((k,), len(mb.csrc_info.MINIMAL_ACTIONS[g]), dtype=torch.int64) for g, _, k in groups]).to(dev)
    total = args.warmup + args.steps
    acts = (torch.randint(0, 1 << 30, (2 * total, n), device=dev, generator=gen) % n_act).to(torch.int32)
    reps = torch.randint(0, cfg["nb_choices"], (2 * total, n), device=dev, generator=gen, dtype=torch.int32)
    # Synchronous PAAC across GPUs (north star; paac.py:233-256 is the single-process update it extends): every
    # T = max_local_steps macro steps the flat fp32 gradient of the workload's network is all-reduced, and the K6
    # episode statistics (count, sum of returns, sum of lengths: SUM; min / max return: MAX on (-min, max)) are reduced.
    # Both run on the pool's stream inside the timed region; CUDA events bracket them.
    grad, coll_events, n_params = None, [], 0
    if world > 1:
        from manette_b200.networks import PolicyVNetwork
        with torch.device("meta"):
            net = PolicyVNetwork(ARCH[args.workload], pool.num_actions, cfg["nb_choices"], depth=pool.depth)
        n_params = sum(p.numel() for p in net.parameters())
        grad = torch.randn(n_params, device=dev) * 1e-3
        stat_sum = torch.zeros(4, dtype=torch.float64, device=dev)
        stat_max = torch.zeros(2, dtype=torch.float64, device=dev)
    stream = pool.stream

    def collective_step():
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        dist.all_reduce(grad)
        grad.div_(world)
        ps, pm = pack_episode_stats(rollout.stats)           # count, sum reward, sum length, min, max, global_step
        stat_sum.copy_(ps)
        stat_max.copy_(pm)
        dist.all_reduce(stat_sum)
        dist.all_reduce(stat_max, op=dist.ReduceOp.MAX)
        eb.record(stream)
        coll_events.append((ea, eb))

    def device_step(t):
        with torch.cuda.stream(stream):
            pool.action_idx.copy_(acts[t], non_blocking=True)
            pool.repetition_idx.copy_(reps[t], non_blocking=True)
            pool.step_async(use_indices=True, stream=stream)
            if t % T_LOCAL == 0:
                rollout.begin(stream)
            rollout.record(t % T_LOCAL, pool.rewards, pool.terminals, pool.action_idx, pool.repetition_idx, stream)
            extra_launches[0] += 1
            if (t + 1) % T_LOCAL == 0:
                rollout.returns(boot, 0.99, stream)
                extra_launches[0] += 2                      # mask flip + K5
            if grad is not None and (t + 1) % T_LOCAL == 0:
                collective_step()

    # ---- device-resident leg
    for t in range(args.decorrelate):   # spread the envs over game states (they all start identical)
        with torch.cuda.stream(stream):
            pool.action_idx.copy_((torch.randint(0, 1 << 30, (n,), device=dev, generator=gen) % n_act).to(torch.int32))
            pool.repetition_idx.copy_(torch.randint(0, cfg["nb_choices"], (n,), device=dev, generator=gen, dtype=torch.int32))
            pool.step_async(use_indices=True, stream=stream)
        pool.wait()
    for t in range(args.warmup):
        device_step(t)
    if grad is not None:           # the first NCCL calls set up communicators and load kernels: not part of a step
        with torch.cuda.stream(stream):
            collective_step()
            collective_step()
    pool.wait()
    barrier()
    f0, l0, i0 = pool.total_next_calls(), pool.launch_count(), pool.total_instructions()
    extra_launches[0] = 0
    del coll_events[:]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    pool.profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    step_marks = []
    for t in range(args.warmup, total):
        device_step(t)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(stream)
        step_marks.append(ev)
    e1.record(stream)
    pool.wait()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    per_step = np.diff([0.0] + [e0.elapsed_time(ev) for ev in step_marks])
    step_latency = {"mean": float(per_step.mean()), "p50": float(np.percentile(per_step, 50)),
                    "p99": float(np.percentile(per_step, 99)), "max": float(per_step.max()), "unit": "ms", "rank": 0}
    prof = pool.profile_end()
    frames = pool.total_next_calls() - f0
    launches = pool.launch_count() - l0 + extra_launches[0]
    # K6 alone: microseconds per launch (launch-latency bound, like K4 / K5)
    k6a, k6b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        k6a.record(stream)
        for i in range(50):
            rollout.record(i % T_LOCAL, pool.rewards, pool.terminals, pool.action_idx, pool.repetition_idx, stream)
        k6b.record(stream)
    stream.synchronize()
    k6_us = 1000.0 * k6a.elapsed_time(k6b) / 50.0
    ins_timed = pool.total_instructions() - i0
    stat = torch.tensor([ms, float(frames), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_max, frames_all, launches_all = float(mx[0]), float(sm[1]), int(sm[2])
    else:
        ms_max, frames_all, launches_all = ms, float(frames), int(launches)
    value = frames_all / (ms_max / 1000.0)
    collective = None
    if grad is not None:
        us = [1000.0 * a.elapsed_time(b) for a, b in coll_events]
        cstat = torch.tensor([float(np.mean(us)) if us else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(cstat, op=dist.ReduceOp.MAX)
        collective = {"what": "NCCL all-reduce of the flat fp32 gradient (%s net, %d parameters) + episode statistics "
                              "(4 doubles SUM, 2 doubles MAX), on the pool's stream inside the timed region"
                              % (ARCH[args.workload], n_params),
                      "bytes": int(4 * n_params + 48), "every_steps": T_LOCAL, "count": len(us),
                      "us_per_allreduce": float(cstat[0]),
                      "share_of_step": float(cstat[0]) * 1e-3 * len(us) / ms_max if ms_max > 0 else None,
                      "episodes_all_ranks": float(stat_sum[0]), "global_steps_all_ranks": float(stat_sum[3])}
        del coll_events[:]

    # ---- steady state: many more macro steps, so that episodes end and restart inside the steps (reset memo hits,
    # and misses -- 80 emulated frames on the step's critical path -- where the memo cannot help)
    steady = None
    if args.steady_state > 0:
        m0, l10 = pool.memo_stats(), pool.memo_level1_hits()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steady_state + 1)]
        term = torch.zeros((), device=dev)
        with torch.cuda.stream(stream):
            marks[0].record(stream)
            for i in range(args.steady_state):
                pool.action_idx.copy_((torch.randint(0, 1 << 30, (n,), device=dev, generator=gen) % n_act).to(torch.int32))
                pool.repetition_idx.copy_(torch.randint(0, cfg["nb_choices"], (n,), device=dev, generator=gen, dtype=torch.int32))
                pool.step_async(use_indices=True, stream=stream)
                term += pool.terminals.sum()
                marks[i + 1].record(stream)
        pool.wait()
        stream.synchronize()
        lat = np.array([marks[i].elapsed_time(marks[i + 1]) for i in range(args.steady_state)])
        m1 = pool.memo_stats()
        steady = {"steps": args.steady_state, "ms_p50": float(np.percentile(lat, 50)), "ms_p99": float(np.percentile(lat, 99)),
                  "ms_max": float(lat.max()), "p99_over_p50": float(np.percentile(lat, 99) / np.percentile(lat, 50)),
                  "episodes_ended": float(term), "random_start": bool(args.random_start),
                  "resets_restored_from_memo": int(m1[0] - m0[0]), "resets_emulated": int(m1[1] - m0[1]),
                  "of_which_only_the_start_frames": int(pool.memo_level1_hits() - l10)}

    # ---- end-to-end leg through Runners with host arrays
    e2e = None
    if not args.no_e2e:
        emus = mb.emulators_for_pool(pool)
        host_states = pool.states.cpu().numpy()
        variables = [host_states, np.zeros(n, np.float32), np.zeros(n, np.float32),
                     np.zeros((n, pool.num_actions), np.float32), np.zeros((n, pool.nb_choices), np.float32)]
        runners = mb.Runners(tab_rep, mb.EmulatorRunner, emus, 1, variables)
        runners.start()
        sv = runners.get_shared_variables()
        h_acts, h_reps = acts[total:].cpu().numpy(), reps[total:].cpu().numpy()
        ar = np.arange(n)

        def host_step(t):
            sv[3][...] = 0
            sv[3][ar, h_acts[t]] = 1
            sv[4][...] = 0
            sv[4][ar, h_reps[t]] = 1
            runners.update_environments()
            if grad is not None and (t + 1) % T_LOCAL == 0:
                with torch.cuda.stream(stream):
                    collective_step()
            runners.wait_updated()
            return float(sv[1].sum()) + float(sv[2].sum()) + float(sv[0][0, 0, 0, 0])

        for t in range(args.warmup):
            host_step(t)
        barrier()
        f0 = pool.total_next_calls()
        t0 = time.perf_counter()
        for t in range(args.warmup, total):
            host_step(t)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier()
        fr = pool.total_next_calls() - f0
        st2 = torch.tensor([dt, float(fr)], dtype=torch.float64, device=dev)
        if world > 1:
            mx2 = st2.clone(); dist.all_reduce(mx2, op=dist.ReduceOp.MAX)
            sm2 = st2.clone(); dist.all_reduce(sm2, op=dist.ReduceOp.SUM)
            dt_max, fr_all = float(mx2[0]), float(sm2[1])
        else:
            dt_max, fr_all = dt, float(fr)
        e2e = {"value": fr_all / dt_max, "unit": "frames/s",
               "h2d_bytes_per_step": int(world * n * (pool.num_actions + pool.nb_choices) * 4),
               "d2h_bytes_per_step": int(world * n * (84 * 84 * 4 * pool.depth + 8)), "ms_per_step": 1000.0 * dt_max / args.steps,
               "api": "Runners.update_environments()/wait_updated() with pinned host arrays",
               "states_path": "the pool's kernels write each env's new state into the pinned host array as soon as the env has "
                              "finished its repeats (mn_set_host_states): the D2H bytes cross PCIe underneath the remaining rounds"}
        runners.stop()

    # ---- second end-to-end figure: mn_step_host, the single C call INTEGRATION.md shows a maintainer, with PAGEABLE
    # numpy arrays (actions/repetitions up, states/rewards/terminals down, all inside the call)
    e2e_host = None
    if not args.no_e2e and world == 1:
        pa = np.zeros((n, pool.num_actions), np.float32); pr = np.zeros((n, pool.nb_choices), np.float32)
        ps = np.zeros(tuple(pool.states.shape), np.uint8); prw = np.zeros(n, np.float32); pt = np.zeros(n, np.float32)
        k_host = max(3, min(args.steps, 6))
        f0 = pool.total_next_calls()
        t0 = None
        for t in range(1 + k_host):
            if t == 1:
                f0, t0 = pool.total_next_calls(), time.perf_counter()
            pa[...] = 0; pa[ar, h_acts[t]] = 1
            pr[...] = 0; pr[ar, h_reps[t]] = 1
            pool.step_host(pa, pr, ps, prw, pt)
        dt = time.perf_counter() - t0
        e2e_host = {"value": (pool.total_next_calls() - f0) / dt, "unit": "frames/s", "steps": k_host,
                    "ms_per_step": 1000.0 * dt / k_host, "api": "mn_step_host() with pageable numpy arrays"}

    # ---- roofline of the dominant kernel (k_round), CUDA events around every launch on its stream
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    round_ms, round_launches = prof["round"]
    push_ms, push_launches = prof["push"]
    emit_ms, emit_launches = prof["emit"]
    round_s = round_ms / 1000.0
    hbm_achieved = frames * ROUND_BYTES_PER_NEXT / round_s / 1e9 if round_ms > 0 else 0.0
    k3_achieved = (frames * k3_bytes(pool.depth)) / (push_ms / 1000.0) / 1e9 if push_ms > 0 else 0.0
    rc, rc_why = load_counters(ROUND_COUNTERS)
    k3c, k3_why = load_counters(K3_COUNTERS)
    next_per_launch = frames / max(round_launches, 1)
    # K1 is an integer / control-flow kernel: its roofline is the SM issue rate, HBM is secondary.
    # achieved = warp instructions per next() (ncu smsp__inst_executed.sum of profiled launches of THESE sources)
    #            x next() calls of this run / k_round's CUDA-event time of this run
    inst_rate = rc["warp_inst_per_next"] * frames / round_s if rc and round_ms > 0 else None
    roofline = {"kernel": "k_round (6502+TIA emulation, one next() per listed env)", "bound": "sm_issue",
                "achieved": inst_rate, "peak": PEAK_WARP_INST_PER_S, "unit": "warp-inst/s",
                "frac": inst_rate / PEAK_WARP_INST_PER_S if inst_rate else None,
                "traffic": rc["dram_bytes_per_next"] * next_per_launch if rc else None,
                "counters": ({"file": os.path.relpath(ROUND_COUNTERS, ROOT), "source_hash": rc["source_hash"],
                              "warp_inst_per_next": rc["warp_inst_per_next"], "dram_bytes_per_next": rc["dram_bytes_per_next"],
                              "profiled_issue_active_pct": [l["issue_active_pct"] for l in rc["launches"]],
                              "profiled_lanes_per_warp_inst": [l["thread_inst_per_warp_inst"] for l in rc["launches"]]}
                             if rc else {"stale": rc_why}),
                "peak_source": "148 SMs x 4 sub-partitions x 1 warp instruction per clock x 1.965 GHz (SURVEY.md 8(d))",
                "avg_launch_ms": round_ms / max(round_launches, 1), "launches": int(round_launches),
                "next_calls_per_launch": next_per_launch, "share_of_step": round_ms / ms if ms > 0 else None,
                "note": "one warp per SM sub-partition runs a serial emulated-6502 instruction stream: the limiter is the "
                        "latency of that stream (dependent issue, branches), not HBM and not tensor cores (no contraction)",
                "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                        "algorithmic_bytes_per_next": ROUND_BYTES_PER_NEXT, "peak_source": peak_src,
                        "traffic_per_next": rc["dram_bytes_per_next"] if rc else None}}
    k3_moved = k3c["dram_bytes_per_next"] if (k3c and pool.depth == 1) else None
    extra = {"k3_push_frames": {"bound": "hbm", "achieved": k3_achieved, "peak": peak, "unit": "GB/s",
                                "frac": k3_achieved / peak, "ms": push_ms, "launches": int(push_launches),
                                "share_of_step": push_ms / ms if ms > 0 else None,
                                "algorithmic_bytes_per_next": k3_bytes(pool.depth),
                                "traffic_per_next": k3_moved,
                                # the nearest map never reads 126 of the 210 source rows: on the bytes that actually
                                # move the kernel runs at a lower fraction than on the algorithmic ones
                                "achieved_moved": k3_achieved * k3_moved / k3_bytes(pool.depth) if k3_moved else None,
                                "frac_moved": k3_achieved * k3_moved / k3_bytes(pool.depth) / peak if k3_moved else None,
                                "counters": os.path.relpath(K3_COUNTERS, ROOT) if k3c else {"stale": k3_why}},
             "k_emit": {"ms": emit_ms, "launches": int(emit_launches), "share_of_step": emit_ms / ms if ms > 0 else None,
                        "history_depth": int(cfg.get("history", 0))},
             "k6_rollout_record": {"us_per_launch": k6_us, "bound": "launch latency", "envs": n}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_pool_run(cfg, steps=10 ** 9, warmup=1, budget_s=12.0)
        cpu = cpu_baseline_block(cfg, r, cpp_pool_run(cfg, budget_s=4.0))

    if rank == 0:
        line = {"metric": "preprocessed env frames/sec (FiGAR10)", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                "raw_emulated_frames_per_s": 4.0 * value, "macro_steps_per_s": world * n * args.steps / (ms_max / 1000.0),
                "step_latency": step_latency, "clocks": clocks, "e2e": e2e, "e2e_step_host": e2e_host, "collective": collective, "steady_state": steady,
                "gpu_launches": launches_all, "roofline": roofline, "kernels": extra,
                "cpu_baseline": cpu, "envs_per_warp": int(args.envs_per_warp),
                "reset_memo": dict(zip(("restored", "emulated", "stored"), pool.memo_stats())), "exact_reruns": pool.redo_count(),
                "emulated_6502_instr_per_s": world * ins_timed / (ms_max / 1000.0)}
        emit_line(line)
    pool.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
