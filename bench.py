#!/usr/bin/env python
"""bench.py - env-steps/sec of the batched random-legal rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 9x9|19x19]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE ply on every board of the batch: one launch of the fused kernel gg_rollout_step
(auto-reset finished boards -> uniform-random legal action incl. pass -> ply -> write the packed record and
the float32 6xNxN observation).  Workload at N=1 = BASELINE.json configs[1]: 9x9, 65,536 boards
(`--workload 19x19` = configs[2]: 16,384 boards).  Multi-GPU: the same per-GPU batch on every rank (weak
scaling), boards keyed by global index so trajectories do not depend on the sharding; no collective in the
timed loop, one all-gather of (plies, seconds) at the end.

Rank 0 prints ONE JSON line (see the keys at the bottom).  `--impl reference` times the CPU arm instead: the
reference's own algorithm (the numpy/scipy port in oracle/gogame_np.py - the reference is pure Python and
cannot travel to the GPU box) on every host core.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "9x9": dict(size=9, boards=65536, name="9x9 x 65,536 boards, uniform-random-legal rollout (configs[1])"),
    "19x19": dict(size=19, boards=16384, name="19x19 x 16,384 boards, uniform-random-legal rollout (configs[2])"),
}
SEED = 0


def algorithmic_bytes_per_ply(n, obs_bytes_per_elem):
    """read R + 4 (action) + write R, R = 3P+4, P = 4*ceil(N^2/32); + 6*N^2 observation."""
    p = 4 * ((n * n + 31) // 32)
    r = 3 * p + 4
    return 2 * r + 4 + 6 * n * n * obs_bytes_per_elem


def profiled_traffic(workload, obs, plies_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this configuration (or None)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["%s/%s/%d" % (workload, obs, plies_per_launch)]["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------- clocks sampler
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.2] or self.rows
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                for name, val in zip(names, f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU arm
def _cpu_worker(size, boards, warmup, steps, seed, barrier, out_q):
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy as np
    from oracle import gogame_np as og
    rng = np.random.RandomState(seed)
    states = [og.init_state(size) for _ in range(boards)]

    def ply():
        for i in range(boards):
            s = states[i]
            a = int(rng.choice(np.flatnonzero(og.valid_moves(s))))
            s = og.next_state(s, a)
            states[i] = og.init_state(size) if og.game_ended(s) else s

    for _ in range(warmup):
        ply()
    barrier.wait()
    t0 = time.time()
    for _ in range(steps):
        ply()
    out_q.put(time.time() - t0)


def usable_cores():
    """host threads this process may really use: affinity mask capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:  # noqa: BLE001
        pass
    return n


def cpu_reference_run(size, boards_per_core, warmup, steps, cores=None):
    """every host core steps `boards_per_core` boards with the numpy/scipy port; -> (plies/s, cores, seconds)"""
    cores = cores or usable_cores()
    ctx = mp.get_context("fork")
    barrier, q = ctx.Barrier(cores), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(size, boards_per_core, warmup, steps, 100 + i, barrier, q))
             for i in range(cores)]
    for p in procs:
        p.start()
    times = [q.get() for _ in procs]
    for p in procs:
        p.join()
    secs = max(times)
    return cores * boards_per_core * steps / secs, cores, secs


def c_oracle_rate(size, seconds=2.0):
    from oracle import c_oracle as co
    steps = 20000
    t0 = time.time()
    co.rollout(size, steps, 1)
    dt = time.time() - t0
    steps = max(steps, int(steps * seconds / max(dt, 1e-6)))
    t0 = time.time()
    co.rollout(size, steps, 2)
    return steps / (time.time() - t0)


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    size = wl["size"]
    boards_per_core = 16 if size <= 9 else 8
    # bound the whole run to a few minutes: ~4k (9x9) / 2k (19x19) plies/s/core for the numpy port
    rate = 3500.0 if size <= 9 else 1800.0
    budget_s = 150.0
    max_steps = int(budget_s * rate / boards_per_core)
    steps = min(args.steps, max(1, int(max_steps * args.steps / float(args.steps + args.warmup))))
    warmup = min(args.warmup, max(0, max_steps - steps))
    value, cores, secs = cpu_reference_run(size, boards_per_core, warmup, steps)
    sample = "%d cores x %d boards x %d plies (after %d warm-up plies), numpy/scipy port of gogame.next_state, " \
             "uniform-random-legal incl. pass, restart on game end" % (cores, boards_per_core, steps, warmup)
    line = {
        "impl": "reference", "metric": "env-steps/sec (batched random-legal rollout)", "value": value,
        "unit": "env-steps/s", "n_gpus": 0, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * secs / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "board_size": size, "boards": cores * boards_per_core,
                   "note": "CPU arm: a step = one ply on every board of the bounded sample"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from gymgo_b200.engine import GoEngine
    from gymgo_b200.envs import BatchedGoEnv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    size, boards = wl["size"], args.boards or wl["boards"]
    obs_dtype = {"f32": torch.float32, "u8": torch.uint8, "bf16": torch.bfloat16}[args.obs]
    obs_elem = {"f32": 4, "u8": 1, "bf16": 2}[args.obs]
    eng = GoEngine(size, dev)
    board0 = rank * boards
    K, W = args.steps, args.warmup

    # rotating observation buffers: > L2 (126 MB) in total so no step rewrites lines still in cache
    dense_bytes = boards * 6 * size * size * obs_elem
    nbuf = max(2, int(-(-300e6 // dense_bytes)))
    obs_ring = eng.empty((nbuf, boards, 6, size, size), dtype=obs_dtype)
    rec = eng.new_records(boards)
    # The e2e leg replays the first W + E plies of this very rollout from host memory, so their actions are
    # recorded; later plies log into a reusable chunk (reward / done are always logged, like an RL loop needs).
    E = min(K, 300) if args.e2e_steps is None else min(K, args.e2e_steps)
    CHUNK = 256
    actions_replay = torch.empty((W + E, boards), dtype=torch.int32, device=dev)
    actions_chunk = torch.empty((CHUNK, boards), dtype=torch.int32, device=dev)
    reward_chunk = torch.empty((CHUNK, boards), dtype=torch.float32, device=dev)
    done_chunk = torch.empty((CHUNK, boards), dtype=torch.uint8, device=dev)
    ppl = args.plies_per_launch

    def plies(t0, count):
        """gg_rollout (persistent kernel, `ppl` plies per launch, boards in registers) in calls of <= CHUNK plies"""
        launches, t, end = 0, t0, t0 + count
        while t < end:
            n = min(CHUNK, end - t)
            if t < W + E:
                n = min(n, W + E - t)
                alog = actions_replay[t:]
            else:
                alog = actions_chunk
            eng.rollout(rec, SEED, board0, t, n, plies_per_launch=ppl, actions_log=alog, obs_ring=obs_ring,
                        done_log=done_chunk, reward_log=reward_chunk, reward_mode=1, komi=0.0)
            launches += -(-n // ppl)
            t += n
        return launches

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    plies(0, W)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    n_launches = plies(W, K)
    ev1.record()
    barrier()
    wall1 = time.time()
    secs = ev0.elapsed_time(ev1) / 1e3
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    # reference state for the e2e replay check: the rollout after exactly W + E plies (deterministic re-run, untimed)
    final_rec = eng.new_records(boards)
    eng.rollout(final_rec, SEED, board0, 0, W + E, plies_per_launch=ppl)

    # ---------------- transparency: the same plies with ONE launch per ply (no register residency across plies)
    scratch = rec.clone()
    ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n1 = min(K, 100)
    eng.rollout(scratch, SEED, board0, W + K, 8, plies_per_launch=1, obs_ring=obs_ring)
    torch.cuda.synchronize()
    ev4.record()
    eng.rollout(scratch, SEED, board0, W + K + 8, n1, plies_per_launch=1, obs_ring=obs_ring, done_log=done_chunk,
                reward_log=reward_chunk, actions_log=actions_chunk, reward_mode=1, komi=0.0)
    ev5.record()
    torch.cuda.synchronize()
    one_ply_secs = ev4.elapsed_time(ev5) / 1e3

    # ---------------- e2e: the public BatchedGoEnv.step with HOST buffers, copies inside the timed region
    actions_host = torch.empty((W + E, boards), dtype=torch.int32, pin_memory=True)
    actions_host.copy_(actions_replay)
    env = BatchedGoEnv(boards, size, reward_method="real", device=dev, obs_dtype=obs_dtype, board_offset=board0)
    obs_host = torch.empty((boards, 6, size, size), dtype=obs_dtype, pin_memory=True)
    rew_host = torch.empty((boards,), dtype=torch.float32, pin_memory=True)
    done_host = torch.empty((boards,), dtype=torch.uint8, pin_memory=True)
    a_dev = env.action_buffer                                               # the env's static action tensor

    def e2e_ply(t):
        a_dev.copy_(actions_host[t], non_blocking=True)                     # H2D: this step's actions
        o, r, d, _ = env.step(a_dev, auto_reset=True)                       # reset finished boards + one ply
        obs_host.copy_(o, non_blocking=True)                                # D2H: observation, reward, done
        rew_host.copy_(r, non_blocking=True)
        done_host.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()                           # the host now owns the result

    e2e_steps = E
    for t in range(W):
        e2e_ply(t)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for t in range(W, W + e2e_steps):
        e2e_ply(t)
    ev3.record()
    barrier()
    e2e_secs = ev2.elapsed_time(ev3) / 1e3
    if not torch.equal(env.rec, final_rec):
        raise SystemExit("e2e replay diverged from the device rollout - refusing to report")

    # ---------------- informational: the same host-driven loop when the observation stays on the device (the
    # consumer is a device-resident policy network): actions H2D, reward + done D2H, synchronised every step
    env.reset()

    def e2e_ply_light(t):
        a_dev.copy_(actions_host[t], non_blocking=True)
        _, r, d, _ = env.step(a_dev, auto_reset=True)
        rew_host.copy_(r, non_blocking=True)
        done_host.copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    light_steps = min(e2e_steps, W + E)
    torch.cuda.synchronize()
    ev6, ev7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev6.record()
    for t in range(light_steps):
        e2e_ply_light(t)
    ev7.record()
    torch.cuda.synchronize()
    light_secs = ev6.elapsed_time(ev7) / 1e3

    # ---------------- gather (plies, seconds) of every rank: the only collective of the job
    from gymgo_b200 import sharding
    allr = sharding.gather_counters([float(boards) * K, secs, float(boards) * e2e_steps, e2e_secs], device=dev)
    if rank == 0:
        total_plies, t_max = float(allr[:, 0].sum()), float(allr[:, 1].max())
        e2e_plies, e2e_t = float(allr[:, 2].sum()), float(allr[:, 3].max())
        value = total_plies / t_max
        bytes_per_ply = algorithmic_bytes_per_ply(size, obs_elem)
        peak, peak_src = measured_peak_gbs()
        launch_s = float(allr[0, 1]) / n_launches                      # average duration of one kernel launch
        achieved = boards * bytes_per_ply * (K / float(n_launches)) / launch_s / 1e9
        line = {
            "metric": "env-steps/sec (batched random-legal rollout)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitboards" if size <= 9 else "u64 bitboards",
            "data": "synthetic",
            "config": {"workload": wl["name"], "board_size": size, "boards_per_gpu": boards,
                       "global_boards": boards * world, "obs": args.obs, "policy": "uniform over valid actions incl. pass, "
                       "Philox4x32-10 keyed (seed 0, global board, ply), auto-reset",
                       "l2": "each step writes a %.0f MB observation (> L2) into one of %d rotating buffers; the "
                             "%.1f MB packed state is L2-resident by nature" % (dense_bytes / 1e6, nbuf,
                                                                               boards * eng.rec_bytes / 1e6),
                       "plies_per_launch": args.plies_per_launch,
                       "parallelism": "dp%d (independent boards, no data-path collective)" % world},
            "e2e": {"value": e2e_plies / e2e_t, "unit": "env-steps/s",
                    "h2d_bytes_per_step": boards * 4 * world,
                    "d2h_bytes_per_step": (dense_bytes + boards * 5) * world, "steps": e2e_steps,
                    "api": "BatchedGoEnv.step(actions, auto_reset=True): pinned-host actions in; %s observation, reward, "
                           "done out to pinned host, stream-synchronised every step" % args.obs,
                    "obs_kept_on_device": {"value": boards * light_steps * world / light_secs, "unit": "env-steps/s",
                                           "d2h_bytes_per_step": boards * 5 * world,
                                           "note": "rank-0 timing of the same loop when only reward + done go back to "
                                                   "the host (observation consumed on the device); informational"}},
            "gpu_launches": n_launches,
            "one_launch_per_ply": {"value": boards * n1 * world / one_ply_secs, "ms_per_step": 1e3 * one_ply_secs / n1,
                                   "note": "rank-0 timing of the same kernel with plies_per_launch=1 (records reloaded "
                                           "and stored every ply)"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": profiled_traffic(args.workload, args.obs, args.plies_per_launch),
                         "algorithmic_bytes_per_launch": boards * bytes_per_ply * args.plies_per_launch, "kernel": "gg::%s, Geo<%d>, %d plies per launch" % (eng.lib.gg_rollout_kernel(size, boards).decode(), size,
                                                                              args.plies_per_launch),
                         "bytes_per_ply": bytes_per_ply, "peak_source": peak_src,
                         "launch_us": launch_s * 1e6},
        }
        if world == 1 and not args.no_cpu_baseline:
            # the CPU arm runs in a fresh interpreter (no CUDA context / torch thread pools to fork)
            env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
            cpu_steps = 260 if size <= 9 else 220
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload",
                                args.workload, "--steps", str(cpu_steps), "--warmup", "20"],
                               stdout=subprocess.PIPE, text=True, env=env)
            try:
                ref = json.loads(p.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = dict(ref["cpu_baseline"], c_oracle_1core=c_oracle_rate(size))
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10000,
                    help="timed plies (default 10,000 = about a quarter second per GPU, long enough to sample clocks)")
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="9x9", choices=sorted(WORKLOADS))
    ap.add_argument("--boards", type=int, default=None, help="boards per GPU (default: the workload's)")
    ap.add_argument("--obs", default="f32", choices=["f32", "u8", "bf16"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--plies-per-launch", type=int, default=32,
                    help="plies the persistent rollout kernel plays per launch (boards stay in registers)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "ours" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself as one process per GPU (what the driver does)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                   str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", "29537",
                                   os.path.abspath(__file__)] + sys.argv[1:])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
