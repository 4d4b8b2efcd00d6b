#!/usr/bin/env python
"""bench.py - env-steps/sec of the batched random-legal rollout (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 9x9|19x19]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE ply on every board of the batch (auto-reset finished boards -> uniform-random legal action incl.
pass -> ply -> packed record, float32 6xNxN observation, action, reward, done).  Workload at N=1 = BASELINE.json
configs[1]: 9x9, 65,536 boards (`--workload 19x19` = configs[2]: 16,384 boards).

What is timed, whatever K/W the caller passes (SURVEY.md 8d defines the metric at steady state):
  1. untimed set-up: the boards are pre-rolled PREROLL plies so that game phases are de-synchronised (9x9 games
     last ~124 plies, 19x19 ~700) - declared in `config.preroll_plies`;
  2. W warm-up steps with all outputs;
  3. the timed region: R repetitions of exactly K steps, R chosen (from an untimed trial, same R on every rank)
     so that the region lasts >= MIN_TIMED_MS of device time; `ms_per_step` is the mean over R*K steps.
The R*K plies are played by the persistent kernel in launches of plies_per_launch (128) plies - launch boundaries do
not follow K - into an observation ring with one slot per ply of a launch: every observation of a launch stays
readable until the next launch (ring bytes >> L2).  Launches are dynamically scheduled (gg_rollout_with + workspace).

The JSON line also carries `extra`: the other BASELINE configs measured in the same run - 19x19 x 16,384
(configs[2]; per GPU this is configs[4] at --gpus 8) and children() of 4,096 9x9 parents (configs[3]) - and `e2e`:
the same metric through the public host-buffer API (BatchedGoEnv.host_stepper: pinned actions in, pinned f32
observation + reward + done out, copies inside the timed region) with its cheaper variants next to it.

Multi-GPU: the same per-GPU batch on every rank (weak scaling), boards keyed by global index so trajectories do not
depend on the sharding; no collective in the timed loop, all-gathers of counters before and after it.

`--impl reference` times the CPU arm instead: the reference's own algorithm (the numpy/scipy port in
oracle/gogame_np.py - the reference is pure Python and cannot travel to the GPU box; the port is 1.35-1.5x FASTER
than the real reference measured in the build container, so ratios against it are conservative) on every host core.
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
PREROLL = 256            # untimed set-up plies (>= the 200 SURVEY.md 8d asks for)
MIN_TIMED_MS = 250.0     # the timed region is repeated until it lasts at least this long (>= 20 clock samples)


def algorithmic_bytes_per_ply(n, obs_bytes_per_elem):
    """read R + 4 (action) + write R, R = 3P+4, P = 4*ceil(N^2/32); + 6*N^2 observation."""
    p = 4 * ((n * n + 31) // 32)
    r = 3 * p + 4
    return 2 * r + 4 + 6 * n * n * obs_bytes_per_elem


def profiled_traffic(size, obs, plies_in_launch):
    """DRAM bytes of one launch of the dominant kernel, scaled from the committed ncu capture of this configuration
    (profiles/traffic.json: bytes per ply of a 128-ply launch) to the plies this run puts in a launch; None if no capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            row = json.load(f)["%dx%d/%s" % (size, size, obs)]
        return {"dram_bytes_per_launch": int(row["dram_bytes_per_ply"] * plies_in_launch), "source": row["source"]}
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
                                          "--format=csv,noheader,nounits", "-lms", "10"],
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
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.05] or self.rows
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
def _cpu_worker(size, boards, preroll, warmup, steps, repeats, seed, barrier, out_q):
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

    for _ in range(preroll + warmup):          # same untimed set-up as the GPU arm: game phases de-synchronised
        ply()
    barrier.wait()
    t0 = time.time()
    for _ in range(steps * repeats):
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


def cpu_reference_run(size, boards_per_core, warmup, steps, cores=None, preroll=PREROLL, repeats=1):
    """every host core steps `boards_per_core` boards with the numpy/scipy port, after `preroll` untimed set-up plies
    (the GPU arm's steady-state definition); the timed region is `repeats` x `steps` plies, like the GPU arm's;
    -> (plies/s, cores, seconds)"""
    cores = cores or usable_cores()
    ctx = mp.get_context("fork")
    barrier, q = ctx.Barrier(cores), ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(size, boards_per_core, preroll, warmup, steps, repeats, 100 + i, barrier, q))
             for i in range(cores)]
    for p in procs:
        p.start()
    times = [q.get() for _ in procs]
    for p in procs:
        p.join()
    secs = max(times)
    return cores * boards_per_core * steps * repeats / secs, cores, secs


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
    # like the GPU arm, the K timed steps are repeated so that the region is long enough to time (~10 s of CPU work)
    repeats = max(1, int(round(10.0 * rate / (boards_per_core * steps))))
    value, cores, secs = cpu_reference_run(size, boards_per_core, warmup, steps, repeats=repeats)
    sample = "%d cores x %d boards x (%d repetitions x %d plies) after %d untimed set-up plies + %d warm-up plies (the " \
             "GPU arm's steady-state definition), numpy/scipy port of gogame.next_state, uniform-random-legal incl. " \
             "pass, restart on game end" % (cores, boards_per_core, repeats, steps, PREROLL, warmup)
    line = {
        "impl": "reference", "metric": "env-steps/sec (batched random-legal rollout)", "value": value,
        "unit": "env-steps/s", "n_gpus": 0, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * secs / (steps * repeats),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "board_size": size, "boards": cores * boards_per_core,
                   "preroll_plies": PREROLL, "repeats": repeats,
                   "note": "CPU arm: a step = one ply on every board of the bounded sample; kind=port: the reference is "
                           "pure Python with gym/pyglet imports and cannot travel to the GPU box; the port keeps its "
                           "scipy.ndimage arithmetic and ran 1.35-1.5x faster than the unmodified reference in the "
                           "build container (ratios against this arm are conservative)"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------- GPU arm
def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


class RolloutBench(object):
    """one (board size, batch) rollout workload on this rank's GPU: pre-rolled records + output buffers"""

    def __init__(self, eng, boards, board0, obs_dtype, obs_elem, ppl, preroll=PREROLL):
        import torch
        self.eng, self.boards, self.board0, self.ppl = eng, boards, board0, ppl
        self.obs_elem, n = obs_elem, eng.size
        self.ring = eng.empty((ppl, boards, 6, n, n), dtype=obs_dtype)          # one slot per ply of a launch
        self.actions = eng.empty((ppl, boards), dtype=torch.int32)
        self.reward = eng.empty((ppl, boards), dtype=torch.float32)
        self.done = eng.empty((ppl, boards))
        self.rec = eng.new_records(boards)
        self.t = 0
        eng.rollout(self.rec, SEED, board0, 0, preroll, plies_per_launch=32)     # untimed set-up
        self.t = preroll
        self.ring_bytes = self.ring.numel() * obs_elem

    def plies(self, count):
        """`count` plies, all outputs, <= ppl plies per launch; -> launches"""
        done, launches = 0, 0
        while done < count:
            n = min(self.ppl, count - done)
            self.eng.rollout(self.rec, SEED, self.board0, self.t, n, plies_per_launch=n, actions_log=self.actions,
                             obs_ring=self.ring, done_log=self.done, reward_log=self.reward, reward_mode=1, komi=0.0)
            self.t += n
            done += n
            launches += 1
        return launches

    def timed(self, k, repeats):
        """the timed region: repeats * k consecutive plies (launch boundaries every `ppl` plies, not every k)"""
        import torch
        ev0, ev1 = _events()
        ev0.record()
        launches = self.plies(k * repeats)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / 1e3, launches


def pick_repeats(trial_secs_all_ranks):
    return max(1, int(-(-MIN_TIMED_MS / 1e3 // max(trial_secs_all_ranks, 1e-6))))


def roofline_record(eng, boards, obs_elem, obs_name, k, secs, repeats, launches, peak, peak_src, write_ceiling=None):
    size = eng.size
    bytes_per_ply = algorithmic_bytes_per_ply(size, obs_elem)
    plies_in_launch = repeats * k / float(launches)                   # plies one launch really played
    launch_s = secs / launches
    per_launch = boards * bytes_per_ply * plies_in_launch
    achieved = per_launch / launch_s / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (profiled_traffic(size, obs_name, plies_in_launch) or {}).get("dram_bytes_per_launch"),
            "traffic_source": (profiled_traffic(size, obs_name, plies_in_launch) or {}).get("source"),
            "algorithmic_bytes_per_launch": per_launch, "plies_per_launch": plies_in_launch,
            "kernel": "gg::%s, Geo<%d>, dynamically scheduled in 4-ply blocks" % (
                eng.lib.gg_rollout_kernel(size, boards).decode(), size),
            "bytes_per_ply": bytes_per_ply, "peak_source": peak_src, "launch_us": launch_s * 1e6,
            "plain_fill_probe": None if write_ceiling is None else {
                "gbs": write_ceiling, "achieved_over_it": achieved / write_ceiling,
                "note": "gg_probe_write (nothing but 16-byte streaming stores, best of several run lengths) timed in this "
                        "run.  `peak` is the driver's measured COPY bandwidth, not the hardware limit (HBM3e nominal "
                        "~8 TB/s): frac > 1 means this kernel's store stream - every warp writing long contiguous runs - "
                        "sustains more than that copy and more than the plain fill; ncu confirms the bytes reach DRAM "
                        "(`traffic` ~ algorithmic bytes, dram__cycles_active 86 % / 76 % in profiles/r02_final_k_rollout_*)"},
            "timing": "CUDA events on the launching stream around %d launches" % launches}


def write_ceiling_gbs(eng, buf, runs=(512, 4096, 65536, 1 << 20)):
    """pure-write bandwidth of this GPU: gg_probe_write (16-byte streaming stores only) over `buf` (>> L2), the best over
    a few contiguous-run lengths per warp and 4 repetitions each"""
    import torch
    from gymgo_b200 import _cabi
    nbytes = min(buf.numel() * buf.element_size(), 8 << 30) // 16 * 16
    stream = eng._enter()
    best = 0.0
    for run in runs:
        for _ in range(4):
            ev0, ev1 = _events()
            ev0.record()
            _cabi.check(eng.lib.gg_probe_write(buf.data_ptr(), nbytes, run, stream))
            ev1.record()
            torch.cuda.synchronize()
            best = max(best, nbytes / (ev0.elapsed_time(ev1) / 1e3) / 1e9)
    return best


def measure_children(eng, parents_n, board0, repeats=10):
    """configs[3]: children(padded=True) of `parents_n` 9x9 boards after 40 random-legal plies (seed 0);
    -> (seconds per call, bytes written per call)"""
    import torch
    parents = eng.new_records(parents_n)
    eng.rollout(parents, SEED, board0, 0, 40, plies_per_launch=8)
    eng.reset(parents, (eng.flags(parents) & 4) != 0)                  # a finished parent has no children: restart it
    for _ in range(3):
        out = eng.children(parents, obs_dtype=torch.float32, want_rec=True)
    torch.cuda.synchronize()
    ev0, ev1 = _events()
    ev0.record()
    for _ in range(repeats):
        out = eng.children(parents, obs_dtype=torch.float32, want_rec=True)
    ev1.record()
    torch.cuda.synchronize()
    written = sum(out[k].numel() * out[k].element_size() for k in ("rec", "obs", "valid", "status"))
    return ev0.elapsed_time(ev1) / 1e3 / repeats, written, repeats


def run_ours(args, wl, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gymgo_b200 import hostmem, sharding
    from gymgo_b200.engine import GoEngine
    from gymgo_b200.envs import BatchedGoEnv

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bound_cpus = hostmem.bind_thread_near(local_rank)                  # best effort; 0 = topology not visible
    size, boards = wl["size"], args.boards or wl["boards"]
    obs_dtype = {"f32": torch.float32, "u8": torch.uint8, "bf16": torch.bfloat16}[args.obs]
    obs_elem = {"f32": 4, "u8": 1, "bf16": 2}[args.obs]
    eng = GoEngine(size, dev)
    board0 = rank * boards
    K, W = args.steps, args.warmup
    ppl = max(1, args.plies_per_launch)
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def gather(values):
        return sharding.gather_counters(values, device=dev)

    # ---------------- headline: device-resident rollout
    main = RolloutBench(eng, boards, board0, obs_dtype, obs_elem, ppl)
    main.plies(W)                                                      # warm-up (the argument)
    trial, _ = main.timed(K, 1)                                        # untimed trial: sizes the repetition count
    R = pick_repeats(float(gather([trial])[:, 0].max()))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    wall0 = time.time()
    secs, n_launches = main.timed(K, R)                                # THE timed region: R x K steps
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    write_ceiling = write_ceiling_gbs(eng, main.ring) if (rank == 0 and not args.quick) else None

    # ---------------- transparency: the same plies with ONE launch per ply (records reloaded and stored every ply)
    n1 = min(max(K, 20), 100)
    one = RolloutBench(eng, boards, board0, obs_dtype, obs_elem, 1, preroll=0)
    one.rec.copy_(main.rec)
    one.t = main.t
    one.plies(8)
    one_secs, _ = one.timed(n1, 1)
    del one

    # ---------------- e2e: the public host-buffer step (BatchedGoEnv.host_stepper), copies inside the timed region.
    # The host plays the part of the policy by replaying the device rollout's own actions (recorded below), and the run
    # refuses to report if the replayed boards do not end up identical to the device rollout's.
    E = min(max(K, 100), 300) if args.e2e_steps is None else max(1, args.e2e_steps)
    We = min(W, 20)
    start_rec = main.rec.clone()
    t_start = main.t
    replay = torch.empty((We + E, boards), dtype=torch.int32, device=dev)
    final_rec = start_rec.clone()
    eng.rollout(final_rec, SEED, board0, t_start, We + E, plies_per_launch=32, actions_log=replay)
    replay_host = torch.empty((We + E, boards), dtype=torch.int32, pin_memory=True)
    replay_host.copy_(replay)
    torch.cuda.synchronize()
    replay_np = replay_host.numpy()          # the host-side "policy" hands its actions over with a plain memcpy
    del main

    def e2e_leg(returns, dtype, transport="dense"):
        env = BatchedGoEnv(boards, size, reward_method="real", device=dev, obs_dtype=dtype, board_offset=board0)
        env.rec.copy_(start_rec)
        env.done.copy_(((eng.flags(start_rec) >> 2) & 1).to(torch.uint8))
        hs = env.host_stepper(returns=returns, auto_reset=True, follow_current_stream=False, transport=transport)
        act_np = hs.actions.numpy()
        torch.cuda.synchronize()                                         # set-up above ran on the default stream
        for t in range(We):
            np.copyto(act_np, replay_np[t])
            hs.step()
        barrier()
        ev0, ev1 = _events()
        wall = time.time()
        ev0.record()
        for t in range(We, We + E):
            np.copyto(act_np, replay_np[t])                               # the "policy": host memory -> pinned buffer
            hs.step()                                                     # H2D, one kernel, D2H (+ host codec), wait
        ev1.record()
        torch.cuda.synchronize()
        wall = time.time() - wall
        barrier()
        ok = torch.equal(env.rec, final_rec)
        if hs.host_expanded_bytes:                                        # the last host-expanded observation == the boards
            ok = ok and torch.equal(hs.obs, eng.unpack(final_rec, dtype=dtype).cpu())
        return dict(secs=ev0.elapsed_time(ev1) / 1e3, wall=wall, h2d=hs.h2d_bytes, d2h=hs.d2h_bytes, ok=ok,
                    placement=hs.placement, host_bytes=hs.host_expanded_bytes, threads=hs.threads, transport=hs.transport)

    def device_policy_leg():
        """BatchedGoEnv.step driven from Python with the actions already on the device (a device-resident policy):
        no host copies, launches pipeline, one synchronisation at the end"""
        env = BatchedGoEnv(boards, size, reward_method="real", device=dev, obs_dtype=obs_dtype, board_offset=board0)
        env.rec.copy_(start_rec)
        for t in range(We):
            env.step(replay[t], auto_reset=True)
        barrier()
        ev0, ev1 = _events()
        ev0.record()
        for t in range(We, We + E):
            env.step(replay[t], auto_reset=True)
        ev1.record()
        torch.cuda.synchronize()
        barrier()
        return dict(secs=ev0.elapsed_time(ev1) / 1e3, h2d=0, d2h=0, ok=torch.equal(env.rec, final_rec), placement=None)

    legs = {"f32": e2e_leg("obs", obs_dtype, transport="auto")}           # the public default
    if not legs["f32"]["ok"]:
        raise SystemExit("e2e replay diverged from the device rollout - refusing to report")
    if not args.quick:
        legs["f32_dense_over_pcie"] = e2e_leg("obs", obs_dtype, transport="dense")
        legs["u8"] = e2e_leg("obs", torch.uint8, transport="auto")
        legs["packed"] = e2e_leg("packed", obs_dtype)
        legs["obs_kept_on_device"] = e2e_leg("none", obs_dtype)
        legs["device_policy"] = device_policy_leg()
        for name, leg in legs.items():
            if not leg["ok"]:
                raise SystemExit("e2e variant %s diverged from the device rollout - refusing to report" % name)

    # ---------------- the other BASELINE configs, measured in the same run (rank-local, gathered below)
    extra_local = {}
    if not args.quick:
        other = "19x19" if args.workload == "9x9" else "9x9"
        owl = WORKLOADS[other]
        oeng = GoEngine(owl["size"], dev)
        ob = RolloutBench(oeng, owl["boards"], rank * owl["boards"], torch.float32, 4, ppl)
        ob.plies(ppl)
        ot, _ = ob.timed(ppl, 1)
        oR = pick_repeats(float(gather([ot])[:, 0].max()))
        barrier()
        osecs, olaunches = ob.timed(ppl, oR)
        barrier()
        extra_local["rollout"] = (owl, oeng, osecs, oR, olaunches)
        del ob
        ceng = GoEngine(9, dev)
        csecs, cbytes, creps = measure_children(ceng, 4096, rank * 4096)
        extra_local["children"] = (csecs, cbytes, creps)

    # ---------------- gather the counters of every rank (the only collectives of the job; none is timed)
    vals = [float(boards) * K * R, secs, float(boards) * E]
    leg_names = ("f32", "f32_dense_over_pcie", "u8", "packed", "obs_kept_on_device", "device_policy")
    leg_col = {}
    for name in leg_names:
        leg_col[name] = len(vals)
        vals.append(legs[name]["secs"] if name in legs else 0.0)
    xcol = len(vals)                                                   # first column of the extra configs
    if extra_local:
        owl, oeng, osecs, oR, olaunches = extra_local["rollout"]
        vals += [float(owl["boards"]) * ppl * oR, osecs, extra_local["children"][0]]
    allr = gather(vals)
    if rank == 0:
        total_plies, t_max = float(allr[:, 0].sum()), float(allr[:, 1].max())
        value = total_plies / t_max
        e2e_plies = float(allr[:, 2].sum())

        def e2e_value(col):
            return e2e_plies / float(allr[:, col].max())

        def leg_record(name, api):
            leg = legs[name]
            rec = {"value": e2e_value(leg_col[name]), "unit": "env-steps/s", "h2d_bytes_per_step": leg["h2d"] * world,
                   "d2h_bytes_per_step": leg["d2h"] * world,
                   "pcie_gbs_per_gpu": (leg["h2d"] + leg["d2h"]) * E / leg["secs"] / 1e9, "api": api}
            if leg.get("host_bytes"):
                rec["host_expanded_bytes_per_step"] = leg["host_bytes"] * world
                rec["host_codec"] = {"threads_per_rank": leg["threads"], "path": eng.lib.gg_host_unpack_path().decode(),
                                     "dense_gbs_per_gpu_incl_transfer": leg["host_bytes"] * E / leg["secs"] / 1e9}
            return rec

        e2e = leg_record("f32", "BatchedGoEnv.host_stepper(returns='obs').step() [transport='auto' -> '%s']: pinned-host int32 "
                                "actions in; %s [B,6,N,N] observation + f32 reward + u8 done out in host memory, waited for "
                                "every step.  H2D copy, ONE kernel (gg_step with in-kernel auto-reset) and the D2H copies "
                                "are one CUDA graph.  transport 'packed': the packed records cross PCIe and gg_host_unpack "
                                "(AVX-512 mask moves, streaming stores, persistent worker pool) writes the dense tensor on "
                                "the host cores inside step() - same tensor, bit for bit (checked in this run against "
                                "gg_unpack of the final boards); transport 'dense': the kernel writes the dense tensor on "
                                "the device and it crosses PCIe" % (legs["f32"]["transport"], args.obs))
        e2e["transport"] = legs["f32"]["transport"]
        e2e["steps"] = E
        e2e["replay_check"] = "boards after the host-driven replay == device rollout (bit-exact)"
        e2e["host_placement"] = dict(legs["f32"]["placement"], thread_bound_to_cpus=bound_cpus)
        if not args.quick:
            e2e["variants"] = {
                "f32_dense_over_pcie": leg_record("f32_dense_over_pcie", "host_stepper(returns='obs', transport='dense'): the "
                                                  "dense f32 tensor is written on the device and crosses PCIe (round 1's e2e)"),
                "u8_observation": leg_record("u8", "the default call with obs_dtype=uint8 [transport -> '%s']" % legs["u8"]["transport"]),
                "packed_records": leg_record("packed", "host_stepper(returns='packed'): the packed records only, no dense "
                                                       "tensor anywhere"),
                "obs_kept_on_device": leg_record("obs_kept_on_device", "host_stepper(returns='none'): host actions in, "
                                                 "reward + done out, waited for every step; observation stays on the device"),
                "device_policy": leg_record("device_policy", "BatchedGoEnv.step(actions on the device, auto_reset=True) "
                                            "driven from Python: one gg_step launch per ply, no host copies, one "
                                            "synchronisation at the end"),
            }
        line = {
            "metric": "env-steps/sec (batched random-legal rollout)", "value": value, "unit": "env-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_max / (K * R), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitboards" if size <= 9 else "u64 bitboards",
            "data": "synthetic",
            "config": {"workload": wl["name"], "board_size": size, "boards_per_gpu": boards,
                       "global_boards": boards * world, "obs": args.obs,
                       "policy": "uniform over valid actions incl. pass, Philox4x32-10 keyed (seed 0, global board, "
                                 "ply), auto-reset",
                       "preroll_plies": PREROLL, "repeats": R,
                       "timed_region": "%d repetitions x %d steps = %d plies per GPU (>= %.0f ms of device time), after "
                                       "%d untimed set-up plies and %d warm-up steps" % (R, K, R * K, MIN_TIMED_MS,
                                                                                        PREROLL, W),
                       "l2": "every launch writes %d observations of %.0f MB into a %d-slot ring (%.1f GB >> 126 MB "
                             "L2; all observations of a launch stay readable); the %.1f MB packed state is L2-resident "
                             "by nature" % (ppl, boards * 6 * size * size * obs_elem / 1e6, ppl,
                                            ppl * boards * 6 * size * size * obs_elem / 1e9,
                                            boards * eng.rec_bytes / 1e6),
                       "plies_per_launch": ppl,
                       "parallelism": "dp%d (independent boards, no data-path collective)" % world},
            "e2e": e2e,
            "gpu_launches": n_launches,
            "one_launch_per_ply": {"value": boards * n1 * world / one_secs, "ms_per_step": 1e3 * one_secs / n1,
                                   "note": "rank-0 timing of the same kernel with plies_per_launch=1 (records reloaded "
                                           "and stored every ply)"},
            "clocks": clocks,
            "roofline": roofline_record(eng, boards, obs_elem, args.obs, K, float(allr[0, 1]), R, n_launches, peak,
                                        peak_src, write_ceiling),
        }
        if extra_local:
            owl, oeng, osecs, oR, olaunches = extra_local["rollout"]
            o_plies, o_t = float(allr[:, xcol].sum()), float(allr[:, xcol + 1].max())
            c_t = float(allr[:, xcol + 2].max())
            csecs, cbytes, creps = extra_local["children"]
            line["extra"] = {
                "rollout_" + ("19x19" if owl["size"] == 19 else "9x9"): {
                    "workload": owl["name"] + (" per GPU; x%d GPUs = %d boards%s" % (
                        world, owl["boards"] * world, " = configs[4]" if owl["size"] == 19 and world == 8 else "")
                        if world > 1 else ""),
                    "value": o_plies / o_t, "unit": "env-steps/s", "ms_per_step": 1e3 * o_t / (ppl * oR),
                    "steps": ppl * oR, "preroll_plies": PREROLL, "obs": "f32",
                    "roofline": roofline_record(oeng, owl["boards"], 4, "f32", ppl, float(allr[0, xcol + 1]), oR, olaunches,
                                                peak, peak_src, write_ceiling)},
                "children_9x9": {
                    "workload": "gogame.children(padded=True) of 4,096 9x9 parents%s after 40 random-legal plies (seed 0): "
                                "82 child slots per parent, packed records + f32 dense states + valid mask "
                                "(configs[3])" % (" per GPU" if world > 1 else ""),
                    "parent_expansions_per_s": 4096 * world / c_t, "child_states_per_s": 4096 * 82 * world / c_t,
                    "us_per_call": c_t * 1e6, "calls_timed": creps,
                    "roofline": {"bound": "hbm", "achieved": cbytes / csecs / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": cbytes / csecs / 1e9 / peak, "traffic": None,
                                 "algorithmic_bytes_per_launch": cbytes, "kernel": "gg::k_step<Geo<9>, CHILDREN>",
                                 "peak_source": peak_src}},
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
    ap.add_argument("--steps", type=int, default=2000, help="timed steps per repetition (the region is repeated until it "
                                                             "lasts >= %d ms)" % MIN_TIMED_MS)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="9x9", choices=sorted(WORKLOADS))
    ap.add_argument("--boards", type=int, default=None, help="boards per GPU (default: the workload's)")
    ap.add_argument("--obs", default="f32", choices=["f32", "u8", "bf16"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--plies-per-launch", type=int, default=128,
                    help="plies the persistent rollout kernel plays per launch = slots of the observation ring (128 = one "
                         "PPO-style rollout segment per launch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline + f32 e2e only (no extra configs, no e2e variants, no plain-fill probe)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.steps < 1:
        args.steps = 1
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
