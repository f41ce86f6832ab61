#!/usr/bin/env python
"""bench.py — creature-steps/sec of the batched REM2D evaluation path (BASELINE.json metric).

A "step" is one pass of the hot path over one population: every creature is reset and simulated for
its whole episode (wall-of-death termination on, REM2D_main.py:350-378), i.e. one call of what
``toolbox.map(toolbox.evaluate, population)`` does in the reference. creature-steps = sum of ticks
actually simulated.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pop P]

N > 1 is launched by torchrun (one rank per GPU); every rank evaluates its own shard of P creatures
(weak scaling: the population shards by individual with no data-path collective) and only the
fitness vector is gathered over NCCL.
``--impl reference`` times the CPU path on the host cores: pybox2d is not installable here, so this
is the oracle port (oracle/rem2d_oracle.c, "kind": "port") on a bounded sample of the same workload.
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

from gym_rem2d_b200 import constants as K  # noqa: E402
from gym_rem2d_b200 import terrain  # noqa: E402
from gym_rem2d_b200.population import random_population  # noqa: E402

METRIC = "creature-steps/sec"
WORKLOAD = "pop %d %s creatures (1-21 modules), rough BipedalWalker-style terrain (env.seed(4)), " \
           "full episodes with wall-of-death termination, dt 1/50, 180 velocity + 60 position iterations"
ENC_NAMES = {"lsystem": "L-system", "direct": "direct-encoding", "ce": "cellular-encoding", "cppn": "CPPN"}


def flops_from_counters(c):
    """Algorithmic FLOPs: 51/joint velocity solve, 72/1-point and 155/2-point manifold solve,
    70/joint and 65/contact-point position solve, 30/body integration, 160/narrow phase, 400/TOI query."""
    return (51 * c["joint_vsolves"] + 72 * c["p1_vsolves"] + 155 * c["m2_vsolves"] + 70 * c["joint_psolves"]
            + 65 * c["point_psolves"] + 30 * c["body_ticks"] + 160 * c["narrow"] + 400 * c["toi_calls"])


def bytes_from_counters(c, pop_bodies_per_tick):
    """Algorithmic bytes: 48 B per body + 40 B per joint + 24 B per contact point, read + written
    once per creature-tick, as if only the dynamic state round-tripped HBM every tick."""
    joints_ticks = c["joint_vsolves"] / 180.0
    points_ticks = (c["p1_vsolves"] + 2 * c["m2_vsolves"]) / 180.0
    return 48.0 * c["body_ticks"] + 40.0 * joints_ticks + 24.0 * points_ticks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(self.samples)}


def cpu_baseline(pop, ys, sample_creatures, threads):
    """Oracle port on the host cores over a bounded sample of the same population."""
    from oracle.oracle import OracleEngine
    sub = pop.select(np.arange(min(sample_creatures, pop.n_creatures)))
    e = OracleEngine(threads=threads)
    e.set_terrain(ys, K.TERRAIN_STEP)
    t0 = time.perf_counter()
    fit, ticks = e.evaluate(sub, K.EVALUATION_STEPS)
    dt = time.perf_counter() - t0
    return float(ticks.sum()) / dt, dt, sub.n_creatures, int(ticks.sum()), fit, ticks


def ea_bench(args, rank, local_rank, world, cores):
    """BASELINE config 5: the EA generation loop at population --pop. Rank 0 runs selection / variation / expansion (persistent
    worker pool) and drives the evaluation; with N > 1 ranks the generation's table is broadcast and every rank evaluates its
    shard (distributed.evaluate_broadcast). Prints per-generation expand / evaluate seconds and creature-steps/s."""
    import random
    from gym_rem2d_b200 import ea
    cfg = ea.default_config(enc=args.encoding.split(",")[0], mr=0.01, mmr=0.01, ms=0.1)      # 0.cfg-like rates
    cfg["ea"]["batch_size"] = str(args.pop)
    run = None
    if rank == 0:
        random.seed(args.seed)
        np.random.seed(args.seed)
        run = ea.run2D(cfg, "", workers=max(2, cores - 2), distributed=world > 1)     # pool first: before CUDA exists here
        run.materialize_result = False           # the final population stays packed (nothing reads it here)
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    if rank != 0:
        from gym_rem2d_b200 import distributed as rdist
        from gym_rem2d_b200.env import BatchedModular2D
        env = BatchedModular2D(device=local_rank)
        env.seed(K.TERRAIN_SEED)
        rdist.serve_evaluations(env.engine, K.EVALUATION_STEPS, gather_ticks=True)
        dist.destroy_process_group()
        return
    t0 = time.perf_counter()
    run.run_deap(cfg, n_generations=args.generations)
    total = time.perf_counter() - t0
    if world > 1:
        from gym_rem2d_b200 import distributed as rdist
        rdist.broadcast_table(None, 0)              # stop signal for the serving ranks
    gens = run.generation_log
    steps = sum(g["creature_steps"] for g in gens)
    secs = sum(g["seconds"] for g in gens)
    print(json.dumps({"metric": "creature-steps/sec (EA loop, config 5)", "value": steps / secs,
                      "unit": "creature-steps/s", "n_gpus": world, "population": args.pop, "generations": len(gens), "host_workers": max(2, cores - 2),
                      "seconds_per_generation": secs / max(1, len(gens)),
                      "per_generation": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in g.items()} for g in gens],
                      "note": "expand_s = the part of selection + clone + mutate + genome.create + flatten (worker pool, host Python) that is "
                              "not hidden behind an evaluation; evaluate_s = GPU evaluation incl. upload and read-back (one per generation, or "
                              "two halves with the second half expanding meanwhile: single device, REM2D_EA_PIPELINE != 0)",
                      "initial_population_s": round(total - secs, 2)}))
    run.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pop", type=int, default=65536, help="creatures per GPU")
    ap.add_argument("--encoding", default="lsystem")
    ap.add_argument("--seed", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ea", action="store_true", help="config 5: generations of the batched EA loop (REM2D_main.py:280-348) at --pop")
    ap.add_argument("--generations", type=int, default=2)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    cache = os.environ.get("REM2D_CACHE", "/tmp/rem2d_cache")
    encodings = tuple(args.encoding.split(","))
    config = {"workload": WORKLOAD % (args.pop, "/".join(ENC_NAMES.get(e, e) for e in encodings)), "population_per_gpu": args.pop, "population_total": args.pop * max(world, 1),
              "encoding": args.encoding, "terrain": "rough seed 4", "episode": "full (WOD on, <= 10000 ticks)",
              "l2_policy": "state blocks (>= 0.5 GB per 65536 creatures) exceed the 126 MB L2; no explicit flush",
              "shard": "by individual, one shard per rank, fitness all_gather over NCCL"}

    # ------------------------------------------------------------------ reference arm: CPU port
    if args.impl == "reference":
        if rank != 0:
            return
        xs, ys = terrain.generate_terrain()
        sample = min(args.pop, max(256, 1800 * cores))      # ~10 s of CPU work per step on the box's cores
        pop = random_population(sample, encodings, seed=args.seed, workers=max(1, cores // 2), cache_dir=cache)
        vals = []
        for i in range(args.warmup + args.steps):
            v, dt, n, cs, _, _ = cpu_baseline(pop, ys, sample, cores)
            if i >= args.warmup:
                vals.append((v, dt))
        value = float(np.mean([v for v, _ in vals])) if vals else 0.0
        ms = float(np.mean([d for _, d in vals])) * 1e3 if vals else 0.0
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": value, "unit": "creature-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "creature-steps/s", "cores": cores, "kind": "port",
                             "sample": "first %d creatures of the seeded population, whole episodes; pybox2d (Box2D==2.3.10) "
                                       "is not installable, this is the float32 C restatement oracle/rem2d_oracle.c" % sample},
            "e2e": {"value": value, "unit": "creature-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ config 5: the EA generation loop
    if args.ea:
        return ea_bench(args, rank, local_rank, world, cores)

    # ------------------------------------------------------------------ our arm
    # host-side expansion first (process pool; must happen before CUDA is initialised in this process)
    pop = random_population(args.pop, encodings, seed=args.seed + 7919 * rank,
                            workers=max(1, cores // max(world, 1)), cache_dir=cache)
    xs, ys = terrain.generate_terrain()

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from gym_rem2d_b200.capi import Engine
    stream = torch.cuda.current_stream()
    eng = Engine(device=local_rank, stream=stream.cuda_stream)
    eng.set_terrain(ys, K.TERRAIN_STEP)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -- device-resident measurement ("value"): population table already in HBM, timed region = rem2d_run_episodes
    eng.upload(pop)
    fit_all = None
    for _ in range(args.warmup):
        eng.run_episodes(K.EVALUATION_STEPS)
    barrier()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = 0.0
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            eng.run_episodes(K.EVALUATION_STEPS)       # world build + whole episodes, all on the device
            kernel_ms += eng.last_step_ms()
        ev1.record(stream)
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    counters = eng.counters()          # of the last evaluation
    fit = eng.fitness()
    ticks_gpu = eng.ticks()
    creature_steps = counters["ticks"]
    t = torch.tensor([ms_total, float(creature_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, total_steps = float(tmax[0]), float(tsum[1])
        # the only collective of the path: gather the fitness vector
        ft = torch.from_numpy(fit.astype(np.float32)).to(dev)
        out = torch.empty(world * ft.numel(), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(out, ft)
        fit_all = out.cpu().numpy()
    else:
        total_steps = float(creature_steps)
        fit_all = fit
    ms_per_step = ms_total / args.steps
    value = total_steps / (ms_per_step * 1e-3)

    # -- end to end through the public call with HOST buffers: upload (H2D) + reset + step + fitness/ticks (D2H)
    h2d = sum(getattr(pop, k).nbytes for k in ("body_off", "shape", "hx", "hy", "x0", "y0", "a0", "joint_parent", "anchor_a",
                                                "anchor_b", "lower", "upper", "max_torque", "ctrl")) + (pop.n_bodies - pop.n_creatures)
    d2h = pop.n_creatures * (8 + 4 + 4 + 4)
    eng.evaluate(pop, K.EVALUATION_STEPS)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = 0
    for _ in range(args.steps):
        f2, tk = eng.evaluate(pop, K.EVALUATION_STEPS)
        e2e_steps += int(tk.sum())
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, float(e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        a = te.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = te.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        e2e_s, e2e_steps = float(a[0]), float(b[1])
    e2e_value = e2e_steps / e2e_s

    # -- strong scaling (N > 1): ONE population of args.pop creatures sharded over the ranks through the product API
    # (distributed.shard_population -> rem2d_evaluate -> distributed.gather_fitness over NCCL), everything inside the timed
    # region: what replaces pool.map(evaluate, population, chunksize=...) of REM2D_main.py:256-262 at BASELINE's stated size.
    strong = None
    if world > 1:
        from gym_rem2d_b200 import distributed as rdist
        whole = random_population(args.pop, encodings, seed=args.seed, workers=1, cache_dir=cache)    # rank 0's population (cached)
        rdist.evaluate_sharded(whole, eng, K.EVALUATION_STEPS, device=dev)                               # warm-up
        barrier()
        t0 = time.perf_counter()
        s_steps = 0
        for _ in range(args.steps):
            fit_strong, st_ = rdist.evaluate_sharded(whole, eng, K.EVALUATION_STEPS, device=dev)
            s_steps += st_
        barrier()
        s_s = time.perf_counter() - t0
        ts = torch.tensor([s_s, float(s_steps)], dtype=torch.float64, device=dev)
        a = ts.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX)
        b = ts.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM)
        strong = {"value": float(b[1]) / float(a[0]), "unit": "creature-steps/s", "ms_per_step": float(a[0]) / args.steps * 1e3,
                  "population_total": args.pop, "population_per_gpu": args.pop // world, "scaling": "strong",
                  "path": "distributed.shard_population -> rem2d_evaluate (host buffers) -> distributed.gather_fitness (NCCL all_gather) "
                          "inside the timed region",
                  # rank 0's weak-scaling population IS this population: the gathered vector must equal its single-GPU result
                  "identical_to_single_gpu": bool(np.array_equal(fit_strong, fit.astype(np.float32))) if rank == 0 else None}

    if rank == 0:
        # roofline of the dominant kernel (step_kernel): algorithmic work of one evaluation / its device time
        step_ms = kernel_ms / args.steps
        flops = flops_from_counters(counters)
        abytes = bytes_from_counters(counters, None)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        fp32_peak = eng.measure_fp32_peak()
        # The step is a chain of dependent fp32 operations on shared-memory state: neither HBM- nor tensor-bound. The
        # binding resource is FP32 ISSUE: algorithmic FLOPs (8d formula fed by the kernels' work counters) over kernel time against
        # the non-fused FMUL/FADD issue peak measured on this device (the kernels are built with -fmad=false for bit parity).
        # The HBM view the contract defines (algorithmic bytes per tick / kernel time against MEASURED_PEAKS.json) is kept under "hbm".
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json"))) if os.path.exists(os.path.join(ROOT, "profiles", "r2_traffic.json")) else {}
        roofline = {"bound": "fp32_issue", "achieved": flops / (step_ms * 1e-3) / 1e12, "peak": fp32_peak / 1e3, "unit": "TFLOP/s",
                    "frac": flops / (step_ms * 1e-3) / 1e9 / fp32_peak if fp32_peak else None,
                    "peak_source": "rem2d_measure_fp32_peak (non-fused FMUL/FADD issue microbenchmark, same process)",
                    "flops_per_creature_step": flops / max(1, creature_steps), "kernel_ms": step_ms,
                    # ncu dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch (queue mode, class NB=22 of this workload)
                    "traffic": traffic.get("dram_bytes") if args.pop == 65536 and args.encoding == "lsystem" else None,
                    "traffic_scope": traffic.get("scope"),
                    "hbm": {"bound": "hbm", "achieved": abytes / (step_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                            "frac": abytes / (step_ms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                            "algorithmic_bytes": "48 B/body + 40 B/joint + 24 B/contact point per creature-tick (SURVEY 8d)"}}
        fp32 = dict(roofline)      # kept under its round-1 name for round-over-round comparison
        fp32.pop("hbm")
        line = {"metric": METRIC, "value": value, "unit": "creature-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "clocks": clocks.summary(), "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "creature-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "roofline": roofline, "fp32_issue": fp32,
                "strong_scaling": strong, "counters": counters, "fitness": {"mean": float(np.mean(fit_all)), "max": float(np.max(fit_all)), "n": int(len(fit_all))},
                "mean_ticks_per_creature": creature_steps / pop.n_creatures}
        if not args.no_cpu_baseline and world == 1:
            sample = min(args.pop, max(256, 1800 * cores))      # ~10 s of CPU work
            v, dt, n, cs, fit_cpu, ticks_cpu = cpu_baseline(pop, ys, sample, cores)
            # parity of the benchmarked configuration: the oracle's per-creature fitness and lifetime of the sampled
            # creatures against what the timed GPU run (device-resident leg) and the e2e leg returned, bit for bit
            bad = int(np.count_nonzero((fit[:n] != fit_cpu) | (ticks_gpu[:n] != ticks_cpu)))
            bad_e2e = int(np.count_nonzero((f2[:n] != fit_cpu) | (tk[:n] != ticks_cpu)))
            line["parity"] = {"n": int(n), "mismatches": bad, "mismatches_e2e": bad_e2e,
                              "checked": "fitness (float64) and ticks of the first n creatures, exact equality vs oracle"}
            line["cpu_baseline"] = {"value": v, "unit": "creature-steps/s", "cores": cores, "kind": "port",
                                    "sample": "first %d creatures of the same population, whole episodes, %.1f s; pybox2d is not "
                                              "installable: float32 C restatement oracle/rem2d_oracle.c" % (n, dt)}
        print(json.dumps(line))
        if line.get("parity", {}).get("mismatches", 0) or line.get("parity", {}).get("mismatches_e2e", 0):
            sys.stderr.write("bench: GPU results differ from the oracle on the benchmarked configuration\n")
            sys.exit(3)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
