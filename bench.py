#!/usr/bin/env python
"""Benchmark of the batched TTV hot path (BASELINE.json metric: FP64 N-body system-steps/s and
TTV sims/s at 1/2/4/8 B200 beside the host CPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (config C2 of BASELINE.json): random 2-planet systems from the priors of
randomize() (nbody/simulation.py:54-129) with the omega sweep of
simulations/generate_simulations.py:83-107 — 10 000 draws x (1 + 6 omega variants) = 70 000
simulations per GPU per step; 365 d sampled hourly (8760 outputs); products per planet: transit
times, linear ephemeris, O-C, maxavg, O-C periodogram, means of a/e/inc/omega.  A step is one
pass of the whole path over that batch.  Weak scaling: every rank owns its own 70 000 systems.

`value` is whole-job TTV sims/s with inputs resident in HBM (torch device tensors, kernels enqueued
through the C ABI with NBT_F_DEVICE_PTRS); `e2e` is the same through nbt_run with HOST buffers
(H2D of the parameter block and plan, D2H of every product inside the timed call);
`system_steps_per_s` counts kick+drift stages.  The roofline is the FP64 vector pipe, measured
live with a DFMA-chain kernel (MEASURED_PEAKS.json has no FP64 figure).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DAYS, N_OUTPUTS = 365.0, 8760
WORKLOAD = ("C2: random 2-planet systems (randomize() priors, generate_simulations.py omega sweep), "
            "365 d @ 1 h, tt + ephemeris + O-C + maxavg + O-C periodogram + element means")


def f_step(np_, stages=3, forces=4, grad_kicks=1):
    """Algorithmic FLOPs per system-step = one Kepler-drift stage of one system (FMA = 2,
    div/sqrt = 1): drift 230 Np + kick 6 Np + the force evaluation (Jacobi<->inertial 21 Np, pairs
    12.5 Np (Np+1) - 25, indirect term 15 (Np-1)).  SURVEY's figure has one force evaluation per stage
    (272 Np + 12.5 Np (Np+1) - 40); the default SBABC3 scheme has `forces` = 4 evaluations and one
    displaced-position set-up (6 Np + 7 (Np-1)) per `stages` = 3 drift stages."""
    force = 21 * np_ + 12.5 * np_ * (np_ + 1) - 25 + 15 * (np_ - 1)
    return 236 * np_ + (forces * force + grad_kicks * (6 * np_ + 7 * (np_ - 1))) / stages


F_OUT = 100  # per planet per output sample (Jacobi a, e, inc + crossing test)
F_OMEGA = 75  # + argument of pericentre (eccentricity vector 20, node 7, acos 35, quadrant and sum 13), NBT_F_MEAN_OMEGA

# (drift stages, force evaluations, displaced-position set-ups) per composition step of each NBT_SCHEME_*
SCHEME_COST = {0: (1, 1, 0), 1: (3, 3, 0), 2: (4, 4, 0), 3: (4, 4, 0), 4: (2, 3, 1), 5: (3, 4, 1)}


def plan_flops(plan):
    """Algorithmic FLOPs of the integration stages of one launch of `plan` (sampling not included)."""
    per, np_, sch = plan.cfg.n_outputs - 1, plan.Np, plan.cfg.scheme
    if plan.substeps is None:
        k_main, k_alt = plan.cfg.substeps * plan.B, 0
    else:
        k_main, k_alt = int(plan.substeps[plan.substeps > 0].sum()), -int(plan.substeps[plan.substeps < 0].sum())
    s, f, g = SCHEME_COST[5 if sch == 6 else sch]
    fl = k_main * per * s * f_step(np_, s, f, g)
    s, f, g = SCHEME_COST[4]   # NBT_SCHEME_AUTO: negative counts are C4 steps
    return fl + k_alt * per * s * f_step(np_, s, f, g)


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.rows, self.stop_flag, self.gpu = [], False, gpu_index
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(np_, b):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_integrate launch on this workload, from the
    committed `ncu --set full` capture (profiles/r1_traffic.json); None if no capture matches."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if t.get("n_planets") == np_ and t.get("systems") == b:
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def hbm_side(traffic_bytes, launch_ms):
    """The HBM view of the same launch (it is not the binding roof: the kernel is register-resident)."""
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "MEASURED_PEAKS.json"
    except Exception:
        peak, src = 6650.0, "fallback (B200_PROFILING.md)"
    if traffic_bytes is None:
        return None
    ach = traffic_bytes / (launch_ms * 1e-3) * 1e-9
    return {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": src}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def cpu_path(params, n_threads=0):
    """The reference algorithm on host cores: oracle (IAS15 restatement + analyze semantics)."""
    import oracle
    from nbody_ai_b200 import _abi, engine
    plan = engine.Plan(params, N_DAYS, N_OUTPUTS, substeps=1, flags=_abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC)
    buf = engine.Buffers(plan)
    nt = n_threads or oracle.max_threads()
    t0 = time.perf_counter()
    ias_steps = oracle.run(plan.cfg, plan.params, buf, n_threads=nt)
    return time.perf_counter() - t0, nt, ias_steps


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  rebound/astropy are not
    installable here, so this is the oracle port (cpu_baseline.kind = "port"), all host threads."""
    if rank != 0:
        return
    from nbody_ai_b200 import workloads
    per_step = args.ref_sample
    need = per_step * (args.steps + args.warmup)
    params = workloads.c2_params(max(1, (need + 6) // 7), 6, seed=0)[:need]
    times, nt = [], 0
    for s in range(args.steps + args.warmup):
        dt, nt, _ = cpu_path(params[s * per_step:(s + 1) * per_step])
        if s >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    sims = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": "ttv_sims_per_sec", "value": sims, "unit": "sims/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_systems_per_step": per_step, "n_days": N_DAYS, "n_outputs": N_OUTPUTS},
        "system_steps_per_s": sims * (N_OUTPUTS - 1),
        "cpu_baseline": {"value": sims, "unit": "sims/s", "cores": nt, "kind": "port", "cpu_model": cpu_model(),
                         "host_cpus": os.cpu_count(),
                         "sample": "%d systems of the C2 batch per step, oracle (IAS15 + analyze) on %d host threads" % (per_step, nt)},
        "e2e": {"value": sims, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def bind_to_gpu_numa_node(gpu_index):
    """One process per GPU: run on the CPUs next to this rank's GPU, so that the pinned staging buffers
    (first touch) and the 340 MB of D2H per step stay on the GPU's own socket.  Returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(gpu_index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    from nbody_ai_b200 import _abi, engine, workloads
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _abi.load()
    dev = torch.device("cuda", local_rank)

    if world > 1:
        # one global batch (a C2 draw per rank's worth of seeds), sharded by system with equal cost per rank:
        # systems sorted by the planner's stage count are dealt to the ranks in snake order
        everyone = np.concatenate([workloads.c2_params(args.draws, args.omegas, seed=r) for r in range(world)])
        mine = engine.shard_balanced(engine.system_cost(everyone, N_DAYS, N_OUTPUTS), rank, world)
        params = everyone[mine]
        del everyone
    else:
        params = workloads.c2_params(args.draws, args.omegas, seed=0)
    B, Np = params.shape[0], params.shape[1] - 1
    flags = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
    plan = engine.Plan(params, N_DAYS, N_OUTPUTS, flags=flags, device=local_rank)
    plan.pin()
    dbuf = engine.DeviceBuffers(plan)
    hbuf = engine.Buffers(plan, pinned=True)

    # FP64 roofline denominator, measured on this GPU right now
    tf, ms = C.c_double(0), C.c_double(0)
    _abi.check(lib.nbt_fp64_peak(local_rank, 1 << 16, C.byref(tf), C.byref(ms)), lib)
    fp64_peak = tf.value
    _abi.check(lib.nbt_fp64_peak(local_rank, -(1 << 16), C.byref(tf), C.byref(ms)), lib)
    fp64_peak_3reg = tf.value   # DFMA whose three operands are distinct registers (register-bandwidth bound)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        engine.run_device(plan, dbuf)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)

    # ---- timed region: K steps, device-resident ------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_ms = np.zeros(4)
    barrier()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                       # L2 flush, outside the event pair
        ev[k][0].record()
        engine.run_device(plan, dbuf, timing=True)
        ev[k][1].record()
        ev[k][1].synchronize()
        stage_ms += engine.stage_timing()
    barrier()
    dev_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    clocks = sampler.finish()

    # ---- end to end through the host-pointer C ABI ---------------------------------------
    for _ in range(min(args.warmup, 2)):
        engine.run(plan, hbuf)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        engine.run(plan, hbuf)                      # H2D + kernels + D2H + stream sync inside
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    h2d = sum(a.nbytes for a in (plan.params, plan.substeps, plan.order, plan.tt_offsets, plan.ls_offsets) if a is not None)
    d2h = sum(getattr(hbuf, n).nbytes for n in ("tt", "ttv", "n_tt", "period", "t0", "ttv_max", "mean_a", "mean_e",
                                               "mean_inc", "mean_omega", "status", "ls_oc_power", "ls_oc_nfreq"))

    # sanity: the timed device path and the host path produced the same transits
    chk = dbuf.to_host()
    assert np.array_equal(chk.n_tt, hbuf.n_tt), "device-resident and host-pointer paths disagree"
    n_tt_total = int(np.minimum(hbuf.n_tt, np.diff(plan.tt_offsets)).sum())

    t = torch.tensor([dev_ms, e2e_ms, stage_ms[0]], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(B), float(plan.n_steps), float(n_tt_total)], dtype=torch.float64, device=dev)
    # per-rank view (diagnosis of the max-over-ranks figure): step time, e2e time, SM clock, power-cap flag
    mine = torch.tensor([dev_ms / args.steps, e2e_ms / args.steps, clocks["sm_mhz"] or 0.0,
                         1.0 if "sw_power_cap" in clocks["reasons"] else 0.0], dtype=torch.float64, device=dev)
    per_rank = [mine]
    if dist is not None:
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    per_rank = [[round(float(x), 3) for x in r.cpu()] for r in per_rank]
    dev_ms, e2e_ms, integ_ms = (float(x) for x in t.cpu())
    sims_all, steps_all, ntt_all = (float(x) for x in tot.cpu())

    if rank == 0:
        sims_per_s = sims_all * args.steps / (dev_ms * 1e-3)
        steps_per_s = steps_all * args.steps / (dev_ms * 1e-3)
        # roofline of the dominant kernel (k_integrate), per launch on this rank
        flops = plan_flops(plan) + B * N_OUTPUTS * (F_OUT + F_OMEGA) * Np   # the bench sets NBT_F_MEAN_OMEGA
        integ_launch_ms = stage_ms[0] / args.steps
        achieved = flops / (integ_launch_ms * 1e-3) * 1e-12
        line = {
            "metric": "ttv_sims_per_sec", "value": sims_per_s, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "systems_per_gpu": B, "draws": args.draws, "omega_variants": args.omegas,
                       "n_planets": Np, "n_days": N_DAYS, "n_outputs": N_OUTPUTS, "scheme": "auto: sbabc3 (3-stage order (6,4) Wisdom-Holman composition with force-gradient end kicks); c4 (2-stage force-gradient) where P_min >= 250 output steps",
                       "substeps_hist": np.bincount(np.abs(plan.substeps)).tolist(), "c4_systems": int((plan.substeps < 0).sum()), "parallelism": "shard-by-system x%d%s" % (world, ", cost-balanced (engine.shard_balanced)" if world > 1 else ""), "cpu_affinity": affinity,
                       "l2": "256 MiB device fill between steps, outside the per-step event pairs"},
            "system_steps_per_s": steps_per_s,
            "system_steps_per_step": steps_all,
            "transits_per_step": ntt_all,
            "stage_ms_per_step": {"k_integrate": stage_ms[0] / args.steps, "k_fit": stage_ms[1] / args.steps,
                                  "k_ls_oc": stage_ms[2] / args.steps},
            "roofline": {"bound": "fp64", "kernel": "k_integrate<%d>" % Np, "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak > 0 else None,
                         "peak_source": "nbt_fp64_peak DFMA-chain kernel measured in this run (nominal 37.2 at 1965 MHz)",
                         "peak_three_register_operands": fp64_peak_3reg,
                         "flops_per_launch": flops, "launch_ms": integ_launch_ms, "traffic": ncu_traffic(Np, B),
                         "hbm": hbm_side(ncu_traffic(Np, B), integ_launch_ms)},
            "e2e": {"value": sims_all * args.steps / (e2e_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": 3 * args.steps,   # k_integrate, k_fit, k_ls_oc
            "clocks": clocks,
            "ranks": {"columns": ["ms_per_step", "e2e_ms_per_step", "sm_mhz", "sw_power_cap"], "rows": per_rank},
        }
        if world == 1 and not args.no_cpu:
            n_s = min(args.cpu_sample, B)
            dt, nt, _ = cpu_path(params[:n_s])
            line["cpu_baseline"] = {"value": n_s / dt, "unit": "sims/s", "cores": nt, "kind": "port", "cpu_model": cpu_model(),
                                    "host_cpus": os.cpu_count(),
                                    "sample": "first %d systems of the same batch, oracle (IAS15 + analyze) on %d host threads, %.1f s" % (n_s, nt, dt)}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


class StdoutToStderr:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr, so that the
    ONE JSON line is the only thing on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


OUT = None


def emit(line):
    text = json.dumps(line)
    if OUT is not None:
        OUT.emit(text)
    else:
        print(text, flush=True)


def ensure_built():
    """The in-tree libraries normally travel with the repo snapshot; compile them if they did not
    (local rank 0 builds, the other ranks of a torchrun job wait for the file)."""
    from nbody_ai_b200 import _abi
    if os.path.exists(_abi.lib_path()):
        return
    if env_int("LOCAL_RANK", 0) == 0:
        import __graft_entry__
        __graft_entry__.build()
    else:
        t0 = time.time()
        while not os.path.exists(_abi.lib_path()) and time.time() - t0 < 600:
            time.sleep(1.0)
        time.sleep(2.0)   # let the linker finish writing


def main():
    global OUT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--draws", type=int, default=10000)
    ap.add_argument("--omegas", type=int, default=6)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--ref-sample", type=int, default=1024)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    with StdoutToStderr() as OUT:
        ensure_built()
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
