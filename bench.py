#!/usr/bin/env python
"""Benchmark of the batched TTV hot path (BASELINE.json metric: FP64 N-body system-steps/s and
TTV sims/s at 1/2/4/8 B200 beside the host CPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Headline workload (`value`, `e2e`, `roofline`): config C2 of BASELINE.json (configs[1]) — random
2-planet systems from the priors of randomize() (nbody/simulation.py:54-129) with the omega sweep of
simulations/generate_simulations.py:83-107: 10 000 draws x (1 + 6 omega variants) = 70 000 simulations per GPU per
step; 365 d sampled hourly (8760 outputs); products per planet: transit times, linear ephemeris, O-C, maxavg, O-C
periodogram, means of a/e/inc/omega.  A step is one pass of the whole path over that batch.  Weak scaling: every rank
owns 70 000 systems.

`value` is whole-job TTV sims/s with inputs resident in HBM (torch device tensors, kernels enqueued through the C ABI
with NBT_F_DEVICE_PTRS), K steps with two in flight: two CUDA streams and two buffer sets alternate, so that the tail of
one heterogeneous launch — its few long systems — is filled by the blocks of the next step.  `value_single_stream`
times the same steps one at a time; that region also gives the per-kernel CUDA-event times of `stage_ms_per_step` and
the roofline's launch duration.  `e2e` is the same through the public host-buffer call a user makes: every step PLANS the
batch anew (engine.Plan -> nbt_plan: sub-steps, launch order, CSR layout — a sampler's or the training-set
generator's parameters change from call to call), copies the parameter block and plan H2D from page-locked memory,
runs the kernels and copies every product D2H into page-locked buffers (nbt_run) — with several steps in flight
(engine.Pipeline: 4 host threads for N <= 2, else 3, one stream and one buffer pool each), so that one step's planning and copies
overlap the next one's kernels; for N > 1 the buffers are shared-memory segments and a reader on rank 0 consumes every
rank's every step in place (the final host gather, inside the timed region).  `e2e_sequential` is the same one step at
a time, `e2e_peaks` with the periodograms reduced to their peaks on the device.

`north_star` block (BASELINE.json's target: >= 1e11 system-steps/s on 8 GPUs for 3-planet systems at >= 50 % of the
FP64 roofline; config 5 sharded over 1/2/4/8): three more device-timed workloads, each with system-steps/s, sims/s and
the SURVEY §8(d) roofline fraction — (i) the README 3-planet system replicated, (ii) random 3-planet systems,
(iii) C5 at full size, 1 M random 4-planet systems over 1095 d, sharded by system over the ranks (strong scaling).

Roofline: the FP64 vector pipe.  achieved = SURVEY §8(d)'s algorithmic FLOPs — system-steps x (272 Np + 12.5 Np (Np+1)
- 40) + output samples x 100 Np — / the CUDA-event duration of k_integrate; peak = a DFMA-chain kernel measured in the
same run (MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s).
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
            "365 d @ 1 h, tt + ephemeris + O-C + maxavg + O-C periodogram + element means "
            "[north_star block: README 3-planet x 227 328 and random 3-planet systems, 365 d @ 1 h; "
            "C5 = 1 M random 4-planet systems, 1095 d @ 1 h, sharded over the ranks]")
FP64_NOMINAL = 37.2


def base_config(args, world):
    """The workload description, identical in both arms (`--impl ours` / `--impl reference`); arm-specific facts go
    into `details` (ours) or `cpu_baseline.sample` (reference)."""
    return {"workload": WORKLOAD, "systems_per_gpu": args.draws * (args.omegas + 1), "draws": args.draws,
            "omega_variants": args.omegas, "n_planets": 2, "n_days": N_DAYS, "n_outputs": N_OUTPUTS,
            "parallelism": "shard-by-system x%d" % world,
            "l2": "256 MiB device fill (> 126 MB L2) in front of every timed step; every step also writes 340 MB of products"}


def f_step(np_):
    """Algorithmic FLOPs per system-step (one Kepler-drift + kick stage of one system), SURVEY.md §8(d):
    272 Np + 12.5 Np (Np + 1) - 40 = 579 / 926 / 1298 for Np = 2 / 3 / 4 (FMA = 2, div / sqrt = 1)."""
    return 272 * np_ + 12.5 * np_ * (np_ + 1) - 40


F_OUT = 100  # per planet per output sample (Jacobi a, e, inc + crossing test), SURVEY.md §8(d)


def survey_flops(plan):
    """§8(d) FLOPs of one k_integrate launch of `plan`: stages + per-sample work (the latter only when the element
    means are on: without NBT_F_MEANS a sample is just the crossing test)."""
    from nbody_ai_b200 import _abi
    samples = int(plan.n_outputs_sys.sum()) if plan.n_outputs_sys is not None else plan.B * plan.cfg.n_outputs
    out = F_OUT * plan.Np * samples if plan.cfg.flags & _abi.F_MEANS else 0
    return plan.n_steps * f_step(plan.Np) + out


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
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_integrate launch on this workload, from the committed
    `ncu --set full` capture (profiles/r2_traffic.json, else r1_traffic.json); None if no capture matches."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
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


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on host cores (oracle port).  Plans with numpy only — this arm never loads
# libnbt.so.
# ---------------------------------------------------------------------------------------------
class HostPlan:
    """What engine.Buffers and oracle.run need of a plan, computed with numpy: CSR capacities of the transit products
    min(ceil(1.05 Ndays / P) + 3, Noutputs) and of the O-C periodograms (astropy's autofrequency count on a baseline of
    capacity - 1 epochs) — the layout include/nbt.h documents for nbt_plan_tt / nbt_plan_ls."""

    def __init__(self, params, n_days, n_outputs, flags):
        from nbody_ai_b200 import _abi
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        B, N, _ = self.params.shape
        self.B, self.N, self.Np = B, N, N - 1
        self.substeps = self.order = self.n_days_sys = self.n_outputs_sys = self.likelihood = None
        nd, no = np.asarray(n_days, dtype=np.float64), np.asarray(n_outputs)
        if nd.ndim or no.ndim:     # per-system grids (generate_simulations.py:45-50)
            self.n_days_sys = np.ascontiguousarray(np.broadcast_to(nd, (B,)), dtype=np.float64)
            self.n_outputs_sys = np.ascontiguousarray(np.broadcast_to(no, (B,)), dtype=np.int32)
            nd, no = self.n_days_sys[:, None], self.n_outputs_sys[:, None].astype(np.float64)
            n_days, n_outputs = float(self.n_days_sys.max()), int(self.n_outputs_sys.max())
        else:
            nd, no = float(n_days), float(n_outputs)
        self.cfg = _abi.make_config(B, N, n_days, n_outputs, substeps=1, scheme=_abi.SCHEME_SBABC3, flags=flags)
        self.rv_nfreq = 0
        P = self.params[:, 1:, _abi.P_P]
        ok = np.isfinite(P) & (P > 0)
        cap = np.where(ok, np.minimum(np.ceil(1.05 * nd / np.where(ok, P, 1.0)) + 3.0, no), 3.0).astype(np.int64)
        if flags & _abi.F_PLANET1_ONLY:
            cap[:, 1:] = 1
        self.tt_offsets = np.concatenate([[0], np.cumsum(cap.ravel())]).astype(np.int64)
        self.cfg.tt_cap_max = int(cap.max()) if B else 1
        self.ls_offsets = None
        if flags & _abi.F_LS_OC:
            c = cap.ravel().astype(np.float64)
            df = 1.0 / np.maximum(c - 1.0, 1.0) / self.cfg.ls_samples_per_peak
            nf = np.where(c >= 3, 1 + np.rint((self.cfg.ls_oc_maxfreq - self.cfg.ls_oc_minfreq) / df), 0).astype(np.int64)
            self.ls_offsets = np.concatenate([[0], np.cumsum(nf)]).astype(np.int64)
        self.n_steps = int(self.n_outputs_sys.sum() - B) if self.n_outputs_sys is not None else B * (n_outputs - 1)


def cpu_path(params, n_days=N_DAYS, n_outputs=N_OUTPUTS, flags=None, n_threads=0):
    """The reference algorithm on host cores: oracle (IAS15 restatement + analyze semantics)."""
    import oracle
    from nbody_ai_b200 import _abi, engine
    if flags is None:
        flags = _abi.F_MEANS | _abi.F_MEAN_OMEGA | _abi.F_LS_OC
    plan = HostPlan(params, n_days, n_outputs, flags)
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
        "config": base_config(args, world),
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
    (first touch) and the D2H traffic of every step stay on the GPU's own socket.  Returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(gpu_index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


class DeviceTimer:
    """K device-timed passes of one plan (CUDA events on torch's current stream = the launch stream, L2 flushed
    between passes outside the event pairs)."""

    def __init__(self, torch, dev, flush, barrier):
        self.torch, self.dev, self.flush, self.barrier = torch, dev, flush, barrier

    def __call__(self, plan, dbuf, steps, warmup):
        from nbody_ai_b200 import engine
        torch = self.torch
        for _ in range(warmup):
            engine.run_device(plan, dbuf)
        torch.cuda.synchronize(self.dev)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        stage_ms = np.zeros(4)
        self.barrier()
        for k in range(steps):
            self.flush.fill_(k & 0xFF)                  # L2 flush, outside the event pair
            ev[k][0].record()
            engine.run_device(plan, dbuf, timing=True)
            ev[k][1].record()
            ev[k][1].synchronize()
            stage_ms += engine.stage_timing()
        self.barrier()
        return float(sum(a.elapsed_time(b) for a, b in ev)), stage_ms


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
    dbuf = engine.DeviceBuffers(plan)

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

    def reduce_max_sum(mx, sm):
        """max over ranks of `mx`, sum over ranks of `sm` (lists of floats)."""
        a = torch.tensor(mx, dtype=torch.float64, device=dev)
        b = torch.tensor(sm, dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(a, op=dist.ReduceOp.MAX)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
        return [float(x) for x in a.cpu()], [float(x) for x in b.cpu()]

    timer = DeviceTimer(torch, dev, flush, barrier)
    for _ in range(args.warmup):
        engine.run_device(plan, dbuf)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)

    # ---- timed region 1: K steps one at a time on one stream (per-kernel CUDA events: the roofline's launch time) ------
    seq_ms, stage_ms = timer(plan, dbuf, args.steps, 0)

    # ---- timed region 2 (`value`): the same K steps with TWO in flight — two streams, two buffer sets alternating — so
    # that the tail of one heterogeneous launch (its few long systems) is filled by the next step's blocks.  N = 1: the
    # second set is another seed of the workload; N > 1: the rank's shard again.  One event pair around the region.
    plan_b = plan if world > 1 else engine.Plan(workloads.c2_params(args.draws, args.omegas, seed=1), N_DAYS, N_OUTPUTS,
                                                flags=flags, device=local_rank)
    main_sets = [(plan, dbuf), (plan_b, engine.DeviceBuffers(plan_b))]
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]

    def overlapped(n, sets=None, streams_=None):
        sets = sets or main_sets
        streams_ = streams_ or streams
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream(dev)
        e0.record(cur)
        for st_ in streams_:
            st_.wait_event(e0)
        for k in range(n):
            with torch.cuda.stream(streams_[k % 2]):
                flush.fill_(k & 0xFF)                  # L2 flush in front of every step, on the step's own stream
                engine.run_device(*sets[k % 2])
        for st_ in streams_:
            cur.wait_stream(st_)
        e1.record(cur)
        e1.synchronize()
        return e0.elapsed_time(e1)

    overlapped(4)
    barrier()
    dev_ms = overlapped(args.steps)
    barrier()
    steps_timed = sum(main_sets[k % 2][0].n_steps for k in range(args.steps))
    clocks = sampler.finish()
    del main_sets

    # ---- end to end through the host-pointer C ABI: plan + H2D + kernels + D2H every step --------------------
    # N = 1: a fresh parameter block every step (other seeds of the same workload), so nothing of a plan can be reused;
    # N > 1: the rank's shard is planned anew every step.
    #   e2e            the throughput form: engine.Pipeline keeps --e2e-depth steps in flight (host threads, each with its own
    #                  stream and page-locked pool), so one step's planning and copies overlap the next one's kernels.
    #                  N > 1 also times the final host gather: the pools are shared-memory segments that rank 0 maps
    #                  (engine.SharedHostPool); a reader thread on rank 0 consumes every rank's every step in place (it
    #                  sums the transit counts as proof) and acknowledges it before the segment is reused
    #   e2e_sequential one step at a time (plan, nbt_run, next), per-rank pinned pool, no gather — a sampler's view
    #   e2e_peaks      sequential, periodograms reduced to their peaks on the device (NBT_F_LS_PEAKS)
    n_sets = 1 if world > 1 else min(max(args.steps, 1), 4)
    param_sets = [params] + [workloads.c2_params(args.draws, args.omegas, seed=100 + i) for i in range(1, n_sets)]
    depth = max(1, args.e2e_depth if args.e2e_depth > 0 else (4 if world <= 2 else 3))
    plan_threads = max(1, min(8, (os.cpu_count() or 8) // (world * depth)))      # nbt_plan threads per call: no oversubscription
    names = ("tt", "ttv", "n_tt", "period", "t0", "ttv_max", "mean_a", "mean_e", "mean_inc", "mean_omega", "status",
             "ls_oc_power", "ls_oc_nfreq", "ls_oc_peaks")

    # -- sequential
    pool = engine.BufferPool(pinned=True)

    def seq_step(k, fl, skip=()):
        tp = time.perf_counter()
        p = engine.Plan(param_sets[k % n_sets], N_DAYS, N_OUTPUTS, flags=fl, device=local_rank, pool=pool, threads=plan_threads)
        tp = time.perf_counter() - tp
        return p, engine.run(p, pool.buffers_for(p, skip=skip)), tp     # H2D + kernels + D2H + stream sync inside

    for k in range(max(min(args.warmup, 2), n_sets)):        # every set once: the pool reaches its final size
        seq_step(k, flags)
    barrier()
    plan_s = 0.0
    t0 = time.perf_counter()
    for k in range(args.steps):
        p, hbuf, tp = seq_step(k, flags)
        plan_s += tp
    torch.cuda.synchronize(dev)
    e2e_seq_ms = (time.perf_counter() - t0) * 1e3
    h2d = sum(a.nbytes for a in (p.params, p.substeps, p.order, p.tt_offsets, p.ls_offsets) if a is not None)
    d2h = sum(getattr(hbuf, n).nbytes for n in names if getattr(hbuf, n) is not None)
    # -- sequential, periodograms reduced to peaks on the device: the power arrays (two thirds of the D2H bytes) stay there
    lean = ("ls_oc_power", "ls_oc_nfreq")
    for k in range(n_sets):
        seq_step(k, flags | _abi.F_LS_PEAKS, lean)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, pbuf, _ = seq_step(k, flags | _abi.F_LS_PEAKS, lean)
    torch.cuda.synchronize(dev)
    e2e_peaks_ms = (time.perf_counter() - t0) * 1e3
    d2h_peaks = sum(getattr(pbuf, n).nbytes for n in names if getattr(pbuf, n) is not None)
    del p, hbuf, pbuf, pool

    # -- pipelined (+ shared-memory gather for N > 1)
    gather_note, pools, seg = "single rank: results are in the caller's buffers", None, None
    if world > 1:
        seg = "nbt_bench_%s_%%d_%%d" % os.environ.get("MASTER_PORT", "0")
        try:
            need = sum(a.nbytes for a in (plan.params, plan.substeps, plan.tt_offsets, plan.ls_offsets) if a is not None) * 2
            need += 8 * (2 * int(plan.tt_offsets[-1]) + int(plan.ls_offsets[-1]) + 16 * B * Np) + (4 << 20)
            pools = []
            for w in range(depth):
                try:
                    os.unlink(os.path.join(engine.SharedHostPool.DIR, seg % (rank, w)))
                except OSError:
                    pass
                pools.append(engine.SharedHostPool(seg % (rank, w), int(1.2 * need)))
            gather_note = ("zero-copy: every rank's buffers are page-locked shared-memory segments (engine.SharedHostPool, %d per rank); a reader "
                           "thread on rank 0 consumes every step of every rank in place and acknowledges it" % depth)
        except Exception as exc:      # no /dev/shm or registration refused: per-rank pinned buffers, gather not timed
            for q in pools or []:
                q.close()
            pools, gather_note = None, "unavailable (%s): per-rank pinned buffers only" % exc
        ok = torch.tensor([1.0 if pools is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0 and pools is not None:
            for q in pools:
                q.close()
            pools, gather_note = None, "unavailable on another rank: per-rank pinned buffers only"
    shared = pools is not None
    consumed, stop_reader, reader, reader_error, gather_failed = [0, 0], [False], None, [], []
    if shared:
        dist.barrier()                                  # every segment exists
        if rank == 0:
            def read_all():
                try:
                    read_loop()
                except BaseException as exc:            # surfaced by the wait below
                    reader_error.append(exc)

            def read_loop():
                views = {}
                while not stop_reader[0]:
                    idle = True
                    for r in range(world):
                        for w in range(depth):
                            v = views.get((r, w))
                            if v is None or v["_hdr"][1] > v["_hdr"][2]:
                                v = engine.SharedHostPool.attach(seg % (r, w), ack=True)
                                views[(r, w)] = v
                            if v["seq"] > int(v["_hdr"][2]):
                                consumed[1] += int(v["n_tt"].sum()) + int(v["status"].shape[0]) + int(v["ls_oc_power" if "ls_oc_power" in v else "ls_oc_peaks"].size > 0)
                                engine.SharedHostPool.acknowledge(v)
                                consumed[0] += 1
                                idle = False
                    if idle:
                        time.sleep(2e-4)
            reader = threading.Thread(target=read_all, daemon=True)
            reader.start()

    def publish(pl, bf, pl_pool):
        pl_pool.publish(pl, bf, extra={"index": mine.astype(np.int64)})

    pipe = engine.Pipeline(depth=depth, device=local_rank, pools=pools, threads=plan_threads)

    def pipelined(n, fl, skip):
        futs = [pipe.submit(param_sets[k % n_sets], N_DAYS, N_OUTPUTS, flags=fl, skip=skip, on_done=publish if shared else None) for k in range(n)]
        for f in futs:
            f.result()

    published = [0]

    def timed_pipeline(fl, skip=(), repeats=3):
        """Warm-up, then `repeats` times K pipelined steps (N > 1: each until rank 0 has read every step of every rank —
        the gather); returns the times [ms] of the repeats.  The boxes are shared virtual machines: a host-side hiccup
        of a few hundred ms now and then lands in one repeat, so the median is reported and all three are listed."""
        n_warm = depth * max(2, -(-n_sets // depth))        # a multiple of depth: every worker keeps its parameter sets
        pipelined(n_warm, fl, skip)
        published[0] += n_warm
        out = []
        for _ in range(repeats):
            barrier()
            t0 = time.perf_counter()
            pipelined(args.steps, fl, skip)
            published[0] += args.steps
            if shared:
                if rank == 0:
                    t_wait = time.perf_counter()
                    while consumed[0] < world * published[0] and not gather_failed:
                        if reader_error or time.perf_counter() - t_wait > 60.0:      # never hang the job: report it instead
                            gather_failed.append(str(reader_error[0]) if reader_error else "rank 0 did not see every step within 60 s")
                            consumed[0] = world * published[0]
                            for r in range(world):                  # release every producer for good
                                for w in range(depth):
                                    try:
                                        engine.SharedHostPool.attach(seg % (r, w), ack=True)["_hdr"][2] = 1 << 62
                                    except Exception:
                                        pass
                            break
                        time.sleep(2e-4)
                dist.barrier()
            torch.cuda.synchronize(dev)
            out.append((time.perf_counter() - t0) * 1e3)
        return out

    e2e_runs = timed_pipeline(flags)
    e2e_peaks_runs = timed_pipeline(flags | _abi.F_LS_PEAKS, lean)      # the same with the periodograms reduced to peaks
    e2e_ms, e2e_peaks_pipe_ms = float(np.median(e2e_runs)), float(np.median(e2e_peaks_runs))
    pipe.close()
    stop_reader[0] = True
    if reader is not None:
        reader.join()
    if shared:
        dist.barrier()
        for q in pools:
            q.close()
    del pipe, pools

    # sanity: the timed device path and the host path produced the same transits
    chk = dbuf.to_host()
    ref_run = engine.run(plan)
    assert np.array_equal(chk.n_tt, ref_run.n_tt), "device-resident and host-pointer paths disagree"
    n_tt_total = int(np.minimum(ref_run.n_tt, np.diff(plan.tt_offsets)).sum())
    del chk, ref_run

    my_dev_ms, my_e2e_ms = dev_ms, e2e_seq_ms      # this rank's own times (the `ranks` table), before the max over ranks
    (dev_ms, e2e_ms, integ_ms, e2e_peaks_ms, e2e_seq_ms, seq_ms, e2e_peaks_pipe_ms), (sims_all, steps_all, ntt_all, steps_one) = reduce_max_sum(
        [dev_ms, e2e_ms, stage_ms[0], e2e_peaks_ms, e2e_seq_ms, seq_ms, e2e_peaks_pipe_ms], [float(B), float(steps_timed), float(n_tt_total), float(plan.n_steps)])
    # per-rank view (diagnosis of the max-over-ranks figure): step time, e2e time, SM clock, power-cap flag
    mine_t = torch.tensor([my_dev_ms / args.steps, my_e2e_ms / args.steps, clocks["sm_mhz"] or 0.0,
                         1.0 if "sw_power_cap" in clocks["reasons"] else 0.0], dtype=torch.float64, device=dev)
    per_rank = [mine_t]
    if dist is not None:
        per_rank = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(per_rank, mine_t)
    per_rank = [[round(float(x), 3) for x in r.cpu()] for r in per_rank]
    main_flops, main_launch_ms = survey_flops(plan), stage_ms[0] / args.steps
    substeps_hist = np.bincount(np.abs(plan.substeps)).tolist()
    c4_systems = int((plan.substeps < 0).sum())
    del dbuf, param_sets
    torch.cuda.empty_cache()

    # ---- north-star block -------------------------------------------------------------------------------------
    north = {}
    launches_extra = 0
    if not args.no_north_star:
        def block(name, par, n_days, n_outputs, fl, note, scaling, steps=2, warmup=1, cpu_sample=0, overlap=True):
            nonlocal launches_extra
            t_plan = time.perf_counter()
            pl = engine.Plan(par, n_days, n_outputs, flags=fl, device=local_rank)
            t_plan = time.perf_counter() - t_plan
            db = engine.DeviceBuffers(pl)
            ms_tot, st = timer(pl, db, steps, warmup)
            status = db.status.cpu().numpy()
            bad = int(((status & (_abi.S_NONFINITE | _abi.S_UNBOUND | _abi.S_KEPLER | _abi.S_TT_OVERFLOW)) != 0).sum())
            ms_two = 0.0
            if overlap:      # the same passes with two in flight (two streams, two buffer sets), like the headline `value`
                db2 = engine.DeviceBuffers(pl)
                pair = [(pl, db), (pl, db2)]
                if warmup:
                    overlapped(2, pair)
                barrier()
                ms_two = overlapped(2 * steps, pair)
                barrier()
                del db2, pair
            del db
            torch.cuda.empty_cache()
            (ms_tot, integ, ms_two), (nsys, nsteps, flops, nbad) = reduce_max_sum(
                [ms_tot, st[0], ms_two], [float(pl.B), float(pl.n_steps), float(survey_flops(pl)), float(bad)])
            launches_extra += (2 + (1 if fl & _abi.F_LS_OC else 0)) * steps * (3 if overlap else 1)
            ach = survey_flops(pl) / (st[0] / steps * 1e-3) * 1e-12       # this rank's kernel against this rank's peak
            entry = {
                "workload": note, "systems": nsys, "n_planets": pl.Np, "n_days": n_days, "n_outputs": n_outputs, "scaling": scaling,
                "passes_timed": steps, "ms_per_pass": ms_tot / steps, "k_integrate_ms_per_pass_max_rank": integ / steps,
                "system_steps_per_pass": nsteps, "system_steps_per_s": nsteps * steps / (ms_tot * 1e-3),
                "sims_per_s": nsys * steps / (ms_tot * 1e-3),
                "roofline": {"bound": "fp64", "kernel": "k_integrate<%d>" % pl.Np, "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": ach / fp64_peak, "frac_of_nominal_37.2": ach / FP64_NOMINAL, "flops_per_launch_rank0": survey_flops(pl),
                             "flops_per_system_step": f_step(pl.Np), "launch_ms_rank0": st[0] / steps},
                "systems_broken_or_overflowing": nbad, "host_plan_s_rank0": t_plan,
            }
            if overlap and ms_two > 0:
                rate = 2 * steps / (ms_two * 1e-3)
                entry["two_in_flight"] = {
                    "passes_timed": 2 * steps, "ms_per_pass": ms_two / (2 * steps), "system_steps_per_s": nsteps * rate, "sims_per_s": nsys * rate,
                    "roofline_frac": flops * rate * 1e-12 / (fp64_peak * world),
                    "note": "two passes in flight on two streams (the next pass fills the tail of the current launch); frac = SURVEY 8(d) FLOPs of the passes / elapsed time of the region (k_fit and k_ls_oc inside) / measured FP64 peak"}
            if cpu_sample and rank == 0 and world == 1 and not args.no_cpu:
                dt, nt, _ = cpu_path(par[:cpu_sample], n_days, n_outputs, fl)
                entry["cpu_port"] = {"sims_per_s": cpu_sample / dt, "system_steps_per_s": cpu_sample * (n_outputs - 1) / dt, "cores": nt,
                                     "sample": "first %d systems, oracle on %d host threads, %.1f s" % (cpu_sample, nt, dt)}
            north[name] = entry

        full = _abi.F_MEANS | _abi.F_MEAN_OMEGA
        block("three_planet_readme", workloads.c1_params(args.np3_systems), N_DAYS, N_OUTPUTS, full,
              "README.md:23-38 system (1.12 Msun + 3 planets, P = 3.29 / 7 / 12 d) x %d per GPU, tt + ephemeris + O-C + maxavg + element means" % args.np3_systems,
              "weak", cpu_sample=64)
        block("three_planet_random", workloads.random_systems(args.np3_random, 3, seed=300 + rank), N_DAYS, N_OUTPUTS, full | _abi.F_LS_OC,
              "random 3-planet systems (randomize() priors extended outward), %d per GPU, all analyze products incl. O-C periodogram" % args.np3_random,
              "weak", cpu_sample=256)
        if args.c5_systems > 0:
            lo, hi = engine.shard_range(args.c5_systems, rank, world)
            # every rank draws the whole-job stream in 125 000-system chunks (seeded per chunk) and keeps its own range
            chunk = 125000
            parts = []
            for c0 in range(0, args.c5_systems, chunk):
                c1 = min(c0 + chunk, args.c5_systems)
                if c1 <= lo or c0 >= hi:
                    continue
                blk = workloads.c5_params(c1 - c0, seed=2 + c0 // chunk)
                parts.append(blk[max(lo - c0, 0):min(hi, c1) - c0])
            block("c5_full", np.concatenate(parts), 1095.0, 26280, _abi.F_MEANS,
                  "BASELINE config 5: %d random 4-planet systems in total, 1095 d @ 1 h, tt + ephemeris + O-C + maxavg + mean a/e/inc, sharded by system over %d rank(s)" % (args.c5_systems, world),
                  "strong", steps=args.c5_passes, warmup=0, cpu_sample=128)

    if rank == 0:
        sims_per_s = sims_all * args.steps / (dev_ms * 1e-3)
        steps_per_s = steps_all / (dev_ms * 1e-3)          # steps_all: system-steps of all K timed steps, all ranks
        achieved = main_flops / (main_launch_ms * 1e-3) * 1e-12
        line = {
            "metric": "ttv_sims_per_sec", "value": sims_per_s, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": base_config(args, world),
            "details": {"scheme": "auto: sbabc3 (3-stage order (6,4) Wisdom-Holman composition with force-gradient end kicks); c4 (2-stage force-gradient) where P_min >= 250 output steps",
                        "substeps_hist": substeps_hist, "c4_systems": c4_systems,
                        "value": "K steps with two in flight: two CUDA streams and two buffer sets alternate (the next step's blocks fill the tail of the current launch); a 256 MiB fill precedes every step on its stream; one event pair around the region.  value_single_stream: one at a time",
                        "shards": "cost-balanced (engine.shard_balanced) of one global batch" if world > 1 else "one batch", "cpu_affinity": affinity,
                        "e2e": "every step: engine.Plan (nbt_plan host planner, %d threads) on %s + nbt_run with page-locked host buffers; e2e = engine.Pipeline, %d steps in flight; e2e_sequential = one at a time" % (plan_threads, "a fresh parameter block (%d seeds in rotation)" % n_sets if n_sets > 1 else "the rank's shard", depth)},
            "steps_in_flight": 2,
            "value_single_stream": {"value": sims_all * args.steps / (seq_ms * 1e-3), "unit": "sims/s", "ms_per_step": seq_ms / args.steps,
                                    "system_steps_per_s": steps_one * args.steps / (seq_ms * 1e-3),
                                    "note": "the same K steps one at a time on one stream (what `stage_ms_per_step` and `roofline` time)"},
            "system_steps_per_s": steps_per_s,
            "system_steps_per_step": steps_one,
            "transits_per_step": ntt_all,
            "stage_ms_per_step": {"k_integrate": stage_ms[0] / args.steps, "k_fit": stage_ms[1] / args.steps,
                                  "k_ls_oc": stage_ms[2] / args.steps},
            "roofline": {"bound": "fp64", "kernel": "k_integrate<%d>" % Np, "achieved": achieved, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak > 0 else None,
                         "frac_of_nominal_37.2": achieved / FP64_NOMINAL,
                         "model": "SURVEY 8(d): system-steps x (272 Np + 12.5 Np (Np+1) - 40) + output samples x 100 Np",
                         "peak_source": "nbt_fp64_peak DFMA-chain kernel measured in this run (nominal 37.2 at 1965 MHz)",
                         "peak_three_register_operands": fp64_peak_3reg,
                         "flops_per_launch": main_flops, "launch_ms": main_launch_ms, "traffic": ncu_traffic(Np, B),
                         "launch_timing": "CUDA events around k_integrate on its launch stream (NBT_F_TIMING) in the single-stream region: one launch at a time",
                         "two_steps_in_flight": {"achieved": main_flops * args.steps / (my_dev_ms * 1e-3) * 1e-12,
                                                 "frac": main_flops * args.steps / (my_dev_ms * 1e-3) * 1e-12 / fp64_peak,
                                                 "note": "k_integrate's FLOPs of the K steps / the elapsed time of the overlapped region (which also holds k_fit and k_ls_oc): the rate the FP64 pipe sustains once the tail of a launch is filled by the next"},
                         "hbm": hbm_side(ncu_traffic(Np, B), main_launch_ms)},
            "e2e": {"value": sims_all * args.steps / (e2e_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps, "steps_in_flight": depth,
                    "repeats_ms_per_step_rank0": [round(x / args.steps, 3) for x in e2e_runs], "statistic": "median of 3 repeats of K steps (max over ranks of each rank's median)",
                    "final_host_gather": gather_note if not gather_failed else "INCOMPLETE (%s) — %s" % (gather_failed[0], gather_note)},
            "e2e_sequential": {"value": sims_all * args.steps / (e2e_seq_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": int(h2d),
                               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_seq_ms / args.steps,
                               "host_plan_ms_per_step_rank0": 1e3 * plan_s / args.steps,
                               "note": "one step at a time: plan, nbt_run (H2D + kernels + D2H + sync), next; per-rank pinned buffers, no gather"},
            "e2e_peaks": {"value": sims_all * args.steps / (e2e_peaks_ms * 1e-3), "unit": "sims/s", "h2d_bytes_per_step": int(h2d),
                          "d2h_bytes_per_step": int(d2h_peaks), "ms_per_step": e2e_peaks_ms / args.steps,
                          "note": "as e2e_sequential, with NBT_F_LS_PEAKS: each O-C periodogram reduced to its 3 highest peaks on the device, power arrays not copied back",
                          "pipelined": {"value": sims_all * args.steps / (e2e_peaks_pipe_ms * 1e-3), "unit": "sims/s", "ms_per_step": e2e_peaks_pipe_ms / args.steps,
                                        "note": "as e2e (steps in flight, gather for N > 1), with NBT_F_LS_PEAKS"}},
            "gpu_launches": 6 * args.steps + launches_extra,   # k_integrate, k_fit, k_ls_oc per step of the two device-timed regions (+ the north-star passes)
            "clocks": clocks,
            "ranks": {"columns": ["ms_per_step", "e2e_sequential_ms_per_step", "sm_mhz", "sw_power_cap"], "rows": per_rank},
            "north_star": north,
        }
        if world == 1 and not args.no_cpu:
            n_s = min(args.cpu_sample, B)
            dt, nt, _ = cpu_path(params[:n_s])
            line["cpu_baseline"] = {"value": n_s / dt, "unit": "sims/s", "cores": nt, "kind": "port", "cpu_model": cpu_model(),
                                    "host_cpus": os.cpu_count(),
                                    "sample": "first %d systems of the same batch, oracle (IAS15 + analyze) on %d host threads, %.1f s" % (n_s, nt, dt)}
        if world == 1 and not args.no_blocks:
            line["blocks"] = extra_blocks(args, local_rank)
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def extra_blocks(args, local_rank):
    """Three more end-to-end figures of BASELINE configs on the public API, beside the CPU port (N = 1 only; C4 — the TTV
    grid of simulations/simulation_grid.py at full size through grid.simulation_grid — is the third):
    C3 — the nested-sampling likelihood batch (nbody/nested.py:85-95): loglike/s through nested.NestedLikelihood
         .loglike_batch, i.e. planning + H2D + forward model at 30-min cadence + ephemeris + chi^2 on the device +
         D2H of loglike[B] per call;
    C2 generator — simulations/generate_simulations.py:41-110 at its own settings (periods = 1000, 6 omega variants,
         every draw on its own grid of P1 x 1000 days at 1 h): accepted draws/s through dataset.generate_training_set."""
    import oracle
    from nbody_ai_b200 import _abi, dataset, nested, workloads
    out = {}
    # ---- C3 ----------------------------------------------------------------------------------------------
    n3 = args.c3_batch
    objects = [{'m': 1.22}, {'m': 10.4 * 0.000954588, 'P': 0.94145, 'inc': np.pi / 2, 'e': 0, 'omega': 0},
               {'m': 3e-4, 'P': 2.1, 'inc': np.pi / 2, 'e': 0.02, 'omega': 1.0}]
    bounds = [[0.0, 1.0], {'P': [0.94135, 0.94155]}, {'P': [1.9, 2.4], 'm': [1.5e-5, 4.2e-4], 'e': [0, 0.1], 'omega': [0, 2 * np.pi]}]
    rng = np.random.default_rng(1)
    xx = np.arange(56, dtype=float)
    _, _, tt0 = nested.get_ttv(objects, ndays=79)
    yy = tt0[:56] + rng.normal(0, 0.5 / 1440, 56)
    yerr = np.full(56, 0.5 / 1440)
    L = nested.NestedLikelihood(xx, yy, yerr, objects, bounds, device=local_rank)
    def cubes(n):
        return np.column_stack([rng.normal(tt0[0], 1e-4, n), rng.normal(0.94145, 1e-6, n), rng.uniform(0.94135, 0.94155, n),
                                rng.uniform(1.9, 2.4, n), rng.uniform(1.5e-5, 4.2e-4, n), rng.uniform(0, 0.1, n), rng.uniform(0, 2 * np.pi, n)])
    sets = [cubes(n3) for _ in range(4)]
    for c in sets[:2]:
        L.loglike_batch(c)
    t0 = time.perf_counter()
    for k in range(args.c3_calls):
        ll = L.loglike_batch(sets[k % 4])
    dt3 = time.perf_counter() - t0
    entry = {"workload": "BASELINE config 3: %d parameter vectors per call (WASP-18-like star + hot Jupiter + perturber), %d d @ 30 min, 56 observed epochs" % (n3, L.ndays),
             "api": "nested.NestedLikelihood.loglike_batch (nbt_plan + nbt_run with NBT_F_LOGLIKE: only loglike[B], status[B] come back)",
             "calls_timed": args.c3_calls, "ms_per_call": 1e3 * dt3 / args.c3_calls, "loglike_per_s": n3 * args.c3_calls / dt3,
             "finite_frac": float(np.isfinite(ll).mean())}
    if not args.no_cpu:
        ns = min(256, n3)
        par = L.params_for(sets[0][:ns])
        plan = HostPlan(par, L.ndays, int(L.ndays * 48), _abi.F_PLANET1_ONLY)
        from nbody_ai_b200 import engine
        buf = engine.Buffers(plan)
        nt = oracle.max_threads()
        t0 = time.perf_counter()
        oracle.run(plan.cfg, plan.params, buf, n_threads=nt)
        for b in range(ns):
            oracle.loglike(buf.ttv_of(b, 0), sets[0][b, 0], sets[0][b, 1], xx, yy, yerr)
        dtc = time.perf_counter() - t0
        entry["cpu_port"] = {"loglike_per_s": ns / dtc, "cores": nt, "sample": "%d of the vectors, oracle (IAS15 + analyze + chi^2) on %d host threads, %.1f s" % (ns, nt, dtc)}
    out["c3_likelihood"] = entry
    # ---- training-set generator -----------------------------------------------------------------------------
    st = {}
    dataset.generate_training_set(samples=8, omegas=6, periods=1000, rng=np.random.default_rng(5), device=local_rank)      # warm-up
    t0 = time.perf_counter()
    X, y = dataset.generate_training_set(samples=args.gen_draws, omegas=6, periods=1000, rng=np.random.default_rng(6), device=local_rank, stats=st)
    dtg = time.perf_counter() - t0
    entry = {"workload": "simulations/generate_simulations.py:41-110 at its defaults: periods = 1000, 6 omega variants per accepted draw, every simulation on its own grid (P1 x 1000 d @ 1 h)",
             "api": "dataset.generate_training_set (candidate draws + variants in one nbt_run with per-system grids, rejection on the host)",
             "accepted_draws": args.gen_draws, "rows": len(X), "seconds": dtg, "draws_per_s": args.gen_draws / dtg, "simulations_per_s": st["simulations"] / dtg,
             "launches": st["launches"], "candidates": st["candidates"], "acceptance": st["accepted"] / max(st["candidates"], 1)}
    if not args.no_cpu:
        blk = dataset.draw_block(4, 6, np.random.default_rng(7))
        par = blk.reshape(-1, 3, _abi.NBT_NPAR)
        nd, no = dataset.grids(par, 1000)
        plan = HostPlan(par, nd, no, _abi.F_PLANET1_ONLY)
        from nbody_ai_b200 import engine
        buf = engine.Buffers(plan)
        nt = oracle.max_threads()
        t0 = time.perf_counter()
        oracle.run(plan.cfg, plan.params, buf, n_threads=nt)
        dtc = time.perf_counter() - t0
        acc = st["accepted"] / max(st["candidates"], 1)
        entry["cpu_port"] = {"simulations_per_s": len(par) / dtc, "draws_per_s": len(par) / dtc / (1.0 / max(acc, 1e-9) + 6.0), "cores": nt,
                             "sample": "4 candidate draws x 7 simulations, oracle on %d host threads, %.1f s; draws/s = simulations/s / (1 / acceptance + 6)" % (nt, dtc)}
    out["training_set_generator"] = entry
    # ---- C4: the reference's TTV grid at full size -----------------------------------------------------------------
    from nbody_ai_b200 import grid
    npts = args.c4_npts
    objs = workloads.c4_objects()
    mass_grid, period_grid = workloads.c4_grids(npts)
    grid.simulation_grid(objs, mass_grid[:8, :8], period_grid[:8, :8], epochs=500, n_order=4, device=local_rank)      # warm-up
    t0 = time.perf_counter()
    g = grid.simulation_grid(objs, mass_grid, period_grid, epochs=500, n_order=4, device=local_rank)
    dt4 = time.perf_counter() - t0
    amp = g["ttv_grid"] * 1440
    entry = {"workload": "BASELINE config 4: simulations/simulation_grid.py:88-153 — %d x %d cells (perturber mass 0.5-325 Mearth x period ratio 1.41-2.41 around a 10 Mearth planet at 1.42 d, 1 Msun), 500 epochs @ 30 min (34 080 outputs); per cell: transit times, O-C, exact periodogram, find_peaks, harmonic least squares, softmax amplitude, RV semi-amplitude" % (npts, npts),
             "api": "grid.simulation_grid (one nbt_run + one nbt_grid_post for the whole grid, host buffers)",
             "seconds": dt4, "cells_per_s": npts * npts / dt4, "ttv_amplitude_min_max_minutes": [float(np.nanmin(amp)), float(np.nanmax(amp))],
             "cells_with_a_peak": int((g["per_epoch_grid"][0] > 0).sum())}
    if not args.no_cpu:
        from scipy.signal import find_peaks
        sel = np.linspace(0, npts * npts - 1, 48).astype(int)
        par, meta = workloads.c4_params(npts)
        plan = HostPlan(par[sel], meta["ndays"], meta["noutputs"], _abi.F_PLANET1_ONLY)
        from nbody_ai_b200 import engine
        buf = engine.Buffers(plan)
        nt = oracle.max_threads()
        t0 = time.perf_counter()
        oracle.run(plan.cfg, plan.params, buf, n_threads=nt)
        for b in range(len(sel)):      # the post-processing of generate_sim (:33-69) with the reference's own calls
            Tc = buf.tt_of(b, 0)
            n = len(Tc)
            if n < 3:
                continue
            orbit = np.arange(n)
            t0_, per = np.linalg.lstsq(np.vstack([np.ones(n), orbit]).T, Tc, rcond=None)[0]
            ttv = Tc - per * orbit - t0_
            freq, power = oracle.lomb_scargle(orbit.astype(float), ttv, 1. / (1.5 * 500), 1. / 2.1, 20)
            pk, am = find_peaks(power, height=0.05)
            pk = pk[np.argsort(am['peak_heights'])[::-1]][:4]
            basis = [np.ones(n), orbit] + [f(2 * np.pi * orbit * freq[k]) for k in pk for f in (np.sin, np.cos)]
            np.linalg.lstsq(np.vstack(basis).T, Tc, rcond=None)
        dtc = time.perf_counter() - t0
        entry["cpu_port"] = {"cells_per_s": len(sel) / dtc, "cores": nt,
                             "sample": "48 cells spread over the grid: oracle (IAS15) on %d host threads + numpy / scipy post-processing on one, %.1f s" % (nt, dtc)}
    out["c4_grid"] = entry
    return out


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


def ensure_built(need_cuda_lib=True):
    """The in-tree libraries normally travel with the repo snapshot; compile them if they did not
    (local rank 0 builds, the other ranks of a torchrun job wait for the file)."""
    from nbody_ai_b200 import _abi
    if not need_cuda_lib or os.path.exists(_abi.lib_path()):
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
    ap.add_argument("--no-north-star", action="store_true")
    ap.add_argument("--np3-systems", type=int, default=227328, help="README 3-planet replicas per GPU (148 SMs x 1536)")
    ap.add_argument("--np3-random", type=int, default=75776, help="random 3-planet systems per GPU (148 x 512)")
    ap.add_argument("--c5-systems", type=int, default=1000000, help="C5 systems in total (sharded over the ranks); 0 = skip")
    ap.add_argument("--c5-passes", type=int, default=2)
    ap.add_argument("--e2e-depth", type=int, default=0, help="steps the pipelined e2e keeps in flight (host threads / streams / buffer pools); 0 = 4 for N <= 2, else 3")
    ap.add_argument("--no-blocks", action="store_true", help="skip the C3 likelihood and training-set generator blocks (N = 1)")
    ap.add_argument("--c3-batch", type=int, default=5000)
    ap.add_argument("--c3-calls", type=int, default=10)
    ap.add_argument("--gen-draws", type=int, default=512)
    ap.add_argument("--c4-npts", type=int, default=256)
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    with StdoutToStderr() as OUT:
        ensure_built(need_cuda_lib=args.impl != "reference")
        if args.impl == "reference":
            run_reference(args, rank, world)
        else:
            run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
