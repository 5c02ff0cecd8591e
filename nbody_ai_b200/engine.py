"""Host side of the batched TTV engine: planning, caller-owned buffers and the nbt_run call.

This file holds no arithmetic of the path — every number comes out of libnbt.so (CUDA,
sm_100a).  It sizes the CSR buffers with the library's own host helpers (nbt_plan_*), owns
the numpy / torch storage the C ABI writes into, and shards batches by system index over
the ranks of a torch.distributed job (no collective on the data path).

Reference path replaced: the per-simulation Python loops around nbody/simulation.py:27-52
(generate), :131-160 (integrate), :163-177 (transit_times, TTV), :179-231 (analyze), as
driven by simulations/generate_simulations.py:41-110, simulations/simulation_grid.py:136-153
and nbody/nested.py:85-95.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import (F_DEVICE_PTRS, F_LS_OC, F_LS_RV, F_MEAN_OMEGA, F_MEANS, F_ORBIT, F_PLANET1_ONLY,
                   F_RV, F_SERIES, NBT_NPAR, SCHEME_AUTO, SCHEMES, TT_GRID_LINEAR, TT_REFINED)


def _as_params(params):
    p = np.ascontiguousarray(params, dtype=np.float64)
    if p.ndim == 2:
        p = p[None]
    if p.ndim != 3 or p.shape[2] != NBT_NPAR:
        raise ValueError("params must be [B, n_bodies, %d] (m, P, e, inc, Omega, omega, f)" % NBT_NPAR)
    if not (2 <= p.shape[1] <= _abi.NBT_MAX_BODIES):
        raise ValueError("n_bodies must be in 2..%d" % _abi.NBT_MAX_BODIES)
    return p


class Likelihood:
    """The likelihood nbt_run evaluates behind the forward model (NBT_F_LOGLIKE): either
    nested.py:85-95 (`nested`: planet-1 O-C against epochs x, times y, errors yerr, linear term c1*x + c0 per
    forward model) or ttv_fit.py:58-94 (`ttvfit`: observed mid-transit times of every planet with data)."""

    def __init__(self, kind, **kw):
        self.kind = kind
        f8 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i4 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        if kind == "nested":
            self.x, self.y, self.yerr = f8(kw["x"]), f8(kw["y"]), f8(kw["yerr"])
            self.c0, self.c1 = f8(kw["c0"]), f8(kw["c1"])
            self.loss = _abi.LOSSES[kw.get("loss", "linear")]
            self.n_data = len(self.x)
            assert len(self.y) == len(self.yerr) == self.n_data and self.c0.shape == self.c1.shape
        elif kind == "ttvfit":
            self.has_data, self.orbit = i4(kw["has_data"]), i4(kw["orbit"])
            self.tc, self.tc_err = f8(kw["tc"]), f8(kw["tc_err"])
            self.loss = 0
            self.n_data = len(self.orbit)
            assert self.tc.shape == self.tc_err.shape == (len(self.has_data), self.n_data)
        else:
            raise ValueError("likelihood kind must be 'nested' or 'ttvfit'")

    ARRAYS = (("c0", C.c_double), ("c1", C.c_double), ("x", C.c_double), ("y", C.c_double), ("yerr", C.c_double),
              ("has_data", C.c_int32), ("tc", C.c_double), ("tc_err", C.c_double), ("orbit", C.c_int32))

    def struct(self, pointer_of=None):
        """nbt_like over this object's host arrays (or over `pointer_of(name, array)` for device copies)."""
        L = _abi.nbt_like()
        L.kind = _abi.LIKE_NESTED if self.kind == "nested" else _abi.LIKE_TTVFIT
        L.loss, L.n_data = int(self.loss), int(self.n_data)
        for name, ct in self.ARRAYS:
            a = getattr(self, name, None)
            if a is not None:
                setattr(L, name, _abi.ptr(a, ct) if pointer_of is None else pointer_of(name, a, ct))
        return L


class Plan:
    """Configuration + CSR layout of one nbt_run call (host planning through libnbt's helpers).

    n_days / n_outputs: scalars (one output grid for the whole batch) or arrays [B] — every system on its own grid
    np.linspace(0, n_days[b], n_outputs[b]) (simulations/generate_simulations.py:45-50), still one launch.
    pool: a BufferPool whose page-locked storage holds the plan's arrays (params copied in), for callers that plan every
    call anew — the plan is then valid until the pool serves the next one.  threads: host threads of nbt_plan (0 = all)."""

    def __init__(self, params, n_days, n_outputs, substeps=None, scheme=SCHEME_AUTO,
                 tt_mode=TT_GRID_LINEAR, flags=F_MEANS, device=0, sort=True, lanes=None, likelihood=None, threads=0, pool=None):
        lib = _abi.load()
        self.params = _as_params(params)
        B, N, _ = self.params.shape
        new = (lambda name, n, dt: np.zeros(n, dtype=dt)) if pool is None else (lambda name, n, dt: pool._take("plan/" + name, (n,), dt))
        if pool is not None:
            pinned = pool._take("plan/params", self.params.shape, np.float64)
            np.copyto(pinned, self.params)
            self.params = pinned
        if isinstance(scheme, str):
            scheme = SCHEMES[scheme]
        self.n_days_sys = self.n_outputs_sys = None
        if np.ndim(n_days) > 0 or np.ndim(n_outputs) > 0:
            self.n_days_sys = np.ascontiguousarray(np.broadcast_to(np.asarray(n_days, dtype=np.float64), (B,)))
            self.n_outputs_sys = np.ascontiguousarray(np.broadcast_to(np.asarray(n_outputs, dtype=np.int32), (B,)))
            n_days = float(self.n_days_sys.max()) if B else 1.0
            n_outputs = int(self.n_outputs_sys.max()) if B else 2
            if B and int(self.n_outputs_sys.min()) < 2:
                raise ValueError("Noutputs must be >= 2")
        if int(n_outputs) < 2:
            raise ValueError("Noutputs must be >= 2")
        self.likelihood = likelihood
        if likelihood is not None:
            flags |= _abi.F_LOGLIKE
        self.cfg = _abi.make_config(B, N, n_days, n_outputs, substeps=0, scheme=scheme,
                                    tt_mode=tt_mode, flags=flags, device=device)
        self.B, self.N, self.Np = B, N, N - 1
        # the whole plan in one call of the library's threaded host planner (nbt_plan): composition steps per output
        # interval; launch order (heavy systems first, equal trip counts within a warp, unstable ones grouped or spread;
        # NBT_SCHEME_AUTO: the systems marked for C4 steps last — sorted or not, scheme uniform per block); which
        # systems run one planet per lane; CSR capacities of the transit products and of the O-C periodograms
        sub_out = None
        if substeps is None:
            self.substeps = None
            sub_out = new("substeps", B, np.int32)
        elif np.isscalar(substeps):
            self.cfg.substeps = int(substeps)
            self.substeps = None
        else:
            self.substeps = np.ascontiguousarray(substeps, dtype=np.int32)
            assert self.substeps.shape == (B,)
        self.order = new("order", B, np.int32)
        self.tt_offsets = new("tt_offsets", B * self.Np + 1, np.int64)
        self.ls_offsets = new("ls_offsets", B * self.Np + 1, np.int64) if flags & F_LS_OC else None
        summary = np.zeros(5, dtype=np.int64)
        _abi.check(lib.nbt_plan(C.byref(self.cfg), C.byref(self._inputs()), 1 if sort else 0, _abi.ptr(sub_out, C.c_int32),
                                _abi.ptr(self.order, C.c_int32), _abi.ptr(self.tt_offsets, C.c_int64),
                                _abi.ptr(self.ls_offsets, C.c_int64), _abi.ptr(summary, C.c_int64), int(threads)), lib)
        if sub_out is not None:
            self.substeps = sub_out
        self.n_steps = int(summary[0])
        self.cfg.tt_cap_max = int(summary[1])
        self.cfg.alt_systems = int(summary[2])
        if summary[3]:
            self.order = None
        if lanes is None:
            lanes = int(summary[4])
        self.cfg.lanes_systems = int(min(max(int(lanes), 0), B))
        self.rv_nfreq = 0
        if flags & F_LS_RV:
            step = float(n_days) / (int(n_outputs) - 1)
            self.rv_nfreq = int(lib.nbt_ls_nfreq(float(n_days) - step, self.cfg.ls_rv_minfreq, self.cfg.ls_rv_maxfreq,
                                                 self.cfg.ls_samples_per_peak))

    def _inputs(self):
        """nbt_inputs over whatever host planning arrays exist so far (for the nbt_plan_* helpers)."""
        inp = _abi.nbt_inputs()
        inp.params = _abi.ptr(self.params)
        inp.substeps = _abi.ptr(getattr(self, "substeps", None), C.c_int32)
        inp.order = _abi.ptr(getattr(self, "order", None), C.c_int32)
        inp.n_days_sys = _abi.ptr(self.n_days_sys)
        inp.n_outputs_sys = _abi.ptr(self.n_outputs_sys, C.c_int32)
        return inp

    def uncalibrated(self):
        """Mask of the systems outside the envelope the sub-step rule was calibrated on (what nbt_run reports as
        NBT_S_UNCALIBRATED: a pair inside 4 mutual Hill radii, e > 0.1 or m_planet / m_star > 0.02)."""
        lib = _abi.load()
        return np.array([lib.nbt_uncalibrated(_abi.ptr(self.params[b]), self.N) for b in range(self.B)], dtype=bool)

    def pin(self):
        """Move the plan's input arrays into page-locked host memory (torch), so that the H2D
        copies inside nbt_run are true asynchronous DMA."""
        import torch
        self._pinned = []
        for name in ("params", "substeps", "order", "tt_offsets", "ls_offsets", "n_days_sys", "n_outputs_sys"):
            a = getattr(self, name)
            if a is not None:
                t = torch.from_numpy(a).pin_memory()
                self._pinned.append(t)
                setattr(self, name, t.numpy())
        return self

    @property
    def times(self):
        return np.linspace(0.0, self.cfg.n_days, self.cfg.n_outputs)  # simulation.py:135

    def times_of(self, b):
        if self.n_days_sys is None:
            return self.times
        return np.linspace(0.0, self.n_days_sys[b], self.n_outputs_sys[b])


OUTPUT_NAMES = (("tt", C.c_double), ("ttv", C.c_double), ("n_tt", C.c_int32), ("period", C.c_double),
                ("t0", C.c_double), ("ttv_max", C.c_double), ("mean_a", C.c_double),
                ("mean_e", C.c_double), ("mean_inc", C.c_double), ("mean_omega", C.c_double),
                ("status", C.c_int32), ("rv", C.c_double), ("rv_max", C.c_double),
                ("orbit_xz", C.c_double), ("series", C.c_double), ("ls_oc_power", C.c_double),
                ("ls_oc_nfreq", C.c_int32), ("ls_rv_power", C.c_double), ("ls_oc_peaks", C.c_double),
                ("loglike", C.c_double))


class Buffers:
    """Caller-owned host output buffers of one Plan (numpy; pinned when torch+CUDA is present).

    skip: names of products the caller does not want back (tt, ttv, n_tt, period, t0, ttv_max, ls_oc_power,
    ls_oc_nfreq): they are left NULL in nbt_outputs — computed on the device, never copied to the host."""

    def __init__(self, plan, pinned=False, skip=()):
        self.plan = plan
        B, Np, NO = plan.B, plan.Np, plan.cfg.n_outputs
        fl = plan.cfg.flags
        self._keep = []
        skip = set(skip)
        z = lambda shape, dt=np.float64: self._alloc(shape, dt, pinned)
        zs = lambda name, shape, dt=np.float64: None if name in skip else self._alloc(shape, dt, pinned)
        ntot = int(plan.tt_offsets[-1])
        self.tt = zs("tt", ntot); self.ttv = zs("ttv", ntot)
        self.n_tt = zs("n_tt", B * Np, np.int32)
        self.period = zs("period", B * Np); self.t0 = zs("t0", B * Np); self.ttv_max = zs("ttv_max", B * Np)
        self.mean_a = z(B * Np); self.mean_e = z(B * Np); self.mean_inc = z(B * Np)
        self.mean_omega = z(B * Np)
        self.status = z(B, np.int32)
        self.rv = z((NO - 1, B)) if fl & F_RV else None
        self.rv_max = z(B) if fl & F_RV else None
        nq = (NO + plan.cfg.orbit_stride - 1) // plan.cfg.orbit_stride
        self.orbit_xz = z((nq, B, Np, 2)) if fl & F_ORBIT else None
        self.series = z((NO, B, 3 + 10 * Np)) if fl & F_SERIES else None
        self.ls_oc_power = zs("ls_oc_power", int(plan.ls_offsets[-1])) if fl & F_LS_OC else None
        self.ls_oc_nfreq = zs("ls_oc_nfreq", B * Np, np.int32) if fl & F_LS_OC else None
        self.ls_rv_power = z((B, plan.rv_nfreq)) if fl & F_LS_RV else None
        self.ls_oc_peaks = z((B * Np, _abi.LS_NPEAKS, 2)) if fl & _abi.F_LS_PEAKS else None
        self.loglike = z(B) if fl & _abi.F_LOGLIKE else None

    def _alloc(self, shape, dtype, pinned):
        pool = getattr(self, "_pool_alloc", None)
        if pool is not None:
            self._n_alloc = getattr(self, "_n_alloc", 0) + 1
            return pool("a%d" % self._n_alloc, shape if isinstance(shape, tuple) else (int(shape),), dtype)
        if pinned:
            import torch
            t = torch.zeros(shape, dtype=torch.float64 if dtype == np.float64 else torch.int32, pin_memory=True)
            self._keep.append(t)
            return t.numpy()
        return np.zeros(shape, dtype=dtype)

    def as_structs(self, params=None):
        p = self.plan
        par = p.params if params is None else _as_params(params)
        inp = _abi.nbt_inputs()
        inp.params = _abi.ptr(par)
        inp.substeps = _abi.ptr(p.substeps, C.c_int32)
        inp.order = _abi.ptr(p.order, C.c_int32)
        inp.tt_offsets = _abi.ptr(p.tt_offsets, C.c_int64)
        inp.ls_offsets = _abi.ptr(p.ls_offsets, C.c_int64)
        inp.n_days_sys = _abi.ptr(p.n_days_sys)
        inp.n_outputs_sys = _abi.ptr(p.n_outputs_sys, C.c_int32)
        like = None
        if p.likelihood is not None:
            like = p.likelihood.struct()
            inp.like = C.pointer(like)
        out = _abi.nbt_outputs()
        for name, ct in OUTPUT_NAMES:
            setattr(out, name, _abi.ptr(getattr(self, name), ct))
        self._structs = (inp, out, par, like)  # keep alive
        return inp, out

    # ---- views -------------------------------------------------------------------------
    def n_transits(self, b, j):
        p = self.plan
        sp = b * p.Np + j
        cap = int(p.tt_offsets[sp + 1] - p.tt_offsets[sp])
        return min(int(self.n_tt[sp]), cap)

    def tt_of(self, b, j):
        off = int(self.plan.tt_offsets[b * self.plan.Np + j])
        return self.tt[off:off + self.n_transits(b, j)]

    def ttv_of(self, b, j):
        """O-C of planet j; the reference's `<3 transits` fallback is ttv=[0] (simulation.py:207-210)."""
        off = int(self.plan.tt_offsets[b * self.plan.Np + j])
        n = self.n_transits(b, j)
        return self.ttv[off:off + (n if n >= 3 else 1)]

    def ls_oc_of(self, b, j):
        """(freq, power) of the O-C periodogram of planet j (tools.py:97-104 on arange(n), ttv)."""
        p = self.plan
        sp = b * p.Np + j
        nf = int(self.ls_oc_nfreq[sp])
        off = int(p.ls_offsets[sp])
        n = max(self.n_transits(b, j), 2)
        df = 1.0 / (n - 1.0) / p.cfg.ls_samples_per_peak
        return p.cfg.ls_oc_minfreq + df * np.arange(nf), self.ls_oc_power[off:off + nf]

    def ls_rv_of(self, b):
        p = self.plan
        step = p.cfg.n_days / (p.cfg.n_outputs - 1)
        T = p.cfg.n_days - step
        df = 1.0 / T / p.cfg.ls_samples_per_peak
        return p.cfg.ls_rv_minfreq + df * np.arange(p.rv_nfreq), self.ls_rv_power[b]


class BufferPool:
    """Page-locked host storage that outlives the plans it serves: a caller whose parameters change from call to call
    (a sampler, the training-set generator) plans every call anew, and the CSR sizes move with the parameters — but
    pinning 100s of MB per call costs more than the call.  `buffers_for(plan)` hands out a Buffers whose arrays are
    views into the pool (grown when a plan needs more)."""

    def __init__(self, pinned=True, headroom=1.1):
        self.pinned, self.headroom, self.store = pinned, headroom, {}

    def _take(self, name, shape, dtype):
        n = int(np.prod(shape))
        name = "%s/%s" % (name, np.dtype(dtype).name)
        cur = self.store.get(name)
        if cur is None or cur.size < n:
            m = max(int(n * self.headroom), 1)
            if self.pinned:
                import torch
                t = torch.zeros(m, dtype={"float64": torch.float64, "int32": torch.int32, "int64": torch.int64}[np.dtype(dtype).name],
                                pin_memory=True)
                self.store[name + "/t"] = t
                cur = t.numpy()
            else:
                cur = np.zeros(m, dtype=dtype)
            self.store[name] = cur
        return cur[:n].reshape(shape)

    def buffers_for(self, plan, skip=()):
        buf = Buffers.__new__(Buffers)
        buf._pool_alloc = self._take
        Buffers.__init__(buf, plan, pinned=False, skip=skip)
        return buf


class SharedHostPool(BufferPool):
    """The final host gather of a sharded batch on one node, without a copy.

    One process per GPU: each rank's results land in its own host buffers.  Here those buffers live in a
    POSIX shared-memory segment per rank, page-locked for CUDA (cudaHostRegister), so that the D2H copies of nbt_run are
    DMA straight into memory that every other process of the job can map: after a barrier rank 0 (or any rank) reads
    every shard in place through `SharedHostPool.attach(name)`.  The "gather" is the barrier.

    A pool serves one plan at a time: call reset(), then Plan(..., pool=pool) and pool.buffers_for(plan), run, and
    publish(plan, buffers) — which writes the array directory into the segment's header for the readers."""

    HEADER = 1 << 16
    DIR = "/dev/shm"      # tmpfs: the segment is ordinary shared memory, named so that the other ranks can map it

    def __init__(self, name, capacity_bytes, register=True):
        import mmap
        import os
        self.name, self.capacity = name, int(capacity_bytes)
        self.path = os.path.join(self.DIR, name)
        fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
        try:
            os.ftruncate(fd, self.HEADER + self.capacity)
            self.map = mmap.mmap(fd, self.HEADER + self.capacity)
        finally:
            os.close(fd)
        self.base = np.frombuffer(self.map, dtype=np.uint8)
        self.base[:] = 0                                   # first touch by the owner: pages on this rank's NUMA node
        self.registered = False
        if register:
            import torch
            rc = torch.cuda.cudart().cudaHostRegister(self.base.ctypes.data, self.base.nbytes, 0)
            if int(rc) != 0:
                raise RuntimeError("cudaHostRegister of the shared segment failed (%s)" % rc)
            self.registered = True
        self.pinned, self.headroom, self.store = False, 1.0, {}
        self.used = 0

    def reset(self):
        self.used = 0

    def _take(self, name, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        off = (self.used + 255) & ~255
        if off + n > self.capacity:
            raise MemoryError("SharedHostPool %s: %d bytes needed, capacity %d" % (self.name, off + n, self.capacity))
        self.used = off + n
        return self.base[self.HEADER + off:self.HEADER + off + n].view(dtype).reshape(shape)

    def buffers_for(self, plan, skip=()):
        buf = Buffers.__new__(Buffers)
        buf._pool_alloc = self._take
        Buffers.__init__(buf, plan, pinned=False, skip=skip)
        return buf

    def publish(self, plan, buffers, extra=None):
        """Write the directory of this step's arrays (results, the CSR offsets they are indexed by, the global system
        indices of the shard if given in `extra`) into the header."""
        import pickle
        entries = {}
        arrays = {n: getattr(buffers, n) for n, _ in OUTPUT_NAMES}
        arrays.update(tt_offsets=plan.tt_offsets, ls_offsets=plan.ls_offsets)
        arrays.update(extra or {})
        lo = self.base.ctypes.data + self.HEADER
        for n, a in arrays.items():
            if a is None:
                continue
            if not (lo <= a.ctypes.data and a.ctypes.data + a.nbytes <= lo + self.capacity):   # not from the pool: copy in
                b = self._take("pub/" + n, a.shape, a.dtype)
                np.copyto(b, a)
                a = b
            entries[n] = (a.ctypes.data - lo, a.shape, a.dtype.str)
        blob = pickle.dumps(dict(B=plan.B, Np=plan.Np, flags=int(plan.cfg.flags), arrays=entries))
        assert len(blob) + 64 <= self.HEADER
        self.base[64:64 + len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        hdr = self.base[:24].view(np.int64)          # [0] directory bytes, [1] steps published, [2] steps consumed (reader's ack)
        hdr[0] = len(blob)
        hdr[1] += 1                                  # after the directory and the data: a reader that sees it sees them

    def wait_consumed(self, timeout=10.0):
        """Block until the reader has acknowledged everything published so far (before the segment is reused).  Returns
        False if it has not within `timeout` seconds — the caller decides whether overwriting an unread step is an error
        (a dead reader must not hang the producers)."""
        import time
        hdr = self.base[:24].view(np.int64)
        t0 = time.perf_counter()
        while hdr[2] < hdr[1]:
            if time.perf_counter() - t0 > timeout:
                return False
            time.sleep(2e-4)
        return True

    def close(self):
        """Unregister and unlink the segment (the mapping itself goes with the last view of it)."""
        import os
        if self.registered:
            import torch
            torch.cuda.cudart().cudaHostUnregister(self.base.ctypes.data)
            self.registered = False
        self.base = self.map = None
        try:
            os.unlink(self.path)
        except OSError:
            pass

    @staticmethod
    def attach(name, ack=False):
        """View of another rank's published step: dict of arrays (views into the shared segment) + B, Np, and `seq`, the
        number of steps published so far.  ack=True maps the segment writable so that acknowledge() can tell the owner
        that the step has been consumed (see wait_consumed)."""
        import mmap
        import os
        import pickle
        fd = os.open(os.path.join(SharedHostPool.DIR, name), os.O_RDWR if ack else os.O_RDONLY)
        try:
            shm = mmap.mmap(fd, 0, prot=(mmap.PROT_READ | mmap.PROT_WRITE) if ack else mmap.PROT_READ)
        finally:
            os.close(fd)
        base = np.frombuffer(shm, dtype=np.uint8)
        hdr = base[:24].view(np.int64)
        out = dict(seq=int(hdr[1]), _shm=shm, _hdr=hdr)
        if out["seq"] == 0:
            return out
        n = int(hdr[0])
        meta = pickle.loads(base[64:64 + n].tobytes())
        out.update(B=meta["B"], Np=meta["Np"], flags=meta["flags"])
        for k, (off, shape, dt) in meta["arrays"].items():
            cnt = int(np.prod(shape)) * np.dtype(dt).itemsize
            out[k] = base[SharedHostPool.HEADER + off:SharedHostPool.HEADER + off + cnt].view(np.dtype(dt)).reshape(shape)
        return out

    @staticmethod
    def acknowledge(view):
        """Reader side: everything published up to view['seq'] has been consumed (needs attach(..., ack=True))."""
        view["_hdr"][2] = view["seq"]


def run(plan, buffers=None, stream=None):
    """One nbt_run over host buffers (H2D + kernels + D2H inside the call).  Raises without CUDA."""
    lib = _abi.load()
    if buffers is None:
        buffers = Buffers(plan)
    inp, out = buffers.as_structs()
    cfg = plan.cfg
    cfg.flags &= ~F_DEVICE_PTRS
    rc = lib.nbt_run(C.byref(cfg), C.byref(inp), C.byref(out), C.c_void_p(stream or 0))
    _abi.check(rc, lib)
    return buffers


class Pipeline:
    """`depth` nbt_run calls in flight from `depth` host threads: ctypes releases the GIL inside libnbt (nbt_plan and
    nbt_run alike), every worker has its own CUDA stream and its own page-locked pool, so the planning and the H2D / D2H
    copies of one batch overlap the kernels of the next — the throughput form of the host-buffer path for callers whose
    batches do not depend on each other (the training-set generator, grids, a sampler's independent walkers).

        pipe = Pipeline(depth=2, device=0)
        futures = [pipe.submit(params_k, n_days, n_outputs, flags=...) for params_k in batches]
        plan, buffers = futures[0].result()

    A result stays valid until its worker serves another batch, i.e. until `depth` more submissions; `on_done(plan,
    buffers, worker_pool)` runs in the worker right after nbt_run (e.g. SharedHostPool.publish).  pools: one pool per
    worker (default: pinned BufferPools)."""

    def __init__(self, depth=2, device=0, pools=None, threads=0):
        import queue
        import threading
        import torch
        self.depth, self.device, self.threads = depth, device, threads
        self.pools = pools if pools is not None else [BufferPool(pinned=True) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(device=device) for _ in range(depth)]
        self.jobs = [queue.Queue() for _ in range(depth)]
        self.next = 0
        self.unacknowledged = 0      # steps of a SharedHostPool overwritten before their reader acknowledged them
        self.workers = [threading.Thread(target=self._work, args=(w,), daemon=True) for w in range(depth)]
        for t in self.workers:
            t.start()

    def _work(self, w):
        pool, stream = self.pools[w], self.streams[w].cuda_stream
        while True:
            job = self.jobs[w].get()
            if job is None:
                return
            fut, params, n_days, n_outputs, kw, skip, on_done = job
            try:
                if hasattr(pool, "wait_consumed"):
                    if not pool.wait_consumed():
                        self.unacknowledged += 1          # the reader is gone or stuck: overwrite rather than hang
                    pool.reset()
                plan = Plan(params, n_days, n_outputs, device=self.device, pool=pool, threads=self.threads, **kw)
                buf = run(plan, pool.buffers_for(plan, skip=skip), stream=stream)
                if on_done is not None:
                    on_done(plan, buf, pool)
                fut.set_result((plan, buf))
            except BaseException as exc:      # surfaced by Future.result() in the caller's thread
                fut.set_exception(exc)

    def submit(self, params, n_days, n_outputs, skip=(), on_done=None, **plan_kw):
        """Queue one batch (round-robin over the workers); returns a concurrent.futures.Future of (plan, buffers)."""
        from concurrent.futures import Future
        fut = Future()
        self.jobs[self.next].put((fut, params, n_days, n_outputs, plan_kw, skip, on_done))
        self.next = (self.next + 1) % self.depth
        return fut

    def close(self):
        for q in self.jobs:
            q.put(None)
        for t in self.workers:
            t.join()


def run_batch(params, n_days, n_outputs, **kw):
    plan = Plan(params, n_days, n_outputs, **kw)
    return run(plan)


# -------------------------------------------------------------------------------------------
# Verified mode.  The sub-step rule is calibrated on well-separated systems; for the ones nbt_run flags
# NBT_S_UNCALIBRATED (a pair inside 4 mutual Hill radii, e > 0.1, a heavy planet) the truncation error of a
# fixed-step composition is amplified by the system's own sensitivity, which nothing predicts a priori.  So they
# are checked the way one checks any integration: again at twice the resolution.  Where the two runs agree to
# half a parity tolerance the finer one is kept; the rest (close encounters, chaotic pairs: a few per thousand
# systems of the reference's priors) are re-integrated with the reference's own integrator, IAS15, on the GPU.
# -------------------------------------------------------------------------------------------
VERIFY_TOL = dict(tt_s=1.0, oc_min=1e-3, a_rel=1e-8, e_rel=1e-8)      # BASELINE.json:north_star


def _grids_of(plan, idx):
    if plan.n_days_sys is None:
        return plan.cfg.n_days, plan.cfg.n_outputs
    return plan.n_days_sys[idx], plan.n_outputs_sys[idx]


def _csr_index(offsets, Np, rows):
    """Flat indices of the CSR slices of systems `rows` (all their planets), in order."""
    lo = offsets[rows * Np]
    hi = offsets[rows * Np + Np]
    n = hi - lo
    return np.repeat(lo - np.concatenate(([0], np.cumsum(n)[:-1])), n) + np.arange(int(n.sum()))


def splice(dst, rows, src):
    """Overwrite the results of systems `rows` of Buffers `dst` with systems 0..len(rows)-1 of Buffers `src` (a run of
    the same systems on the same output grids, so every CSR capacity matches)."""
    rows = np.asarray(rows, dtype=np.int64)
    dp, sp = dst.plan, src.plan
    Np = dp.Np
    q = np.arange(len(rows), dtype=np.int64)
    d_tt, s_tt = _csr_index(dp.tt_offsets, Np, rows), _csr_index(sp.tt_offsets, Np, q)
    assert len(d_tt) == len(s_tt), "splice: transit capacities differ"
    for name in ("tt", "ttv"):
        if getattr(dst, name) is not None and getattr(src, name) is not None:
            getattr(dst, name)[d_tt] = getattr(src, name)[s_tt]
    if dst.ls_oc_power is not None and src.ls_oc_power is not None:
        dst.ls_oc_power[_csr_index(dp.ls_offsets, Np, rows)] = src.ls_oc_power[_csr_index(sp.ls_offsets, Np, q)]
    d_sp = (rows[:, None] * Np + np.arange(Np)[None]).ravel()
    s_sp = (q[:, None] * Np + np.arange(Np)[None]).ravel()
    for name in ("n_tt", "period", "t0", "ttv_max", "mean_a", "mean_e", "mean_inc", "mean_omega", "ls_oc_nfreq", "ls_oc_peaks"):
        if getattr(dst, name) is not None and getattr(src, name) is not None:
            getattr(dst, name)[d_sp] = getattr(src, name)[s_sp]
    for name in ("status", "rv_max", "loglike", "ls_rv_power"):
        if getattr(dst, name) is not None and getattr(src, name) is not None:
            getattr(dst, name)[rows] = getattr(src, name)[q]
    for name in ("rv", "orbit_xz", "series"):
        if getattr(dst, name) is not None and getattr(src, name) is not None:
            getattr(dst, name)[:, rows] = getattr(src, name)[:, q]


def disagreement(a, rows_a, b, tol=VERIFY_TOL):
    """Worst difference, in units of the parity tolerances, between systems `rows_a` of Buffers `a` and systems
    0.. of Buffers `b` (same systems, same grids): max over transit times [s], O-C [min], mean a and mean e
    (relative, e with the 1e-4 floor of the parity tests).  inf where transit counts differ or a run broke."""
    pa, pb = a.plan, b.plan
    Np = pa.Np
    out = np.zeros(len(rows_a))
    bad = _abi.S_NONFINITE | _abi.S_UNBOUND | _abi.S_KEPLER | _abi.S_TT_OVERFLOW
    for q, r in enumerate(rows_a):
        w = 0.0
        if (a.status[r] | b.status[q]) & bad:
            out[q] = np.inf
            continue
        for j in range(Np):
            if a.n_tt[r * Np + j] != b.n_tt[q * Np + j]:
                w = np.inf
                break
            n = a.n_transits(r, j)
            if n:
                w = max(w, np.abs(a.tt_of(r, j) - b.tt_of(q, j)).max() * 86400 / tol["tt_s"])
            if n >= 3:
                w = max(w, np.abs(a.ttv_of(r, j) - b.ttv_of(q, j)).max() * 1440 / tol["oc_min"])
            if pa.cfg.flags & F_MEANS:
                sa, sb = r * Np + j, q * Np + j
                w = max(w, abs(a.mean_a[sa] / b.mean_a[sb] - 1) / tol["a_rel"],
                        abs(a.mean_e[sa] - b.mean_e[sb]) / max(b.mean_e[sb], 1e-4) / tol["e_rel"])
        out[q] = w if np.isfinite(w) else np.inf
    return out


def run_verified(plan, buffers=None, accept=0.5, fallback="ias15", report=None):
    """nbt_run + verification of the systems flagged NBT_S_UNCALIBRATED (see above).  Returns the Buffers of `plan`
    with the verified systems' results replaced by the finer / IAS15 run; a system that could not be settled keeps
    NBT_S_UNRESOLVED in its status (fallback=None).  report: dict that receives the counts."""
    if plan.likelihood is not None:
        raise ValueError("run_verified compares transit products: run it without a fused likelihood")
    buffers = run(plan, buffers)
    risky = np.flatnonzero(buffers.status & _abi.S_UNCALIBRATED)
    info = dict(systems=int(plan.B), flagged=int(len(risky)), accepted_doubling=0, ias15=0, unresolved=0)
    if report is not None:
        report.update(info)
    if len(risky) == 0 or plan.cfg.scheme in (_abi.SCHEME_IAS15, _abi.SCHEME_WHFAST):
        return buffers
    nd, no = _grids_of(plan, risky)
    flags = plan.cfg.flags & ~(F_DEVICE_PTRS | _abi.F_TIMING)
    k = np.abs(plan.substeps[risky]) if plan.substeps is not None else np.full(len(risky), max(plan.cfg.substeps, 1))
    scheme = _abi.SCHEME_SBABC3 if plan.cfg.scheme == SCHEME_AUTO else plan.cfg.scheme
    twin = run(Plan(plan.params[risky], nd, no, substeps=(2 * k).astype(np.int32), scheme=scheme, tt_mode=plan.cfg.tt_mode,
                    flags=flags, device=plan.cfg.device))
    d = disagreement(buffers, risky, twin)
    ok = d <= accept
    splice(buffers, risky[ok], _subset(twin, np.flatnonzero(ok)))
    info["accepted_doubling"] = int(ok.sum())
    rest = risky[~ok]
    if len(rest):
        if fallback == "ias15":
            nd, no = _grids_of(plan, rest)
            ias = run(Plan(plan.params[rest], nd, no, scheme=_abi.SCHEME_IAS15, tt_mode=plan.cfg.tt_mode, flags=flags,
                           device=plan.cfg.device, sort=False, lanes=0))
            splice(buffers, rest, ias)
            info["ias15"] = int(len(rest))
        else:
            buffers.status[rest] |= _abi.S_UNRESOLVED
            info["unresolved"] = int(len(rest))
    if report is not None:
        report.update(info)
    return buffers


def _subset(buf, rows):
    """Buffers holding systems `rows` of `buf` re-packed as systems 0..len(rows)-1 (for splice)."""
    rows = np.asarray(rows, dtype=np.int64)
    p = buf.plan
    nd, no = _grids_of(p, rows)
    sub = Buffers(Plan(p.params[rows], nd, no, substeps=1, scheme=_abi.SCHEME_SBABC3, tt_mode=p.cfg.tt_mode,
                       flags=p.cfg.flags & ~(F_DEVICE_PTRS | _abi.F_TIMING), device=p.cfg.device, sort=False, lanes=0))
    Np = p.Np
    q = np.arange(len(rows), dtype=np.int64)
    s_tt, d_tt = _csr_index(p.tt_offsets, Np, rows), _csr_index(sub.plan.tt_offsets, Np, q)
    for name in ("tt", "ttv"):
        if getattr(buf, name) is not None:
            getattr(sub, name)[d_tt] = getattr(buf, name)[s_tt]
    if buf.ls_oc_power is not None:
        sub.ls_oc_power[_csr_index(sub.plan.ls_offsets, Np, q)] = buf.ls_oc_power[_csr_index(p.ls_offsets, Np, rows)]
    s_sp = (rows[:, None] * Np + np.arange(Np)[None]).ravel()
    for name in ("n_tt", "period", "t0", "ttv_max", "mean_a", "mean_e", "mean_inc", "mean_omega", "ls_oc_nfreq", "ls_oc_peaks"):
        if getattr(buf, name) is not None and getattr(sub, name) is not None:
            getattr(sub, name)[...] = getattr(buf, name)[s_sp]
    for name in ("status", "rv_max", "loglike", "ls_rv_power"):
        if getattr(buf, name) is not None and getattr(sub, name) is not None:
            getattr(sub, name)[...] = getattr(buf, name)[rows]
    for name in ("rv", "orbit_xz", "series"):
        if getattr(buf, name) is not None and getattr(sub, name) is not None:
            getattr(sub, name)[...] = getattr(buf, name)[:, rows]
    return sub


# -------------------------------------------------------------------------------------------
# Multi-GPU: shard by system index, one process per GPU, no collective on the data path.
# -------------------------------------------------------------------------------------------
def shard_range(n_systems, rank, world_size):
    """Contiguous, balanced [lo, hi) slice of the batch owned by `rank`."""
    base, rem = divmod(int(n_systems), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def system_cost(params, n_days, n_outputs, scheme=SCHEME_AUTO):
    """Relative integration cost of every system of a batch: Kepler-drift stages per output interval under
    the planner's sub-step rule (3 per SBABC3 step, 2 per C4 step, ...) plus one for the per-sample work."""
    lib = _abi.load()
    params = np.ascontiguousarray(params, dtype=np.float64)
    B, N = params.shape[0], params.shape[1]
    if isinstance(scheme, str):
        scheme = SCHEMES[scheme]
    cfg = _abi.make_config(B, N, n_days, n_outputs, substeps=0, scheme=scheme)
    k = np.zeros(B, dtype=np.int32)
    inp = _abi.nbt_inputs()
    inp.params = _abi.ptr(params)
    _abi.check(lib.nbt_plan_substeps(C.byref(cfg), C.byref(inp), _abi.ptr(k, C.c_int32)), lib)
    stages = {0: 1, 1: 3, 2: 4, 3: 4, 4: 2, 5: 3, 6: 3}[int(scheme)]
    return np.where(k < 0, -2.0 * k, float(stages) * k) + 1.0


def shard_balanced(cost, rank, world_size):
    """Cost-balanced shard of a batch: systems sorted by decreasing cost are dealt to the ranks in
    snake order (0..W-1, W-1..0, ...), so that every rank gets the same count (+-1) and the same mix of
    long and short systems.  Returns the sorted indices owned by `rank`; pass them to gather_results."""
    cost = np.asarray(cost, dtype=np.float64)
    order = np.argsort(-cost, kind="stable")
    pos = np.arange(len(order))
    lap, slot = np.divmod(pos, int(world_size))
    owner = np.where(lap % 2 == 0, slot, int(world_size) - 1 - slot)
    return np.sort(order[owner == int(rank)])


def gather_results(local, n_systems, group=None, index=None):
    """Final host gather of per-rank result dicts (arrays whose first axis is the local system
    index) onto every rank, via torch.distributed.all_gather_object (gloo or nccl groups).
    index: the global indices of this rank's systems (shard_balanced); None = contiguous shards in
    rank order (shard_range)."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    parts = [None] * ws
    dist.all_gather_object(parts, (local, None if index is None else np.asarray(index)), group=group)
    out = {}
    for k in local:
        cat = np.concatenate([p[0][k] for p in parts], axis=0)
        assert cat.shape[0] == n_systems
        if index is not None:
            where = np.concatenate([p[1] for p in parts])
            assert np.array_equal(np.sort(where), np.arange(n_systems))
            full = np.empty_like(cat)
            full[where] = cat
            cat = full
        out[k] = cat
    return out


# -------------------------------------------------------------------------------------------
# Device-resident form: inputs and outputs live in HBM as torch tensors, nbt_run only enqueues
# kernels on torch's current stream (NBT_F_DEVICE_PTRS).  torch is plumbing here (memory + streams).
# -------------------------------------------------------------------------------------------
class DeviceBuffers:
    _OUT = (("tt", "f8", "ntt"), ("ttv", "f8", "ntt"), ("n_tt", "i4", "nsp"), ("period", "f8", "nsp"),
            ("t0", "f8", "nsp"), ("ttv_max", "f8", "nsp"), ("mean_a", "f8", "nsp"), ("mean_e", "f8", "nsp"),
            ("mean_inc", "f8", "nsp"), ("mean_omega", "f8", "nsp"), ("status", "i4", "B"))

    def __init__(self, plan, device=None):
        import torch
        self.torch = torch
        self.plan = plan
        self.device = torch.device("cuda", plan.cfg.device) if device is None else torch.device(device)
        B, Np, NO = plan.B, plan.Np, plan.cfg.n_outputs
        fl = plan.cfg.flags
        sizes = {"ntt": int(plan.tt_offsets[-1]), "nsp": B * Np, "B": B}
        dev = self.device
        up = lambda a: None if a is None else torch.from_numpy(a).to(dev)
        self.params, self.substeps, self.order = up(plan.params), up(plan.substeps), up(plan.order)
        self.tt_offsets, self.ls_offsets = up(plan.tt_offsets), up(plan.ls_offsets)
        self.n_days_sys, self.n_outputs_sys = up(plan.n_days_sys), up(plan.n_outputs_sys)
        self.like_arrays, self.like = {}, None
        if plan.likelihood is not None:      # device copies of the likelihood's arrays; the struct itself stays on the host
            def dev_ptr(name, a, ct):
                self.like_arrays[name] = torch.from_numpy(a).to(dev)
                return self._p(self.like_arrays[name], ct)
            self.like = plan.likelihood.struct(dev_ptr)
        for name, dt, key in self._OUT:
            setattr(self, name, torch.zeros(max(sizes[key], 1), dtype=torch.float64 if dt == "f8" else torch.int32, device=dev))
        f8 = lambda *shape: torch.zeros(*shape, dtype=torch.float64, device=dev)
        self.rv = f8(NO - 1, B) if fl & F_RV else None
        self.rv_max = f8(B) if fl & F_RV else None
        nq = (NO + plan.cfg.orbit_stride - 1) // plan.cfg.orbit_stride
        self.orbit_xz = f8(nq, B, Np, 2) if fl & F_ORBIT else None
        self.series = f8(NO, B, 3 + 10 * Np) if fl & F_SERIES else None
        self.ls_oc_power = f8(max(int(plan.ls_offsets[-1]), 1)) if fl & F_LS_OC else None
        self.ls_oc_nfreq = torch.zeros(B * Np, dtype=torch.int32, device=dev) if fl & F_LS_OC else None
        self.ls_rv_power = f8(B, plan.rv_nfreq) if fl & F_LS_RV else None
        self.ls_oc_peaks = f8(B * Np, _abi.LS_NPEAKS, 2) if fl & _abi.F_LS_PEAKS else None
        self.loglike = f8(B) if fl & _abi.F_LOGLIKE else None

    def set_params(self, params):
        """New parameter values for the same plan (a sampler's next batch): one H2D copy."""
        self.params.copy_(self.torch.from_numpy(_as_params(params)), non_blocking=True)

    @staticmethod
    def _p(t, ct):
        return C.cast(C.c_void_p(0 if t is None else t.data_ptr()), C.POINTER(ct))

    def as_structs(self):
        inp = _abi.nbt_inputs()
        inp.params = self._p(self.params, C.c_double)
        inp.substeps = self._p(self.substeps, C.c_int32)
        inp.order = self._p(self.order, C.c_int32)
        inp.tt_offsets = self._p(self.tt_offsets, C.c_int64)
        inp.ls_offsets = self._p(self.ls_offsets, C.c_int64)
        inp.n_days_sys = self._p(self.n_days_sys, C.c_double)
        inp.n_outputs_sys = self._p(self.n_outputs_sys, C.c_int32)
        if self.like is not None:
            inp.like = C.pointer(self.like)
        out = _abi.nbt_outputs()
        for name, ct in OUTPUT_NAMES:
            setattr(out, name, self._p(getattr(self, name), ct))
        return inp, out

    def to_host(self):
        """Copy every output into a host Buffers (for checks)."""
        hb = Buffers(self.plan)
        for name, _ in OUTPUT_NAMES:
            t = getattr(self, name)
            h = getattr(hb, name)
            if t is not None and h is not None:
                h[...] = t.cpu().numpy().reshape(-1)[:h.size].reshape(h.shape)
        return hb


def run_device(plan, dbuf, timing=False):
    """Enqueue the whole path on torch's current stream over device-resident buffers (no copies,
    no synchronisation)."""
    lib = _abi.load()
    inp, out = dbuf.as_structs()
    cfg = plan.cfg
    saved = cfg.flags
    cfg.flags = saved | F_DEVICE_PTRS | (_abi.F_TIMING if timing else 0)
    try:
        stream = dbuf.torch.cuda.current_stream(dbuf.device).cuda_stream
        rc = lib.nbt_run(C.byref(cfg), C.byref(inp), C.byref(out), C.c_void_p(stream))
    finally:
        cfg.flags = saved
    _abi.check(rc, lib)
    return dbuf


def stage_timing():
    """Device milliseconds of the last timed nbt_run on this thread: (front end, fit, ls_oc, rv)."""
    lib = _abi.load()
    ms = np.zeros(4)
    _abi.check(lib.nbt_timing(_abi.ptr(ms)), lib)
    return ms


# -------------------------------------------------------------------------------------------
# Single-process multi-GPU: one host thread per device (ctypes releases the GIL inside nbt_run),
# each integrating its contiguous shard of the batch — the in-process form of SURVEY.md §8e.
# -------------------------------------------------------------------------------------------
class ShardedBuffers:
    """Results of a batch that was split over several devices: routes the per-system views of
    Buffers (tt_of, ttv_of, n_transits, ls_oc_of, ...) to the shard that owns the system."""

    def __init__(self, shards, bounds):
        self.shards, self.bounds = shards, bounds      # bounds[i] = (lo, hi) of shard i
        self.B = bounds[-1][1]
        self.Np = shards[0].plan.Np

    def locate(self, b):
        for sh, (lo, hi) in zip(self.shards, self.bounds):
            if lo <= b < hi:
                return sh, b - lo
        raise IndexError(b)

    def n_transits(self, b, j):
        sh, q = self.locate(b)
        return sh.n_transits(q, j)

    def tt_of(self, b, j):
        sh, q = self.locate(b)
        return sh.tt_of(q, j)

    def ttv_of(self, b, j):
        sh, q = self.locate(b)
        return sh.ttv_of(q, j)

    def ls_oc_of(self, b, j):
        sh, q = self.locate(b)
        return sh.ls_oc_of(q, j)

    def gathered(self, name):
        """Per-(system, planet) or per-system array `name` (period, t0, ttv_max, mean_a, ..., status,
        n_tt) concatenated over the shards in system order."""
        return np.concatenate([getattr(sh, name) for sh in self.shards])


def run_multi_gpu(params, n_days, n_outputs, devices=None, **plan_kw):
    """Shard `params` by system index over `devices` (default: every visible GPU) and run the shards
    concurrently, one host thread per device.  Returns ShardedBuffers."""
    import threading
    lib = _abi.load()
    params = _as_params(params)
    if devices is None:
        devices = list(range(lib.nbt_device_count()))
    if not devices:
        raise RuntimeError("run_multi_gpu: no CUDA device (there is no CPU fallback)")
    B = params.shape[0]
    devices = devices[:max(1, min(len(devices), B))]
    bounds = [shard_range(B, r, len(devices)) for r in range(len(devices))]
    shards, errors = [None] * len(devices), []

    def work(r):
        try:
            lo, hi = bounds[r]
            shards[r] = run(Plan(params[lo:hi], n_days, n_outputs, device=devices[r], **plan_kw))
        except Exception as exc:      # surfaced in the caller's thread below
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(r,)) for r in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return ShardedBuffers(shards, bounds)
