"""Batched, device-resident front end of the EMC hot path.

``Context``   one EmcCtx (C-ABI handle) + the parameter snapshot of ``init_`` + uploaded tables
``Ensemble``  structure-of-arrays float64 particle state in torch-owned HBM buffers;
              ``step`` replaces the particle loop of main_.py:360-397 with ONE kernel launch
``Mesh``      charge mesh (int64 fixed point), rho, potential and field on the device;
              deposit -> (all-reduce) -> rho -> SOR -> field  (poisson_.py)
``HostStepper`` the same timestep for callers that hold the reference's host arrays
              ((N,6) coords, valley, dtau): chunked pinned H2D / kernel / D2H pipeline

PyTorch is used for buffer ownership, streams and torch.distributed only; every compute
step is a call through the ctypes binding (:mod:`.abi`) into hand-written CUDA.  There is
no CPU fallback: constructing a Context without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi


def _torch():
    import torch
    return torch


def params_from_reference_modules(efield=None, num_carr=None, z_cyclic=False):
    """Snapshot init_.Parameters / init_.GaAs / init_.Device (init_.py:58-157) into EmcParams."""
    from . import init_, scatter_
    P, G, D = init_.Parameters, init_.GaAs, init_.Device
    p = abi.EmcParams()
    p.dt, p.gamma, p.hbar, p.q = P.dt, G.tot_scat, P.hbar, P.q
    ef = D.elec_field if efield is None else efield
    n_carr = D.num_carr if num_carr is None else num_carr
    for v in range(3):
        p.mass[v] = G.mass[v]
        p.nonparab[v] = G.nonparab[v]
        p.tot_dim[v] = float(D.tot_dim[v])
        p.efield[v] = float(ef[v])
        p.seg[v] = int(D.tot_seg[v])
        p.dl[v] = float(D.dl[v])
    p.e_phonon = G.E_phonon
    p.ion_eb = float(scatter_.ion_eb(0))
    p.kT, p.mean_energy = P.kT, D.mean_energy
    # poisson_.py:27 background, :36 increment
    p.rho_background = float(G.dope * np.prod(D.dl) * (-100 ** 3))
    p.rho_increment = float((G.dope * D.vol * (100 ** 3) / n_carr)[0])
    p.eps_s = G.eps_0
    p.surf_val = 0.6
    p.z_cyclic = int(bool(z_cyclic))          # extension: cyclic z for bulk runs (default: specular, like the reference)
    return p


def params_from_dict(phys: dict, device: dict | None = None):
    """EmcParams from plain dicts (the layout of tests/golden/params.json)."""
    p = abi.EmcParams()
    p.dt, p.gamma, p.hbar, p.q = phys["dt"], phys["gamma"], phys["hbar"], phys["q"]
    for v in range(3):
        p.mass[v] = phys["mass"][v]
        p.nonparab[v] = phys["nonparab"][v]
        p.tot_dim[v] = phys["tot_dim"][v]
        p.efield[v] = float(phys["efield"][v])
    p.e_phonon, p.ion_eb = phys["e_phonon"], phys["ion_eb"]
    p.kT, p.mean_energy = phys.get("kT", 0.02585), phys.get("mean_energy", 0.1)
    dev = device or {}
    seg = dev.get("seg", [1, 1, 1])
    dl = dev.get("dl", [phys["tot_dim"][a] / seg[a] for a in range(3)])
    for v in range(3):
        p.seg[v] = int(seg[v])
        p.dl[v] = float(dl[v])
    p.rho_background = dev.get("rho_background", 0.0)
    p.rho_increment = dev.get("charge", dev.get("rho_increment", 0.0))
    p.eps_s = phys.get("eps_0", 1.0)
    p.surf_val = dev.get("surf_val", 0.6)
    p.z_cyclic = int(bool(phys.get("z_cyclic", 0)))
    return p


class Context:
    """Owns one EmcCtx.  ``tables`` = (ek, [tab_G, tab_L, tab_X], mechanism_arrays)."""

    def __init__(self, params: abi.EmcParams | None = None, device: int = 0, tables=None):
        torch = _torch()
        self.lib = abi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the EMC engine has no CPU fallback")
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)                 # make torch create the primary context
        self.params = params if params is not None else params_from_reference_modules()
        h = C.c_void_p()
        st = self.lib.emc_create(C.byref(self.params), self.device_index, C.byref(h))
        if st < 0:
            msg = self.lib.emc_last_error(None)
            raise abi.EmcError(st, msg.decode() if msg else "?")
        self.h = h
        self._keep = []
        self.obs_before = False
        if tables is None:
            from . import init_, scatter_
            dfEk = init_.init_energy_df(2, 10000)
            tables = (dfEk["Ek"].to_numpy(),
                      [t.to_numpy() for t in scatter_.calc_scatter_table(dfEk, 0)],
                      scatter_.mechanism_arrays(0))
        self.upload_tables(*tables)

    # -- plumbing ---------------------------------------------------------------------
    def check(self, status: int) -> None:
        abi.check(self.lib, self.h, status)

    @property
    def stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    @property
    def seg(self):
        return tuple(int(v) for v in self.params.seg)

    @property
    def padded_shape(self):
        return tuple(int(v) + 2 for v in self.params.seg)

    def launch_count(self) -> int:
        return int(self.lib.emc_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.emc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001  (interpreter shutdown)
            pass

    # -- tables / fields --------------------------------------------------------------
    def upload_tables(self, ek, tabs, mech) -> None:
        """scatter_.calc_scatter_table output -> device (scatter_.py:954-984, :892-951)."""
        ek = np.ascontiguousarray(ek, dtype=np.float64)
        tabs = [np.ascontiguousarray(t, dtype=np.float64) for t in tabs]
        n_mech = np.array([t.shape[1] for t in tabs], dtype=np.int32)
        dE = np.zeros((3, abi.EMC_MAX_MECH), dtype=np.float64)
        trans = np.zeros((3, abi.EMC_MAX_MECH), dtype=np.int32)
        sampler = np.zeros((3, abi.EMC_MAX_MECH), dtype=np.int32)
        for v in range(3):
            _names, d, t, s = mech[v]
            dE[v, :len(d)], trans[v, :len(t)], sampler[v, :len(s)] = d, t, s
        self.check(self.lib.emc_upload_tables(
            self.h, ek.ctypes.data, len(ek), tabs[0].ctypes.data, tabs[1].ctypes.data,
            tabs[2].ctypes.data, n_mech.ctypes.data, dE.ctypes.data, trans.ctypes.data,
            sampler.ctypes.data))
        self.n_mech = n_mech

    def _table_args(self, ek, tabs, mech):
        ek = np.ascontiguousarray(ek, dtype=np.float64)
        tabs = [np.ascontiguousarray(t, dtype=np.float64) for t in tabs]
        n_mech = np.array([t.shape[1] for t in tabs], dtype=np.int32)
        dE = np.zeros((3, abi.EMC_MAX_MECH), dtype=np.float64)
        trans = np.zeros((3, abi.EMC_MAX_MECH), dtype=np.int32)
        sampler = np.zeros((3, abi.EMC_MAX_MECH), dtype=np.int32)
        for v in range(3):
            _names, d, t, s = mech[v]
            dE[v, :len(d)], trans[v, :len(t)], sampler[v, :len(s)] = d, t, s
        keep = (ek, tabs, n_mech, dE, trans, sampler)
        return keep, (ek.ctypes.data, len(ek), tabs[0].ctypes.data, tabs[1].ctypes.data, tabs[2].ctypes.data,
                      n_mech.ctypes.data, dE.ctypes.data, trans.ctypes.data, sampler.ctypes.data)

    def set_layers(self, layers=None, init_weights=None, tables=None) -> None:
        """Multi-layer device (dev.layers, init_.py:127-130): the kernels then look the layer of a
        particle up from its z (init_.where_am_i) before every drift, scatter and observable and
        use that layer's material and table, as main_.py:366-396 does.

        ``layers``: list of ``init_.Layer`` (default ``init_.Device.layers``); the material of
        layer i is ``layers[i].matl`` and its tables are ``scatter_.calc_scatter_table(dfEk, i)``
        with ``Device.Materials[i] = layers[i].matl`` (init_.py:129), unless ``tables`` gives one
        ``(ek, [tab_G, tab_L, tab_X], mechanism_arrays)`` per layer.  ``init_weights``: share of the
        initial ensemble per layer for ``Ensemble.init`` (extension; default uniform z)."""
        from . import init_, scatter_
        layers = list(init_.Device.layers if layers is None else layers)
        n = len(layers)
        assert 1 <= n <= abi.EMC_MAX_LAYERS
        w = [0.0] * n if init_weights is None else [float(x) for x in init_weights]
        if tables is None:
            saved = init_.Device.Materials
            init_.Device.Materials = [l.matl for l in layers]            # init_.py:129
            try:
                dfEk = init_.init_energy_df(2, 10000)
                ek = dfEk["Ek"].to_numpy()
                tables, ebs = [], []
                for i in range(n):
                    first = next(j for j in range(i + 1) if layers[j].matl is layers[i].matl)
                    if first < i:
                        tables.append(tables[first])
                    else:
                        tables.append((ek, [t.to_numpy() for t in scatter_.calc_scatter_table(dfEk, i)],
                                       scatter_.mechanism_arrays(i)))
                    ebs.append(float(scatter_.ion_eb(i)))
            finally:
                init_.Device.Materials = saved
        else:
            saved = init_.Device.Materials
            init_.Device.Materials = [l.matl for l in layers]
            try:
                ebs = [float(scatter_.ion_eb(i)) for i in range(n)]
            finally:
                init_.Device.Materials = saved
        arr = (abi.EmcLayer * n)()
        for i, l in enumerate(layers):
            arr[i].lz = float(l.lz)
            for v in range(3):
                arr[i].mass[v] = l.matl.mass[v]
                arr[i].nonparab[v] = l.matl.nonparab[v]
            arr[i].e_phonon = l.matl.E_phonon
            arr[i].ion_eb = ebs[i]
            arr[i].init_weight = w[i]
        self.check(self.lib.emc_set_layers(self.h, arr, n))
        for i in range(n):
            keep, args = self._table_args(*tables[i])
            self.check(self.lib.emc_upload_tables_layer(self.h, i, *args))
        self.layers = layers

    def set_obs_phase(self, before_step: bool) -> None:
        """``True``: ``Ensemble.step(observe=True)`` reduces the observables of the ensemble as the
        step FOUND it (= after the previous step) in the kernel's fully occupied load phase;
        ``engine.run_steps`` shifts the rows back.  Default ``False``: after the step."""
        self.check(self.lib.emc_set_obs_phase(self.h, int(bool(before_step))))
        self.obs_before = bool(before_step)

    def set_deposit_mode(self, mode: int) -> None:
        """NGP kernel for meshes beyond one block's shared memory: 0 auto, 1 direct atomics, 2 tiled
        counting sort + shared-memory partial meshes (``emc_set_deposit_mode``); identical results."""
        self.check(self.lib.emc_set_deposit_mode(self.h, int(mode)))

    def set_exact_libm(self, on: bool) -> None:
        """``True``: compose acos / asin / atan exactly like scatter_.py:1061-1062 and the samplers
        (:811-872) instead of their algebraic equivalents (``emc_set_exact_libm``)."""
        self.check(self.lib.emc_set_exact_libm(self.h, int(bool(on))))

    def describe(self) -> dict:
        """Run-time knobs that select kernels on this context (``emc_describe``) as a dict."""
        buf = C.create_string_buffer(512)
        self.lib.emc_describe(self.h, buf, 512)
        return dict(kv.split("=", 1) for kv in buf.value.decode().split())

    def select_layer(self, layer: int) -> None:
        """Material used by ``emc_calc_energy`` / ``emc_scatter`` (their ``mat`` / ``mat_i`` argument)."""
        self.check(self.lib.emc_select_layer(self.h, int(layer)))

    def set_background_profile(self, background_z) -> None:
        """Per-z-cell background charge for ``Mesh.to_rho`` (doping profiles); None = uniform."""
        if background_z is None:
            self.check(self.lib.emc_set_background_profile(self.h, None, 0))
            return
        b = np.ascontiguousarray(background_z, dtype=np.float64)
        self.check(self.lib.emc_set_background_profile(self.h, b.ctypes.data, len(b)))

    def set_efield(self, efield) -> None:
        ef = np.ascontiguousarray(efield, dtype=np.float64)
        assert ef.shape == (3,)
        self.check(self.lib.emc_set_efield(self.h, ef.ctypes.data))

    def set_field_groups(self, efields, group_size: int) -> None:
        ef = np.ascontiguousarray(efields, dtype=np.float64)
        assert ef.ndim == 2 and ef.shape[1] == 3
        self.check(self.lib.emc_set_field_groups(self.h, ef.ctypes.data, ef.shape[0], int(group_size)))


def _obs_from_tensor(t) -> dict:
    return _obs_from_raw(t.cpu().numpy())


def _obs_from_raw(raw) -> dict:
    raw = np.ascontiguousarray(raw)
    d = raw[:5].view(np.float64)
    c = raw[5:]
    return dict(sum_z=float(d[0]), sum_vz=float(d[1]), sum_e=float(d[2]), sum_vz2=float(d[3]),
                sum_e2=float(d[4]), n_valley=[int(c[0]), int(c[1]), int(c[2])], n_fault=int(c[3]),
                n_segments=int(c[4]), n_events=int(c[5]))


class Ensemble:
    """Particle state: one (7, n) float64 tensor (rows kx ky kz x y z dtau, each row a
    contiguous n-vector: structure of arrays) plus an int32 valley vector."""

    ROWS = ("kx", "ky", "kz", "x", "y", "z", "dtau")

    def __init__(self, ctx: Context, n: int, pid_offset: int = 0, capacity: int | None = None):
        """``capacity`` >= n reserves room for carriers injected later (``inject_photoex``); the
        ``state`` / ``valley`` attributes are always views of the first ``n`` particles."""
        torch = _torch()
        self.ctx, self.n, self.pid_offset = ctx, int(n), int(pid_offset)
        self.capacity = max(int(capacity or 0), self.n)
        self._store = torch.zeros((7, self.capacity), dtype=torch.float64, device=ctx.device)
        self._vstore = torch.zeros(self.capacity, dtype=torch.int32, device=ctx.device)
        self._n_candidates = 0                       # photo-excitation candidates drawn so far (Philox ids)
        self._views()
        self.obs = torch.zeros(11, dtype=torch.int64, device=ctx.device)    # one EmcObs (88 B)
        self._obs_cache = None

    def _views(self):
        self.state = self._store[:, :self.n]
        self.valley = self._vstore[:self.n]

    def _ptrs(self, first: int = 0):
        base, stride = self._store.data_ptr(), self.capacity * 8
        return [base + r * stride + first * 8 for r in range(7)] + [self._vstore.data_ptr() + first * 4]

    def reserve(self, capacity: int):
        """Grow the buffers (device-to-device copy) so that ``capacity`` particles fit."""
        torch = _torch()
        if capacity <= self.capacity:
            return self
        store = torch.zeros((7, capacity), dtype=torch.float64, device=self.ctx.device)
        vstore = torch.zeros(capacity, dtype=torch.int32, device=self.ctx.device)
        store[:, :self.n].copy_(self._store[:, :self.n])
        vstore[:self.n].copy_(self._vstore[:self.n])
        self._store, self._vstore, self.capacity = store, vstore, int(capacity)
        self._views()
        return self

    def inject_photoex(self, n_candidates: int, seed: int, energy: float | None = None,
                       z_scale: float | None = None, xy_std: float | None = None,
                       gamma_first: float = 1E15) -> int:
        """Append photo-excited carriers (init_.init_photoex, init_.py:244-280; main_.py:345-358) on
        the device: ``n_candidates`` are drawn, the ones beyond the slab are dropped, the survivors
        get valley 0 and a first free flight from ``free_flight(1E15)``.  Defaults: the reference's
        laser (``1240/laser_ex - EG`` eV, depth scale ``min(1/alpha, slab)``, ``laser_std``).
        Returns the number of carriers added (the reference adds ``n_candidates`` to its count and
        breaks when a candidate was dropped: main_.py:347-358)."""
        from . import init_
        c = self.ctx
        n_candidates = int(n_candidates)
        if n_candidates <= 0:
            return 0
        G, las = init_.Device.Materials[0], init_.laser
        energy = (1240 / las.laser_ex - G.EG) if energy is None else energy
        z_scale = min(1 / G.alpha, float(c.params.tot_dim[2])) if z_scale is None else z_scale
        xy_std = las.laser_std if xy_std is None else xy_std
        self.reserve(self.n + n_candidates)
        got = C.c_int64(0)
        c.check(c.lib.emc_init_photoex(c.h, *self._ptrs(self.n), self.capacity - self.n, n_candidates,
                                       float(energy), float(z_scale), float(xy_std), float(gamma_first),
                                       seed, self._n_candidates, C.byref(got), c.stream))
        self._n_candidates += n_candidates
        self.n += int(got.value)
        self._views()
        self._obs_cache = None
        return int(got.value)

    def row(self, name: str):
        return self.state[self.ROWS.index(name)]

    # -- initialisation ---------------------------------------------------------------
    def init(self, seed: int):
        """init_.init_coords + valley 0 + first free flights, on the device (init_.py:194-234,
        main_.py:53,337)."""
        c = self.ctx
        c.check(c.lib.emc_init_ensemble(c.h, *self._ptrs(), self.n, seed, self.pid_offset, c.stream))
        return self

    def load(self, coords, valley, dtau):
        """From the reference's host arrays: coords (N,6), valley (N,), dtau (N,)."""
        torch = _torch()
        c = self.ctx
        co = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).to(c.device)
        assert co.shape == (self.n, 6)
        dtau = np.ascontiguousarray(dtau, dtype=np.float64)
        if dtau.shape != (self.n,):              # None would broadcast NaN into every free flight
            raise ValueError("dtau must have shape (%d,), got %s" % (self.n, dtau.shape))
        p = self._ptrs()
        c.check(c.lib.emc_aos_to_soa(c.h, co.data_ptr(), self.n, *p[:6], c.stream))
        self.state[6].copy_(torch.from_numpy(np.ascontiguousarray(dtau, dtype=np.float64)))
        self.valley.copy_(torch.from_numpy(np.ascontiguousarray(valley).astype(np.int32)))
        torch.cuda.current_stream(c.device).synchronize()      # `co` may be freed after this
        return self

    def coords(self):
        """Back to the reference's layout: ((N,6) coords, valley, dtau) numpy arrays."""
        torch = _torch()
        c = self.ctx
        out = torch.empty((self.n, 6), dtype=torch.float64, device=c.device)
        p = self._ptrs()
        c.check(c.lib.emc_soa_to_aos(c.h, *p[:6], self.n, out.data_ptr(), c.stream))
        return out.cpu().numpy(), self.valley.cpu().numpy(), self.state[6].cpu().numpy()

    # -- the timestep -----------------------------------------------------------------
    def step(self, step: int, seed: int, field_mode: int = abi.FIELD_CONSTANT, mesh: "Mesh | None" = None,
             observe: bool = True, deposit: bool = False, obs_out=None):
        """One timestep of main_.py:360-397 for the whole ensemble (asynchronous).  ``obs_out``: an
        11-element int64 device tensor that receives the EmcObs instead of ``self.obs``."""
        c = self.ctx
        obs_t = self.obs if obs_out is None else obs_out
        if obs_out is None:
            self._obs_cache = None
        ex = ey = ez = None
        if field_mode == abi.FIELD_MESH:
            ex, ey, ez = mesh.ex.data_ptr(), mesh.ey.data_ptr(), mesh.ez.data_ptr()
        elif field_mode == abi.FIELD_MESH_PACKED:
            ex = mesh.e4.data_ptr()
        elif field_mode == abi.FIELD_MESH_XZ:
            ex = mesh.e2.data_ptr()
        c.check(c.lib.emc_flight_scatter(
            c.h, *self._ptrs(), self.n, field_mode, ex, ey, ez, seed, self.pid_offset, step,
            obs_t.data_ptr() if observe else None,
            mesh.counts.data_ptr() if deposit else None, c.stream))

    def run_steps(self, first_step: int, n_steps: int, seed: int, step_fn=None, **step_kw):
        """``n_steps`` timesteps with the per-step observables of main_.py:394-403, no host round
        trip inside the loop.  Uses the before-step observable phase (``emc_set_obs_phase``): launch
        c reduces the state that step c-1 left behind while it loads it (all lanes busy), one
        stand-alone ``emc_observables`` reads the final state, and the rows are shifted back so
        that ``rows[c]`` describes the ensemble AFTER step ``first_step + c`` exactly as with
        ``step(); last_obs()``.  ``step_fn(c, obs_out)`` replaces the plain ``step`` (coupled loops)."""
        torch = _torch()
        c = self.ctx
        hist = torch.zeros((n_steps + 1, 11), dtype=torch.int64, device=c.device)
        was = c.obs_before
        c.set_obs_phase(True)
        try:
            if step_fn is None and not step_kw.get("deposit", False) and step_kw.get("observe", True):
                # the whole loop in one C call (no Python / ctypes overhead per timestep)
                field_mode, mesh = step_kw.get("field_mode", abi.FIELD_CONSTANT), step_kw.get("mesh")
                ex = ey = ez = None
                if field_mode == abi.FIELD_MESH:
                    ex, ey, ez = mesh.ex.data_ptr(), mesh.ey.data_ptr(), mesh.ez.data_ptr()
                elif field_mode == abi.FIELD_MESH_PACKED:
                    ex = mesh.e4.data_ptr()
                elif field_mode == abi.FIELD_MESH_XZ:
                    ex = mesh.e2.data_ptr()
                self._obs_cache = None
                c.check(c.lib.emc_flight_scatter_steps(
                    c.h, *self._ptrs(), self.n, field_mode, ex, ey, ez, seed, self.pid_offset,
                    first_step, n_steps, hist.data_ptr(), c.stream))
            else:
                for k in range(n_steps):
                    if step_fn is not None:
                        step_fn(first_step + k, hist[k])
                    else:
                        self.step(first_step + k, seed, obs_out=hist[k], **step_kw)
        finally:
            c.set_obs_phase(was)
        p = self._ptrs()
        final = abi.EmcObs()
        c.check(c.lib.emc_observables(c.h, p[0], p[1], p[2], p[5], p[7], self.n, C.byref(final), c.stream))
        raw = hist.cpu().numpy()
        rows = []
        for k in range(n_steps):
            work = _obs_from_raw(raw[k])                      # n_segments / n_events of step k
            state = _obs_from_raw(raw[k + 1]) if k + 1 < n_steps else final.as_dict()
            state["n_segments"], state["n_events"] = work["n_segments"], work["n_events"]
            rows.append(state)
        self._obs_cache = dict(rows[-1]) if rows else None      # what last_obs() reports after the loop
        return rows

    def advance(self, first_step: int, n_steps: int, seed: int, field_mode: int = abi.FIELD_CONSTANT,
                mesh: "Mesh | None" = None, check: bool = False):
        """``n_steps`` timesteps WITHOUT observables (burn-in phases, sweeps that sample every m-th
        timestep): ``emc_flight_scatter_steps`` with no observable rows, which runs all of them in one
        persistent launch -- particles stay in the warps' queues from timestep to timestep -- whenever
        the lean conditions hold (constant / group field along z, one layer, default libm), else one
        launch per timestep.  The final state is the same bit for bit either way.  ``check=True``
        synchronises and verifies the fused kernel's self-check flag."""
        c = self.ctx
        ex = ey = ez = None
        if field_mode == abi.FIELD_MESH:
            ex, ey, ez = mesh.ex.data_ptr(), mesh.ey.data_ptr(), mesh.ez.data_ptr()
        elif field_mode == abi.FIELD_MESH_PACKED:
            ex = mesh.e4.data_ptr()
        elif field_mode == abi.FIELD_MESH_XZ:
            ex = mesh.e2.data_ptr()
        self._obs_cache = None
        c.check(c.lib.emc_flight_scatter_steps(c.h, *self._ptrs(), self.n, field_mode, ex, ey, ez, seed,
                                               self.pid_offset, first_step, n_steps, None, c.stream))
        if check:
            c.check(c.lib.emc_flight_steps_check(c.h, c.stream))
        return self

    def step_replay(self, uniforms, cursor, field_mode: int = abi.FIELD_CONSTANT, mesh: "Mesh | None" = None,
                    observe: bool = True, deposit: bool = False):
        """Replay mode: uniforms (n, L) float64 device tensor, cursor (n,) int32 device tensor."""
        c = self.ctx
        self._obs_cache = None
        assert uniforms.is_contiguous() and uniforms.shape[0] == self.n
        ex = ey = ez = None
        if field_mode == abi.FIELD_MESH:
            ex, ey, ez = mesh.ex.data_ptr(), mesh.ey.data_ptr(), mesh.ez.data_ptr()
        elif field_mode == abi.FIELD_MESH_PACKED:
            ex = mesh.e4.data_ptr()
        elif field_mode == abi.FIELD_MESH_XZ:
            ex = mesh.e2.data_ptr()
        c.check(c.lib.emc_flight_scatter_replay(
            c.h, *self._ptrs(), self.n, field_mode, ex, ey, ez, self.pid_offset,
            uniforms.data_ptr(), uniforms.shape[1], cursor.data_ptr(),
            self.obs.data_ptr() if observe else None,
            mesh.counts.data_ptr() if deposit else None, c.stream))

    def last_obs(self) -> dict:
        """Observables written by the last ``step(observe=True)`` (synchronises), or the last row
        of the last ``run_steps``."""
        if getattr(self, "_obs_cache", None) is not None:
            return dict(self._obs_cache)
        return _obs_from_tensor(self.obs)

    def observables(self) -> dict:
        """main_.py:394-403 on the current state (synchronises)."""
        c = self.ctx
        o = abi.EmcObs()
        p = self._ptrs()
        c.check(c.lib.emc_observables(c.h, p[0], p[1], p[2], p[5], p[7], self.n, C.byref(o), c.stream))
        return o.as_dict()


def summarize(obs: dict) -> dict:
    """EmcObs sums -> the per-step row of main_.py:398-403 (<z> nm, <v_z>, <E>, valley counts)."""
    n_ok = sum(obs["n_valley"])
    n = max(n_ok, 1)
    mz, mv, me = obs["sum_z"] / n, obs["sum_vz"] / n, obs["sum_e"] / n
    var_v = max(obs["sum_vz2"] / n - mv * mv, 0.0)
    var_e = max(obs["sum_e2"] / n - me * me, 0.0)
    return dict(z_nm=mz * 1e9, vz=mv, e=me, vz_se=(var_v / n) ** 0.5, e_se=(var_e / n) ** 0.5,
                n_valley=obs["n_valley"], n_fault=obs["n_fault"], n=n_ok,
                n_segments=obs["n_segments"], n_events=obs["n_events"])


class Mesh:
    """Device-resident charge mesh, rho, potential and field of one Context's geometry."""

    def __init__(self, ctx: Context, scheme: int = abi.DEPOSIT_NGP, p2p_group=None):
        """``p2p_group``: a torch.distributed process group (or ``True`` for the default group)
        whose ranks share one node: the charge mesh is then placed in NVLink symmetric memory and
        ``allreduce`` runs this package's own one-kernel reduce-scatter + all-gather over peer
        pointers (``emc_mesh_allreduce_p2p``) instead of an NCCL all-reduce."""
        torch = _torch()
        self.ctx, self.scheme = ctx, scheme
        nx, ny, nz = ctx.seg
        cells = nx * ny * nz
        self.p2p = None
        self.p2p_error = None
        if p2p_group is not None and p2p_group is not False:
            try:
                self._setup_p2p(cells, p2p_group)
            except Exception as exc:      # noqa: BLE001  (no symmetric memory on this system: NCCL path)
                self.p2p, self.p2p_error = None, repr(exc)
        if self.p2p is None:
            self.counts = torch.zeros(cells, dtype=torch.int64, device=ctx.device)
            self.total = self.counts                       # after allreduce(): summed in place
        pad = ctx.padded_shape
        if self.p2p is not None and "rho" in self.p2p:
            self.rho_pad = self.p2p["rho"]                 # symmetric memory: peers store their slice of rho here
        else:
            self.rho_pad = torch.zeros(pad, dtype=torch.float64, device=ctx.device)
        self.u_pad = torch.zeros(pad, dtype=torch.float64, device=ctx.device)
        self.ex = torch.zeros(pad, dtype=torch.float64, device=ctx.device)
        self.ey = torch.zeros(pad, dtype=torch.float64, device=ctx.device)
        self.ez = torch.zeros(pad, dtype=torch.float64, device=ctx.device)
        self.e4 = torch.zeros(pad + (4,), dtype=torch.float64, device=ctx.device)   # (ex, ey, ez, 0) in V/cm
        # ny = 1: (ex, ez) per interior cell, unpadded (FIELD_MESH_XZ)
        self.e2 = torch.zeros((nx, nz, 2), dtype=torch.float64, device=ctx.device) if ny == 1 else None
        self._rho_halo_done = self._e4_halo_done = False
        self.set_initial_guess(None)

    def set_initial_guess(self, u0):
        """poisson_.py:47-52: u = guess (default zeros; the reference draws np.random.random),
        zero Dirichlet halo, z-min face = surf_val."""
        torch = _torch()
        self.u_pad.zero_()
        if u0 is not None:
            self.u_pad[1:-1, 1:-1, 1:-1] = torch.as_tensor(np.asarray(u0, dtype=np.float64)).to(self.ctx.device)
        self.u_pad[:, :, 0] = self.ctx.params.surf_val
        return self

    def zero_counts(self):
        self.counts.zero_()

    def deposit(self, ens: Ensemble):
        c = self.ctx
        p = ens._ptrs()
        c.check(c.lib.emc_deposit(c.h, p[3], p[4], p[5], ens.n, self.counts.data_ptr(), self.scheme, c.stream))

    def _setup_p2p(self, cells: int, group):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        torch = _torch()
        if group is True:
            group = dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2:
            raise RuntimeError("single rank: nothing to reduce")
        if world > 16:
            raise RuntimeError("more than EMC_MAX_PEERS ranks")
        even = cells + (cells & 1)
        buf = symm.empty((2, even), dtype=torch.int64, device=self.ctx.device)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        base = [int(p) for p in hdl.buffer_ptrs]
        self.p2p = dict(hdl=hdl, world=world, rank=rank, cells=cells, buf=buf,
                        in_ptrs=(C.c_void_p * world)(*base),
                        out_ptrs=(C.c_void_p * world)(*[b + even * 8 for b in base]))
        self.counts = buf[0, :cells]          # this rank's deposits
        self.total = buf[1, :cells]           # sum over the ranks, written by all peers
        if os.environ.get("EMC_P2P_RHO", "1") != "0":
            # rho in symmetric memory as well: allreduce_to_rho() lets every rank compute rho for its slice only
            pad = self.ctx.padded_shape
            rbuf = symm.empty(pad[0] * pad[1] * pad[2], dtype=torch.float64, device=self.ctx.device)
            rbuf.zero_()
            rhdl = symm.rendezvous(rbuf, group)
            self.p2p.update(rho_hdl=rhdl, rho_buf=rbuf, rho=rbuf.view(pad),
                            rho_ptrs=(C.c_void_p * world)(*[int(p) for p in rhdl.buffer_ptrs]))

    def allreduce(self, group=None):
        """Sum the integer charge mesh over the ranks; the result is ``self.total``.  NVLink peer
        memory kernel when the mesh lives in symmetric memory, else NCCL (gloo in CPU tests)."""
        if self.p2p is not None:
            c, p = self.ctx, self.p2p
            p["hdl"].barrier(channel=0)                      # every rank's deposits are complete
            c.check(c.lib.emc_mesh_allreduce_p2p(c.h, p["in_ptrs"], p["out_ptrs"], p["world"], p["rank"],
                                                 p["cells"], c.stream))
            p["hdl"].barrier(channel=1)                      # every peer has written its slice of my totals
            return
        from . import parallel
        parallel.allreduce_counts(self.counts, group)

    def allreduce_to_rho(self, group=None):
        """``allreduce`` + ``to_rho``.  With the mesh AND rho in NVLink symmetric memory both happen in one
        kernel (``emc_mesh_allreduce_rho_p2p``): every rank accumulates rho for its 1/world slice of the cells
        only and stores it into all ranks' rho arrays -- bit-identical to ``to_rho`` of the totals."""
        if self.p2p is not None and "rho_ptrs" in self.p2p:
            c, p = self.ctx, self.p2p
            p["hdl"].barrier(channel=0)                      # every rank's deposits are complete (and rho is free)
            c.check(c.lib.emc_mesh_allreduce_rho_p2p(c.h, p["in_ptrs"], p["out_ptrs"], p["rho_ptrs"], p["world"],
                                                     p["rank"], p["cells"], self.scheme, c.stream))
            p["hdl"].barrier(channel=1)                      # every peer has written its slice of my totals and rho
            self._rho_halo_done = True                       # the halo of the symmetric buffer was zeroed at set-up
            return
        self.allreduce(group)
        self.to_rho()

    def to_rho(self):
        """rho_pad from the (all-reduced) charge mesh.  The first call writes the zero halo too, later
        ones only the interior cells (the halo of this Mesh's own buffer stays zero)."""
        c = self.ctx
        fn = c.lib.emc_mesh_to_rho_interior if self._rho_halo_done else c.lib.emc_mesh_to_rho
        c.check(fn(c.h, self.total.data_ptr(), self.scheme, self.rho_pad.data_ptr(), c.stream))
        self._rho_halo_done = True

    def solve(self, omega0: float = 1.5, tol: float = 1e-6, max_sweeps: int = 10000,
              order: int = abi.SOR_RED_BLACK, history: bool = False):
        """poisson_.py:63-84.  Returns (sweeps, err[, err_hist]); synchronises."""
        c = self.ctx
        n, err = C.c_int32(0), C.c_double(0.0)
        hist = np.zeros(max_sweeps, dtype=np.float64) if history else None
        c.check(c.lib.emc_poisson_sor(c.h, self.u_pad.data_ptr(), self.rho_pad.data_ptr(), omega0, tol,
                                      max_sweeps, order, C.byref(n), C.byref(err),
                                      hist.ctypes.data if history else None, c.stream))
        if history:
            return n.value, err.value, hist[:n.value].copy()
        return n.value, err.value

    def solve_async(self, omega0: float = 1.5, tol: float = 1e-6, max_sweeps: int = 10000,
                    order: int = abi.SOR_RED_BLACK):
        """Enqueue the solve without a host round trip (the time loop); ``last_solve`` reports."""
        c = self.ctx
        c.check(c.lib.emc_poisson_sor(c.h, self.u_pad.data_ptr(), self.rho_pad.data_ptr(), omega0, tol,
                                      max_sweeps, order, None, None, None, c.stream))

    def solve_mg(self, tol_residual: float = 1e-8, max_cycles: int = 30, nu1: int = 2, nu2: int = 2,
                 history: bool = False, sync: bool = True):
        """Geometric multigrid on the same discrete problem (``emc_poisson_mg``): V(nu1, nu2) cycles
        until the RMS of the TRUE residual (volts) <= ``tol_residual``.  Returns (cycles, rms_residual
        [, per-cycle history]); ``sync=False`` only enqueues the solve (``last_solve`` reports)."""
        c = self.ctx
        if not sync:
            c.check(c.lib.emc_poisson_mg(c.h, self.u_pad.data_ptr(), self.rho_pad.data_ptr(), tol_residual, max_cycles,
                                         nu1, nu2, None, None, None, c.stream))
            return None
        n, res = C.c_int32(0), C.c_double(0.0)
        hist = np.zeros(max_cycles, dtype=np.float64) if history else None
        c.check(c.lib.emc_poisson_mg(c.h, self.u_pad.data_ptr(), self.rho_pad.data_ptr(), tol_residual, max_cycles,
                                     nu1, nu2, C.byref(n), C.byref(res), hist.ctypes.data if history else None, c.stream))
        if history:
            return n.value, res.value, hist[:n.value].copy()
        return n.value, res.value

    def residual(self, out=None):
        """(RMS, max) of the true residual ``dl0^2 rho/eps - A u`` (volts) of the current potential
        (``emc_poisson_residual``): what the reference's stopping rule does not bound.  With ``out`` (a
        2-element float64 device tensor) the numbers are written there in stream order, no sync."""
        c = self.ctx
        if out is not None:
            c.check(c.lib.emc_poisson_residual(c.h, self.u_pad.data_ptr(), self.rho_pad.data_ptr(), None, None,
                                               out.data_ptr(), c.stream))
            return None
        rms, mx = C.c_double(0.0), C.c_double(0.0)
        c.check(c.lib.emc_poisson_residual(c.h, self.u_pad.data_ptr(), self.rho_pad.data_ptr(), C.byref(rms),
                                           C.byref(mx), None, c.stream))
        return rms.value, mx.value

    def last_solve(self):
        """(sweeps, err) of the most recent solve; raises EmcError(ERR_DIVERGED) like poisson_.py:78-80."""
        c = self.ctx
        n, err = C.c_int32(0), C.c_double(0.0)
        c.check(c.lib.emc_poisson_last(c.h, C.byref(n), C.byref(err), c.stream))
        return n.value, err.value

    def field(self):
        c = self.ctx
        c.check(c.lib.emc_field(c.h, self.u_pad.data_ptr(), self.ex.data_ptr(), self.ey.data_ptr(),
                                self.ez.data_ptr(), c.stream))

    def field_xz(self):
        """ny = 1 meshes: the field as one 16-byte (ex, ez) record per interior cell (FIELD_MESH_XZ)."""
        c = self.ctx
        c.check(c.lib.emc_field_xz(c.h, self.u_pad.data_ptr(), self.e2.data_ptr(), c.stream))

    def field_packed(self):
        """The same field as one 32-byte record per cell (the flight kernel's FIELD_MESH_PACKED)."""
        c = self.ctx
        fn = c.lib.emc_field_packed_interior if self._e4_halo_done else c.lib.emc_field_packed
        c.check(fn(c.h, self.u_pad.data_ptr(), self.e4.data_ptr(), c.stream))
        self._e4_halo_done = True


class CoupledStepper:
    """Self-consistent timestep (BASELINE configs 3/4): flight in the mesh field with the
    end-of-step NGP deposit fused into the same kernel -> integer all-reduce of the charge
    mesh over the ranks -> rho -> red-black SOR (warm-started from the previous potential) ->
    field.  Everything stays on the device and in stream order; no host synchronisation."""

    def __init__(self, ctx: Context, ens: Ensemble, mesh: Mesh | None = None, seed: int = 12345,
                 tol: float = 1e-6, max_sweeps: int = 10000, order: int = abi.SOR_RED_BLACK,
                 group=None, fuse_deposit: bool = False, packed_field: bool = True, xz_field: bool = True):
        # fuse_deposit: the flight kernel can deposit its end-of-step positions itself (saves
        # re-reading 24 B/particle) but the assignment then runs on half-filled warps inside the
        # register-heavy kernel: measured 0.45 ms fused vs 0.19 ms for the separate deposit kernel
        # at 1.25e7 particles on a 1024x1024 mesh (profiles/coupled_probe.py), so the default is off.
        self.fuse_deposit = fuse_deposit
        # packed_field: gather one (ex, ey, ez, 0) record per flight segment instead of three
        # separate arrays; identical values (Mesh.field() still gives the reference's ef_x/y/z)
        self.packed_field = packed_field
        self.field_mode = abi.FIELD_MESH_PACKED if packed_field else abi.FIELD_MESH
        if packed_field and ctx.seg[1] == 1 and xz_field:
            # ny = 1: half-size (ex, ez) records without halo (identical trajectories: ey is -0 there)
            self.field_mode = abi.FIELD_MESH_XZ
        self.ctx, self.ens, self.seed = ctx, ens, seed
        if mesh is None:
            # multi-rank runs on large meshes sum the charge with the NVLink peer-memory kernel
            # (measured 0.050 vs 0.069 ms NCCL for 1024x1024 at 8 GPUs; on small meshes the two
            # barriers cost more than NCCL's latency).  EMC_P2P=0 forces NCCL.
            import os
            from . import parallel
            _rank, world = parallel.world_info(group)
            cells = ctx.seg[0] * ctx.seg[1] * ctx.seg[2]
            want = world > 1 and cells >= 65536 and os.environ.get("EMC_P2P", "1") != "0"
            mesh = Mesh(ctx, p2p_group=(group if group is not None else True) if want else None)
        self.mesh = mesh
        self.tol, self.max_sweeps, self.order, self.group = tol, max_sweeps, order, group
        self.n_steps = 0

    def prime(self, max_sweeps: int | None = None):
        """Field of the initial ensemble (the reference would call poisson() before the loop).
        The cold solve may need far more sweeps than the warm-started ones of the time loop
        (thousands on a 1024 x 1024 mesh): ``max_sweeps`` overrides the per-step cap here."""
        m = self.mesh
        m.zero_counts()
        m.deposit(self.ens)
        self._solve(max_sweeps)
        return self

    def _solve(self, max_sweeps: int | None = None):
        m = self.mesh
        m.allreduce_to_rho(self.group)
        m.solve_async(tol=self.tol, max_sweeps=max_sweeps or self.max_sweeps, order=self.order)
        self.update_field()

    def update_field(self):
        m = self.mesh
        if self.field_mode == abi.FIELD_MESH_XZ:
            m.field_xz()
        elif self.packed_field:
            m.field_packed()
        else:
            m.field()

    def step(self, observe: bool = True, obs_out=None):
        m = self.mesh
        m.zero_counts()
        self.ens.step(self.n_steps, self.seed, field_mode=self.field_mode, mesh=m, observe=observe,
                      deposit=self.fuse_deposit, obs_out=obs_out)
        if not self.fuse_deposit:
            m.deposit(self.ens)
        self._solve()
        self.n_steps += 1

    def run_steps(self, n_steps: int, check_every: int = 0):
        """``n_steps`` coupled timesteps; rows[c] = observables after step c (``Ensemble.run_steps``:
        reduced in the load phase of the next launch, no host round trip inside the loop).

        The Poisson solves are only enqueued, so a diverged solve (the reference's ``'Too large du'``
        ValueError, poisson_.py:78-80 / main_.py:303) would go unnoticed: the solver keeps a STICKY
        divergence flag on the device, which is read back (one synchronisation) every
        ``check_every`` steps if that is > 0 and always at the end of the loop; ``EmcError``
        (``ERR_DIVERGED``) is raised then."""
        def one(c, out):
            self.step(observe=True, obs_out=out)
            if check_every > 0 and (c + 1 - first) % check_every == 0:
                self.check_solver()
        first = self.n_steps
        rows = self.ens.run_steps(self.n_steps, n_steps, self.seed, step_fn=one)
        self.check_solver()
        return rows

    def check_solver(self):
        """(sweeps, err) of the latest solve; raises ``EmcError(ERR_DIVERGED)`` if ANY solve since
        the last check diverged (synchronises the stream)."""
        return self.mesh.last_solve()


def chunk_schedule(n: int, chunk: int, ramp_from: int | None = None) -> list:
    """[(first particle, count)] of ``HostStepper.run``'s chunks.  Without ``ramp_from``: equal chunks.
    With it: ramp_from, 2 ramp_from, 4 ramp_from ... up to ``chunk``, chunks of ``chunk``, and the same
    ramp down again (the remainder goes before the ramp down), e.g. 2.4e7 particles, chunk 2^23,
    ramp_from 2^20: 1, 2, 4, 8, 4.9, 2, 1 Mi particles."""
    n, chunk = int(n), max(int(chunk), 1)
    if n <= 0:
        return []
    sizes = []
    if ramp_from is None or ramp_from <= 0 or ramp_from >= chunk:
        sizes = [chunk] * (n // chunk) + ([n % chunk] if n % chunk else [])
    else:
        up, c = [], int(ramp_from)
        while c < chunk:
            up.append(c)
            c *= 2
        down = up[::-1][1:] if len(up) > 1 else up[::-1]       # the top step of the ramp only once
        while up and sum(up) + sum(down) > n:                  # small ensembles: shorten both ramps from the top
            if len(up) >= len(down) + 1 or not down:
                up.pop()
            else:
                down.pop(0)
        rest = n - sum(up) - sum(down)
        mid = [chunk] * (rest // chunk) + ([rest % chunk] if rest % chunk else [])
        sizes = up + mid + down
    out, lo = [], 0
    for m in sizes:
        out.append((lo, m))
        lo += m
    assert lo == n and all(0 < m <= chunk for _, m in out)
    return out


class HostStepper:
    """The timestep for callers that keep the reference's HOST arrays (main_.py:51-53,337):
    coords (N,6) float64, valley (N,) int32, dtau (N,) float64, all in pinned memory.

    Each call moves the whole state host -> device -> host, as a drop-in replacement of the
    Python particle loop has to when the caller owns numpy arrays.  The ensemble is cut into
    chunks and the three stages (H2D, kernels, D2H) of consecutive chunks overlap on three
    CUDA streams, so the call runs at PCIe speed in both directions at once."""

    def __init__(self, ctx: Context, n: int, chunk: int = 1 << 21, pid_offset: int = 0, ramp_from: int | None = None):
        """``chunk``: particles per chunk (and per device buffer).  ``ramp_from``: ``run`` then cuts the
        ensemble into GRADUATED chunks (``chunk_schedule``): the first and the last copy of a loop are the
        part of the PCIe traffic that no kernel overlaps, so the loop starts and ends with chunks of
        ``ramp_from`` particles and doubles up to ``chunk`` in between, where long launches amortise the
        fixed cost of a launch (~20 us of a 0.18-ms launch at 2^22 particles).  Results do not depend on
        the chunking (Philox counters are global particle ids)."""
        torch = _torch()
        self.ctx, self.n, self.pid_offset = ctx, int(n), int(pid_offset)
        self.chunk = int(min(chunk, max(self.n, 1)))
        self.ramp_from = None if ramp_from is None else int(ramp_from)
        self.schedule = chunk_schedule(self.n, self.chunk, self.ramp_from)
        # pinned host arrays the caller reads / writes (numpy views)
        self.coords_t = torch.zeros((self.n, 6), dtype=torch.float64).pin_memory()
        self.valley_t = torch.zeros(self.n, dtype=torch.int32).pin_memory()
        self.dtau_t = torch.zeros(self.n, dtype=torch.float64).pin_memory()
        self.coords, self.valley, self.dtau = (self.coords_t.numpy(), self.valley_t.numpy(),
                                               self.dtau_t.numpy())
        self.obs_host = torch.zeros((0, 11), dtype=torch.int64)
        nbuf = 3
        dev = ctx.device
        self.bufs = []
        for _ in range(nbuf):
            self.bufs.append(dict(
                aos=torch.empty((self.chunk, 6), dtype=torch.float64, device=dev),
                soa=torch.empty((7, self.chunk), dtype=torch.float64, device=dev),
                valley=torch.empty(self.chunk, dtype=torch.int32, device=dev),
                obs=torch.zeros(11, dtype=torch.int64, device=dev),
                done=torch.cuda.Event(), free=torch.cuda.Event()))
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(dev), torch.cuda.Stream(dev),
                                             torch.cuda.Stream(dev))
        self.n_chunks = max((self.n + self.chunk - 1) // self.chunk, len(self.schedule))
        self.obs_pinned = torch.zeros((self.n_chunks, 11), dtype=torch.int64).pin_memory()

    @property
    def h2d_bytes(self) -> int:
        return self.n * (48 + 4 + 8)

    @property
    def d2h_bytes(self) -> int:
        return self.n * (48 + 4 + 8) + self.n_chunks * 88

    def run(self, first_step: int, n_steps: int, seed: int, field_mode: int = abi.FIELD_CONSTANT) -> list:
        """The WHOLE time loop of main_.py:339-403 on the pinned host arrays, in place: the reference
        needs ``coords`` on the host only before and after the loop and seven scalars per timestep,
        so every chunk crosses PCIe once in each direction -- H2D, ``n_steps`` launches that keep the
        chunk in HBM (``emc_flight_scatter_steps``, per-step EmcObs rows reduced on the device in the
        load phase of the next launch + one enqueued reduction of the final state), D2H -- and the
        three stages of consecutive chunks overlap on three streams.  Returns one observable dict
        per timestep (``rows[c]`` = the ensemble after step ``first_step + c``, like
        ``Ensemble.run_steps``).  Bytes moved per timestep: ``run_h2d_bytes(n_steps) / n_steps``."""
        torch = _torch()
        c, lib = self.ctx, self.ctx.lib
        n_steps = int(n_steps)
        if n_steps < 1:
            return []
        rows_dev = getattr(self, "_rows_dev", None)
        if rows_dev is None or rows_dev.shape[1] < n_steps + 1:
            self._rows_dev = rows_dev = torch.zeros((len(self.bufs), n_steps + 1, 11), dtype=torch.int64, device=c.device)
            self._rows_pinned = torch.zeros((self.n_chunks, n_steps + 1, 11), dtype=torch.int64).pin_memory()
        rows_pinned = self._rows_pinned
        cur = torch.cuda.current_stream(c.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        was = c.obs_before
        c.set_obs_phase(True)
        try:
            for ci, (lo, m) in enumerate(self.schedule):
                bi = ci % len(self.bufs)
                b = self.bufs[bi]
                with torch.cuda.stream(self.s_in):
                    if ci >= len(self.bufs):
                        self.s_in.wait_event(b["free"])
                    b["aos"][:m].copy_(self.coords_t[lo:lo + m], non_blocking=True)
                    b["soa"][6, :m].copy_(self.dtau_t[lo:lo + m], non_blocking=True)
                    b["valley"][:m].copy_(self.valley_t[lo:lo + m], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(ev_in)
                    st = C.c_void_p(self.s_run.cuda_stream)
                    base, stride = b["soa"].data_ptr(), self.chunk * 8
                    ptrs = [base + r * stride for r in range(7)] + [b["valley"].data_ptr()]
                    rows = rows_dev[bi]
                    c.check(lib.emc_aos_to_soa(c.h, b["aos"].data_ptr(), m, *ptrs[:6], st))
                    c.check(lib.emc_flight_scatter_steps(c.h, *ptrs, m, field_mode, None, None, None, seed,
                                                         self.pid_offset + lo, first_step, n_steps,
                                                         rows.data_ptr(), st))
                    c.check(lib.emc_observables_async(c.h, ptrs[0], ptrs[1], ptrs[2], ptrs[5], ptrs[7], m,
                                                      rows.data_ptr() + n_steps * 88, st))
                    c.check(lib.emc_soa_to_aos(c.h, *ptrs[:6], m, b["aos"].data_ptr(), st))
                    b["done"].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(b["done"])
                    self.coords_t[lo:lo + m].copy_(b["aos"][:m], non_blocking=True)
                    self.dtau_t[lo:lo + m].copy_(b["soa"][6, :m], non_blocking=True)
                    self.valley_t[lo:lo + m].copy_(b["valley"][:m], non_blocking=True)
                    rows_pinned[ci, :n_steps + 1].copy_(rows[:n_steps + 1], non_blocking=True)
                    b["free"].record(self.s_out)
        finally:
            c.set_obs_phase(was)
        self.s_out.synchronize()
        cur.wait_stream(self.s_out)
        raw = rows_pinned.numpy()[:len(self.schedule), :n_steps + 1]
        d = raw[:, :, :5].copy().view(np.float64).sum(axis=0)         # (n_steps + 1, 5): sums over the chunks
        k = raw[:, :, 5:].sum(axis=0)
        out = []
        for s in range(n_steps):
            # row s holds the WORK of step s (segments, events) and the state BEFORE it; the state after
            # step s is what row s + 1 (or the final reduction) saw
            out.append(dict(sum_z=float(d[s + 1, 0]), sum_vz=float(d[s + 1, 1]), sum_e=float(d[s + 1, 2]),
                            sum_vz2=float(d[s + 1, 3]), sum_e2=float(d[s + 1, 4]),
                            n_valley=[int(k[s + 1, 0]), int(k[s + 1, 1]), int(k[s + 1, 2])],
                            n_fault=int(k[s + 1, 3]), n_segments=int(k[s, 4]), n_events=int(k[s, 5])))
        return out

    def run_h2d_bytes(self, n_steps: int) -> int:
        return self.h2d_bytes

    def run_d2h_bytes(self, n_steps: int) -> int:
        return self.n * (48 + 4 + 8) + len(self.schedule) * 88 * (int(n_steps) + 1)

    def step(self, step: int, seed: int, field_mode: int = abi.FIELD_CONSTANT) -> dict:
        """One timestep on the pinned host arrays, in place; returns the observables."""
        torch = _torch()
        c, lib = self.ctx, self.ctx.lib
        if c.obs_before:
            raise RuntimeError("HostStepper returns the observables AFTER the step: call "
                               "Context.set_obs_phase(False) first")
        cur = torch.cuda.current_stream(c.device)
        for s in (self.s_in, self.s_run, self.s_out):
            s.wait_stream(cur)
        for ci in range((self.n + self.chunk - 1) // self.chunk):
            lo = ci * self.chunk
            m = min(self.chunk, self.n - lo)
            b = self.bufs[ci % len(self.bufs)]
            with torch.cuda.stream(self.s_in):
                if ci >= len(self.bufs):
                    self.s_in.wait_event(b["free"])            # previous user of this buffer drained
                b["aos"][:m].copy_(self.coords_t[lo:lo + m], non_blocking=True)
                b["soa"][6, :m].copy_(self.dtau_t[lo:lo + m], non_blocking=True)
                b["valley"][:m].copy_(self.valley_t[lo:lo + m], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(ev_in)
                st = C.c_void_p(self.s_run.cuda_stream)
                base, stride = b["soa"].data_ptr(), self.chunk * 8
                ptrs = [base + r * stride for r in range(7)] + [b["valley"].data_ptr()]
                c.check(lib.emc_aos_to_soa(c.h, b["aos"].data_ptr(), m, *ptrs[:6], st))
                c.check(lib.emc_flight_scatter(c.h, *ptrs, m, field_mode, None, None, None, seed,
                                               self.pid_offset + lo, step,
                                               b["obs"].data_ptr(), None, st))
                c.check(lib.emc_soa_to_aos(c.h, *ptrs[:6], m, b["aos"].data_ptr(), st))
                b["done"].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(b["done"])
                self.coords_t[lo:lo + m].copy_(b["aos"][:m], non_blocking=True)
                self.dtau_t[lo:lo + m].copy_(b["soa"][6, :m], non_blocking=True)
                self.valley_t[lo:lo + m].copy_(b["valley"][:m], non_blocking=True)
                self.obs_pinned[ci].copy_(b["obs"], non_blocking=True)
                b["free"].record(self.s_out)
        self.s_out.synchronize()
        cur.wait_stream(self.s_out)
        raw = self.obs_pinned.numpy()[:(self.n + self.chunk - 1) // self.chunk]
        d = raw[:, :5].copy().view(np.float64).sum(axis=0)
        k = raw[:, 5:].sum(axis=0)
        return dict(sum_z=float(d[0]), sum_vz=float(d[1]), sum_e=float(d[2]), sum_vz2=float(d[3]),
                    sum_e2=float(d[4]), n_valley=[int(k[0]), int(k[1]), int(k[2])], n_fault=int(k[3]),
                    n_segments=int(k[4]), n_events=int(k[5]))
