#!/usr/bin/env python
"""bench.py -- descriptor + energy + force atoms/s on the synthetic Cu fcc cell of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W     (the reference's dense CPU algorithm, bounded sample)

A step is one pass of the hot path over the whole cell: K0 wrap -> K1 cell-list neighbour list -> K2/K3 radial
and angular symmetry functions with analytic summed gradients -> K5 normalise -> K6 atomic MLP (e, dE/dg) ->
K4c force contraction.  N=1 workload: configs[3] of BASELINE.json, 63^3 fcc cells = 1,000,188 Cu atoms, rc = 6 A,
32 G2 + 64 G4 (the literal table is nnmd_b200.workloads.bench_table), fp32, reference (wrap-only) PBC semantics.
N>1: the same cell sharded spatially into boxes (2x2x2 on 8 GPUs) with a halo of rc (nnmd_b200.dist), strong scaling.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "descriptor+energy+force atoms/s (Cu fcc, 1M atoms)"
HIDDEN = [64, 64]


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (nvml, 20 ms period)."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        except Exception:
            self._nv = None
        return self

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in {**self.BAD, **self.NOTE}.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_net(nf: int, device, dtype=torch.float32):
    from nnmd_b200.workloads import bench_net
    return bench_net(nf, device, dtype)


# ------------------------------------------------------------------------------------------------- reference arm
def _reference_step_fn(ncell, sfd, net):
    """One descriptor+energy+force pass of the reference's own CPU implementation on an ncell^3 block of the bench lattice:
    the UNMODIFIED package staged in oracle/_ref (nnmd.features.calculate_sf, then bpnn.py:537-544: AtomicNN forward,
    autograd dE/dg, einsum -- BPNN.predict itself raises TypeError as shipped) when it is staged, else
    the dense restatement oracle/dense_oracle.py."""
    from nnmd_b200.workloads import fcc_cell
    pos, cell, pbc = fcc_cell(ncell, dtype=torch.float32)
    n = pos.shape[0]
    from oracle.build_ref import available, load_reference
    if available():
        load_reference()
        from nnmd.features import calculate_sf as ref_calculate_sf
        from nnmd.nn import AtomicNN as RefAtomicNN
        ref_net = RefAtomicNN(len(sfd["features"]), list(HIDDEN))
        ref_net.load_state_dict(net.state_dict())

        def step():
            g, dg = ref_calculate_sf(pos[None].clone(), cell, pbc, sfd, disable_tqdm=True)
            torch.autograd.set_detect_anomaly(False)  # calculate_sf.py:42 switches it on globally
            g = g.squeeze(0).requires_grad_(True)
            e = ref_net(g)
            dE = torch.autograd.grad(e.sum(), g)[0]
            f = -torch.einsum("ij,ijk->ik", dE, dg.squeeze(0))       # bpnn.py:541-544, single frame
            return float(e.sum()), f
        return step, n, "reference"
    from oracle import dense_oracle as O2
    Ws = [m.weight.detach() for m in net.model]
    bs = [m.bias.detach() for m in net.model]

    def step():
        g, dg = O2.descriptors(pos[None], cell, pbc, sfd, pow_mode="complex64")
        e, dE, f = O2.energy_force(g, dg, Ws, bs)
        return float(e.sum()), f
    return step, n, "port"


def run_reference(args):
    """`--impl reference`: the reference's dense O(N^2)/O(N^3) CPU path on all host cores, on the largest block of the
    bench lattice (108 / 256 / 500 atoms) whose K + W steps fit the time budget; the reference cannot hold more than
    ~500 atoms (N^3 temporaries), let alone the 1M-atom cell."""
    from nnmd_b200.workloads import bench_table
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sfd = bench_table()
    net = make_net(len(sfd["features"]), "cpu")
    # calibrate on 108 atoms, then take the largest block that keeps (K + W) steps inside the budget (t ~ N^3)
    ncell, warmup, t108 = 3, args.warmup, None
    if args.ref_cells:
        ncell = args.ref_cells
    else:
        step, n, kind = _reference_step_fn(3, sfd, net)
        step()
        t0 = time.perf_counter()
        step()
        t108 = time.perf_counter() - t0
        for cand in (5, 4):
            est = t108 * ((4 * cand ** 3) / 108.0) ** 3
            w = args.warmup if est < 5.0 else 1   # a 20 s CPU step needs no three warm-ups; the line reports what was done
            if est * (args.steps + w) <= args.ref_budget_s:
                ncell, warmup = cand, w
                break
    if ncell != 3 or t108 is None:
        step, n, kind = _reference_step_fn(ncell, sfd, net)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = n / dt
    what = ("unmodified reference from oracle/_ref: nnmd.features.calculate_sf + AtomicNN + autograd + einsum (bpnn.py:537-544)"
            if kind == "reference" else "dense torch restatement oracle/dense_oracle.py (complex64 pow); oracle/_ref is not staged")
    sample = (f"{n}-atom ({ncell}^3 cells) block of the same jittered Cu fcc lattice, same 96-function table and network, fp32, "
              f"{what}; " + (f"108-atom step {t108:.2f} s, " if t108 is not None else "") + "cost grows as N^3, the reference cannot hold the 1M-atom cell")
    sizes = [{"atoms": n, "functions": "32 G2 + 64 G4", "s_per_pass": dt, "atoms_per_s": value, "timed_steps": args.steps}]
    if not args.no_ref_sizes:
        sizes += reference_size_sweep(sfd, net, t108 if t108 is not None else dt * (108.0 / n) ** 3, skip_atoms=n, budget_s=max(args.ref_budget_s * 1.25 - dt * (args.steps + warmup), 30.0))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Cu fcc, rc=6 A, 32 G2 + 64 G4, MLP 96-64-64-1: descriptor+energy+force", "atoms": n},
            "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": torch.get_num_threads(), "kind": kind, "sample": sample,
                             "host_cpus": cores, "sizes": sizes,
                             "law": "t ~ F_angular * N^3 (dense (B,N,N,N) temporaries, symm_funcs.py:181-219): the 1M-atom cell is "
                                    "out of reach of the reference by ~9 orders of magnitude in memory alone"},
            "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def reference_size_sweep(sfd, net, t108, skip_atoms=0, budget_s=150.0):
    """the reference's calculate_sf path at the sizes that fit, one pass each, same table and
    seeded lattice -- the full 96-function set at N = 256; at N = 500 the 32 radial functions plus ONE angular function
    (the full set is ~10 min per pass there) with the F * N^3 extrapolation; radial-only at N = 4000."""
    from nnmd_b200.workloads import fcc_cell
    from oracle.build_ref import available, load_reference
    if not available():
        return []
    load_reference()
    from nnmd.features import calculate_sf as ref_sf
    out = []
    t_start = time.perf_counter()

    def one(ncell, feats, params, label):
        pos, cell, pbc = fcc_cell(ncell, dtype=torch.float32)
        t0 = time.perf_counter()
        ref_sf(pos[None].clone(), cell, pbc, {"features": feats, "params": params}, disable_tqdm=True)
        torch.autograd.set_detect_anomaly(False)
        dt = time.perf_counter() - t0
        return {"atoms": int(pos.shape[0]), "functions": label, "s_per_pass": dt, "atoms_per_s": pos.shape[0] / dt}

    F, P = sfd["features"], sfd["params"]
    rad = one(5, F[:32], P[:32], "32 G2")
    r = one(5, F[:33], P[:33], "32 G2 + 1 G4")
    r["extrapolated_full_set_s"] = rad["s_per_pass"] + 64 * (r["s_per_pass"] - rad["s_per_pass"])
    out += [rad, r]
    if time.perf_counter() - t_start < budget_s * 0.3:
        out.append(one(10, F[:32], P[:32], "32 G2"))
    # the full set at N = 256 costs ~18x the 108-atom pass (measured): only when the budget allows
    if skip_atoms != 256 and t108 * 18.0 < budget_s - (time.perf_counter() - t_start):
        out.append(one(4, F, P, "32 G2 + 64 G4"))
    return out


# --------------------------------------------------------------------------------------------------------- ours
def cpu_baseline_port(sfd, seconds_target=20.0):
    """The O(N) cell-list C restatement (oracle/nl_oracle.c, OpenMP, fp64) on all host cores, bounded sample."""
    from nnmd_b200.workloads import fcc_cell
    from oracle import nl_oracle as O3
    threads = O3.max_threads()
    ncell = 8
    pos, cell, pbc = fcc_cell(ncell, dtype=torch.float64)
    t0 = time.perf_counter()
    O3.raw_descriptors(pos[None], cell, pbc, sfd)
    dt = time.perf_counter() - t0
    if dt < seconds_target / 8 and ncell < 24:  # grow once if the box is fast (about 10 s of work on 16 cores)
        ncell = 24
        pos, cell, pbc = fcc_cell(ncell, dtype=torch.float64)
        t0 = time.perf_counter()
        O3.raw_descriptors(pos[None], cell, pbc, sfd)
        dt = time.perf_counter() - t0
    n = pos.shape[0]
    return {"value": n / dt, "unit": "atoms/s", "cores": threads, "kind": "port",
            "sample": f"{n}-atom ({ncell}^3 cells) block of the same lattice and table; descriptors + summed gradients only "
                      f"(the >99% share), O(N) cell-list C restatement oracle/nl_oracle.c, fp64, OpenMP x{threads}, one pass {dt:.1f} s"}


def run_ours(args):
    from nnmd_b200 import _lib, dist
    from nnmd_b200.engine import energy_forces
    from nnmd_b200.nn import BPNN
    from nnmd_b200.workloads import bench_table, fcc_cell

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as td
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    sfd = bench_table()
    nf = len(sfd["features"])
    dt = torch.float64 if args.dtype == "f64" else torch.float32   # f64: the parity mode of north_star, measured on request
    pos, cell, pbc = fcc_cell(args.cells, dtype=dt)
    n_atoms = pos.shape[0]
    net = make_net(nf, dev, dt)
    lib = _lib.load()

    # ---- spatial shard of this rank (whole cell for N=1): owned atoms first, halo atoms after
    shard = dist.SlabShard.from_global(pos, cell, pbc, rc=6.0, rank=rank, world=world, grid=args.grid).to(dev)
    local_pos = shard.exchange_halo(dev)           # (n_local, 3) on the GPU: owned + halo (NCCL send/recv for N>1)
    n_owned = shard.n_owned

    # capacity-bound neighbour build: after the first step nothing in a step waits for the host (skin 0: the list is rebuilt
    # from scratch EVERY step -- a step is one complete pass of the hot path over the structure)
    from nnmd_b200.neighbor import NeighborWorkspace
    ws = None if args.sync_build else NeighborWorkspace(1, local_pos.shape[0], 6.0, dt, dev, n_centres=n_owned)

    graphed = None
    if world > 1 and not args.no_graph and ws is not None:
        # sharded steps are a few ms: replay the whole local evaluation from ONE CUDA graph (the halo exchange writes into the
        # persistent buffer the graph reads; NCCL stays outside the graph)
        from nnmd_b200.engine import GraphedEnergyForces
        graphed = GraphedEnergyForces(local_pos, cell, pbc, sfd, net, n_centres=n_owned, mode=args.mode, mlp_precision=args.mlp)
        ws = graphed.ws

    def step_device():
        # N>1: the halo exchange is part of every step (positions change every MD step)
        lp = shard.exchange_halo(dev) if world > 1 else local_pos
        if graphed is not None:
            assert lp.data_ptr() == local_pos.data_ptr()
            res = graphed()
        else:
            # cell and pbc stay on the host: K0 takes them as kernel arguments (a device copy would cost a read-back per step)
            res = energy_forces(lp[None], cell, pbc, sfd, net, mode=args.mode, n_centres=n_owned,
                                mlp_precision=args.mlp, check=False, workspace=ws)
        if world > 1:
            dist.allreduce_sum(res.energy)  # total energy of the structure (peer-window kernel; NCCL where unavailable)
        return res

    flush_buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as td
            td.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res = step_device()
    torch.cuda.synchronize()
    assert int(res.flag.item()) == 0, "NaN / overflow flag raised in the bench workload"
    mean_nbrs = float(ws.true_sizes()[0] if ws is not None else res.nlist.total) / max(n_owned, 1)

    # ---- timed region: K steps, device events, L2 flushed between steps
    lib.nnmd_profile_enable(1)
    _lib.profile_read()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    with ClockSampler(local_rank) as clocks:
        # the barrier comes AFTER the sampler has started: NVML initialisation takes a different time on every rank (tens of
        # ms), and at N>1 a rank that enters the loop early spends that skew waiting in the first halo exchange -- inside its
        # first timed step (seen once at 8 GPUs: 8.48 ms per step over kernel sums of 7.35 ms)
        barrier()
        for k in range(args.steps):
            flush_buf.zero_()
            starts[k].record()
            res = step_device()
            ends[k].record()
        barrier()
    launches = _lib.launch_count() - launches0
    assert int(res.flag.item()) == 0, "NaN / overflow flag raised during the timed steps"
    prof = _lib.profile_read()
    prof_steps = args.steps
    if graphed is not None:
        # a graph replay carries no per-kernel events and no host-side launch calls: the stage breakdown (and the launch
        # count) comes from two extra, UNTIMED steps launched kernel by kernel on the same buffers
        launches0 = _lib.launch_count()
        for _ in range(2):
            energy_forces(local_pos[None], cell, pbc, sfd, net, mode=args.mode, n_centres=n_owned, mlp_precision=args.mlp,
                          check=False, workspace=ws)
        launches = (_lib.launch_count() - launches0) * args.steps // 2
        prof = _lib.profile_read()
        prof_steps = 2
    lib.nnmd_profile_enable(0)
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as td
        td.all_reduce(t, op=td.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_atoms / (ms_per_step * 1e-3)

    # ---- end to end: pinned host positions in, energy + forces back on the host, every step
    #      N=1: through the public API a user calls (BPNN.predict); N>1: the sharded engine call with the H2D of the
    #      owned positions, the halo exchange and the D2H of the owned forces inside the timed region
    e2e = None
    n_e2e = max(2, min(args.steps, 5))
    if world == 1:
        bp = BPNN(dtype=dt)
        bp.species, bp.atomic_nets, bp.pbc = ["Cu"], [net], None
        host = pos.clone().pin_memory()

        out_e = torch.empty(n_atoms, dtype=dt).pin_memory()     # per-atom energies (single frame: bpnn.py:532-554)
        out_f = torch.empty(n_atoms, 3, dtype=dt).pin_memory()

        def step_e2e():
            E, F = bp.predict({"Cu": host}, cell, {"Cu": sfd}, pbc=pbc)
            out_e.copy_(E, non_blocking=True)
            out_f.copy_(F, non_blocking=True)
            torch.cuda.synchronize()
            return out_e, out_f
        api = "nnmd_b200.nn.BPNN.predict"
    else:
        host = shard.owned_pos.cpu().pin_memory()

        out_e = torch.empty(1, dtype=dt).pin_memory()
        out_f = torch.empty(1, n_owned, 3, dtype=dt).pin_memory()

        def step_e2e():
            shard.set_owned_positions(host.to(dev, non_blocking=True))
            r = step_device()
            out_e.copy_(r.energy, non_blocking=True)
            out_f.copy_(r.forces, non_blocking=True)
            torch.cuda.synchronize()
            return out_e, out_f
        api = "nnmd_b200.engine.energy_forces on nnmd_b200.dist.SlabShard"
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        E, F = step_e2e()
    barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        td.all_reduce(dt, op=td.ReduceOp.MAX)
    dt = float(dt.item())
    e2e = {"value": n_atoms / dt, "unit": "atoms/s", "h2d_bytes_per_step": int(host.numel() * host.element_size()) * world,
           "d2h_bytes_per_step": int(F.numel() * F.element_size() + E.numel() * E.element_size()) * world, "api": api, "ms_per_step": dt * 1e3}

    if getattr(shard, "_window", None) is not None:
        shard._window.check()  # raises if a peer ever failed to arrive at a halo exchange
    # ---- checksum of the structure the LAST timed step computed: must agree across N = 1, 2, 4, 8 to fp32 round-off
    chk = torch.stack([res.forces.double().abs().sum(), *res.forces.double().sum(dim=(0, 1))])
    if world > 1:
        td.all_reduce(chk)  # res.energy was already reduced inside the step
    checksum = {"energy": float(res.energy.double().sum()), "sum_abs_F": float(chk[0]), "sum_F": [float(v) for v in chk[1:]],
                "atoms": n_atoms}

    # ---- training-step throughput of configs[2] and configs[4] (frames sharded; one flat gradient all-reduce per step)
    train = None
    if not args.no_train:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_train
        del flush_buf
        torch.cuda.empty_cache()
        train = {}
        for cfg_name in ("li", "bin"):
            # weak scaling keeps the work PER GPU fixed: --train-frames frames (and one batch of 500) on every rank
            dataset = bench_train.build_dataset(cfg_name, args.train_frames * world, dev, rank, world)
            entry = bench_train.train_bench(cfg_name, dev, rank, world, args.train_frames * world, 500, args.train_steps, 5, "weak", dataset)
            if world > 1:  # the reference's optimisation: global batch stays 500 frames
                strong = bench_train.train_bench(cfg_name, dev, rank, world, args.train_frames * world, 500, args.train_steps, 5, "strong", dataset)
                entry["strong"] = {k: strong[k] for k in ("global_batch_frames", "ms_per_step", "frames_per_s", "atoms_per_s")}
            if world == 1 and rank == 0 and not args.no_cpu_baseline:
                entry["cpu_baseline"] = bench_train.reference_train_baseline(cfg_name, *dataset)
            train[cfg_name] = entry
            del dataset
    # per-rank stage times (device events): the step of a sharded run is set by the slowest rank
    stage_names = sorted(prof)
    stage_vec = torch.tensor([prof[k][0] / prof_steps for k in stage_names] + [float(n_owned), float(local_pos.shape[0])],
                             dtype=torch.float64, device=dev)
    if world > 1:
        all_stage = [torch.empty_like(stage_vec) for _ in range(world)]
        td.all_gather(all_stage, stage_vec)
    else:
        all_stage = [stage_vec]
    per_rank = [{**{k: float(v[i]) for i, k in enumerate(stage_names)}, "owned_atoms": int(v[-2]), "local_atoms": int(v[-1])} for v in all_stage]
    if rank != 0:
        return
    # ---- roofline of the dominant kernel (angular K3e): algorithmic bytes per centre atom
    peak, peak_src = load_peaks()
    n_ang = sum(1 for k in sfd["features"] if k in (4, 5))
    ang_ms, ang_n = prof["angular"]
    ang_ms_per_launch = ang_ms / max(ang_n, 1)
    # own position + neighbour row (int32) + each neighbour position once + the angular columns of G and dG
    bytes_per_atom = 16 + 4 * mean_nbrs + 16 + n_ang * 4 + (n_ang * 12 if args.mode == "materialize" else 0)
    achieved = bytes_per_atom * n_owned / (ang_ms_per_launch * 1e-3) / 1e9 if ang_ms_per_launch > 0 else 0.0
    traffic = None
    try:  # DRAM bytes of ONE launch of this kernel from the committed ncu --set full capture (scaled if the cell differs)
        with open(os.path.join(ROOT, "profiles", "r2_k3m_traffic.json")) as fh:
            tr = json.load(fh)
        traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * (n_owned / tr["atoms"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "angular_edgem_kernel<true,true,uint16_t> (K3e, moment form)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "profiles/r2_k3m_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
                "ms_per_launch": ang_ms_per_launch, "algorithmic_bytes_per_atom": bytes_per_atom,
                "note": "K3e is bound by the FP32 FMA pipe and the SFU (ex2), not by HBM: see fp32_pipe/sfu_pipe, DESIGN.md, profiles/",
                "stage_ms_per_step": {k: v[0] / prof_steps for k, v in prof.items()}}
    # compute-side view of the same kernel against the MEASURED pipe peaks (tools/microbench.cu -> profiles/microbench_r1.json):
    # one (edge, third atom, function) costs 9 FMA-pipe lane-ops in the moment form of K3e (3 ladder multiplications + 6
    # accumulations; the round-1 form needed 11) and about one ex2 (3 per record and lane for 4 functions, plus the record's
    # own 15); records per atom = closed ordered (j,k) pairs / 2 = triangles-per-atom (3 edges x 1/3 ownership)
    try:
        with open(os.path.join(ROOT, "profiles", "microbench_r1.json")) as fh:
            mb = json.load(fh)
    except Exception:
        mb = None
    if mb is not None and ang_ms_per_launch > 0 and args.tri_per_atom > 0:
        rec = args.tri_per_atom * (mean_nbrs / 78.0) ** 2 * n_owned  # free surfaces of the wrap-only cell thin the rows
        fma_ops = rec * n_ang * 9.0
        sfu_ops = rec * n_ang * 1.0
        roofline["fp32_pipe"] = {"achieved_lane_ops_per_s": fma_ops / (ang_ms_per_launch * 1e-3), "peak": mb["ffma_lane_ops_per_s"],
                                 "frac": fma_ops / (ang_ms_per_launch * 1e-3) / mb["ffma_lane_ops_per_s"],
                                 "ops_per_record_and_function": 9,
                                 "frac_in_round1_op_model": rec * n_ang * 11.0 / (ang_ms_per_launch * 1e-3) / mb["ffma_lane_ops_per_s"],
                                 "note": "frac counts the operations this kernel executes (9); frac_in_round1_op_model prices the same "
                                         "records at the 11 operations of the round-1 kernel (0.456 then), i.e. the speed-up in its units",
                                 "peak_source": "measured, tools/microbench.cu (FFMA and FFMA2 give the same lane rate)"}
        roofline["sfu_pipe"] = {"achieved_lane_ops_per_s": sfu_ops / (ang_ms_per_launch * 1e-3), "peak": mb["mufu_ex2_lane_ops_per_s"],
                                "frac": sfu_ops / (ang_ms_per_launch * 1e-3) / mb["mufu_ex2_lane_ops_per_s"]}
    line = {"metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "synthetic Cu fcc 1M-atom cell, rc=6 A, 32 G2 + 64 G4: descriptor+force eval"
                                   if args.cells == 63 else f"synthetic Cu fcc {args.cells}^3 cells",
                       "atoms": n_atoms, "cells": args.cells, "mlp": [nf] + HIDDEN + [1], "mode": args.mode,
                       "mlp_precision": args.mlp, "pbc": "reference wrap-only (SURVEY Q1)", "mean_neighbours": mean_nbrs,
                       "l2": "256 MiB buffer written between timed steps; per-step working set (neighbour list + dG) is >1 GB",
                       "parallelism": "single GPU" if world == 1 else
                       f"{'x'.join(str(v) for v in shard.grid)} boxes (equal-work boundaries), halo rc, "
                       + ("halo exchange and energy sum as kernels over NVLink peer memory (CUDA IPC windows)"
                          if getattr(shard, "_window", None) is not None else "NCCL send/recv"),
                       "launch": "one CUDA-graph replay per step + NCCL" if graphed is not None else "kernel by kernel"},
            "clocks": clocks.summary(), "gpu_launches": launches, "roofline": roofline}
    line["checksum"] = checksum
    if world > 1:
        line["per_rank_stage_ms"] = per_rank
    if train is not None:
        line["train"] = train
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_port(sfd)
        try:  # the reference's own dense algorithm beside it, on the largest block that takes about a second
            step, n_ref, kind = _reference_step_fn(3, sfd, make_net(nf, "cpu"))
            step()
            t0 = time.perf_counter()
            step()
            line["cpu_baseline"]["reference_dense"] = {"value": n_ref / (time.perf_counter() - t0), "unit": "atoms/s", "atoms": n_ref,
                                                       "kind": kind, "note": "see `bench.py --impl reference` for N = 256 / 500"}
        except Exception as exc:
            line["cpu_baseline"]["reference_dense"] = {"unavailable": repr(exc)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=63, help="fcc conventional cells per edge (63 -> 1,000,188 atoms)")
    ap.add_argument("--mode", default="materialize", choices=["materialize", "fused"])
    ap.add_argument("--mlp", default="auto", choices=["auto", "exact", "tf32"],
                    help="atomic MLP kernel: auto = tensor cores (3xTF32, parity-tested to 1e-5) when the shape allows")
    ap.add_argument("--ref-cells", type=int, default=0, help="reference arm sample: cells per edge (3 -> 108, 4 -> 256, 5 -> 500 atoms; 0 = largest that fits --ref-budget-s)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="reference arm: time budget of the K + W steps (the single passes at other sizes get what is left of 1.25x this)")
    ap.add_argument("--no-ref-sizes", action="store_true", help="reference arm: skip the N = 256 / 500 / 4000 single passes")
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput block")
    ap.add_argument("--train-frames", type=int, default=8000)
    ap.add_argument("--train-steps", type=int, default=40)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="working precision (f32 = the reference's training/eval precision and the headline; f64 = parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grid", default="auto", type=lambda v: v if v == "auto" else tuple(int(x) for x in v.split("x")),
                    help="N>1: decomposition, 'auto' (least boundary area, 2x2x2 on 8 GPUs) or e.g. 8x1x1 for x-slabs")
    ap.add_argument("--no-graph", action="store_true", help="N>1: launch the kernels of a step one by one instead of replaying a CUDA graph")
    ap.add_argument("--sync-build", action="store_true", help="neighbour build with a host sync per step (the round-1 form)")
    ap.add_argument("--tri-per-atom", type=float, default=1350.0,
                    help="closed triangles per atom of the workload (bulk fcc, rc=6: 1350) for the pipe-utilisation estimate")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as td
        if td.is_initialized():
            td.destroy_process_group()


if __name__ == "__main__":
    main()
