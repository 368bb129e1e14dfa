#!/usr/bin/env python
"""bench.py - processor edge-updates/sec of the message-passing hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
                  [--precision P]

A step is one pass of the hot path (encoder -> 15 GN_Blocks -> decoder [-> loss -> backward for the
training workload]) over one batch of synthetic meshes.  Default workload = BASELINE.json configs[1]:
FvgnA, batch 8 x 20k-cell cylinder meshes, on each GPU (weak scaling, meshes are independent units,
no data-path collective in the forward; gradient all-reduce in the training workload).
Prints ONE JSON line (rank 0).  `--impl reference` times the oracle port of the reference's CPU
implementation on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

MP_NUM = 15
WORKLOADS = {
    # name: (model, meshes per GPU, cells per mesh, obstacle, training step?)
    "fvgn_fwd_8x20k": ("FvgnA", 8, 20000, "cylinder", False),
    "fvgn_train_8x20k": ("FvgnA", 8, 20000, "cylinder", True),
    # BASELINE.json configs[4]: VertPot / StreamFunc data-parallel training, 8 meshes per GPU (64 over 8 GPUs)
    "vertpot_train_8x20k": ("VertPotA", 8, 20000, "cylinder", True),
    "streamfunc_train_8x20k": ("StreamFuncA", 8, 20000, "cylinder", True),
    "mgn_fwd_2k": ("MgnA", 1, 2048, "none", False),
    "mgn_fwd_200k": ("MgnA", 1, 200000, "airfoil", False),
    # autoregressive rollouts (second half of the metric: rollout steps/sec); "rollout" in the training slot
    "mgn_rollout_2k": ("MgnA", 1, 2048, "cylinder", "rollout"),        # BASELINE.json configs[0] shape
    "flux_rollout_200k": ("FluxA", 1, 200000, "cylinder", "rollout"),  # configs[2]
    "cons_rollout_200k": ("ConservativeA", 1, 200000, "cylinder", "rollout"),  # configs[2], Conservative variant
    "mgn_rollout_4m": ("MgnA", 1, 4000000, "airfoil", "rollout"),      # configs[3]: domain-decomposed over --gpus N
}
DEFAULT_WORKLOAD = "fvgn_train_8x20k"   # BASELINE.json configs[1]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def _collate(samples):
    from gnn_fluid_dynamics_b200.graph import collate_triplet
    return collate_triplet(samples) if len(samples) > 1 else _with_batch([g.clone() for g in samples[0]])


def build_batch(model_name, n_meshes, n_cells, kind, seed0=0):
    """Synthetic batch of independent meshes, PyG-style concatenation (training-time edge flips on)."""
    return _collate(build_samples(model_name, n_meshes, n_cells, kind, seed0))


def build_samples(model_name, n_meshes, n_cells, kind, seed0=0):
    """The per-mesh graph triplets of the synthetic batch (what a dataset hands to the collation)."""
    from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs
    samples = []
    for i in range(n_meshes):
        g = mesh_graphs(make_mesh(n_cells, kind, seed=seed0 + i), seed=100 + seed0 + i, flip_edges=True)
        if model_name == "MgnA":
            g[0].y = torch.cat([g[0].y, torch.zeros(g[0].x.shape[0], 1)], 1)
            g[1].y = g[1].y[:, :2].contiguous()
        elif model_name in ("FvgnA", "FluxA", "ConservativeA"):
            g[1].y = g[1].y[:, :3].contiguous()
        else:                               # the model's own targets / extra inputs, as the parity tests build them
            from helpers import finish_graphs
            g = finish_graphs(model_name, g)
            for t in g[:2]:
                t._store.pop("batch", None)          # re-created by the collation
        samples.append(g)
    return samples


def workload_config(workload, n_cells_total, n_faces, n_vertices):
    """The `config` object of the JSON line: ONLY what defines the workload, identical in both arms
    (`--impl ours` and `--impl reference`), so the driver can check that they measured the same thing."""
    model_name, n_meshes, n_cells, kind, train = WORKLOADS[workload]
    working_set = 4 * 128 * (2 * n_faces + 3 * n_cells_total) + 4 * 64 * n_vertices
    if train == "rollout":
        timed = "one autoregressive step: normalise + encoder + 15 GN_Blocks + decoder (+ integrator) + state advance"
    elif train:
        timed = "forward (encoder + 15 GN_Blocks + decoder + integrator) + loss + backward + grad clip + Adam step"
    else:
        timed = "encoder + 15 GN_Blocks + decoder"
    return {"workload": workload, "model": model_name, "mp_num": MP_NUM, "hidden": 128,
            "meshes_per_gpu": n_meshes, "cells_per_mesh": n_cells, "obstacle": kind, "cells": n_cells_total,
            "faces": n_faces, "vertices": n_vertices, "timed": timed,
            "l2": "working set > L2" if working_set >= 256e6 else "flushed between iterations"}


def _with_batch(g):
    g[0].batch = torch.zeros(g[0].x.shape[0], dtype=torch.long)
    g[1].batch = torch.zeros(g[1].pos.shape[0], dtype=torch.long)
    return g


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:  # noqa: BLE001
            self.err = repr(exc)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self.t.start()

    def stop(self):
        self._stop.set()
        if self.ok and self.t.is_alive():
            self.t.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def ncu_traffic(which):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of a kernel at the bench shape, from the committed
    `ncu --set full` captures (scripts/refresh_profiles.sh): profiles/r02_fwd_edge_fast_ncu_summary.json (inference edge
    block, scripts/prof_fwd_edge.py fast) and profiles/r02_train_kernels_ncu_summary.json (scripts/prof_train_kernels.py:
    forward + stash, dgrad chain, the lean weight-gradient launch, ...), looked up by kernel instantiation."""
    try:
        if which in ("edge_chain", "edge_stash"):
            name = "r02_train_kernels_ncu_summary.json"
            d = json.load(open(os.path.join(ROOT, "profiles", name)))
            # (instantiation prefix: <FP16, NA, NW, BWD, EPI[, LEAN]>)
            tag = "mlp_tc_kernel<0, 2, 2, 1, 0" if which == "edge_chain" else "mlp_tc_kernel<0, 2, 2, 0, 0"
            hits = [l for l in d["launches"] if tag in l.get("kernel", "")]
            return int(max(hits, key=lambda l: l.get("duration_us", 0))["dram_traffic_bytes"]), name
        name = "r02_fwd_edge_fast_ncu_summary.json"
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        return int(d["launches"][0]["dram_traffic_bytes"]), name
    except Exception:  # noqa: BLE001
        return None, None


def algorithmic_bytes_forward(E, N, V, family="fvgn"):
    """Whole processor pass, SURVEY.md section 8d: 512 (3E + 4N + V) + 16E + 12N bytes per GN_Block (FVGN order)."""
    return MP_NUM * (512 * (3 * E + 4 * N + V) + 16 * E + 12 * N)


def algorithmic_bytes_edge_kernel(E, N):
    """Fused edge block, FVGN order: read e (E rows) + gather x' (N unique rows) + write e+e' (E rows),
    512 B per fp32 row, + row/col int32 indices."""
    return 512 * (2 * E + N) + 8 * E


def _oracle_step_fn(model_name, sd, graphs, train):
    """One step of the CPU arm: the oracle port of the reference's forward (+ loss + backward + Adam step for the
    training workload) on ONE mesh.  Imports nothing from the product package that loads the CUDA library."""
    from oracle import model as omodel
    from fixtures import default_stats
    from helpers import LOSS_W
    stats = default_stats()
    params = {k: v.clone().requires_grad_(bool(train) and v.is_floating_point() and "running" not in k
                                          and not k.startswith("normalizer.")) for k, v in sd.items()}
    opt = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=1e-4) if train else None

    def step():
        g = [x.clone() for x in graphs]
        if train:
            opt.zero_grad(set_to_none=True)
            out, _ = omodel.model_forward(model_name, params, stats, g, MP_NUM, mode="train", training=True)
            loss = omodel.fvgn_loss(params, out, g, LOSS_W, training=True)["total_log_loss"]
            loss.backward()
            opt.step()
        else:
            with torch.no_grad():
                omodel.model_forward(model_name, sd, stats, g, MP_NUM, mode="train")
    return step


def run_reference(args, world, rank):
    """Reference arm: the oracle port of the reference's CPU path on the host cores (the Python reference itself
    cannot travel to the GPU box).  The CUDA library is never loaded here: parameters come from
    tests/fixtures.py state_dict_from_keys (reference-generated key list + the deterministic fill)."""
    if rank != 0:
        return
    from fixtures import state_dict_from_keys
    model_name, n_meshes, n_cells, kind, train = WORKLOADS[args.workload]
    if train == "rollout":
        train = False        # the CPU arm times the forward of the step (state advance is negligible on the CPU)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = build_batch(model_name, n_meshes, n_cells, kind)      # for the config's totals (same batch as our arm)
    N, E_total, V = batch[0].x.shape[0], batch[0].edge_index.shape[1], batch[2].pos.shape[0]
    graphs = build_batch(model_name, 1, n_cells, kind)      # bounded sample: ONE mesh of the batch
    sd = state_dict_from_keys(model_name)
    E = graphs[0].edge_index.shape[1]
    step = _oracle_step_fn(model_name, sd, graphs, train)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = E * MP_NUM / dt
    sample = f"1 of {n_meshes} meshes ({n_cells}-cell {kind}), whole forward{'+loss+backward+Adam step' if train else ''} per step"
    line = {
        "impl": "reference", "metric": "processor edge-updates/sec", "value": value, "unit": "edge-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, N, E_total, V),
        "cpu_baseline": {"value": value, "unit": "edge-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "edge-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    assert "gnn_fluid_dynamics_b200._lib" not in sys.modules, "the CPU reference arm must not load the CUDA library"
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-4m", default="auto", choices=["auto", "on", "off"],
                    help="also run BASELINE configs[3] (4M-cell MGN rollout, domain-decomposed over the run's GPUs) and report it "
                         "under the `strong_4m` key; auto = with the default workload")
    ap.add_argument("--halo", default="nccl", choices=["peer", "nccl"],
                    help="domain-decomposed workloads: ghost latents via NCCL send/receive (default, measured faster) or "
                         "via peer-memory gathers inside the edge kernel")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from helpers import build_model
    from gnn_fluid_dynamics_b200 import ops
    from gnn_fluid_dynamics_b200.precisions import available
    from gnn_fluid_dynamics_b200.topology import get_topology
    from gnn_fluid_dynamics_b200.dist import allreduce_gradients

    model_name, n_meshes, n_cells, kind, train = WORKLOADS[args.workload]
    if train == "rollout":
        return run_rollout(args, world, rank, dev, dist)
    from gnn_fluid_dynamics_b200.models.base import DEFAULT_PRECISION
    prec = args.precision or DEFAULT_PRECISION
    assert prec in available(), prec
    model = build_model(model_name, precision=prec).to(dev)
    model.train() if train else model.eval()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4) if train else None      # reference src/train.py:83
    host_samples = build_samples(model_name, n_meshes, n_cells, kind, seed0=rank * n_meshes)      # per-mesh triplets
    host_graphs = [g.pin_memory() for g in _collate(host_samples)]
    host_samples = [[g.pin_memory() for g in smp] for smp in host_samples]
    N, E = host_graphs[0].x.shape[0], host_graphs[0].edge_index.shape[1]
    V = host_graphs[2].pos.shape[0]

    # ---- device-resident leg (`value`): inputs already normalised and in HBM ------------------------
    gd = [g.to(dev) for g in host_graphs]
    if model_name == "FvgnA":
        gd = model.normalizer.input(gd)
    topo = get_topology(gd).validate()
    if train:
        topo.build_row_col_interleaved_csr(); topo.build_vf_csr()
        gd[0].topology = gd[2].topology = topo
    working_set = 4 * 128 * (2 * E + 3 * N) + 4 * 64 * V
    flush = torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev) if working_set < 256e6 else None

    has_fn = model_name == "FvgnA"       # (subclasses inherit the method but not its contract)

    def train_step(graphs_in):
        """forward (encoder + 15 GN_Blocks + decoder, integrator) + loss + backward + clip + Adam step
        (reference Trainer._train_step, src/train.py:245-272).  Models with ``forward_normalised`` take the already
        normalised resident batch; the others normalise inside ``forward`` (in place, like the reference), so they get
        a fresh copy of the raw batch every step - what the reference's loader hands its trainer."""
        opt.zero_grad(set_to_none=True)
        if has_fn:
            graphs_norm = graphs_in
            out = model.forward_normalised(graphs_norm, mode="train")
        else:
            graphs_norm = [g.clone() for g in graphs_in]
            out = model(graphs_norm, mode="train")
        loss = model.loss(out, graphs_norm)["total_log_loss"]
        loss.backward()
        if world > 1:   # data-parallel: meshes are independent units, one gradient all-reduce per step
            allreduce_gradients(list(model.parameters()), world)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    def step_resident():
        if train:
            return train_step(gd)
        with torch.no_grad():
            return model.encode_process_decode(gd[0].x, gd[1].x, topo)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ops.LAUNCHES = 0
    for s, e in ev:
        if flush is not None:
            flush.zero_()                       # L2 flush between timed iterations (not timed)
        s.record()
        step_resident()
        e.record()
    barrier()
    launches = ops.LAUNCHES
    ms = sum(s.elapsed_time(e) for s, e in ev) / args.steps
    clocks = sampler.stop()
    if os.environ.get("GNNFD_BENCH_PROFILE_STEP") and rank == 0:
        # diagnostic (never part of a reported number): one more step under torch.profiler, top operators by device time
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step_resident()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), file=sys.stderr)

    # forward-only time of the same batch (reported next to the training number)
    fwd_ms = None
    if train:
        model.eval()
        with torch.no_grad():
            for _ in range(3):
                model.encode_process_decode(gd[0].x, gd[1].x, topo)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                model.encode_process_decode(gd[0].x, gd[1].x, topo)
            b.record()
        torch.cuda.synchronize()
        fwd_ms = a.elapsed_time(b) / 5
        model.train()

    # ---- end-to-end leg (`e2e`): host graphs -> public API -> host result ---------------------------
    # The public data path: per-mesh host samples (pinned) -> GraphCache.fetch (static geometry / connectivity of each
    # mesh resident in HBM after its first use, batch collated on the device, only the per-sample attributes - state,
    # targets, the re-flipped c_graph.edge_index - travel every step) -> model.  SURVEY.md section 8f row 4.
    from gnn_fluid_dynamics_b200.graph_cache import GraphCache
    cache = GraphCache(dev)
    mesh_keys = [("bench", rank, i) for i in range(n_meshes)]

    def step_e2e():
        g = cache.fetch(mesh_keys, host_samples)
        before = torch.cuda.Event()
        before.record()
        if train:
            loss = train_step(model.normalizer.input(g) if has_fn else g)              # enqueued, not waited for
            cache.prefetch(mesh_keys, host_samples, after=before)      # the next step's inputs travel while this one computes
            return float(loss.item())                                                  # D2H of the loss
        with torch.no_grad():
            out = model(g, mode="train")
            cache.prefetch(mesh_keys, host_samples, after=before)
            return out["cell_velocity_change"].to("cpu", non_blocking=False)

    h2d_full = sum(t.numel() * t.element_size() for g in host_graphs for t in g._store.values() if torch.is_tensor(t))
    d2h = 4 if train else N * 2 * 4
    for _ in range(3):
        step_e2e()
    h2d = cache.h2d_bytes                       # a steady-state step: the per-sample attributes only
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps

    # ---- dominant kernels timed alone with CUDA events on their stream (L2 flushed between launches) -------
    model.eval()
    blk = model.processer_list[7]
    from gnn_fluid_dynamics_b200 import processor as P
    from gnn_fluid_dynamics_b200.ops import Seg
    from gnn_fluid_dynamics_b200._lib import SEG_GATHER
    x_lat = torch.randn(N, 128, device=dev)
    e_lat = torch.randn(E, 128, device=dev)
    g_lat = torch.randn(E, 128, device=dev)
    flush_k = flush if flush is not None else torch.empty(160 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def time_kernel(fn):
        ts = []
        with torch.no_grad():
            for i in range(args.warmup + args.steps):
                flush_k.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                if i >= args.warmup:
                    ts.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ts) / len(ts)

    hbm_peak, peak_kind = peaks()
    face_mlp = (blk.face_block if hasattr(blk, "face_block") else blk.edge_block).face_mlp
    w_edge = P.weights_of(face_mlp)
    esegs = [Seg(e_lat), Seg(x_lat, SEG_GATHER, (topo.row,)), Seg(x_lat, SEG_GATHER, (topo.col,))]
    # the inference edge block exactly as the model's forward launches it: x'[row] / x'[col] TMA-gathered from the split
    # shadow the node block's epilogue wrote, e updated in place by the TMA reduce-store epilogue
    fast = P.Fast(N, model.prec, dev)
    hi = x_lat.to(fast.dtype)
    fast.xs[:, :128] = hi
    fast.xs[:, 128:] = (x_lat - hi.float()).to(fast.dtype)
    k_ms = time_kernel(lambda: P.edge_mlp_concat(face_mlp, e_lat, None, topo, model.prec, want_raw=False, fast=fast))
    e_lat = torch.randn(E, 128, device=dev)          # (the in-place runs above accumulated into it)
    alg = algorithmic_bytes_edge_kernel(E, N)
    traffic, traffic_src = ncu_traffic("edge_fwd")
    fwd_roof = {"kernel": "mlp_tc_kernel<EPI=1>: fused edge block forward (TMA gather4 + 3-layer MLP + LayerNorm + in-place residual), inference",
                "bound": "hbm", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg / (k_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": peak_kind, "kernel_ms": k_ms, "algorithmic_bytes": alg}
    roofline, other = fwd_roof, []
    if train:
        _, _, st = ops.mlp_forward(esegs, w_edge, E, model.prec, residual=e_lat, want_raw=False, want_sum=True, stash=True)
        packs = {}
        c_ms = time_kernel(lambda: ops.dgrad_chain(w_edge, st, g_lat, 128, model.prec, residual=g_lat, pack_cache=packs))
        c_alg = 512 * 7 * E                       # read dy, a2, a1, residual; write dA2, dA1, dIn0
        c_traffic, c_src = ncu_traffic("edge_chain")
        chain_roof = {"kernel": "mlp_tc_kernel<BWD=1>: dgrad chain of the fused edge block",
                      "bound": "hbm", "achieved": c_alg / (c_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": c_alg / (c_ms * 1e-3) / 1e9 / hbm_peak, "traffic": c_traffic, "traffic_source": c_src,
                      "peak_kind": peak_kind, "kernel_ms": c_ms, "algorithmic_bytes": c_alg}
        s_ms = time_kernel(lambda: ops.mlp_forward(esegs, w_edge, E, model.prec, residual=e_lat, want_raw=False, want_sum=True, stash=True))
        s_alg = alg + 512 * 3 * E + 4 * E          # + stash a1, a2, x-hat, rstd
        s_traffic, s_src = ncu_traffic("edge_stash")
        # the line's `roofline` = the dominant kernel family of the training step by launch-list share
        # (profiles/r02*_train_launch_shares.md: mlp_tc_kernel<BWD=0> forward + stash launches)
        roofline = {"kernel": "mlp_tc_kernel<BWD=0>: fused edge block forward + training stash (dominant family of the step)",
                    "bound": "hbm", "achieved": s_alg / (s_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": s_alg / (s_ms * 1e-3) / 1e9 / hbm_peak, "traffic": s_traffic, "traffic_source": s_src,
                    "peak_kind": peak_kind, "kernel_ms": s_ms, "algorithmic_bytes": s_alg}
        wout = torch.empty(128, 128, device=dev)
        w_ms = time_kernel(lambda: ops.wgrad(Seg(g_lat), [Seg(st.a1)], E, wout, b_act=1))
        w_alg = 512 * 2 * E
        v_ms = time_kernel(lambda: P.vertex_half_sum(e_lat, topo))
        v_alg = 512 * E + 256 * V + 4 * (2 * E + V + 1)      # read e once, write vsum, CSR offsets + perm
        other = [fwd_roof, chain_roof,
                 {"kernel": "wgrad_lean_kernel: dW = dA^T SiLU(a) (tcgen05 split-bf16, MN-major operands; incl. launch + split-K reduction)", "kernel_ms": w_ms,
                  "algorithmic_bytes": w_alg, "achieved": w_alg / (w_ms * 1e-3) / 1e9, "frac": w_alg / (w_ms * 1e-3) / 1e9 / hbm_peak},
                 {"kernel": "segment_sum_kernel<16>: deterministic edge->vertex half-sums over the receiver-sorted CSR (gather / segment-sum phase)",
                  "kernel_ms": v_ms, "algorithmic_bytes": v_alg, "achieved": v_alg / (v_ms * 1e-3) / 1e9,
                  "frac": v_alg / (v_ms * 1e-3) / 1e9 / hbm_peak}]

    # ---- max over ranks, aggregate -------------------------------------------------------------------
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(E)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(t[0]), float(t[1])
    E_total = float(tot[0])

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and model_name in ("FvgnA", "MgnA"):   # models the oracle trains
        cpu_baseline = time_cpu_baseline(model_name, n_meshes, n_cells, kind, train)

    strong, rollouts = None, None
    if args.strong_4m == "on" or (args.strong_4m == "auto" and args.workload == DEFAULT_WORKLOAD):
        del model, opt, gd, cache
        torch.cuda.empty_cache()
        strong = strong_scaling_4m(world, rank, dev, dist, steps=max(args.steps, 20) if world > 1 else max(5, min(args.steps, 10)),
                                   halo=args.halo)
        if world == 1:
            # the second half of BASELINE.json's metric (rollout steps/sec) for configs[0] and configs[2], same run
            rollouts = {}
            for wl, n_steps in (("mgn_rollout_2k", 100), ("flux_rollout_200k", 20), ("cons_rollout_200k", 20)):
                m = measure_rollout(wl, 1, 0, dev, dist, steps=n_steps, warmup=5)
                rollouts[wl] = {"model": m["model"], "cells": m["N"], "faces": m["E"], "ms_per_step": m["ms"],
                                "rollout_steps_per_s": 1e3 / m["ms"], "edge_updates_per_s": m["E"] * MP_NUM / (m["ms"] * 1e-3),
                                "e2e_ms_per_step": m["ms_e2e"], "execution": "CUDA-graph replay, velocity read back per step in e2e",
                                "clocks": m["clocks"]}

    if rank == 0:
        line = {
            "metric": "processor edge-updates/sec", "value": E_total * MP_NUM / (ms * 1e-3),
            "unit": "edge-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"f32": "f32", "bf16x3": "bf16x3 (split-bf16 operands, fp32 accumulate)"}.get(prec, prec),
            "data": "synthetic",
            "config": workload_config(args.workload, N, E, V),
            "details": {"precision": prec, "forward_only_ms": fwd_ms,
                        "backward_precision": "dgrad and wgrad split-bf16 (bf16x3), fp32 accumulate" if train else None},
            "e2e": {"value": E_total * MP_NUM / (ms_e2e * 1e-3), "unit": "edge-updates/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                    "h2d_bytes_first_step": h2d_full,
                    "api": ("GraphCache.fetch(per-mesh pinned host samples: device-side collation, static mesh data resident "
                            "after the first step) -> " +
                            ("model.forward + model.loss + backward + Adam step, loss read back"
                             if train else "model.forward(graphs, mode='train'), result read back"))},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline, "roofline_other_kernels": other,
            "roofline_forward": whole_pass_roofline(algorithmic_bytes_forward(E, N, V), fwd_ms if train else ms, hbm_peak,
                                                    "whole forward (encoder + 15 GN_Blocks + decoder): 3 096 B per edge-update, SURVEY.md 8d"),
            "roofline_step": (whole_pass_roofline(3 * algorithmic_bytes_forward(E, N, V), ms, hbm_peak,
                                                  "whole training step, counted as 3 x the forward's algorithmic bytes "
                                                  "(forward + input-gradient pass + weight-gradient pass)") if train else None),
            "cpu_baseline": cpu_baseline,
        }
        if strong is not None:
            line["strong_4m"] = strong
        if rollouts is not None:
            line["rollouts"] = rollouts
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def whole_pass_roofline(alg_bytes, ms, peak, what):
    if ms is None:
        return None
    ach = alg_bytes / (ms * 1e-3) / 1e9
    return {"what": what, "bound": "hbm", "algorithmic_bytes": alg_bytes, "ms": ms, "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak}


def measure_rollout(workload, world, rank, dev, dist, steps, warmup, halo="nccl", prec="bf16x3", single_gpu_too=False):
    """K autoregressive steps of one mesh.  world == 1: RolloutEngine (CUDA-graph replay); world > 1: the mesh is
    domain-decomposed, one partition per GPU, halo exchange per GN_Block (strong scaling).  ``single_gpu_too``: rank 0
    also times the WHOLE mesh on its own GPU first (the N = 1 point of the strong-scaling curve, same run, same box).
    Returns a dict of measurements (times are the max over ranks)."""
    from helpers import build_model
    from gnn_fluid_dynamics_b200 import ops
    from gnn_fluid_dynamics_b200.mesh import make_mesh, mesh_graphs
    from gnn_fluid_dynamics_b200.rollout import RolloutEngine
    model_name, _, n_cells, kind, _ = WORKLOADS[workload]
    model = build_model(model_name, precision=prec).to(dev).eval()
    mesh = make_mesh(n_cells, kind, seed=0)
    cons = model_name.startswith("Conservative")
    g = mesh_graphs(mesh, seed=100, flavour="conservative" if cons else "fvgn")
    if model_name == "MgnA":
        g[0].y = torch.cat([g[0].y, torch.zeros(g[0].x.shape[0], 1)], 1)
        g[1].y = g[1].y[:, :2].contiguous()
    elif cons:
        g[1].y = g[1].y[:, :3].contiguous()
    g = _with_batch(g)
    N, E, V = g[0].x.shape[0], g[0].edge_index.shape[1], g[2].pos.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step, n_steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_steps):
            step()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n_steps

    n1_ms = None
    if world > 1 and single_gpu_too:
        if rank == 0:
            eng1 = RolloutEngine(model, [t.to(dev) for t in g], cuda_graph=True, need_cell_csr=cons, two_hop=not cons)
            for _ in range(3):
                eng1.step()
            torch.cuda.synchronize()
            n1_ms = timed(eng1.step, max(3, min(steps, 10)))
            del eng1
            torch.cuda.empty_cache()
        barrier()
    halo_bytes = 0
    if world > 1:
        from gnn_fluid_dynamics_b200.dist import PartitionedRollout, TorchDistTransport
        from gnn_fluid_dynamics_b200.partition import local_graphs, partition_mesh
        parts = partition_mesh(g[0].edge_index, g[2].edge_index, g[2].face, g[0].pos[:, 0], world, f_face=g[1].face)
        part = parts[rank]
        transport = TorchDistTransport()
        peer = None
        if halo == "peer":
            from gnn_fluid_dynamics_b200.dist import PeerBuffers, peer_indices
            bufs = PeerBuffers(max(p.n_owned for p in parts), 128, dev, world, rank)
            peer = (bufs,) + peer_indices(part, {a: parts[a].send[rank] for a in part.recv}, dev)
        eng = PartitionedRollout(model, [part], [[t.to(dev) for t in local_graphs(g, part)]], transport, peer=peer)
        out_rows = part.n_owned
    else:
        eng = RolloutEngine(model, [t.to(dev) for t in g], cuda_graph=True, need_cell_csr=cons, two_hop=not cons)
        out_rows = N
    step = eng.step
    host_out = torch.empty(out_rows, 2).pin_memory()
    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    ops.LAUNCHES = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    barrier()
    ms = a.elapsed_time(b) / steps
    launches = ops.LAUNCHES
    clocks = sampler.stop()
    if world > 1:
        halo_bytes = transport.bytes_sent // (steps + warmup)
    # e2e: every step's velocity field is read back to pinned host memory (what the reference's writer consumes)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        v = step()
        host_out.copy_(v[0] if isinstance(v, list) else v, non_blocking=False)
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del eng
    torch.cuda.empty_cache()
    return {"model": model_name, "N": N, "E": E, "V": V, "ms": float(t[0]), "ms_e2e": float(t[1]), "launches": launches,
            "clocks": clocks, "halo_bytes": halo_bytes, "out_rows": out_rows, "n1_ms": n1_ms, "halo": halo, "prec": prec}


def strong_scaling_4m(world, rank, dev, dist, steps, halo):
    """BASELINE.json configs[3] next to the headline number: the 4M-cell MGN rollout, domain-decomposed over the run's
    N GPUs (strong scaling), so the driver's 1/2/4/8-GPU runs record it.  On N > 1 rank 0 first times the whole mesh on
    one GPU, which makes the efficiency self-contained (same box, same clocks)."""
    m = measure_rollout("mgn_rollout_4m", world, rank, dev, dist, steps=steps, warmup=3, halo=halo, single_gpu_too=True)
    out = {"workload": "mgn_rollout_4m", "cells": m["N"], "faces": m["E"], "n_gpus": world, "steps": steps,
           "ms_per_step": m["ms"], "edge_updates_per_s": m["E"] * MP_NUM / (m["ms"] * 1e-3),
           "rollout_steps_per_s": 1e3 / m["ms"], "scaling": "strong",
           "halo": ("NCCL send/receive of ghost-cell latents, once per GN_Block + once per step" if world > 1 else None),
           "halo_bytes_per_step_rank0": m["halo_bytes"], "clocks": m["clocks"], "e2e_ms_per_step": m["ms_e2e"]}
    if m["n1_ms"] is not None:
        out["n1_ms_per_step"] = m["n1_ms"]
        out["efficiency_vs_n1"] = m["n1_ms"] / (world * m["ms"])
    return out


def run_rollout(args, world, rank, dev, dist):
    """Rollout workloads as the bench line (--workload *_rollout_*)."""
    prec = args.precision or "bf16x3"
    m = measure_rollout(args.workload, world, rank, dev, dist, args.steps, args.warmup, args.halo, prec)
    model_name, N, E, V, ms, ms_e2e = m["model"], m["N"], m["E"], m["V"], m["ms"], m["ms_e2e"]
    halo_bytes, out_rows, launches, clocks = m["halo_bytes"], m["out_rows"], m["launches"], m["clocks"]
    if rank == 0:
        line = {
            "metric": "processor edge-updates/sec", "value": E * MP_NUM / (ms * 1e-3), "unit": "edge-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "bf16x3 (split-bf16 operands, fp32 accumulate)" if prec == "bf16x3" else prec, "data": "synthetic",
            "config": {"workload": args.workload, "model": model_name, "mp_num": MP_NUM, "hidden": 128, "cells": N,
                       "faces": E, "vertices": V, "precision": prec, "rollout_steps_per_s": 1e3 / ms,
                       "timed": "one autoregressive step: normalise + encoder + 15 GN_Blocks + decoder (+ integrator) + state advance",
                       "execution": (("domain-decomposed, one partition per GPU; ghost latents gathered from the owner's HBM over "
                                      "NVLink inside the fused edge kernel (peer-memory), 1 barrier per GN_Block; "
                                      if args.halo == "peer" else
                                      "domain-decomposed, one partition per GPU, 1 NCCL halo exchange per GN_Block + 1 per step; ")
                                     + f"{halo_bytes} bytes sent through NCCL per step by rank 0") if world > 1 else "CUDA-graph replay",
                       "l2": "working set > L2" if N > 100000 else "small mesh: L2-resident by nature of the workload"},
            "e2e": {"value": E * MP_NUM / (ms_e2e * 1e-3), "unit": "edge-updates/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": out_rows * 8, "ms_per_step": ms_e2e,
                    "api": "RolloutEngine.step / PartitionedRollout.step + velocity read back to pinned host memory"},
            "gpu_launches": launches if world > 1 else "CUDA graph (kernels of one captured step replayed)",
            "clocks": clocks, "roofline": None, "cpu_baseline": None,
            "roofline_forward": whole_pass_roofline(
                MP_NUM * ((512 * (3 * E + 4 * N + V) + 16 * E + 12 * N) if not model_name.startswith("Conservative")
                          else (512 * (3 * E + 4 * N) + 16 * E)), ms, peaks()[0],
                "one rollout step over the processor's algorithmic bytes (SURVEY.md 8d; encoder / decoder / glue not counted)"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_cpu_baseline(model_name, n_meshes, n_cells, kind, train=False):
    """Oracle port of the reference's CPU path on this box's host cores, bounded sample (1 mesh)."""
    from fixtures import state_dict_from_keys
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    graphs = build_batch(model_name, 1, n_cells, kind)
    E = graphs[0].edge_index.shape[1]
    step = _oracle_step_fn(model_name, state_dict_from_keys(model_name), graphs, train)
    best = None
    for i in range(4):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i > 0:
            best = dt if best is None else min(best, dt)
    what = "forward + loss + backward + Adam step" if train else "whole forward"
    return {"value": E * MP_NUM / best, "unit": "edge-updates/s", "cores": cores, "kind": "port",
            "sample": f"1 of {n_meshes} meshes ({n_cells}-cell {kind}), {what}, best of 3 after 1 warm-up"}


if __name__ == "__main__":
    main()
