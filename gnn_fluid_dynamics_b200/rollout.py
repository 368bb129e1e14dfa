"""Autoregressive rollout of the hot path (reference ``src/rollout.py:304-369``, the inner loop only: the
error metrics, HDF5 writer and dataloader around it are out of scope).

One step = ``model([g.clone() ...], mode='rollout')`` -> ``cell_velocity = x[:, :2] + cell_velocity_change``
-> ``model.update_features``.  The mesh is static, so its CSRs are built once (``attach_topology``) and - since
every kernel of the path is stream-ordered, allocation-free and host-sync-free - the whole step is captured
once in a CUDA graph and replayed: on small meshes (BASELINE.json config 1, 2k cells) the step is
launch-latency bound (~60 kernels of a few microseconds), which the graph removes.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from ._lib import choose_launch_overlap
from .topology import attach_topology


class _FusedStep:
    """One rollout step of FvgnA / MgnA / FluxA without the per-step graph clones and per-column (de)normalisation kernels:

      forward on RESIDENT normalised inputs  ->  [integrator]  ->  de-normalise the predicted change (one kernel)  ->
      ``fvm_ops.state_advance`` (two kernels: rollout.py:340 + update_features, Fvgn.py:133-148 / Mgn.py:139-151, + the
      z-scoring of the next step's inputs, normalisation.py:255-278)

    The reference's loop body re-clones the three graphs, z-scores every input column and inverts every output column
    with a handful of tiny tensor kernels each (~100 launches of a few microseconds per step); here the raw state
    (``c_graph.x``, ``f_graph.x[:, :2]``) and its normalised copy are both kept in HBM and advanced together."""

    KINDS = ("FvgnA", "MgnA", "FluxA")

    @staticmethod
    def supports(model, graphs) -> bool:
        c, f, _ = graphs
        if type(model).__name__ not in _FusedStep.KINDS or c.x.shape[1] != 2 or "x" not in f:
            return False
        keys = ("cell_velocity_x", "cell_velocity_y", "face_velocity_difference_x", "face_velocity_difference_y",
                "cell_velocity_change_x", "cell_velocity_change_y")
        nz = model.normalizer
        return all(nz.kinds.get(k) == "z_score" and hasattr(nz, f"{k}_mean") and hasattr(nz, f"{k}_std") for k in keys)

    def __init__(self, model, graphs, topo):
        from .mesh import NODE_INFLOW, NODE_WALL
        self.model, self.graphs, self.topo = model, graphs, topo
        c, f, v = graphs
        nz = model.normalizer
        ms = lambda k: (float(getattr(nz, f"{k}_mean")), float(torch.clamp(getattr(nz, f"{k}_std"), min=1e-8) + 1e-8))
        self.cell_stats = ms("cell_velocity_x") + ms("cell_velocity_y")
        self.face_stats = ms("face_velocity_difference_x") + ms("face_velocity_difference_y")
        dev = c.x.device
        (m0, s0), (m1, s1) = ms("cell_velocity_change_x"), ms("cell_velocity_change_y")
        self.out_mean = torch.tensor([m0, m1], dtype=torch.float32, device=dev)
        self.out_scale = torch.tensor([s0, s1], dtype=torch.float32, device=dev)
        # normalised copies of the inputs: the static columns are z-scored once, the state columns every step
        gn = nz.input([g.clone() for g in graphs])
        self.x_norm, self.f_norm = gn[0].x.contiguous(), gn[1].x.contiguous()
        self.c_norm = gn[0]                       # normals / volumes / dt for the integrator (untouched by the z-scoring)
        self.c_norm.topology = topo
        self.f_static = gn[1]
        if type(model).__name__ in ("FvgnA", "FluxA"):      # FluxA inherits FvgnA.update_features
            mask = ((f.type == NODE_INFLOW) | (f.type == NODE_WALL)).reshape(-1)
        else:
            mask = f.boundary_mask.reshape(-1)
        self.mask = mask.to(torch.uint8).contiguous()
        self.vel = torch.empty(c.x.shape[0], 2, dtype=torch.float32, device=dev)
        if not (c.x.is_contiguous() and f.x.is_contiguous()):
            c.x, f.x = c.x.contiguous(), f.x.contiguous()
        if type(model).__name__ == "FluxA":
            # FluxA's integrator (Flux.py:166-206): the mesh, dt and the evaluation-mode BatchNorm statistics are fixed
            # over a rollout, so the normalised mean(dt) / face-volume coefficient and the normalised face area are
            # computed ONCE here; per step the integrator is then one kernel (fvm_ops.flux_integrate)
            from .fvm_ops import cell_faces
            from .models.Flux import normalize_vol_dt
            from .models.Fvgn import normalize_face_area
            ig = model.integrator
            with torch.no_grad():
                self.flux_coeff = normalize_vol_dt(gn[0].volume, gn[0].edge_index, gn[0].dt, ig.vol_dt_norm, topo=topo).reshape(-1).contiguous()
                self.flux_area = normalize_face_area(gn[1].area, gn[0].volume, gn[0].edge_index, gn[0].dt, ig.face_area_norm,
                                                     topo=topo).reshape(-1).contiguous()
            self.flux_cf = cell_faces(topo, gn[1].face)

    @torch.no_grad()
    def step(self) -> torch.Tensor:
        from . import fvm_ops
        model, (c, f, _), topo = self.model, self.graphs, self.topo
        x, e, dec = model.encode_process_decode(self.x_norm, self.f_norm, topo)
        if type(model).__name__ == "FvgnA":
            self.c_norm.x = self.x_norm
            change = model.integrator(dec, self.c_norm, self.f_static, self.c_norm.dt)      # normalised change [N, 2]
        elif type(model).__name__ == "FluxA":
            change = fvm_ops.flux_integrate(dec, self.flux_coeff, self.flux_area, self.c_norm.normal, self.flux_cf,
                                            topo.row, topo.col, model.integrator.rho)
        else:
            change = dec[:, 0:2]
        delta = torch.addcmul(self.out_mean, change, self.out_scale)                       # normalizer.output(inverse)
        fvm_ops.state_advance(c.x, delta, True, topo.row, topo.col, f.x, self.mask, f.y, x_norm=self.x_norm,
                              cell_stats=self.cell_stats, f_norm=self.f_norm, face_stats=self.face_stats, vel_out=self.vel)
        return self.vel


class RolloutEngine:
    def __init__(self, model, graphs, cuda_graph: bool = True, need_cell_csr: bool = False, two_hop: bool = True,
                 fused_step: bool = True):
        """``graphs`` = [c_graph, f_graph, v_graph] on the GPU; its ``c_graph.x`` / ``f_graph.x`` are the rollout
        state and are advanced in place."""
        if graphs[0].x.device.type != "cuda":
            raise RuntimeError("RolloutEngine needs the graphs on a CUDA device (no CPU fallback on this path)")
        self.model = model.eval()
        self.graphs = graphs
        self.topo = attach_topology(graphs, need_cell_csr=need_cell_csr, two_hop=two_hop)
        # forwards that synchronise with the host (FvgnK: boolean-mask indexing / torch.unique on the Reynolds
        # groups, Fvgn.py:1290-1340) cannot be captured: they step eagerly
        self.use_graph = cuda_graph and not getattr(model, "host_sync_in_forward", False)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._vel: Optional[torch.Tensor] = None
        self.face_attr = "x_asym" if "x_asym" in graphs[1] else "x"
        # FvgnA / MgnA: state advance and (de)normalisation as fused kernels on resident buffers (``fused_step=False``
        # steps through the reference's loop body literally: forward on cloned graphs + update_features)
        self._fused = _FusedStep(self.model, graphs, self.topo) if fused_step and _FusedStep.supports(self.model, graphs) else None

    @torch.no_grad()
    def _step_eager(self) -> torch.Tensor:
        """One pass of the reference loop body (rollout.py:313-369).  A model that returns ``cell_velocity`` is taken
        at its word (rollout.py:336-337: MgnB / MgnC / StreamFunc); otherwise velocity = x[:, :2] + change
        (rollout.py:340).  Temporally bundled outputs [N, k, 2] (FvgnC) yield k velocities per forward and the LAST
        one feeds ``update_features`` (rollout.py:319-332, 369); the returned tensor is then [N, k, 2]."""
        if self._fused is not None:
            return self._fused.step()
        c, f, v = self.graphs
        cx = c.x                                                       # static state buffer
        out = self.model([g.clone() for g in self.graphs], mode="rollout")           # rollout.py:313
        if "cell_velocity" in out:
            vel = out["cell_velocity"]
        elif "cell_velocity_change" in out:
            dv = out["cell_velocity_change"]
            vel = cx[:, None, :2] + dv if dv.dim() == 3 else cx[:, :2] + dv
        else:
            raise RuntimeError(f"{type(self.model).__name__}.forward returned neither 'cell_velocity' nor "
                               f"'cell_velocity_change' (keys: {sorted(out)})")
        last = vel[:, -1] if vel.dim() == 3 else vel
        self.model.update_features({"cell_velocity": last}, self.graphs)                # rollout.py:369
        # update_features rebinds c_graph.x to the new tensor; keep the state in the static buffer instead
        cx[:, :2].copy_(last)
        c.x = cx
        return vel

    def _capture(self):
        c, f, _ = self.graphs
        live = [c.x, getattr(f, self.face_attr)]      # everything a step advances in place
        if self._fused is not None:
            live += [self._fused.x_norm, self._fused.f_norm]
        state = [t.clone() for t in live]

        def restore():
            for t, s0 in zip(live, state):
                t.copy_(s0)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                 # warm-up outside capture (lazy kernel attributes, packs)
            self._step_eager()
        torch.cuda.current_stream().wait_stream(s)
        restore()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._vel = self._step_eager()
        restore()

    def step(self) -> torch.Tensor:
        """Advance the state by one timestep; returns the new cell velocity [N, 2] (a static buffer when the
        step is graph-replayed: clone it to keep it)."""
        choose_launch_overlap(self.topo.n_faces, False)     # (a captured graph keeps the policy it was captured with)
        if not self.use_graph:
            return self._step_eager()
        if self._graph is None:
            self._capture()
        self._graph.replay()
        return self._vel

    def run(self, n_steps: int, keep: bool = False) -> List[torch.Tensor]:
        outs = []
        for _ in range(n_steps):
            v = self.step()
            if keep:
                outs.append(v.clone())
        return outs
