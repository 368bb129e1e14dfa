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

from .topology import attach_topology


class RolloutEngine:
    def __init__(self, model, graphs, cuda_graph: bool = True, need_cell_csr: bool = False, two_hop: bool = True):
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

    @torch.no_grad()
    def _step_eager(self) -> torch.Tensor:
        """One pass of the reference loop body (rollout.py:313-369).  A model that returns ``cell_velocity`` is taken
        at its word (rollout.py:336-337: MgnB / MgnC / StreamFunc); otherwise velocity = x[:, :2] + change
        (rollout.py:340).  Temporally bundled outputs [N, k, 2] (FvgnC) yield k velocities per forward and the LAST
        one feeds ``update_features`` (rollout.py:319-332, 369); the returned tensor is then [N, k, 2]."""
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
        state = (c.x.clone(), getattr(f, self.face_attr).clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):                 # warm-up outside capture (lazy kernel attributes, packs)
            self._step_eager()
        torch.cuda.current_stream().wait_stream(s)
        c.x.copy_(state[0]); getattr(f, self.face_attr).copy_(state[1])
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._vel = self._step_eager()
        c.x.copy_(state[0]); getattr(f, self.face_attr).copy_(state[1])

    def step(self) -> torch.Tensor:
        """Advance the state by one timestep; returns the new cell velocity [N, 2] (a static buffer when the
        step is graph-replayed: clone it to keep it)."""
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
