"""Host-side mirror of the reference's model interface (``src/models/Model.py``).

Same constructor ``Model(config, loss_func, dataset, stats)``, same ``forward(graphs, mode) -> dict``,
``loss``, ``update_features``, classmethods ``get_feature_sizes`` / ``get_normalisation_map`` and - the
hard constraint - the same ``state_dict`` keys and shapes, so the reference's checkpoints load and
``config.model.module = "gnn_fluid_dynamics_b200.models.Fvgn"`` is the whole switch-over
(``src/train.py:348-349``).  The parameters live in ordinary ``nn.Sequential`` / ``nn.LayerNorm``
containers (``monitoring.py:18`` iterates ``decoder.face_mlp``); the arithmetic of
encoder/processor/decoder is done by the CUDA kernels through ``processor.py``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch import nn

from .._lib import PRECISIONS

# default GEMM arithmetic of the hot path; "f32" = exact CUDA-core path, others = tcgen05
DEFAULT_PRECISION = "bf16x3"


def build_mlp(config, in_size, hidden_size, out_size, norm_layer=True):
    """Parameter container with the reference's layout (Model.py:12-40): ``Sequential(Linear, SiLU,
    Linear, SiLU, Linear)`` (indices 0,2,4), wrapped as ``Sequential(mlp, LayerNorm)`` when normalised.
    With ``config.training.dropout_rate > 0`` the reference inserts a Dropout after each SiLU (Linear
    indices 0,3,6); the same layout is built here so that its checkpoints load.  Dropout is the identity
    in eval mode, which is all the fused kernels implement: a TRAINING forward of such a model raises
    (``processor.weights_of``); 0.0 in every shipped config."""
    rate = getattr(getattr(config, "training", None), "dropout_rate", 0.0) or 0.0
    layers = [nn.Linear(in_size, hidden_size), nn.SiLU()]
    if rate > 0:
        layers.append(nn.Dropout(p=rate))
    layers += [nn.Linear(hidden_size, hidden_size), nn.SiLU()]
    if rate > 0:
        layers.append(nn.Dropout(p=rate))
    layers.append(nn.Linear(hidden_size, out_size))
    mlp = nn.Sequential(*layers)
    if norm_layer:
        return nn.Sequential(mlp, nn.LayerNorm(normalized_shape=out_size))
    return mlp


def build_mlp_antisym(config, in_size, hidden_size, out_size):
    """Bias-free Tanh MLP, no LayerNorm (Conservative.py:31-43; every call site passes
    norm_layer=False)."""
    return nn.Sequential(nn.Linear(in_size, hidden_size, bias=False), nn.Tanh(),
                         nn.Linear(hidden_size, hidden_size, bias=False), nn.Tanh(),
                         nn.Linear(hidden_size, out_size, bias=False))


class Normalizer(nn.Module):
    """Per-column affine (de)normalisation with the reference's buffer names ``{key}_{stat}``
    (normalisation.py:207-278).  ``inputs`` rows are (graph index, attribute, column slice, stats
    key); ``outputs`` rows are (output index, column slice, stats key); ``kinds`` maps a stats key to
    z_score | mean_scale | std_scale | max_scale | min_max (normalisation.py:281-322).
    """

    def __init__(self, stats, kinds, inputs, outputs):
        super().__init__()
        self.kinds = dict(kinds)
        self.inputs = list(inputs)
        self.outputs = list(outputs)
        for key, stat in stats.items():
            for name, value in stat.items():
                self.register_buffer(f"{key}_{name}", torch.tensor(value, dtype=torch.float))

    def _apply_one(self, data, key, inverse):
        kind = self.kinds[key]
        g = lambda s: getattr(self, f"{key}_{s}")
        if kind == "z_score":
            scale = torch.clamp(g("std"), min=1e-8) + 1e-8
            return data * scale + g("mean") if inverse else (data - g("mean")) / scale
        if kind == "mean_scale":
            scale = g("mean") + 1e-8
            return data * scale if inverse else data / scale
        if kind == "std_scale":
            scale = g("std") + 1e-8
            return data * scale if inverse else data / scale
        if kind == "max_scale":
            scale = g("max") + 1e-8
            return data * scale if inverse else data / scale
        if kind == "min_max":
            rng = g("max") - g("min") + 1e-8
            return data * rng + g("min") if inverse else (data - g("min")) / rng
        raise ValueError(kind)

    # --- fused path: one kernel per tensor instead of ~4 tensor kernels per normalised column --------------------
    def _shift_scale(self, key):
        """(shift, scale) 0-dim tensors of a stats key: forward = (x - shift) / scale, inverse = x * scale + shift."""
        kind = self.kinds[key]
        g = lambda s: getattr(self, f"{key}_{s}")
        if kind == "z_score":
            return g("mean"), torch.clamp(g("std"), min=1e-8) + 1e-8
        if kind in ("mean_scale", "std_scale", "max_scale"):
            scale = g({"mean_scale": "mean", "std_scale": "std", "max_scale": "max"}[kind]) + 1e-8
            return torch.zeros_like(scale), scale
        if kind == "min_max":
            return g("min"), g("max") - g("min") + 1e-8
        raise ValueError(kind)

    def _spec(self, rows, width):
        """Device arrays (columns, shift, scale) of the table rows that apply to a tensor of ``width`` columns, cached until
        a statistics buffer is replaced or modified.  None: overlapping columns (the sequential tensor path handles them)."""
        bufs = tuple(self.buffers())
        key = (tuple(rows), width, tuple((b.data_ptr(), b._version) for b in bufs))
        cache = self.__dict__.setdefault("_spec_cache", {})
        hit = cache.get(key)
        if hit is None:
            cols, shift, scale = [], [], []
            for c, k in rows:
                if width < c.stop:
                    continue
                a, b = self._shift_scale(k)
                for j in range(c.start, c.stop):
                    cols.append(j); shift.append(a); scale.append(b)
            if len(set(cols)) != len(cols):
                hit = (None,)
            elif not cols:
                hit = ((None, None, None, 0),)
            else:
                dev = shift[0].device
                hit = ((torch.tensor(cols, dtype=torch.int32, device=dev), torch.stack(shift).float().contiguous(),
                        torch.stack(scale).float().contiguous(), len(cols)),)
            if len(cache) > 64:
                cache.clear()
            cache[key] = hit
        return hit[0]

    def _apply_fused(self, t, rows, inverse) -> bool:
        """True if the tensor was (de)normalised by ``gnnfd_affine_columns`` (CUDA fp32 [R, C] with contiguous columns)."""
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] >= 1
                and (t.stride(1) == 1 or t.shape[1] == 1) and not (torch.is_grad_enabled() and t.requires_grad)):
            return False
        spec = self._spec(rows, t.shape[1])
        if spec is None or spec[0] is not None and spec[0].device != t.device:
            return False
        cols, shift, scale, n = spec
        if n and t.shape[0]:
            from .. import ops
            from .._lib import check, lib
            check(lib.gnnfd_affine_columns(t.data_ptr(), t.shape[0], t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1]),
                                           n, cols.data_ptr(), shift.data_ptr(), scale.data_ptr(), int(inverse), ops._stream()),
                  "gnnfd_affine_columns")
            ops._count(1)
        return True

    def input(self, graphs, inverse=False):
        """In place on the passed graphs, like the reference (normalisation.py:255-264)."""
        groups = {}
        for gi, attr, cols, key in self.inputs:
            groups.setdefault((gi, attr), []).append((cols, key))
        for (gi, attr), rows in groups.items():
            t = getattr(graphs[gi], attr, None)
            if t is None or t.dim() < 2:
                continue
            if self._apply_fused(t, rows, inverse):
                continue
            for cols, key in rows:
                if t.shape[-1] >= cols.stop:
                    t[..., cols] = self._apply_one(t[..., cols], key, inverse)      # last dim: [R, C] or bundled [R, k, C]
        return graphs

    def output(self, outputs, inverse=False):
        groups = {}
        for oi, cols, key in self.outputs:
            groups.setdefault(oi, []).append((cols, key))
        for oi, rows in groups.items():
            t = outputs[oi]
            if t is None:
                continue
            if self._apply_fused(t, rows, inverse):
                continue
            for cols, key in rows:
                if t.shape[-1] >= cols.stop:
                    t[..., cols] = self._apply_one(t[..., cols], key, inverse)
        return outputs


def col(i, j=None):
    return slice(i, i + 1 if j is None else j)


class Model(nn.Module, ABC):
    """Base class: configuration, feature sizes, normaliser, precision switch."""
    cell_grad_weights_use = False
    face_grad_weights_use = False
    pushforward_use = False
    family = None  # processor data-flow, see processor.py

    def __init__(self, config, loss_func, dataset, stats):
        super().__init__()
        self.config = config
        self.loss_func = loss_func
        self.hidden_size = config.model.hidden_width
        if self.hidden_size != 128:
            raise NotImplementedError("the B200 kernels are built for hidden_width = 128 "
                                      "(every shipped reference config, config/train.json:27)")
        self.input_sizes, self.output_sizes = self.get_feature_sizes(dataset)
        kinds, inputs, outputs = self.normalisation_tables()
        self.normalizer = Normalizer(stats, kinds, inputs, outputs)
        self.precision = getattr(config.model, "precision", None) or DEFAULT_PRECISION

    @property
    def prec(self) -> int:
        return PRECISIONS[self.precision]

    def set_precision(self, name: str):
        if name not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = name
        return self

    def mse_term(self, output, target, mask, batch=None):
        """One loss term.  The reference's loss callables (``utils/loss.py:16-60``: element-wise masked MSE, which
        ``train.py:372`` passes) run as ONE fused kernel without boolean-mask indexing (``fvm_ops.masked_mse``); any
        other callable is applied as given."""
        from ..fvm_ops import is_plain_mse, masked_mse
        if output.is_cuda and is_plain_mse(self.loss_func) and output.dtype == torch.float32:
            return masked_mse(output, target, mask)
        return self.loss_func(output, target, mask, batch)

    def wants_grad(self) -> bool:
        """True when the hot path must record its backward (training step, reference src/train.py:253-256)."""
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.processer_list.parameters())

    def training_plan(self):
        """MLP sites of encoder / processor / decoder for the kernel-scheduled backward (training.py)."""
        from ..training import Plan, Site
        plan = getattr(self, "_gnnfd_plan", None)
        if plan is None:
            blocks = [(Site(b.cell_block.cell_mlp), Site(b.face_block.face_mlp)) for b in self.processer_list]
            plan = Plan(self.family, Site(self.encoder.face_mlp), Site(self.encoder.cell_mlp), blocks,
                        Site(self.decoder.face_mlp))
            object.__setattr__(self, "_gnnfd_plan", plan)
        return plan

    @classmethod
    @abstractmethod
    def get_feature_sizes(cls, dataset):
        ...

    @classmethod
    @abstractmethod
    def normalisation_tables(cls):
        """(kinds {stats key: kind}, inputs [(graph, attr, cols, key)], outputs [(out, cols, key)])."""
        ...

    @classmethod
    def transform_features(cls, dataset, graphs, mesh_id=None):
        """Raw time series -> network inputs / targets, as the reference's dataset calls it on the model class
        (``src/train.py:350``); see ``models/features.py``."""
        from .features import transform_for
        return transform_for(cls, dataset, graphs)

    @classmethod
    def get_normalisation_map(cls):
        """The reference's (registry, inputs, outputs) dict-of-lambdas form of the same tables
        (consumed by its statistics accumulator, ``src/datasets/DataSet.py:318``)."""
        kinds, inputs, outputs = cls.normalisation_tables()
        registry = {}
        for gi, attr, cols, key in inputs:
            registry.setdefault(key, ((lambda g, gi=gi, attr=attr, cols=cols: getattr(g[gi], attr)[:, cols]),
                                      kinds[key]))
        for key, fn in getattr(cls, "_registry_overrides", {}).items():
            registry[key] = (fn, kinds[key])
        for key, kind in kinds.items():
            registry.setdefault(key, ((lambda g: None), kind))
        ins = {f"in{i}_{key}": ((lambda g, gi=gi, attr=attr, cols=cols: getattr(g[gi], attr)[:, cols]), key)
               for i, (gi, attr, cols, key) in enumerate(inputs)}
        outs = {f"out{i}_{key}": ((lambda o, oi=oi, cols=cols: o[oi][:, cols]), key)
                for i, (oi, cols, key) in enumerate(outputs)}
        return registry, ins, outs

    @abstractmethod
    def forward(self, graphs, mode="train"):
        ...

    def count_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


def n_class_types(dataset) -> int:
    ct = getattr(dataset, "class_types", None)
    return len(ct) if ct is not None else 5
