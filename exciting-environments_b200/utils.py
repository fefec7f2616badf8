"""Host-side helpers mirroring ``exciting_environments/utils.py``: the ``MinMaxNormalization`` container
(reference utils.py:8-17) and the JSON (de)serialisation of simulation properties (utils.py:21-52) used by the
golden-vector fixtures.  Plus the tiny pytree helpers the drop-in classes need (the reference relies on
jax.tree_util / jax_dataclasses; neither is available here)."""
from __future__ import annotations

import dataclasses
import json
from dataclasses import asdict, dataclass, fields, is_dataclass

import torch


@dataclass
class MinMaxNormalization:
    min: object
    max: object

    def normalize(self, denormalized_value):
        return 2 * (denormalized_value - self.min) / (self.max - self.min) - 1

    def denormalize(self, normalized_value):
        return (normalized_value + 1) / 2 * (self.max - self.min) + self.min


def dump_sim_properties_to_json(params, action_normalizations, physical_normalizations, tau, filename):
    data = {
        "params": params,
        "action_normalizations": {k: asdict(v) for k, v in action_normalizations.items()},
        "physical_normalizations": {k: asdict(v) for k, v in physical_normalizations.items()},
        "tau": tau,
    }
    with open(filename, "w") as f:
        json.dump(data, f, indent=4)


def load_sim_properties_from_json(filename):
    with open(filename, "r") as f:
        data = json.load(f)
    an = {k: MinMaxNormalization(**v) for k, v in data["action_normalizations"].items()}
    pn = {k: MinMaxNormalization(**v) for k, v in data["physical_normalizations"].items()}
    return data["params"], an, pn, data["tau"]


# ---- minimal pytree support (dataclass field order == jax_dataclasses flatten order) ---------------
_FIELD_NAMES = {}  # class -> dataclass field names (None for non-dataclasses); reflection once per class


def _field_names(cls):
    try:
        return _FIELD_NAMES[cls]
    except KeyError:
        names = tuple(f.name for f in fields(cls)) if is_dataclass(cls) else None
        _FIELD_NAMES[cls] = names
        return names


def tree_flatten(tree):
    """Leaves in dataclass-field / tuple order; ``None`` is an empty subtree (as in jax.tree_util)."""
    leaves = []

    def rec(x):
        if x is None:
            return ("none",)
        cls = type(x)
        if cls is tuple or cls is list:
            return ("seq", cls, tuple(rec(v) for v in x))
        if cls is not type:
            names = _field_names(cls)
            if names is not None:
                return ("dc", cls, tuple(rec(getattr(x, n)) for n in names))
        if isinstance(x, (tuple, list)):
            return ("seq", cls, tuple(rec(v) for v in x))
        leaves.append(x)
        return ("leaf",)

    struct = rec(tree)
    return leaves, struct


def tree_unflatten(struct, leaves):
    it = iter(leaves)

    def rec(s):
        kind = s[0]
        if kind == "none":
            return None
        if kind == "leaf":
            return next(it)
        if kind == "dc":
            cls, subs = s[1], s[2]
            return cls(*[rec(c) for c in subs])
        cls, subs = s[1], s[2]
        return cls(rec(c) for c in subs)

    return rec(struct)


def tree_structure(tree):
    return tree_flatten(tree)[1]


def tree_map(fn, tree):
    leaves, struct = tree_flatten(tree)
    return tree_unflatten(struct, [fn(x) for x in leaves])


def is_array(x):
    return isinstance(x, torch.Tensor) and x.ndim >= 1
