"""`hydra.utils.instantiate` restated: import `_target_`, call it with the other keys."""
import importlib


class AttrDict(dict):
    """dict with attribute access (the reference reads `TRANSFORMER.d_model`, denoiser.py:41)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:  # pragma: no cover
            raise AttributeError(key) from exc


def to_attr(node):
    if isinstance(node, dict):
        return AttrDict({k: to_attr(v) for k, v in node.items()})
    return node


def instantiate(config, _recursive_=False, **overrides):
    spec = dict(config)
    target = spec.pop("_target_")
    module_name, _, symbol = target.rpartition(".")
    factory = getattr(importlib.import_module(module_name), symbol)
    kwargs = {k: to_attr(v) for k, v in spec.items()}
    kwargs.update(overrides)
    return factory(**kwargs)


def get_original_cwd():
    return "."
