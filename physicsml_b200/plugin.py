"""molflux model-zoo entry points for the B200 path (reference ``pyproject.toml:78-101``,
``docs/source/pages/models/how_to_add_models.md:82-147``).

``mace_model_b200`` / ``nequip_model_b200`` are the reference's own ``MACEModel`` / ``NequipModel`` (same
``MACEModelConfig`` / ``NequipModelConfig``, same datamodule, losses, optimiser wiring and checkpoints) whose
``_instantiate_module`` additionally calls :func:`physicsml_b200.patching.patch` on the Lightning module it built, so

    model = load_from_dict({"name": "mace_model_b200", "config": {...}})      # molflux.modelzoo

trains and predicts through the CUDA kernels.  The stock names keep working too: ``physicsml_b200.patch(model.module)``.

The reference package and molflux are imported lazily (they are not a dependency of the kernels); the classes are built
on first attribute access, which is what ``importlib.metadata`` entry-point loading does.
"""
from __future__ import annotations

_CACHE: dict = {}


def _build(name: str):
    from .patching import patch

    if name == "MACEModelB200":
        from physicsml.models.mace.supervised.mace_model import MACEModel as Base
    elif name == "NequipModelB200":
        from physicsml.models.nequip.supervised.nequip_model import NequipModel as Base
    else:
        raise AttributeError(name)

    class _B200(Base):  # type: ignore[misc, valid-type]
        def _instantiate_module(self):
            return patch(super()._instantiate_module())

    _B200.__name__ = _B200.__qualname__ = name
    return _B200


def __getattr__(name: str):
    if name in ("MACEModelB200", "NequipModelB200"):
        if name not in _CACHE:
            try:
                _CACHE[name] = _build(name)
            except ImportError as exc:
                raise ImportError(f"physicsml_b200.plugin.{name} subclasses the reference's model class and needs the "
                                  f"`physicsml` package and `molflux` to be importable ({exc})") from exc
        return _CACHE[name]
    raise AttributeError(name)
