"""mmcv-free stand-in for ``mmcv.utils.Registry`` / ``build_from_cfg`` (mmcv is not a dependency).

The reference registers its PnP op in a local registry, ``PNP = Registry('pnp')``
(monorun/ops/least_squares/builder.py:3-7), and its heads in mmdet's ``HEADS``.  This module offers the
same three calls the reference code uses -- ``Registry(name)``, ``@REG.register_module()`` and
``build_from_cfg(cfg, registry, default_args)`` -- so ``dict(type='PnPUncert', ...)`` blocks copied from
configs/kitti_*.py build unchanged.  When mmcv is importable the real ``PNP``/``HEADS`` registries can be
used instead: see INTEGRATION.md.
"""
import inspect


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return key in self._module_dict

    def __repr__(self):
        return f'{self.__class__.__name__}(name={self._name}, items={list(self._module_dict)})'

    def get(self, key):
        return self._module_dict.get(key)

    def _register(self, cls, name=None, force=False):
        if not inspect.isclass(cls):
            raise TypeError(f'module must be a class, but got {type(cls)}')
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f'{name} is already registered in {self._name}')
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def _deco(cls):
            self._register(cls, name, force)
            return cls
        return _deco


def build_from_cfg(cfg, registry, default_args=None):
    """Same contract as mmcv.utils.build_from_cfg: ``cfg['type']`` names a registered class (or is the
    class itself); remaining keys, completed by ``default_args``, are constructor kwargs."""
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg:
        if default_args is None or 'type' not in default_args:
            raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}\n{default_args}')
    if not isinstance(registry, Registry):
        raise TypeError(f'registry must be a Registry object, but got {type(registry)}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    return obj_cls(**args)


# registries of the reference that live on the hot path
PNP = Registry('pnp')                            # monorun/ops/least_squares/builder.py:3
HEADS = Registry('head')                         # mmdet.models.builder.HEADS
COORD_CODERS = Registry('coord_coder')           # monorun/core/bbox_3d/builder.py:7
PROJ_ERROR_CODERS = Registry('proj_error_coder') # monorun/core/bbox_3d/builder.py:4
DIM_CODERS = Registry('dim_coder')               # monorun/core/bbox_3d/builder.py:3
ROTATION_CODERS = Registry('rotation_coder')     # monorun/core/bbox_3d/builder.py:5
LOSSES = Registry('loss')                        # mmdet.models.builder.LOSSES (training only; stubs)
ROI_EXTRACTORS = Registry('roi_extractor')       # mmdet.models.builder.ROI_EXTRACTORS


def build_pnp(cfg, **default_args):
    """monorun/ops/least_squares/builder.py:6-7."""
    return build_from_cfg(cfg, PNP, default_args)


def build_head(cfg, **default_args):
    return build_from_cfg(cfg, HEADS, default_args)


def build_roi_extractor(cfg, **default_args):
    return build_from_cfg(cfg, ROI_EXTRACTORS, default_args)


def build_coord_coder(cfg, **default_args):
    return build_from_cfg(cfg, COORD_CODERS, default_args)


def build_proj_error_coder(cfg, **default_args):
    return build_from_cfg(cfg, PROJ_ERROR_CODERS, default_args)


def build_dim_coder(cfg, **default_args):
    return build_from_cfg(cfg, DIM_CODERS, default_args)


def build_rotation_coder(cfg, **default_args):
    return build_from_cfg(cfg, ROTATION_CODERS, default_args)
