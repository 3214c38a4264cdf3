"""Loader for the reference's mmcv-style python config files (configs/kitti_*.py are plain modules that assign
dict literals, so ``runpy`` is enough; mmcv's ``Config.fromfile`` is not needed)."""
import runpy


class ConfigDict(dict):
    """dict with attribute access, like mmcv.ConfigDict (``test_cfg.rcnn.cov_correction``)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return ConfigDict(v) if isinstance(v, dict) and not isinstance(v, ConfigDict) else v


def load_config(path):
    ns = runpy.run_path(path)
    return ConfigDict({k: v for k, v in ns.items() if not k.startswith('__')})


def build_roi_head_from_fixture(name='kitti_multiclass.py', path=None):
    """Build ``MonoRUnRoIHead`` from the committed copy of a reference config's ``roi_head`` / ``test_cfg.rcnn`` blocks
    (tests/golden/roi_head_cfgs.json, written by tests/golden/make_cfg_fixtures.py) -- the same blocks
    ``build_roi_head(load_config('configs/<name>'))`` reads where the reference tree is present."""
    import json
    import os
    from . import heads, pnp  # noqa: F401
    from .registry import build_head
    path = path or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'roi_head_cfgs.json')
    blk = json.load(open(path))[name]
    return build_head(dict(blk['roi_head']), test_cfg=ConfigDict(blk['test_cfg_rcnn']))


def build_roi_head(cfg):
    """Build ``MonoRUnRoIHead`` from a loaded reference config (``cfg.model.roi_head`` + ``cfg.test_cfg.rcnn``)."""
    from . import heads, pnp  # noqa: F401  (registration side effects, like `import monorun`)
    from .registry import build_head
    roi = dict(cfg['model']['roi_head'])
    test_cfg = cfg.get('test_cfg', {}).get('rcnn') if isinstance(cfg.get('test_cfg'), dict) else None
    return build_head(roi, test_cfg=ConfigDict(test_cfg) if test_cfg else None)
