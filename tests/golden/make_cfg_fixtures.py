"""Writes tests/golden/roi_head_cfgs.json: the ``model.roi_head`` block and ``test_cfg.rcnn`` of every shipped reference
config (configs/kitti_*.py), so that the GPU box -- where /root/reference does not exist -- builds MonoRUnRoIHead from
the reference's REAL blocks.  Run in the authoring container:  python tests/golden/make_cfg_fixtures.py"""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from monorun_b200.config import load_config  # noqa: E402

out = {}
for path in sorted(glob.glob('/root/reference/configs/kitti_*.py')):
    cfg = load_config(path)
    out[os.path.basename(path)] = dict(roi_head=cfg['model']['roi_head'], test_cfg_rcnn=cfg['test_cfg']['rcnn'])
json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'roi_head_cfgs.json'), 'w'), indent=1, sort_keys=True)
print({k: sorted(v['roi_head']) for k, v in out.items()})
