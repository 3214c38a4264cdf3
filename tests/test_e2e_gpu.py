"""GPU test (-m gpu) of BASELINE.json configs[3]: kitti_multiclass end to end on synthetic frames with random-init
weights (tools/e2e_config4.py) -- shapes / validity through the whole 3-D branch and PnP-stage parity with the oracle
on the tensors captured at the head -> PnP boundary."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_config4_end_to_end(cuda_lib, tmp_path):
    out = tmp_path / 'e2e.json'
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'e2e_config4.py'), '--frames', '2', '--objects', '12',
                        '--steps', '2', '--out', str(out)], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    d = json.load(open(out))
    tf = d['teacher_forced']
    assert tf['valid_poses'] == 1.0 and tf['oracle_valid'] == 1.0
    assert tf['t_rel_err_vs_oracle_max'] < 1e-3 and tf['yaw_err_vs_oracle_max_rad'] < 1e-3
    assert d['ms_per_frame'] > 0 and d['objects_per_frame'] == 12
