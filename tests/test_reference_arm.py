"""``bench.py --impl reference`` end to end on CPU: the unmodified reference (baseline/_ref) is planned by its own ``clusterize``,
two provider processes train the 3-layer CNN through its own ``Node`` / ``Trainer`` over localhost gRPC, and one JSON line comes back.
(The GPU numbers of the arm are produced by the same code path with ``device='cuda'``.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_runs_the_unmodified_reference(tmp_path):
    if not os.path.isdir('/root/reference/ravnest') and not os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'ravnest')):
        pytest.skip('reference tree not available')
    env = dict(os.environ, TMPDIR=str(tmp_path), RVN_REF_TIMEOUT='400')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--model', 'cnn', '--ref-cpu', '--gpus', '1',
                          '--steps', '4', '--warmup', '3'], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=560)
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert lines, out.stderr[-2000:]
    res = json.loads(lines[-1])
    assert res.get('impl') == 'reference' and 'unavailable' not in res, res
    assert res['steps'] == 4 and res['value'] > 0 and res['gpu_launches'] == 0
    assert res['config']['parallelism'].startswith('pp2xdp1')
    record = json.load(open(os.path.join(ROOT, 'baseline', '_ref', 'INSTALL_RECORD.json')))
    assert record['source_digest'] == record['copy_digest']           # the vendored tree is byte-identical to the reference
