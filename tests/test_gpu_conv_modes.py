"""GPU: the optional conv modes (halo-tile A operand, split-K, CTA pairs) against the CUDA-core direct conv.
Both are selected by environment variables read once per process, so each mode runs in a subprocess."""
import json
import os
import subprocess
import sys

import pytest

from tests import fixtures as fx

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import json, sys
import rm_radar_b200 as rr
out = []
for sh in json.loads(sys.argv[1]):
    d, ref, _ = rr.conv_selftest(*sh, seed=7)
    out.append([d, ref])
print(json.dumps(out))
"""
# (n, h, w, cin, cout, k, stride, act, residual, out_f32)
HALO_SHAPES = [(1, 160, 160, 64, 64, 3, 1, 1, 0, 0), (1, 20, 20, 256, 256, 3, 1, 1, 1, 0), (3, 40, 40, 128, 128, 3, 1, 1, 1, 0),
               (2, 80, 80, 128, 64, 3, 1, 1, 0, 0), (1, 17, 23, 64, 48, 3, 1, 0, 0, 1)]
SPLIT_SHAPES = [(1, 20, 20, 512, 64, 3, 1, 1, 0, 0), (1, 20, 20, 256, 256, 3, 1, 1, 1, 0), (2, 10, 10, 512, 128, 3, 1, 1, 0, 0),
                (1, 20, 20, 1024, 512, 1, 1, 1, 0, 0), (1, 40, 40, 256, 256, 3, 2, 1, 0, 0)]


PAIR_SHAPES = [(1, 160, 160, 64, 64, 3, 1, 1, 0, 0), (1, 320, 320, 32, 64, 3, 2, 1, 0, 0), (3, 80, 80, 128, 12, 1, 1, 0, 0, 1),
               (8, 10, 10, 512, 512, 3, 1, 1, 1, 0), (1, 20, 20, 256, 256, 3, 1, 1, 1, 0), (5, 20, 20, 384, 384, 3, 2, 1, 0, 0),
               (1, 40, 40, 768, 256, 1, 1, 1, 0, 0)]


@pytest.mark.parametrize("env,shapes", [({"RMR_HALO": "1"}, HALO_SHAPES), ({"RMR_SPLITK": "1"}, SPLIT_SHAPES),
                                        ({"RMR_PAIR": "1"}, PAIR_SHAPES)],
                         ids=["halo", "split_k", "cta_pair"])
def test_optional_conv_modes_match_direct(env, shapes):
    e = dict(os.environ, **env)
    e["PYTHONPATH"] = fx.ROOT + os.pathsep + e.get("PYTHONPATH", "")
    out = subprocess.run([sys.executable, "-c", SCRIPT, json.dumps(shapes)], env=e, capture_output=True, text=True,
                         timeout=300, cwd=fx.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    for sh, (diff, ref) in zip(shapes, res):
        tol = (2e-3 if sh[9] else 6e-3) * max(1.0, ref)
        assert diff == diff and diff <= tol, f"{sh}: max|diff| {diff} > {tol}"
