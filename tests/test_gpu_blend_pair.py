"""The CTA-pair blend kernel (HP3D_BLEND=pair, tcgen05.mma.cta_group::2) against the SMPL oracle. Runs in a subprocess
(the variant is fixed when a handle is created) under a timeout; every mbarrier wait in the kernel traps after ~2 s
instead of hanging. First run on a B200 in round 2: `profiles/r02a_unverified.log`."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_pair_blend_matches_oracle(built_lib):
    env = dict(os.environ, HP3D_BLEND="pair")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_blend_pair.py")], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
