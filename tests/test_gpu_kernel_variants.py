"""The kernel families that lost their measurements (thread-per-block K1-T/K2-T, single-thread solo, tagged tables, ...)
and the product's own alternatives (half-warp matchers, K1-S2) stay under test: tests/variants_impl.py forces each of
them through the launcher's switches and compares with the oracle.  The losers only exist in the A/B build of the
library (lz4_flex_b200/liblz4b200_ab.so, -DLZ4B200_AB_VARIANTS), so the suite runs in a child process that loads that
library through LZ4B200_SO_OVERRIDE; the product library of this process is untouched."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_kernel_variant_matches_the_oracle():
    from lz4_flex_b200 import _native
    so = _native.build_ab()
    env = dict(os.environ, LZ4B200_SO_OVERRIDE=so)
    for k in list(env):
        if k.startswith("LZ4B200_") and k != "LZ4B200_SO_OVERRIDE":
            del env[k]
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "variants_impl.py"), "-q", "-p", "no:cacheprovider",
                        "--tb=short", "-m", "gpu"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-6000:] + r.stderr[-2000:])
