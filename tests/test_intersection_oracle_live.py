"""intersection ids against the LIVE reference on seeds that are not in the fixtures, FREE-RUNNING (no state injection
after the common seed): state, reward, flags, observation and the numpy generator words after every step.  Each case
runs in its own subprocess (tests/live_intersection_worker.py): IntersectionEnv._make_vehicles mutates IDMVehicle class
constants process-wide (envs/intersection_env.py:262-265), which is why the in-process live tests skip these ids."""
import os
import subprocess
import sys

import pytest

import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference not mounted")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env_id,obs,seed0,n", [
    ("intersection-v0", "default", 5000, 6), ("intersection-v0", "OccupancyGrid", 5100, 6),
    ("intersection-v2", "default", 5200, 4), ("intersection-multi-agent-v0", "default", 5300, 4),
])
def test_intersection_free_running_vs_live_reference(env_id, obs, seed0, n):
    out = subprocess.run([sys.executable, os.path.join(HERE, "live_intersection_worker.py"), env_id, obs, str(seed0), str(n)],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-1500:]
    last = out.stdout.strip().splitlines()[-1].split()
    assert last[0] == "OK" and int(last[1]) >= 3 * n, out.stdout[-300:]
