"""Host-side multi-process logic of bench.py, world_size 2, gloo backend on the CPU (replicas
only, no data-path collective -- the process group carries the barrier and a max-reduction)."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_gloo_max_over_ranks_and_sharding():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--dist-selftest"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2
    assert out["ms"] == 20.0                      # the slowest rank defines the step time
    assert out["owned"] == [[0, 1, 2, 3], [4, 5, 6]]  # 7 units over 2 ranks: disjoint, complete, remainder first
    assert out["max_units"] == 4.0


def test_shard_of_covers_everything():
    sys.path.insert(0, ROOT)
    import bench

    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 8, 1024):
            got = [u for r in range(world) for u in bench.shard_of(r, world, n)]
            assert got == list(range(n))


def test_reference_arm_extra_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
