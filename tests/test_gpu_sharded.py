"""Sharded engine (one process per GPU, reference `-t N` semantics) against fixtures from the tapped reference run with
N threads: per-worker record streams and the merged tables must be bit-exact.  Needs >= N GPUs on the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


# sync: "device" = fqsk_sync_device (both barriers and the p-mer statistics on the device, no collective library per sync);
#       "host"   = the three-step form with the caller's all-reduce (NCCL) as the second barrier
#       "... grow" = smallest legal tables that double -- all shards of a table together -- several times during the run (FQSK_RESHARD)
@pytest.mark.parametrize("name,world,sync", [("se_orig_gs1_t2", 2, "device"), ("pe_orig_gs1_t2", 2, "device"), ("se_orig_gs16_t3", 3, "device"), ("se_orig_gs1_t2", 2, "host"),
                                             ("se_orig_gs1_t2", 2, "device grow"), ("pe_orig_gs1_t2", 2, "device grow"), ("se_orig_gs1_t2", 2, "host grow"),
                                             ("se_sorted_gs1_t2", 2, "device"), ("pe_sorted_gs1_t2", 2, "device")])      # the reference's default order (-om s) at -t 2
def test_sharded_engine_matches_reference_threads(name, world, sync):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py"), name, *sync.split()]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("bit-exact") == world, r.stdout[-2000:]
