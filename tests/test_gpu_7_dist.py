"""On-GPU data-parallel invariant: N ranks on batch shards == 1 rank on the full batch,
through the REAL path - the CUDA kernels, the NCCL all-reduces of the packed BN statistics in forward and backward,
averaged parameter gradients (monkey-net_b200/dist_check.py).  Spawns `torch.distributed.run` with 2 ranks when the box
has at least 2 GPUs (gpurun --gpus 2); bench.py --gpus N runs the same check as a pre-flight and reports it."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc):
    env = dict(os.environ)
    env.setdefault('NCCL_DEBUG', 'WARN')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr',
           '127.0.0.1', '--master-port', '29611', os.path.join(ROOT, 'monkey-net_b200', 'dist_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith('DIST_CHECK ')]
    assert lines, (r.returncode, r.stdout[-2000:], r.stderr[-3000:])
    return r.returncode, json.loads(lines[0][len('DIST_CHECK '):])


def test_single_rank_self_consistency():
    """world = 1: the 'sharded' and the 'full' run are the same computation - the check itself must read clean."""
    rc, rep = _run(1)
    print(rep)
    assert rc == 0 and rep['ok'] and rep['world'] == 1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_two_ranks_equal_one_rank_on_the_full_batch():
    rc, rep = _run(2)
    print(rep)
    assert rep['world'] == 2
    assert rep['prediction_max_abs'] < 1e-4 and rep['kp_mean_max_abs'] < 1e-5 and rep['loss_terms_rel'] < 1e-4
    assert rep['bn_running_stats_rel'] < 1e-5
    assert rep['grad_cosine_median'] > 0.9999 and rep['grad_cosine_min'] > 0.99
    assert rc == 0 and rep['ok']
