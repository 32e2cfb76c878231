"""bench.py's reference arm (the CPU leg the driver runs beside ours) prints ONE JSON line with the contract's keys;
under a multi-rank launch only rank 0 works.  The GPU arm cannot run here (no CUDA) and must say so loudly."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                          timeout=timeout, env=e)


def test_reference_arm_json_contract():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '1', '--batch', '4', '--config', 'shapes', '--res', '64'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['e2e']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    # the reference's own modules (source tree here, byte-compiled oracle/_ref on the GPU box); the port only as fallback
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    for k in ('metric', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'dtype', 'data', 'config'):
        assert k in d, k


def test_reference_arm_non_zero_rank_exits_quietly():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '1', '--config', 'shapes', '--res', '64'],
             env={'RANK': '1', 'WORLD_SIZE': '2'}, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_without_cuda_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--steps', '1', '--warmup', '1'], timeout=120)
    assert r.returncode != 0 and 'no CPU fallback' in r.stderr
