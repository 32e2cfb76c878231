"""Byte-compile the UNMODIFIED reference into oracle/_ref/ (test infrastructure; never imported by the product).

The reference is pure Python, so "building" it means `py_compile`: every module on the hot path plus the three
drivers that call it is compiled FROM THE SOURCES WHERE THEY LIE under /root/reference into sourceless `.pyc` files
under oracle/_ref/ (git-ignored, NOT gpurun-ignored - it travels to the GPU box like the built .so).  No reference
source text is copied into the repository.  On the GPU box (where /root/reference does not exist) this lets
  * `bench.py --impl reference` time the reference's OWN modules on the host cores (cpu_baseline.kind = "reference"),
  * `tests/test_gpu_6_reference_drivers.py` execute the reference's own `train.py` / `transfer.py` /
    `reconstruction.py` against the drop-in `modules/` + `sync_batchnorm/`,
  * the oracle port be re-pinned against the live reference next to the GPU.
`__graft_entry__.build()` calls this when /root/reference is present; otherwise the prebuilt files are used.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
SRC = os.environ.get('MONKEY_REF_SRC', '/root/reference')

FILES = ['modules/util.py', 'modules/keypoint_detector.py', 'modules/movement_embedding.py',
         'modules/dense_motion_module.py', 'modules/generator.py', 'modules/discriminator.py', 'modules/losses.py',
         'sync_batchnorm/__init__.py', 'sync_batchnorm/batchnorm.py', 'sync_batchnorm/comm.py',
         'sync_batchnorm/replicate.py', 'train.py', 'transfer.py', 'reconstruction.py']


def build(force=False):
    """Returns the output directory, or None when the reference tree is not present (GPU box)."""
    if not os.path.isdir(os.path.join(SRC, 'modules')):
        return OUT if os.path.isdir(os.path.join(OUT, 'modules')) else None
    for rel in FILES:
        src = os.path.join(SRC, rel)
        dst = os.path.join(OUT, rel + 'c')
        if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: tracebacks name the reference-relative path, not this container's absolute one
        py_compile.compile(src, cfile=dst, dfile=os.path.join('<reference>', rel), doraise=True)
    with open(os.path.join(OUT, 'BUILD_INFO'), 'w') as f:
        f.write('py_compile of %s with python %s\n' % (SRC, sys.version.split()[0]))
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
