"""Builds csrc/ -> lib/libssdk.so with nvcc for sm_100a (in-tree, so the .so travels to the GPU box)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False, extra=''):
    env = dict(os.environ)
    if extra:
        env['EXTRA'] = extra
    cmd = ['make', '-C', os.path.join(HERE, 'csrc'), '-j', str(min(8, os.cpu_count() or 1))]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError('building libssdk.so failed')
    return os.path.join(HERE, 'lib', 'libssdk.so')


if __name__ == '__main__':
    print(build(verbose=True))
