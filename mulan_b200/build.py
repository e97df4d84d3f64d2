"""In-tree build of libmulan_b200.so (nvcc, sm_100a only; no torch dependency).

The library is pure CUDA-runtime + C ABI, so it is compiled with nvcc directly rather than
through torch.utils.cpp_extension.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
LIB_PATH = PKG_DIR / 'libmulan_b200.so'

SOURCES = ['mulan_fwd_pre.cu', 'mulan_post.cu', 'mulan_bwd_pre.cu', 'mulan_aux.cu',
           'mulan_optim.cu', 'mulan_peer.cu', 'mulan_sampler.cu', 'mulan_rk45.cu', 'mulan_rng.cu', 'mulan_abi.cu',
           'mulan_host.cu', 'mulan_xla_legacy.cu']

NVCC_FLAGS = [
    '-O3', '-std=c++17', '--threads', '0',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo',
    # Plain a*b+c rounds twice like the reference's op-by-op float32; FMAs only where the
    # kernels spell fmaf() (see csrc/mulan_common.cuh).
    '-fmad=false',
    '-Xcompiler', '-fPIC', '-shared',
]


def _nvcc() -> str:
  nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not os.path.exists(nvcc):
    raise RuntimeError('nvcc not found; cannot build libmulan_b200.so')
  return nvcc


def _stale() -> bool:
  if not LIB_PATH.exists():
    return True
  built = LIB_PATH.stat().st_mtime
  deps = list(CSRC.glob('*.cu')) + list(CSRC.glob('*.cuh')) + list(CSRC.glob('*.h'))
  deps.append(PKG_DIR.parent / 'include' / 'mulan_b200.h')
  deps.append(PKG_DIR.parent / 'include' / 'mulan_b200_xla.h')
  deps.append(Path(__file__))
  return any(d.stat().st_mtime > built for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
  """Compile csrc/*.cu into mulan_b200/libmulan_b200.so if missing or out of date."""
  if not force and not _stale():
    return LIB_PATH
  cmd = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else [])
  cmd += ['-o', str(LIB_PATH)] + [str(CSRC / s) for s in SOURCES]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + res.stdout + res.stderr)
  if verbose:
    print(res.stderr)
  return LIB_PATH


if __name__ == '__main__':
  print(build_library(force=True, verbose=True))
