"""TEST INFRASTRUCTURE — builds the C half of the oracle (oracle/conv_ref.c) with gcc.

Output goes to oracle/_build/libmoe_oracle.so (git-ignored, travels to the GPU box with the
snapshot).  The reference itself is Python + PyTorch (no C sources on this path), so there is no
`oracle/_ref` binary to compile: "reference unbuildable as native code — it is interpreted Python",
see DESIGN.md §oracle.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libmoe_oracle.so')


def build(force=False):
  src = os.path.join(HERE, 'conv_ref.c')
  if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
    return LIB
  os.makedirs(OUT_DIR, exist_ok=True)
  # no -march=native: the .so is built in the container and must also run on the GPU box's host CPU.
  cmd = ['gcc', '-O3', '-mavx2', '-mfma', '-fopenmp', '-shared', '-fPIC', '-o', LIB, src]
  subprocess.run(cmd, check=True)
  return LIB


if __name__ == '__main__':
  print(build(force=True))
