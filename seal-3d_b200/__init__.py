"""seal3d_b200 -- B200-native (sm_100a) hot path of Seal-3D: occupancy ray marching and
compositing, multiresolution hash-grid encoding, SH / frequency encoding, the fused sigma/colour
MLP and the teacher->student proxy mapping + distillation step, behind the reference's own
extension surface (``_raymarching``, ``_gridencoder``, ``_shencoder``, ``_freqencoder``, ``_ffmlp``).

Layout
  csrc/            CUDA kernels + the C-ABI (``include/seal3d_b200.h``) -> libseal3d_b200.so
  _lib.py          ctypes binding of the C-ABI (fails loudly when the library is missing)
  raymarching.py, gridencoder.py, shencoder.py, freqencoder.py, ffmlp.py
                   host-side mirrors of the reference's Python wrappers (same names / arguments)
  network.py, renderer.py, seal.py, trainer.py
                   NGP field, run_cuda / update_extra_state, bbox proxy, fused distillation step
  synth.py         deterministic synthetic workload (host only)
"""
__version__ = "0.1.0"
