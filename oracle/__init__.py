"""CPU oracle for the FNO spectral-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline.  The product path
(``deno4pytorch_b200``) never imports this package and raises when its CUDA
library is missing.

Two independent restatements live here:

* ``fno_port``  - the reference's own algorithm (full ``rfftn`` -> keep corner
  modes -> per-mode complex ``einsum`` -> zero-padded ``irfftn`` -> add 1x1
  conv -> activation) written against ``torch`` CPU ops, differentiated by
  autograd.  Pinned against the real reference (imported from
  ``/root/reference`` in the build container) through the fixtures under
  ``tests/golden/``.
* ``dft_truth`` - the fp64 truncated-DFT "math contract" 
  with explicit adjoint formulas for every gradient, in numpy.  This is what
  the CUDA kernels actually implement, so it pins each kernel stage.
"""
