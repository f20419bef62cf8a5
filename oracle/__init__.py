"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the unet-design hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU baseline.  The
product package (``unet_design_b200``) never imports this package and raises
loudly when its CUDA extension is missing.

Parity status
-------------
* Haar DWT/iDWT (``haar_np``, ``haar_c.c``, ``pytorch_wavelets_restated``):
  **parity unpinned**.  The arithmetic lives in the third-party, un-vendored and
  un-pinned ``pytorch-wavelets`` package (requirements_diff_cifar.txt:41; filter
  bank from ``PyWavelets==1.3.0``, next line of the same file), which is absent
  from /root/reference and from this image.  The restatement follows its published
  algorithm (separable grouped strided convolution, taps ``[s,s]`` / ``[s,-s]``,
  ``s = fl32(1/sqrt 2)``, ``mode='zero'`` appends one zero at the end of an odd
  axis) and is anchored on the reference's own call sites and on algebraic
  identities (orthonormality, perfect reconstruction, ``LL/2 == avg_pool2d``).
* Conv / ResBlock / whole-model paths (``torch_ref``): pinned against golden
  vectors produced here by importing the reference's *own* module classes from
  /root/reference (``tools/make_golden.py``; fixtures in ``tests/golden``).
"""
