"""CPU oracle for the Accel (dff_deeplab) per-frame hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``accel_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may.  The product path fails loudly when its CUDA library is missing; it never falls
back to this code.

PARITY UNPINNED.  The reference (SamvitJ/Accel @ d1d7bb1) ships no tests, golden vectors or
fixtures for this path, and the arithmetic lives in Apache MXNet @ 62ecb60 (README.md:30,105-109),
which is not vendored and cannot be installed here (no mxnet, no python2, no network).  The oracle
therefore restates

* the reference's own graph wiring (``dff_deeplab/symbols/accel_{18,34,50,101}.py``,
  ``dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py``) and keyframe loops
  (``dff_deeplab/demo.py:165-250``, ``dff_deeplab/core/loader.py:259-303``,
  ``dff_deeplab/core/tester.py:158-171,246-256``) -- readable in /root/reference, and
* the published semantics of the MXNet operators those files call (``oracle/ops.py``; every
  function states the rule it assumes),

in plain PyTorch fp32 on the CPU.  It is anchored on the reference's call sites (parameter names,
layer hyper-parameters, output names) and on internal cross-checks (``tests/test_oracle_ops.py``:
hand-written samplers vs ``F.grid_sample``, deformable conv vs dilated conv / torchvision on
interior samples, closed-form bilinear kernels), not on outputs of the reference itself.
"""
