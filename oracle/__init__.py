"""CPU oracle for the Accel (dff_deeplab) per-frame hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``accel_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may.  The product path fails loudly when its CUDA library is missing; it never falls
back to this code.

PARITY: WIRING PINNED, OPERATOR ARITHMETIC UNPINNED.  The reference (SamvitJ/Accel @ d1d7bb1) ships no tests,
golden vectors or fixtures for this path, and the arithmetic lives in Apache MXNet @ 62ecb60
(README.md:30,105-109), which is not vendored and cannot be installed here (no mxnet, no python2, no
network).  The oracle restates

* the reference's own graph wiring (``dff_deeplab/symbols/accel_{18,34,50,101}.py``,
  ``dff_deeplab/symbols/resnet_v1_101_flownet_deeplab.py``) and keyframe loops
  (``dff_deeplab/demo.py:165-250``, ``dff_deeplab/core/loader.py:259-303``,
  ``dff_deeplab/core/tester.py:158-171,246-256``), and
* the published semantics of the MXNet operators those files call (``oracle/ops.py``; every
  function states the rule it assumes, tagged [MXNet-ext]),

in plain PyTorch fp32 on the CPU.

What IS pinned against the reference's own code, by executing it in the build container:
* the wiring: ``oracle/mxstub.py`` stands in for ``mxnet.symbol`` so that the reference's symbol files build their
  key / cur graphs themselves; evaluated with ``oracle/ops.py`` they equal ``oracle/nets.py`` + ``oracle/schedule.py``
  BIT FOR BIT on every frame of the fixture clip, for all five versions, and the committed fixtures
  ``tests/golden/accel_<v>_128x256.npz`` are written from that run (``tests/golden/make_reference_wired.py``,
  ``tests/test_reference_wired.py``).  Layer order, parameter names and shapes, kernel / stride / pad / dilate /
  eps / fix_gamma / no_bias of every node and the output names are therefore the reference's, not a reading of it;
  the same script runs the frame loop of the reference's ``dff_deeplab/demo.py`` (:165-256) itself over those graphs
  and gets the same uint8 label maps;
* the host-side functions around the graphs (``transform``, ``fast_hist``, ``per_class_iu``, ``getpallete``,
  ``im_segment``, ``TestLoader.next/get_batch``, the greedy video -> GPU split, ``load_param``):
  ``tests/golden/make_reference_vectors.py`` executes them and ``tests/test_reference_vectors.py`` compares.

What is NOT pinned: the arithmetic inside each MXNet operator (convolution / deconvolution geometry, BatchNorm,
pooling conventions, GridGenerator + BilinearSampler, DeformableConvolution's border rule, argmax ties).  For those
the oracle relies on the published operator semantics and on internal cross-checks (``tests/test_oracle_ops.py``:
hand-written samplers vs ``F.grid_sample``, deformable conv vs dilated conv / torchvision on interior samples,
closed-form bilinear kernels).
"""
