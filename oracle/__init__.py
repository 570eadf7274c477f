"""CPU oracle for the AIS hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package is a plain-PyTorch (CPU) restatement of the algorithms on the hot
path of lollcat/fab-torch:

* ``oracle.realnvp``  -- the ``normflows`` RealNVP the reference builds in
  ``experiments/make_flow/make_normflow_model.py:11-30,82-96`` and calls through
  ``fab/wrappers/normflows.py:8-31``.  ``normflows`` is a third-party package
  (``requirements.txt:3``, **unpinned**, not vendored, not installed here, no
  network), so this part restates its published algorithm.
  **PARITY UNPINNED for the flow arithmetic**: the reference holds no numeric
  test for it (``fab/wrappers/normflow_test.py:28-34`` checks shapes only); the
  restatement is checked by self-consistency (inverse∘forward = id,
  log_prob(sample) = returned log_q, finite-difference log-det, autograd vs.
  analytic input-gradient) in ``tests/test_oracle_flow.py``.
* ``oracle.sampler``  -- ``Point``/``create_point``/γ/∇γ
  (``fab/sampling_methods/base.py``), ``AnnealedImportanceSampler``
  (``fab/sampling_methods/ais.py``), ``HamiltonianMonteCarlo`` and ``Metropolis``
  (``fab/sampling_methods/transition_operators/{hmc,metropolis}.py``).
  **PINNED**: ``oracle/gen_golden.py`` imports the reference itself from
  ``/root/reference`` (in the build container) and checks bit-for-bit equality
  of the restatement against it under identical seeds, then writes the golden
  fixtures in ``tests/golden``.
* ``oracle.targets``  -- ``ManyWellEnergy.log_prob`` / ``GMM.log_prob`` /
  ``effective_sample_size``.  PINNED the same way, plus the published log-Z
  constants.
* ``oracle.resample`` -- systematic resampling with a fixed-point CDF (a build
  extension; the reference only has multinomial ``resample``,
  ``fab/sampling_methods/base.py:121-124``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package, and only as the checker or
the timed CPU baseline.  Nothing under ``fab_torch_b200/`` imports it.
"""
