"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's InfoNCE training step, used as the parity checker by
``tests/``, by ``__graft_entry__.smoke()`` and as ``bench.py``'s ``cpu_baseline`` / reference arm.
Nothing under ``cl-ica_b200/`` (the product) may import this package.

* ``lpnce_oracle.c``  double-precision pair-walking restatement of ``LpSimCLRLoss`` fwd + bwd
  (reference ``losses.py:443-477,506-510``), loaded through :mod:`oracle.c_oracle`.
* ``torch_port.py``   the same step written with the same torch formulation the reference executes
  (materialised B x M x d broadcast, ``nn.Linear``/``LeakyReLU`` stack, Adam) -- this is what is
  timed as the CPU baseline ("port").
* ``mlp_oracle.py``   numpy float64 forward/backward of the encoder stack.

Parity pin: the reference has no tests or golden vectors; every restatement here is pinned against
outputs of the reference itself (``tests/golden/*.npz``, produced by ``tests/golden/make_golden.py``
importing ``/root/reference``), see ``tests/test_oracle_vs_golden.py``.
"""
