"""CPU oracle for the DPT rollout hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a numpy / C restatement of the
reference's algorithm (titanium-47/decision-pretrained-transformer) used to
verify the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package (``decision-pretrained-transformer_b200``) never does, and
fails loudly when its CUDA extension is missing.

Parity pin: the reference ships no tests or golden vectors, so
the pin is created here -- ``oracle/make_golden.py`` runs the UNMODIFIED
reference from ``/root/reference`` (dev container only) under fixed seeds,
checks that this restatement reproduces its outputs bit-for-bit from the same
global ``np.random`` stream, and commits outputs + the consumed noise under
``tests/golden/``.
"""
