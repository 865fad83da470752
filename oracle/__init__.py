"""CPU oracle for the per-element FEM elasticity hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the algorithm the reference
(otmanon/simkit @ 4e19c36, pure Python) runs for the path in SURVEY.md §8.  It
is the *checker* for the CUDA implementation in ``simkit_b200/``; it is never
the thing shipped or measured as the product:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing under ``simkit_b200/`` imports it, and ``simkit_b200`` raises if its
  CUDA library is missing instead of falling back to this code.

Parity pin: the reference is importable Python, so the oracle is pinned two
ways -- ``oracle/validate_against_reference.py`` compares every function here
with the reference itself (imported read-only from ``/root/reference``), and
``oracle/make_golden.py`` froze reference outputs into ``tests/golden/*.npz``
which ``tests/test_oracle_golden.py`` replays without the reference present.
The reference's own tests hold no golden vectors for this path.
"""
