"""CPU oracle for the OOD-DFQ fake-quantisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ood_dfq_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path runs on
hand-written sm_100a kernels and fails loudly when they are missing.

Parity status: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tools/make_golden.py`` imports ``/root/reference/quantization_utils``
in the build container and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays those vectors through every function
here (bit-exact for codes, dequantised values and range state).
"""
