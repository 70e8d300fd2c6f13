"""CPU oracle for the OOD-DFQ fake-quantisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ood_dfq_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path runs on
hand-written sm_100a kernels and fails loudly when they are missing.

Contents: fq_torch.py (op-by-op torch-CPU restatement; also the kind="port" CPU baseline), fq_numpy.py
(numpy float32), fq_oracle.c + fq_c.py (plain C, gcc -ffp-contract=off, built by build_c.py), bns_torch.py
(BN-statistics loss with autograd), fused_torch.py (eval-BN -> ReLU -> QuantAct chain for the fusion pass).

Parity status: the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``tools/make_golden.py`` imports ``/root/reference/quantization_utils``
in the build container and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays those vectors through every function
here (bit-exact for codes, dequantised values and range state).
"""
