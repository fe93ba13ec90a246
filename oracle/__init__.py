"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the Polyffusion sampling hot path.

Nothing under ``oracle/`` is imported by the product package ``polyffusion_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it, as the checker or the
timed CPU baseline, never as the thing shipped.

Contents
--------
* ``unet_oracle`` / ``sampler_oracle``: a functional fp32 PyTorch-CPU restatement of the reference
  algorithm (each function cites the reference file:line it follows).
* ``reference_loader``: imports the *real* reference from ``/root/reference`` (build container only)
  through three stub modules (``shim/``), used to pin the restatement and to generate the golden
  vectors in ``tests/golden/`` (``make_golden.py``).

Parity pin status: the reference ships no tests or golden vectors, so the
pin is (a) direct comparison of the restatement with the imported reference modules in this
container (tests/test_oracle_vs_reference.py, skipped where /root/reference is absent) and
(b) golden outputs of the imported reference committed under tests/golden/.
"""
