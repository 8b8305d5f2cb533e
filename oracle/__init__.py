"""CPU oracle for the BLSTM -> softmax -> CTC -> decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline.  The product (``mgr_b200`` = the hyphenated package directory) calls
hand-written sm_100a CUDA through the C ABI in ``include/gr_b200.h`` and raises
if that library is missing.

PARITY STATUS: *parity unpinned by reference fixtures*.  The reference
(/root/reference, Python-2 Keras 2.1.4 / TensorFlow 1.12.1 scripts) ships no
tests, golden vectors or weights and cannot be imported in this image (no
tensorflow / keras / python2), and the arithmetic lives in those un-vendored
third-party packages (requirements.txt:4, :8).  The oracle therefore restates
the *published* algorithms of those packages at the reference's call sites
(each function cites the call site it follows) and is pinned by independent
means instead (tests/test_oracle_*.py):
  * exhaustive CTC path enumeration on tiny cases,
  * fp64 finite differences for every gradient,
  * torch-CPU ``F.ctc_loss`` as an independent implementation,
  * hand-derived known answers,
  * literal-vs-closed-form property tests for ``decode_batch``,
  * exhaustive most-probable-labelling search for the beam decoder.
"""
