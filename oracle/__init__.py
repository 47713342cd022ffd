"""CPU oracle for the xcorr cross-correlation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``xcorr_b200/`` may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` (``cpu_baseline`` / ``--impl reference``) use it, and
only as the checker / the timed CPU arm.

Parity pin: ``oracle.cc_oracle`` is checked against fixtures produced by
executing the reference's own ``xcorr/util/cc.py`` (see
``tests/golden/make_golden.py``); the reference ships no golden vectors of its
own (its ``tests/test_cc.py`` holds imports only).  The obspy-based
preprocessing restatement (``oracle.preprocess_oracle``) is **parity
unpinned**: obspy is neither installed nor vendored, so it restates the
published obspy 1.2-1.4 algorithms on scipy and is pinned only by analytic
known-answer tests.
"""
