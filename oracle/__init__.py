"""
CPU oracle for the IMAGINE hot path (fields -> LOS maps -> ensemble log-likelihood).

TEST INFRASTRUCTURE ONLY.  Nothing under ``imagine_b200/`` imports this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / baseline.

Parity status per stage:
  * fields      (oracle/fields.py)     pinned by tutorial goldens
  * TestSimulator sampling (fields.py) pinned end-to-end by -55.581767201053125
  * likelihood  (oracle/likelihood.py) pinned against the reference's own code run
                                       unmodified in the build container through
                                       oracle/ref_shim.py -> tests/golden/*.npz
  * LOS         (oracle/los.py)        PARITY UNPINNED: the reference delegates this
                                       arithmetic to the external hammurabiX v2.4.2
                                       binary (not vendored, not installed).  The
                                       oracle *defines* the result; analytic closed
                                       forms pin it.
"""
