"""CPU oracle for the detect -> crop -> A2J-pose path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``handnet-pipeline_b200/``); only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

Every function restates, in plain torch-CPU / numpy, what the reference computes
and cites the reference file:line it follows.  The reference has no tests or
golden vectors of its own, so parity is pinned by
``oracle/make_golden.py``: it imports the real reference from /root/reference
with stub modules, runs it on seeded inputs, and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays the oracle against
those fixtures.
"""
