"""CPU oracle for the PowerGridworld ``MultiAgentEnv.step`` hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``powergridworld_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline.
"""
