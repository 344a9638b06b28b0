"""CPU oracle for the MESM per-pair inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mesm_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the CPU baseline.
"""
