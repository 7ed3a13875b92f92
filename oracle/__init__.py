"""CPU oracle for the Semi-DETR hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``semi_detr_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` are allowed to import, link or execute code that lives here, and
only as the checker / CPU baseline -- never as the product path.

Each function cites the reference file:line (relative to /root/reference) whose
arithmetic it restates.  Pinning status of each part is stated in its header
and in DESIGN.md.
"""
