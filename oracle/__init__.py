"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the legged_gym per-environment step.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.
The product path (``legged_games_gym_b200``) never imports this package and raises
if its CUDA extension is missing.
"""
