"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's PDHG hot path (the parity oracle).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`
may import anything from this package.  The product (`pdhg-optimal-control_b200/`) never does.
"""
