"""ORACLE — CPU restatement of the reference's forward/detect path (test infrastructure only).

Nothing under maf_yolo_b200/ imports this package.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may use it — as the checker or the timed CPU
baseline, never as the product path.
"""
