"""TEST INFRASTRUCTURE: CPU restatements of the reference algorithms (oracle) and the harness that runs the reference itself.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this package."""
