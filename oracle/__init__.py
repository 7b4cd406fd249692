"""CPU oracle for the spectral_analyzer hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (no JVM in the container; the reference holds no golden vectors).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; the product (spectral_analyzer_b200/) never does.
"""
