"""CPU oracle for the smcdet hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product package ``smcdet_b200`` never does.
"""
