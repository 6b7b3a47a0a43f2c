"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements (plain torch / numpy, fp32) of the reference algorithms on the hot path, used as the parity
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package `renderih_b200` never imports anything from here.
"""
