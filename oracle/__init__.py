"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference path).

Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports it.
"""
