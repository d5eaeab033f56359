"""ORACLE package: CPU restatements used ONLY as checkers by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Product code never imports this."""
