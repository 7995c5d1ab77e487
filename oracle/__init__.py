"""ORACLE — test infrastructure only.  CPU restatement of the reference's PuTransE hot path, used
as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product (openke-putranse_b200/) never imports this package."""
