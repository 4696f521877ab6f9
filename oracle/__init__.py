"""CPU oracle (TEST INFRASTRUCTURE).  Import only from tests/, __graft_entry__.smoke() and bench.py."""
