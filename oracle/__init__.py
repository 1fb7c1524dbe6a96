"""ORACLE — CPU restatement of the reference's detect + locate path.  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
