"""TEST INFRASTRUCTURE.  CPU oracle for the hot path; see oracle/rd_oracle.py.
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm."""
