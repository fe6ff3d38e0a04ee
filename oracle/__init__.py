"""oracle/ — TEST INFRASTRUCTURE ONLY: CPU restatement (paged_ops.py) of the reference's decode path and
the reference's own CPU kernels compiled into oracle/_ref (build_ref.py, ref_lib.py).
Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference)."""
