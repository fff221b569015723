# Bare `pytest` at the repo root: the tests live in tests/ only.  tiny-ram-halo2_b200/test_utils.py is the mirror of the reference's
# src/test_utils.rs (a library module of the package, imported through __graft_entry__.load_package), not a test file.
collect_ignore_glob = ["tiny-ram-halo2_b200/*", "oracle/*", "profiles/*", "rust/*", "gpurun_out/*", "bench.py", "__graft_entry__.py"]
