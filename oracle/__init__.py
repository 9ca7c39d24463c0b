"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's Flash Checkpoint
serialisation algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.  The product
(dlrover_b200/) never does."""
