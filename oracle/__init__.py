"""CPU oracles (test infrastructure only): restatements of the reference algorithms, pinned by tests/golden/."""
