"""Developer tools: golden-fixture generators, GPU check / profiling scripts."""
