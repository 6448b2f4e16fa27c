"""Stand-in for torch-geometric 1.6.3 (only what the reference imports). Test infrastructure only."""
