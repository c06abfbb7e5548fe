"""Stand-in for torch_geometric 1.x -- only the symbols the reference imports."""
