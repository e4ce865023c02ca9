"""CPU oracle -- test infrastructure only (see the header of each module)."""
