"""CPU oracles for dualip_b200 (test infrastructure; never imported by the product package)."""
