"""Host-side mirror of the reference's ``quant.binary`` package (same names and signatures)."""
