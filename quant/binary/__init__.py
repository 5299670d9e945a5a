"""Drop-in for the reference package quant.binary."""
