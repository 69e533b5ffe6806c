"""Parity tests: CPU suite (-m "not gpu") and B200 suite (-m gpu); fixtures under tests/golden/."""
