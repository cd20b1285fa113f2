"""Defaults of the reference's constants.py that the scoring path reads (reference constants.py:5-6)."""
MC_DROPOUT_RATE = 0.25
MC_STEPS = 20
