"""Drop-in mirrors of the reference's `models` namespace: models.mdgat.MDGAT and
models.superglue.SuperGlue (same constructor dict, state-dict layout and forward contract)."""
