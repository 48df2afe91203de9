"""Import-path shim: ``anatomix.model.network.Unet`` and
``anatomix.model.load_from_hf.load_from_hf`` resolve to the B200-backed
implementations in `anatomix_b200` (boundary contract)."""
