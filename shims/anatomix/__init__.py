"""OPT-IN import-path shim (add ``<repo>/shims`` to ``sys.path`` yourself): ``anatomix.model.network.Unet``
and ``anatomix.model.load_from_hf.load_from_hf`` resolve to the B200-backed implementations in
`anatomix_b200` (boundary contract).

For machines WITHOUT the reference package only.  It is not on the repository root on purpose: there it
would shadow an installed reference ``anatomix`` (its ``registration`` / ``segmentation`` subpackages and
``model.vit3d`` would stop importing); with the reference installed use ``anatomix_b200.patch_reference()``."""
