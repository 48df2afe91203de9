"""``anatomix.model.load_from_hf`` import path -> `anatomix_b200.hf`."""
from anatomix_b200.hf import (ANATOMIX_VARIANTS, DEFAULT_REPO,
                              _load_handling_compile, load_from_file,
                              load_from_hf)

__all__ = ["ANATOMIX_VARIANTS", "DEFAULT_REPO", "load_from_hf",
           "load_from_file", "_load_handling_compile"]
