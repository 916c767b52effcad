"""rupphash_b200 -- B200-native (sm_100a) hot paths of phdupes (Safari77/rupphash).

Host-side mirror of the reference's interface for the two hot paths, forwarding to the
CUDA library through its C ABI (include/rupphash_b200.h):

    pdqhash      generate_pdq_features / generate_pdq / PdqFeatures   (src/pdqhash.rs)
    phash        DctPhash + bit-level dihedral operations             (src/phash.rs)
    hamminghash  hamming_distance / MIHIndex / find_groups            (src/hamminghash.rs)
    scanner      group_files_generic / group_with_pdqhash / batching  (src/scanner.rs:1146-1832)

Importing the package does not touch the GPU; the first call that needs the device loads
librupphash_b200.so and fails loudly if it (or a CUDA device) is missing.
"""
from ._lib import (Context, RupphashError, Unsupported, default_context, LAYOUT_RGB8, LAYOUT_RGBA8, LAYOUT_LUMA8,
                   MAX_SIMILARITY_64, MAX_SIMILARITY_256, PDQ_MIN_QUALITY)

__all__ = ["Context", "RupphashError", "Unsupported", "default_context", "LAYOUT_RGB8", "LAYOUT_RGBA8",
           "LAYOUT_LUMA8", "MAX_SIMILARITY_64", "MAX_SIMILARITY_256", "PDQ_MIN_QUALITY"]
__version__ = "0.1.0"
