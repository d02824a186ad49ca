"""commitment::injective_map (R/commitment/injective_map/mod.rs) -- re-export of the compressor over the Pedersen commitment."""
from .pedersen import PedersenCommCompressor  # noqa: F401
