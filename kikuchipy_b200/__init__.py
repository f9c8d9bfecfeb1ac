"""B200-native EBSD dictionary indexing behind kikuchipy's plugin surface.

Only the dictionary-indexing hot path is implemented: the
``SimilarityMetric`` classes, ``dictionary_indexing`` and ``orientation_similarity_map``, plus
the rows section 8f marks next: dictionary generation (``get_patterns``) and ``merge_crystal_maps``.
Everything numerical runs in ``libkdi.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/kdi.h``); importing this package does not need a GPU, calling it does.
"""

from ._lib import Context, KdiError, bind_to_gpu_numa_node, default_context
from .indexing import DictionaryIndexingResult, dictionary_indexing, orientation_similarity_map
from .similarity_metrics import (
    NormalizedCrossCorrelationMetric,
    NormalizedDotProductMetric,
    SimilarityMetric,
)
from .distributed import dictionary_indexing_sharded, gather_topk, shard_bounds
from .io_edax import load_edax_binary
from .io_h5ebsd import H5EBSDScan, load_h5ebsd, save_h5ebsd
from .io_nordif import NordifScan, load, load_nordif
from .io_oxford import load_oxford_binary
from .master_pattern import GeneratedDictionary, direction_cosines, get_patterns
from .merge_maps import MergedCrystalMap, merge_crystal_maps
from .preprocessing import (
    average_neighbour_patterns,
    preprocess,
    remove_dynamic_background,
    remove_static_background,
)
from .refinement import (
    Detector,
    RefinementResult,
    refine_orientation,
    refine_orientation_projection_center,
    refine_projection_center,
)

__all__ = [
    "Context",
    "Detector",
    "DictionaryIndexingResult",
    "GeneratedDictionary",
    "H5EBSDScan",
    "KdiError",
    "MergedCrystalMap",
    "NordifScan",
    "NormalizedCrossCorrelationMetric",
    "NormalizedDotProductMetric",
    "RefinementResult",
    "SimilarityMetric",
    "average_neighbour_patterns",
    "bind_to_gpu_numa_node",
    "default_context",
    "dictionary_indexing",
    "dictionary_indexing_sharded",
    "direction_cosines",
    "get_patterns",
    "load",
    "load_edax_binary",
    "load_h5ebsd",
    "load_nordif",
    "load_oxford_binary",
    "gather_topk",
    "merge_crystal_maps",
    "orientation_similarity_map",
    "preprocess",
    "refine_orientation",
    "refine_orientation_projection_center",
    "refine_projection_center",
    "remove_dynamic_background",
    "remove_static_background",
    "save_h5ebsd",
    "shard_bounds",
]
__version__ = "0.1.0"
