"""moc_b200: B200-native (sm_100a) implementation of the MOC per-slide hot path.

The public names mirror the reference (xmed-lab/MOC): ``slide_process``, the four ``index_*_classifier``
selectors, ``topj_pooling`` and its variants, ``senet``, and the ``train`` / ``evaluation`` /
``zs_evaluation`` / ``ablation_evaluation`` loops.  Everything numeric runs in libmoc_b200.so (CUDA, C ABI in
include/moc_b200.h); importing the package does not need a GPU, calling an op without one raises.
"""
__version__ = "0.1.0"

_LAZY = {
    "slide_process": "slide", "senet": "model",
    "index_topj_classifier": "selectors", "index_delta_softmax_classifier": "selectors",
    "index_delta_diff_classifier": "selectors", "index_bottomk_irrel_classifier": "selectors",
    "topj_pooling": "pooling", "delta_softmax_classifier_pooling": "pooling",
    "delta_diff_classifier_pooling": "pooling", "bottomk_irrel_classifier_pooling": "pooling",
    "train": "loops", "evaluation": "loops", "zs_evaluation": "loops", "ablation_evaluation": "loops",
    "set_prompts": "loops", "main": "loops",
    "MocEngine": "engine", "RaggedBagStore": "bag_store", "BagDataset": "bag_store", "BagLoader": "bag_store",
    "MocError": "_lib",
    "Generic_MIL_Dataset": "datasets", "Generic_Split": "datasets", "Generic_WSI_Classification_Dataset": "datasets",
    "Conch_CLIP_Ada": "mil_heads", "CLAM_SB": "mil_heads", "MIL_fc": "mil_heads", "Attn_Net_Gated": "mil_heads",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError("module 'moc_b200' has no attribute %r" % name)
    import importlib
    return getattr(importlib.import_module("." + mod, __name__), name)
