"""bpmult_b200 -- B200-native (sm_100a) implementation of the BPMulT fusion trunk behind the reference's module API.

Public surface (mirrors /root/reference/bpmult/models): `MultiprojectionMMTransformer3DGMUClf` (= "mmtrvat"),
`TransformerEncoder`, `TransformerEncoderLayer`, `MultiheadAttention`, `SinusoidalPositionalEmbedding`, the GMU modules,
`get_model`, plus `Trainer` (graph-captured data-parallel training step).  Everything on the device runs in our own CUDA
kernels behind the C ABI in include/bpmult_b200.h (libbpmult_b200.so); there is no CPU fallback."""
__version__ = "0.1.0"


_PUBLIC = {
    "modules": ["MultiprojectionMMTransformer3DGMUClf", "MultiprojectionMMTransformerGMUClf", "TransformerEncoder", "TransformerEncoderLayer", "MultiheadAttention",
                "SinusoidalPositionalEmbedding", "GatedMultimodalLayer", "GatedMultimodalLayerFeatures", "TextShifting3Layer",
                "TextShifting4Layer", "TextShiftingNLayer", "AudioEncoder", "buffered_future_mask", "get_model", "MODELS", "manual_seed"],
    "trainer": ["Trainer"],
    "evaluate": ["model_eval", "multilabel_metrics", "weighted_acc"],
}


def __getattr__(name):            # lazy: importing the package must not require a GPU or the built library
    import importlib
    for mod, names in _PUBLIC.items():
        if name in names:
            return getattr(importlib.import_module("." + mod, __name__), name)
    raise AttributeError(name)
